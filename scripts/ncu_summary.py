"""Print the key raw metrics of every kernel in an ncu report.  usage: python scripts/ncu_summary.py rep.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    print('----')
    for k in keys:
        if k in idx: print('  %-70s %s %s' % (k, r[idx[k]], rows[1][idx[k]]))
    st = sorted(((float(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:6]
    print('  top stalls (warps per issue):', ', '.join('%s=%.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in st))
