touch implicitpic1d_b200/csrc/particles.cu implicitpic1d_b200/csrc/abi.cu
make -s -C implicitpic1d_b200/csrc EXTRA="-DIPIC_PUSH_STATS" 2>&1 | grep -E " error"
python scripts/tail_probe.py 62500000 4 2>&1 | grep -v "^\[push stats\] transit\|general\|fresh" | tail -8
touch implicitpic1d_b200/csrc/particles.cu implicitpic1d_b200/csrc/abi.cu; make -s -C implicitpic1d_b200/csrc 2>&1 | grep -E " error"
