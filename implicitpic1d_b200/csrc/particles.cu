// Particle kernels of the B200 implicit-PIC library: the sub-stepping push (depositJ / moveParticles),
// rho deposition, moments, stable compaction, AoS<->SoA transposition, synthetic loader and cell sort.
//
// Reference semantics restated here (never copied): OpenMP/NGP/src/mod_particleMover.f90:21-270 and
// OpenMP/CIC/src/mod_particleMoverPicard.f90:17-317.
#include "ipic_common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <vector>
#include <algorithm>
#include <cstring>

namespace ipic {

#define FULL 0xffffffffu
#define RC_(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

// =====================================================================================================
// Shared-memory fixed-point grid: each cell is three 32-bit limbs (16/16/32-bit digits of the int64 sum)
// so that deposits are native 32-bit shared atomics (ATOMS.ADD) fed by native warp reductions (REDUX).
// Limb counters carry into the next limb when they pass 2^31, so capacity is unbounded.
// =====================================================================================================
struct LimbGrid {
    unsigned *g0, *g1;
    int *g2;
    unsigned long long *gacc;   // global int64 grid, receives top-limb spills and the final flush
};

__device__ __forceinline__ void limb2_add(const LimbGrid &g, int c, int r2) {
    int o = atomicAdd(&g.g2[c], r2);
    int nv = o + r2;
    if (nv >= (1 << 30) && o < (1 << 30)) {
        atomicSub(&g.g2[c], 1 << 30);
        atomicAdd(&g.gacc[c], (unsigned long long)(1ull << 62));
    } else if (nv <= -(1 << 30) && o > -(1 << 30)) {
        atomicAdd(&g.g2[c], 1 << 30);
        atomicAdd(&g.gacc[c], (unsigned long long)(0ull - (1ull << 62)));
    }
}
__device__ __forceinline__ void limb1_add(const LimbGrid &g, int c, unsigned r1) {
    unsigned o = atomicAdd(&g.g1[c], r1);
    if (o < 0x80000000u && o + r1 >= 0x80000000u) {
        atomicSub(&g.g1[c], 0x80000000u);
        limb2_add(g, c, 1 << 15);
    }
}
__device__ __forceinline__ void limb0_add(const LimbGrid &g, int c, unsigned r0) {
    unsigned o = atomicAdd(&g.g0[c], r0);
    if (o < 0x80000000u && o + r0 >= 0x80000000u) {
        atomicSub(&g.g0[c], 0x80000000u);
        limb1_add(g, c, 1u << 15);
    }
}
__device__ __forceinline__ void limb_add(const LimbGrid &g, int c, unsigned r0, unsigned r1, int r2) {
    limb0_add(g, c, r0);
    limb1_add(g, c, r1);
    limb2_add(g, c, r2);
}

// Warp-aggregated deposit of one fixed-point value per lane.  Must be called by all 32 lanes.
// Lanes that share the leader's cell are summed with REDUX and written once; after MAXPASS groups the
// remaining lanes fall back to their own atomics.
template <int MAXPASS>
__device__ __forceinline__ void warp_deposit(const LimbGrid &g, bool has, int cell, long long q, int lane) {
    unsigned pend = __ballot_sync(FULL, has);
    if (!pend) return;
    const unsigned u0 = (unsigned)q & 0xFFFFu, u1 = (unsigned)(q >> 16) & 0xFFFFu;
    const int s2 = (int)(q >> 32);
#pragma unroll
    for (int pass = 0; pass < MAXPASS; ++pass) {
        const int leader = __ffs(pend) - 1;
        const int key = __shfl_sync(FULL, cell, leader);
        const bool mine = has && (cell == key);
        const unsigned grp = __ballot_sync(FULL, mine);
        if (mine) {
            const unsigned r0 = __reduce_add_sync(grp, u0);
            const unsigned r1 = __reduce_add_sync(grp, u1);
            const int r2 = __reduce_add_sync(grp, s2);
            if (lane == leader) limb_add(g, key, r0, r1, r2);
        }
        has = has && !mine;
        pend &= ~grp;
        if (!pend) return;
    }
    if (has) limb_add(g, cell, u0, u1, s2);
}

// Two-point variant: qa goes to `cell`, qb to `cell + 1` (CIC current, NGP charge).
template <int MAXPASS>
__device__ __forceinline__ void warp_deposit2(const LimbGrid &g, bool has, int cell, long long qa, long long qb, int lane) {
    unsigned pend = __ballot_sync(FULL, has);
    if (!pend) return;
    const unsigned a0 = (unsigned)qa & 0xFFFFu, a1 = (unsigned)(qa >> 16) & 0xFFFFu;
    const int a2 = (int)(qa >> 32);
    const unsigned b0 = (unsigned)qb & 0xFFFFu, b1 = (unsigned)(qb >> 16) & 0xFFFFu;
    const int b2 = (int)(qb >> 32);
#pragma unroll
    for (int pass = 0; pass < MAXPASS; ++pass) {
        const int leader = __ffs(pend) - 1;
        const int key = __shfl_sync(FULL, cell, leader);
        const bool mine = has && (cell == key);
        const unsigned grp = __ballot_sync(FULL, mine);
        if (mine) {
            const unsigned ra0 = __reduce_add_sync(grp, a0), ra1 = __reduce_add_sync(grp, a1);
            const int ra2 = __reduce_add_sync(grp, a2);
            const unsigned rb0 = __reduce_add_sync(grp, b0), rb1 = __reduce_add_sync(grp, b1);
            const int rb2 = __reduce_add_sync(grp, b2);
            if (lane == leader) {
                limb_add(g, key, ra0, ra1, ra2);
                limb_add(g, key + 1, rb0, rb1, rb2);
            }
        }
        has = has && !mine;
        pend &= ~grp;
        if (!pend) return;
    }
    if (has) {
        limb_add(g, cell, a0, a1, a2);
        limb_add(g, cell + 1, b0, b1, b2);
    }
}

__device__ __forceinline__ void limb_grid_zero(const LimbGrid &g, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) { g.g0[i] = 0u; g.g1[i] = 0u; g.g2[i] = 0; }
}
// flush entries [1..n] (1-based cells) to the global int64 grid
__device__ __forceinline__ void limb_grid_flush(const LimbGrid &g, int n) {
    for (int i = threadIdx.x + 1; i <= n; i += blockDim.x) {
        unsigned long long v = (unsigned long long)g.g0[i] + ((unsigned long long)g.g1[i] << 16) +
                               ((unsigned long long)(long long)g.g2[i] << 32);
        if (v) atomicAdd(&g.gacc[i], v);
    }
}

// =====================================================================================================
// The push kernel (depositJ sweep / moveParticles).  Persistent CTAs, one per SM, 32 warps each.
//
// Work distribution: the particle array is cut into TILEs of 512 particles that warps claim from a global
// counter (one atomic per tile).  Cell-sorted particles make the cost of a tile depend on where it lies
// (sheath electrons take tens of sub-steps, bulk ions one), so a static partition would leave most of the
// GPU waiting for the few warps that drew the sheath; dynamic tiles keep every warp busy to the end.
//
// Streaming: when a warp claims a tile it prefetches the tile's 2 x 4 KB into L2 (one prefetch instruction
// per array: 32 lanes x one 128-byte line); it then reads the tile in sub-tiles of 32*SUB particles with
// coalesced, L1-bypassing loads issued one sub-tile ahead into registers, so the loop body never waits
// on HBM and the shared memory stays small enough to leave most of the 256 KB to L1 (per-cell tables).
//
// Divergence control: a batch is 32 particles, one per lane, and executes exactly ONE sub-step.
//   fresh batch : the cheap first part of the sub-step only (no square root, no divide).  A particle that
//                 stays in its cell for the whole del_t -- every ion, most bulk electrons at small del_t --
//                 is finished here.  One that reaches a cell face is pushed, unchanged, to the warp's
//                 continuation queue in shared memory.
//   queued batch: as soon as the queue holds 32 entries they are popped and run through the general
//                 sub-step with all lanes active.  Lanes with time left go back to the queue -- unless
//                 at least `stay_min` lanes continue, in which case the batch simply loops in registers
//                 (sheath electrons: tens of consecutive sub-steps without any queue traffic).
//
// Deposition (MOVE == false): deposits are 64-bit fixed-point integers (see ipic_common.cuh), so the sum
// is exact and independent of order, of the tile schedule and of the GPU count.  Every lane keeps
// register accumulators for the warp's current cell `base` and base+1 and a private shared-memory window
// [WIN][lane] for base-WLO .. base-WLO+WIN-1 (conflict-free, no atomics); anything further away goes
// straight to the global int64 grid with RED.64.  When `base` moves on (sorted input: about once per tile)
// the accumulators are summed across the warp with integer REDUX on 21-bit limbs and flushed with one
// RED.64 per cell.
//
// Runs of fully traversed cells (NGP): a fast electron that crosses k whole cells deposits exactly +-1.0
// (= +-2^44 in fixed point) in each of them.  Instead of k deposits the particle carries the first cell of
// the run; when the run ends, [a, b] += s is recorded as two entries of a difference grid, D[a] += s,
// D[b+1] -= s, whose integer prefix sum the field kernel adds back (exactly the same integers, so the
// result is bit-identical to depositing cell by cell).  A sheath electron therefore costs ~4 deposits
// per sweep, not one per sub-step.
//   SCHEME 0: NGP constant-field sub-step   (mod_particleMover.f90:21-63)
//   SCHEME 1: CIC Picard sub-step           (mod_particleMoverPicard.f90:17-65)
//   MOVE false: depositJ sweep (no particle writes, J accumulated)   true: moveParticles (write back)
// =====================================================================================================
#ifndef IPIC_SUB
#define IPIC_SUB 2
#endif
constexpr int SUB = IPIC_SUB;                       // batches per register-staged sub-tile
constexpr int SUBTILE = 32 * SUB;            // 64 particles
#ifndef IPIC_TILE_SUBS
#define IPIC_TILE_SUBS 8
#endif
constexpr int TILE_SUBS = IPIC_TILE_SUBS;   // sub-tiles per scheduled tile for loads that keep every warp busy: 512 particles = one 4 KB L2 prefetch per
                                            // array; small loads get smaller tiles (make_push_args: at least ~3 tiles per resident warp)
constexpr int TILE_MAX = SUBTILE * TILE_SUBS;
constexpr int WIN = 4, WLO = 2;              // window slots hold cells base-2, base-1, base+2, base+3 (base, base+1 live in registers)
constexpr int QCAP = 32 * (SUB + 1);         // continuation queue entries per warp: at most 31 + 32*SUB live

struct alignas(16) WarpSmem {                // per-warp shared memory: 4.4 KB
    unsigned long long win[WIN][32];
    double ql[QCAP], qv[QCAP], qt[QCAP];
    int qcell[QCAP];
    int qpend[QCAP];                         // pending run of fully traversed cells: +first cell (moving right), -first cell (moving left)
    unsigned qidx[QCAP];                     // global particle index; bit 31 = already reflected (MOVE)
};

// exact warp sum of one int64 per lane, one RED.64 to *dst.  Called by all 32 lanes.
__device__ __forceinline__ void warp_flush_value(unsigned long long v, unsigned long long *dst, int lane) {
    if (!__any_sync(FULL, v != 0ull)) return;
    const unsigned lo = __reduce_add_sync(FULL, (unsigned)v & 0x1FFFFFu);
    const unsigned mid = __reduce_add_sync(FULL, (unsigned)(v >> 21) & 0x1FFFFFu);
    const int hi = __reduce_add_sync(FULL, (int)((long long)v >> 42));
    if (lane == 0) {
        const unsigned long long t = (unsigned long long)lo + ((unsigned long long)mid << 21) + ((unsigned long long)(long long)hi << 42);
        if (t) atomicAdd(dst, t);
    }
}

struct DepositState {
    unsigned long long r0 = 0ull, r1 = 0ull;   // accumulators for cells base, base+1
    unsigned dirty = 0u;
    int base = 0;
    int *sd = nullptr;                         // this CTA's shared-memory copy of the difference grid, in units of FX_ONE (entry-stage kernels), or null
    int sd_cells = 0;                          // entries it holds (cells beyond go to the global grid)
};

__device__ __forceinline__ void deposit_flush(WarpSmem &W, DepositState &D, int lane, unsigned long long *gacc) {
    if (D.base > 0) {
        warp_flush_value(D.r0, gacc + D.base, lane);
        warp_flush_value(D.r1, gacc + D.base + 1, lane);
    }
    D.r0 = 0ull; D.r1 = 0ull;
    unsigned d = __reduce_or_sync(FULL, D.dirty);
    while (d) {
        const int k = __ffs(d) - 1;
        d &= d - 1;
        unsigned long long v = 0ull;
        if ((D.dirty >> k) & 1u) { v = W.win[k][lane]; W.win[k][lane] = 0ull; }
        warp_flush_value(v, gacc + (D.base - WLO + k), lane);
    }
    D.dirty = 0u;
}

__device__ __forceinline__ void deposit_add(WarpSmem &W, DepositState &D, int lane, unsigned long long *gacc, int cell, long long q) {
    const int k = cell - D.base;
    if (k == 0) D.r0 += (unsigned long long)q;
    else if (k == 1) D.r1 += (unsigned long long)q;
    else if ((unsigned)(k + WLO) < (unsigned)WIN) {
        W.win[k + WLO][lane] += (unsigned long long)q;
        D.dirty |= 1u << (k + WLO);
    } else {
        atomicAdd(gacc + cell, (unsigned long long)q);
    }
}

constexpr long long FX_ONE = 1ll << FX_BITS;

#ifdef IPIC_PUSH_STATS
#define IPIC_TRIPS_INC "add.u32 %6, %6, 1;\n\t"
#else
#define IPIC_TRIPS_INC
#endif
#ifdef IPIC_PUSH_STATS
// measurement build only (scripts/variant_bench.sh "-DIPIC_PUSH_STATS"): lane utilisation of the deposit sweeps, printed by ipic_destroy
__device__ unsigned long long g_push_stats[16];
#define PSTAT(i, x) do { st_[(i)] += (unsigned)(x); } while (0)
#else
#define PSTAT(i, x) do { } while (0)
#endif

// One end of a run of fully traversed cells: D[cell] += s * FX_ONE, s = +-1.  Run ends are scattered over the domain -- 32 lanes, 32 cells -- and
// a global RED per lane is 26 wavefronts of the LSU data pipe and one L2 atomic each; the kernels of electron-like species therefore count them
// in a per-CTA shared-memory copy of the difference grid (32-bit counts, native shared atomics, a few bank conflicts per warp) that the CTA adds
// to the global grid once, when all its warps are done.  Integer sums either way: the same grid bit for bit.
__device__ __forceinline__ void run_mark(const DepositState &D, unsigned long long *gdacc, int cell, int s) {
    if (D.sd != nullptr && cell < D.sd_cells) atomicAdd(D.sd + cell, s);
    else atomicAdd(gdacc + cell, (unsigned long long)((long long)s * FX_ONE));
}
// ends a run of fully traversed cells; `cend` is the cell the particle is in now (not part of the run)
__device__ __forceinline__ void close_run(WarpSmem &W, DepositState &D, int lane, unsigned long long *gacc, unsigned long long *gdacc, int pend, int cend) {
    if (pend > 0) {                                          // moving right: cells pend .. cend-1 got +1 each
        if (cend - pend == 1) deposit_add(W, D, lane, gacc, pend, FX_ONE);
        else {
            run_mark(D, gdacc, pend, 1);
            run_mark(D, gdacc, cend, -1);
        }
    } else {                                                 // moving left: cells cend+1 .. -pend got -1 each
        const int start = -pend;
        if (start - cend == 1) deposit_add(W, D, lane, gacc, start, -FX_ONE);
        else {
            run_mark(D, gdacc, cend + 1, -1);
            run_mark(D, gdacc, start + 1, 1);
        }
    }
}

// sqrt(s) for a normal positive operand: MUFU.RSQ64H seed (2^-20), one coupled Goldschmidt step on (g ~ sqrt(s), h ~ 1/(2 sqrt(s))) and the
// residual correction g1 + (s - g1^2) h1 -- 6 FP64 instructions, error 2^-81 before the final rounding, so the result is the correctly
// rounded root: bit-identical to sqrt() on 5e9 random and near-square operands (scripts/micro/transit_bench2.cu).  No special-operand
// path: s <= 0 gives NaN (callers select it away or rely on it).  The same sequence is inlined in the PTX of transit_loop.
__device__ __forceinline__ double lean_sqrt(double s) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double hs = __hiloint2double(__double2hiint(y) - 0x100000, 0);
    const double g = __dmul_rn(s, y);
    const double r = fma(-g, hs, 0.5);
    const double g1 = fma(g, r, g), h1 = fma(hs, r, hs);
    const double r1 = fma(-g1, g1, s);
    return fma(r1, h1, g1);
}

// The lean transit loop: face-to-face crossings of interior cells (NGP sub-step, mod_particleMover.f90:36-60, for a particle that sits on
// the entry face and moves inward).  Per crossing: s = u^2 + sdpe, u_f = sqrt(s), t_f = t + (u_f - u) * sinv, committed while t_f < t_lim,
// with u = |v| and {sdpe, sinv} = {+-2 a dx, +-1/a} read from the table of the lane's direction of travel (the table of the left-movers is
// stored mirrored and negated, so every lane walks its table upwards in 16-byte steps).  These are the general sub-step's own operations
// on the same operands: l_sub on the face makes diff_PE = +-two_a_dx exactly, sign flips are exact, and both paths use lean_sqrt and the
// fused time update, so a crossing gives the same bits whichever path runs it.
// The whole loop is one PTX block: 10 FP64 instructions per crossing (DMUL, DADD, 6 for the root, DADD, DFMA), an integer time test
// (positive doubles order like their bit patterns; NaN from a turning particle, a = 0 or a wall cell's NaN entry, and negative values,
// compare high and fail), and real branches instead of selects: the two state copies A and B alternate as source and destination, a lane
// that fails a test leaves the divergent region with its last committed copy intact (A after an even number of crossings, B after an odd
// one) and costs nothing further.  The loop runs while at least `tmin` lanes are in transit.  The table entries of the next two cells
// are loaded one crossing ahead.  u, t: in/out (state of the last committed crossing); off0: byte offset of the lane's first entry in
// `tab`; returns the number of committed crossings.
__device__ __forceinline__ int transit_loop(double &u, double &t, const double2 *tab, unsigned off0, unsigned long long tlim_bits, bool alive, unsigned tmin,
                                            bool &out, double &t_fail, unsigned &trips) {
    // off: byte offset of the lane's current entry in `tab` (a 32-bit add per committed crossing, one IMAD.WIDE per two loads)
    unsigned off = off0;
    double uB, tB;
    int dead;
    unsigned g_transit_trips_ = 0;       // measurement builds: passes of the two-crossing loop body
    asm volatile("{\n\t"
        ".reg .pred pd, pv;\n\t"
        ".reg .f64 s, y, hs, g, ng, r, g1, ng1, h1, r1, d, e0x, e0y, e1x, e1y;\n\t"
        ".reg .b32 lo, hi, hi2, m, c;\n\t"
        ".reg .b64 tb, ad;\n\t"
        "setp.eq.s32 pd, %9, 0;\n\t"                          // pd: the lane is out (not in transit, or a test failed)
        "mov.f64 e0x, 0d0000000000000000;\n\t"
        "mov.f64 e0y, 0d0000000000000000;\n\t"
        "mov.f64 e1x, 0d0000000000000000;\n\t"
        "mov.f64 e1y, 0d0000000000000000;\n\t"
        "mad.wide.u32 ad, %2, 1, %7;\n\t"
        "@!pd ld.global.nc.v2.f64 {e0x, e0y}, [ad];\n\t"
        "@!pd ld.global.nc.v2.f64 {e1x, e1y}, [ad+16];\n\t"
        "mov.f64 %3, %0;\n\t"
        "mov.f64 %4, %1;\n\t"
        "TLOOP:\n\t"
        IPIC_TRIPS_INC
        "@pd bra TSKIP;\n\t"
        // crossing: A -> B with e0
        "mul.rn.f64 s, %0, %0;\n\t"
        "add.rn.f64 s, s, e0x;\n\t"
        "rsqrt.approx.ftz.f64 y, s;\n\t"
        "mov.b64 {lo, hi}, y;\n\t"
        "add.s32 hi2, hi, -1048576;\n\t"
        "mov.b64 hs, {lo, hi2};\n\t"
        "mul.rn.f64 g, s, y;\n\t"
        "neg.f64 ng, g;\n\t"
        "fma.rn.f64 r, ng, hs, 0d3FE0000000000000;\n\t"
        "fma.rn.f64 g1, g, r, g;\n\t"
        "fma.rn.f64 h1, hs, r, hs;\n\t"
        "neg.f64 ng1, g1;\n\t"
        "fma.rn.f64 r1, ng1, g1, s;\n\t"
        "fma.rn.f64 %3, r1, h1, g1;\n\t"
        "sub.rn.f64 d, %3, %0;\n\t"
        "fma.rn.f64 %4, d, e0y, %1;\n\t"
        "mov.b64 tb, %4;\n\t"
        "setp.ge.u64 pd, tb, %8;\n\t"
        "@pd bra TSKIP;\n\t"
        "add.u32 %2, %2, 16;\n\t"
        "mad.wide.u32 ad, %2, 1, %7;\n\t"
        "ld.global.nc.v2.f64 {e0x, e0y}, [ad+16];\n\t"
        // crossing: B -> A with e1
        "mul.rn.f64 s, %3, %3;\n\t"
        "add.rn.f64 s, s, e1x;\n\t"
        "rsqrt.approx.ftz.f64 y, s;\n\t"
        "mov.b64 {lo, hi}, y;\n\t"
        "add.s32 hi2, hi, -1048576;\n\t"
        "mov.b64 hs, {lo, hi2};\n\t"
        "mul.rn.f64 g, s, y;\n\t"
        "neg.f64 ng, g;\n\t"
        "fma.rn.f64 r, ng, hs, 0d3FE0000000000000;\n\t"
        "fma.rn.f64 g1, g, r, g;\n\t"
        "fma.rn.f64 h1, hs, r, hs;\n\t"
        "neg.f64 ng1, g1;\n\t"
        "fma.rn.f64 r1, ng1, g1, s;\n\t"
        "fma.rn.f64 %0, r1, h1, g1;\n\t"
        "sub.rn.f64 d, %0, %3;\n\t"
        "fma.rn.f64 %1, d, e1y, %4;\n\t"
        "mov.b64 tb, %1;\n\t"
        "setp.ge.u64 pd, tb, %8;\n\t"
        "@pd bra TSKIP;\n\t"
        "add.u32 %2, %2, 16;\n\t"
        "ld.global.nc.v2.f64 {e1x, e1y}, [ad+32];\n\t"
        "TSKIP:\n\t"
        "vote.sync.ballot.b32 m, !pd, 0xffffffff;\n\t"
        "popc.b32 c, m;\n\t"
        "setp.ge.u32 pv, c, %10;\n\t"
        "@pv bra TLOOP;\n\t"
        "selp.s32 %5, 1, 0, pd;\n\t"
        "}"
        : "+d"(u), "+d"(t), "+r"(off), "=&d"(uB), "=&d"(tB), "=r"(dead), "+r"(g_transit_trips_)
        : "l"(tab), "l"(tlim_bits), "r"((int)alive), "r"(tmin)
        : "memory");
    const int k = (int)((off - off0) >> 4);
    // the copy that is not the committed one holds the attempt that failed its test
    if (k & 1) { t_fail = t; u = uB; t = tB; } else t_fail = tB;
    out = dead != 0;
#ifdef IPIC_PUSH_STATS
    trips = g_transit_trips_;
#endif
    return k;
}

// The whole sweep of one species by one warp: claims tiles until the species has none left, drains its queue, flushes its deposits
// and counters.  `A` is always a kernel parameter (constant bank), so the body can be instantiated for several species in one kernel.
template <int SCHEME, bool MOVE, int TS = TILE_SUBS, bool ENTRY = false>
__device__ __forceinline__ void push_body(const PushArgs &A, WarpSmem &W, const int lane, int *sd = nullptr, int sd_cells = 0) {
    DepositState D;
    D.sd = sd; D.sd_cells = sd_cells;
    // this CTA's copy of the deposition grids (see Ctx::d_acc)
    const size_t rep_off = MOVE ? 0 : (size_t)(blockIdx.x % (unsigned)A.acc_reps) * A.acc_rep_stride;
    unsigned long long *const gacc = A.acc + rep_off, *const gdacc = A.dacc + rep_off;
    if (!MOVE) {
#pragma unroll
        for (int k = 0; k < WIN; ++k) W.win[k][lane] = 0ull;
    }
    __syncwarp();

    const unsigned n = (unsigned)A.n;                       // n < 2^31 (checked on the host)
    constexpr int tile_subs = TS;                          // compile-time: 8 for loads that fill the GPU, 1 for small ones (make_push_args)
    constexpr unsigned TILE = (unsigned)SUBTILE * (unsigned)TS;
    const unsigned ntiles = (n + TILE - 1) / TILE;
    const double del_t = A.del_t, f_tol = A.f_tol;
    const int stay_min = A.stay_min;
    int qcount = 0;
    unsigned nsub = 0, nfev = 0, nabsL = 0, nabsR = 0, nrefl = 0;
#ifdef IPIC_PUSH_STATS
    unsigned st_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long st_t0_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(st_t0_));
#endif

    // tile claims run one ahead so that the atomic's latency is never exposed
    auto claim = [&]() -> unsigned {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(A.tile_counter, 1u);
        t = __shfl_sync(FULL, t, 0);
        // claims are mapped through a stride permutation so that the tiles in flight are spread over the whole
        // domain: expensive regions are met throughout the kernel and concurrent warps deposit to different cells
        if (A.tile_mode == 2) t = (t < ntiles) ? ((t & 1u) ? ntiles - 1u - (t >> 1) : (t >> 1)) : ntiles;      // from both ends inwards
        else t = (t < ntiles) ? (unsigned)(((unsigned long long)t * A.tile_stride) % ntiles) : ntiles;
        if (t < ntiles) {                                    // pull the tile into L2: 32 lanes x 128 B = 4 KB per array
            const size_t off = (size_t)t * TILE + 16u * lane;
            if (16u * (unsigned)lane < TILE) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A.l + off));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(A.vx + off));
            }
        }
        return t;
    };
    unsigned tile = claim();
    unsigned tile_next = (tile < ntiles) ? claim() : ntiles;
    int sub = 0;
    double nl[SUB], nv[SUB];                                // next sub-tile, in flight
    unsigned next_g0 = tile * (unsigned)TILE;               // global index of its first particle
    bool next_ok = tile < ntiles;
    auto issue_loads = [&]() {
#pragma unroll
        for (int b = 0; b < SUB; ++b) {
            const unsigned gi = next_g0 + 32u * b + lane;   // arrays are padded past n (ensure_capacity)
            nl[b] = __ldcs(A.l + gi);
            nv[b] = __ldcs(A.vx + gi);
        }
    };
    // The next sub-tile's loads are issued one sub-tile ahead into registers in the kernels whose batches are short (ions, the move: every
    // particle one sub-step; ion sweep 0.236 -> 0.213 ms at C5) and right before use in the entry-stage kernels (electrons: the four
    // staging registers cost the long batches more than the exposed L2 latency, 1.171 -> 1.177 ms).  IPIC_REG_PREFETCH = 0 / 1 forces one way.
#ifndef IPIC_REG_PREFETCH
#define IPIC_REG_PREFETCH 2
#endif
    constexpr bool REGPF = (IPIC_REG_PREFETCH == 1) || (IPIC_REG_PREFETCH == 2 && !ENTRY);
    if (REGPF && next_ok) issue_loads();

    // ---------------- the general sub-step: runs batches popped from the queue ----------------
    bool drain_all = !next_ok;
    for (;;) {
        while (qcount >= 32 || (drain_all && qcount > 0)) {
            const int take = min(qcount, 32);
            qcount -= take;
            PSTAT(6, 1);
            bool valid = lane < take;
            const int q = qcount + (valid ? lane : 0);
            double l = W.ql[q], v = W.qv[q], tp = W.qt[q];
            int cell = W.qcell[q];
            int pend = W.qpend[q];
            unsigned idx = W.qidx[q];
            __syncwarp();
            for (;;) {
                // ---- transit phase: face-to-face crossings of interior cells, the sub-step of fast (sheath) electrons.
                // A lane sitting on the entry face of an interior cell and moving inward is advanced face by face in a lean loop
                // that carries only (|v|, t, cell): with l_sub on the face, diff_PE = +-two_a_dx exactly, so the crossing speed
                // sqrt(v^2 +- 2 a dx) and time (v_f - v)/a are the general sub-step's own operations on the same operands (sign
                // flips are exact).  The crossing is committed only when it leaves more than f_tol of the step; every other outcome
                // (turns around, runs out of time in the cell or exactly on the face, a = 0, wall cells) leaves the lane untouched
                // for the general sub-step below.  "Reaches the far face within the time left" is decided from the crossing time
                // instead of from the end-of-step position l_f0 the reference tests first (mod_particleMover.f90:36-44): the same
                // decision except within the rounding of that formula itself.  ~2x fewer instructions than the general sub-step.
                bool requeue = false;
                if (SCHEME != 1) {
                    const bool vpos = v > 0.0;
                    const unsigned tr_span = (unsigned)max(A.NB - 3, 0);          // interior cells 2 .. N_cells-1
                    bool tr = valid && (v != 0.0) && (l == (double)(cell + (vpos ? 0 : 1))) && ((unsigned)(cell - 2) < tr_span);
                    if (!MOVE) tr = tr && (pend == 0 || ((pend > 0) == vpos));
                    if (__popc(__ballot_sync(FULL, tr)) >= A.transit_min) {
                        // right-movers walk the table of their species upwards from entry `cell`, left-movers the mirrored, negated copy
                        // (entry tab_cells - 1 - cell); a lane that is not in transit points at a valid entry and never loads
                        const unsigned toff = 16u * (unsigned)(tr ? (vpos ? cell : 2 * A.Nh + 3 - cell) : 0);
                        double u = fabs(v);
                        PSTAT(7, 1);
                        bool out; double t_fail;
                        unsigned trips = 0;
                        const int k = transit_loop(u, tp, A.ttab, toff, A.tlim_bits, tr, (unsigned)A.transit_min, out, t_fail, trips);
                        PSTAT(0, 2 * trips);
                        if (k) {
                            nsub += (unsigned)k;
                            PSTAT(1, k);
                            if (!MOVE && pend == 0) pend = vpos ? cell : -cell;
                            cell += vpos ? k : -k;
                            v = vpos ? u : -u;
                            l = (double)(cell + (vpos ? 0 : 1));
                        }
                        // a lane still in transit when the loop ended goes on as it is (back to the top in registers, or to the queue)
                        requeue = tr && !out;
                        // ---- finish stage: a lane whose next crossing would outlast the step ends its sweep inside this cell, almost
                        // always: the general sub-step's own first branch (stays in the cell, keeps its direction; mod_particleMover.f90:
                        // 36-47) on the same operands, everything else is left to the general sub-step below
#ifndef IPIC_FINISH_STAGE
#define IPIC_FINISH_STAGE 1
#endif
#ifdef IPIC_PUSH_STATS
                        if (tr && out && !(t_fail == t_fail)) atomicAdd(&g_push_stats[12], 1ull);      // left the loop on a NaN: turning point, wall cell, a = 0
                        if (tr && out && (t_fail == t_fail)) atomicAdd(&g_push_stats[13], 1ull);       // left the loop on the clock
#endif
                        // A lane that left the loop on a NaN inside an interior cell with a /= 0 cannot reach the face ahead (s = v^2 +- 2 a dx <= 0,
                        // i.e. the general sub-step's `diff_PE > -v^2` is false, re-checked below on its own operands): it either ends the step
                        // inside the cell (same formula as above, whatever the sign of v_f) or comes back to the face it entered through --
                        // the general sub-step's `bwd` branch with l on the face: diff_PE = 0, v_f = -v, dtau = -2 v / a -- and goes on
                        // from there into the neighbouring cell, in transit again.
                        const bool outx = tr && out;
                        const bool clockx = outx && (t_fail == t_fail);
                        const bool nanx = outx && !(t_fail == t_fail) && ((unsigned)(cell - 2) < tr_span);
                        bool fin = IPIC_FINISH_STAGE && (clockx || nanx);
                        bool back = false;
                        if (__any_sync(FULL, fin)) {
                            const double2 *tp2 = reinterpret_cast<const double2 *>(reinterpret_cast<const CellTabNGP *>(A.tab) + (fin ? cell : 0));
                            const double2 t0 = __ldg(tp2);
                            const double lc = (SCHEME == 2) ? (double)cell : 0.0;
                            const double lw = l - lc;
                            const double dtau = del_t - tp;
                            const double v_f0 = fma(t0.x, dtau, v);
                            const double l_f0 = fma(__dmul_rn(__dadd_rn(v, v_f0), dtau), t0.y, lw);
                            const bool inCell = (SCHEME == 2) ? ((int)(l_f0 + 1.0) == 1) : ((int)l_f0 == cell);
                            bool turnable = false;
                            if (__any_sync(FULL, fin && nanx)) {
                                const double2 t1 = __ldg(tp2 + 1);                         // two_a_dx, 1 / a
                                const double dfw = vpos ? t1.x : -t1.x;                    // diff_PE towards the face ahead: two_a_dx * (+-1)
                                turnable = fin && nanx && (t0.x != 0.0) && !(dfw > -__dmul_rn(v, v));
                                const double tpb = fma(__dsub_rn(-v, v), t1.y, tp);
                                back = turnable && !inCell && ((unsigned long long)__double_as_longlong(tpb) < A.tlim_bits);
                                if (back) {
                                    nsub += 1;
                                    PSTAT(3, 1);
                                    if (!MOVE && pend != 0) { close_run(W, D, lane, gacc, gdacc, pend, cell); pend = 0; }
                                    tp = tpb; v = -v; cell += vpos ? -1 : 1;               // l stays: the face between the two cells
                                }
                            }
                            fin = fin && inCell && ((clockx && (vpos == (v_f0 > 0.0))) || turnable);
                            if (fin) {
                                nsub += 1;
                                PSTAT(3, 1);
                                if (MOVE) {
                                    const unsigned gi = idx & 0x7fffffffu;
                                    A.l_out[gi] = l_f0 + lc;
                                    A.vx_out[gi] = v_f0;
                                } else {
                                    if (pend != 0) close_run(W, D, lane, gacc, gdacc, pend, cell);
                                    deposit_add(W, D, lane, gacc, cell, fx_from_double(l_f0 - lw));
                                }
                            }
                        }
                        requeue = requeue || back;
                        valid = valid && !fin;
                    }
                }
                bool cont = requeue;
                double l_f = l, v_f = v;
                if (__any_sync(FULL, valid && !requeue)) PSTAT(2, 1);
                if (valid && !requeue) {
                    nsub += 1;
                    PSTAT(3, 1);
                    double dtau = del_t - tp;
                    bool atB;
                    int lb = 0;
                    if (SCHEME != 1) {
                        // particleSubStepNoBField (mod_particleMover.f90:21-63 ; NGP_Periodic/src/mod_particleMover.f90:18-59),
                        // written flat: the three outcomes (reaches the face ahead / leaves through the face behind / stays in
                        // the cell) are selected, not branched on, so the common face-to-face transit runs without divergence.
                        // SCHEME 2 works in cell-local coordinates (l - l_cell in [0,1]) exactly as that tree does.
                        const double2 *tp2 = reinterpret_cast<const double2 *>(reinterpret_cast<const CellTabNGP *>(A.tab) + cell);
                        const double2 t0 = __ldg(tp2), t1 = __ldg(tp2 + 1);
                        const double a = t0.x, hinv = t0.y, two_a_dx = t1.x, inv_a = t1.y;
                        const double lc = (SCHEME == 2) ? (double)cell : 0.0;       // origin of the working coordinate
                        const double lw = l - lc;
                        const double v_f0 = fma(a, dtau, v);
                        const double l_f0 = fma(__dmul_rn(__dadd_rn(v, v_f0), dtau), hinv, lw);
                        const bool inCell = (SCHEME == 2) ? ((int)(l_f0 + 1.0) == 1) : ((int)l_f0 == cell);
                        const bool vpos = v > 0.0;
                        const bool eqSign = (vpos == (v_f0 > 0.0));
                        const double v2 = __dmul_rn(v, v);
                        lb = cell + (vpos ? 1 : 0);
                        double diff_PE = __dmul_rn(two_a_dx, ((double)lb - lc) - lw);
                        const bool maybeB = !(inCell && eqSign);
                        const bool fwd = maybeB && (diff_PE > -v2);
                        const bool bwd = maybeB && !fwd && !inCell;
                        atB = fwd || bwd;
                        double sgn = v;
                        if (bwd) {
                            lb = 2 * cell + 1 - lb;
                            diff_PE = __dmul_rn(two_a_dx, ((double)lb - lc) - lw);
                            sgn = -v;
                        }
                        v_f = v_f0;
                        double l_fw = l_f0;
                        if (atB) {
                            l_fw = (double)lb - lc;
                            v_f = copysign(lean_sqrt(__dadd_rn(diff_PE, v2)), sgn);
                        }
                        // time used: the crossing time (v_f - v) / a is added fused, here and in the transit loop
                        tp = atB ? fma(__dsub_rn(v_f, v), inv_a, tp) : __dadd_rn(tp, dtau);
                        l_f = l_fw + lc;
                        if (!MOVE) {
                            const double dl = l_fw - lw;
                            const bool full = atB && (fabs(dl) == 1.0);      // a whole cell AND moving on: part of a run
                            const bool right = dl > 0.0;
                            if (pend != 0 && !(full && ((pend > 0) == right))) { close_run(W, D, lane, gacc, gdacc, pend, cell); pend = 0; }
                            if (full) pend = (pend != 0) ? pend : (right ? cell : -cell);
                            else deposit_add(W, D, lane, gacc, cell, fx_from_double(dl));
                        }
                    } else {
                        const double2 *tp2 = reinterpret_cast<const double2 *>(reinterpret_cast<const CellTabCIC *>(A.tab) + cell);
                        const double2 t0 = __ldg(tp2), t1 = __ldg(tp2 + 1);
                        const double cap = __ldg(reinterpret_cast<const double *>(tp2 + 2));
                        const double A0 = t0.x, A1 = t0.y, inv_dx = t1.x, dx = t1.y;
                        dtau = fmin(dtau, cap);
                        const double ls = l - (double)cell;
                        const double r = dtau * inv_dx;
                        double l_prev = ls, lf = ls, d_half = ls, accel = 0.0, v_half = v;
                        int it;
                        atB = false;
                        for (it = 1; it <= 50; ++it) {
                            d_half = (ls + l_prev) * 0.5;
                            accel = fma(A1, d_half, A0);
                            v_half = fma(0.5 * accel, r, v);
                            lf = fma(v_half, r, ls);
                            atB = (lf <= 0.0) || (lf >= 1.0);
                            if (atB) lf = (v_half < 0.0) ? 0.0 : 1.0;
                            if (fabs(lf - l_prev) < 1.e-10) break;
                            l_prev = lf;
                        }
                        int numIter = it;   // 51 after an exhausted loop, as the Fortran loop variable
                        if (!atB) {
                            v_f = 2.0 * v_half - v;
                        } else {
                            if (lf != ls) {
                                v_f = copysign(sqrt(2.0 * accel * (lf - ls) + v * v), v_half);
                                dtau = (lf - ls) * dx / v_half;
                                numIter += 1;
                            } else {
                                dtau = -2.0 * v * dx / accel;
                                v_f = -v;
                            }
                            lb = (int)lf + cell;
                        }
                        l_f = lf + (double)cell;
                        if (MOVE) nfev += (unsigned)numIter;
                        if (!MOVE) {
                            const double Jp = l_f - l;
                            deposit_add(W, D, lane, gacc, cell, fx_from_double((1.0 - d_half) * Jp));
                            deposit_add(W, D, lane, gacc, cell + 1, fx_from_double(Jp * d_half));
                        }
                    }
                    bool absorbed = false;
                    if (atB) {
                        // boundaryConditions is zero except at the two walls (mod_domain.f90: walls only at index 1 and N)
                        if (lb != 1 && lb != A.NB) {
                            cell += signbit(v_f) ? -1 : 1;
                        } else {
                            const int b = (lb == 1) ? A.bcL : A.bcR;
                            if (!MOVE && SCHEME != 1 && b != 0 && pend != 0) {
                                // a run that ends on a wall includes the wall cell: the particle does not move on to a new cell
                                close_run(W, D, lane, gacc, gdacc, pend, cell + (pend > 0 ? 1 : -1));
                                pend = 0;
                            }
                            if (b == 0) {
                                cell += signbit(v_f) ? -1 : 1;
                            } else if (b == 1 || b == 4) {
                                absorbed = true;
                            } else if (b == 2) {
                                if (lb == A.NB) { if (v_f > 0.0) v_f = -v_f; }
                                else { if (v_f < 0.0) v_f = -v_f; }
                                if (MOVE && !(idx >> 31)) { idx |= 0x80000000u; nrefl += 1; if (A.refl) A.refl[idx & 0x7fffffffu] = 1; }
                            } else if (b == 3) {
                                l_f = (double)abs(lb - A.NB - 1);
                                cell = abs(cell - A.NB);
                            }
                        }
                    }
                    if (SCHEME == 1) tp = __dadd_rn(tp, dtau);
                    const double rem = del_t - tp;
                    if (absorbed) {
                        if (MOVE) {
                            // mark for the compaction pass; the wall side is recoverable from the marker
                            const bool left = (lb == 1);
                            if (left) nabsL += 1; else nabsR += 1;
                            const unsigned gi = idx & 0x7fffffffu;
                            A.l_out[gi] = left ? -1.0 : -2.0;
                            A.vx_out[gi] = v_f;
                            if (A.abs_list) {
                                const unsigned slot = atomicAdd(A.abs_count, 1u);
                                if (slot < A.abs_cap) A.abs_list[slot] = gi;
                            }
                        }
                    } else if (!(rem > f_tol)) {
                        if (MOVE) {
                            const unsigned gi = idx & 0x7fffffffu;
                            A.l_out[gi] = l_f;
                            A.vx_out[gi] = v_f;
                        }
                        if (!MOVE && SCHEME != 1 && pend != 0) close_run(W, D, lane, gacc, gdacc, pend, cell);
                    } else {
                        cont = true;
                    }
                }
                const unsigned cm = __ballot_sync(FULL, cont);
                if (!cm) break;
                if (__popc(cm) >= stay_min) {       // most of the batch goes on: loop in registers
                    valid = cont; l = l_f; v = v_f;
                    continue;
                }
                if (cont) {
                    const int qq = qcount + __popc(cm & ((1u << lane) - 1u));
                    W.ql[qq] = l_f; W.qv[qq] = v_f; W.qt[qq] = tp; W.qcell[qq] = cell; W.qpend[qq] = pend; W.qidx[qq] = idx;
                }
                qcount += __popc(cm);
                __syncwarp();
                break;
            }
        }
        if (drain_all) break;

        // ---------------- fresh batches of the next sub-tile ----------------
        // rotate the staged sub-tile into the working registers and put the following one in flight
        double cl[SUB], cv[SUB];
        if (!REGPF) issue_loads();                           // load right before use (the tile is already in L2)
#pragma unroll
        for (int b = 0; b < SUB; ++b) { cl[b] = nl[b]; cv[b] = nv[b]; }
        const unsigned g0 = next_g0;
        if (++sub == tile_subs) {
            sub = 0;
            tile = tile_next;
            if (tile < ntiles) tile_next = claim();
        }
        next_ok = tile < ntiles;
        next_g0 = tile * (unsigned)TILE + (unsigned)sub * SUBTILE;
        if (REGPF && next_ok) issue_loads();

#pragma unroll
        for (int b = 0; b < SUB; ++b) {
            const double l = cl[b], v = cv[b];
            const unsigned gi = g0 + 32u * b + lane;
            const bool valid = (gi < n) && (l >= 0.0);      // l < 0 marks a particle absorbed earlier and not yet compacted
            // l_cell = INT(l_sub + SIGN(1.d-12, v_sub)) ; NGP_Periodic: INT(l_sub), minus one for a left-mover sitting exactly on a
            // face (mod_particleMover.f90:84,91-95 there)
            int cell;
            if (SCHEME == 2) { cell = (int)l; if ((double)cell == l && signbit(v)) cell -= 1; }
            else cell = (int)(l + copysign(1.e-12, v));
            if (!MOVE && b == 0) {
                // follow the stream: re-centre the accumulators on the cell of the sub-tile's first particle
                const int nb = __shfl_sync(FULL, cell, 0);
                if (nb != D.base && nb > 0) {
                    deposit_flush(W, D, lane, gacc);
                    D.base = nb;
                }
            }
            bool cont = false;
            double e_l = l, e_v = v, e_t = 0.0;                 // what goes to the queue: the particle as it is, unless the entry stage moves it to a face
            int e_cell = cell, e_pend = 0;
            if (SCHEME != 1) {
                // Branch-free: the whole first (and for most particles only) sub-step in straight-line code.  An invalid lane reads table
                // entry 0 (zeros) and is masked out of every effect.  Sorted input puts a sub-tile inside one or two cells, so the deposit
                // is two warp-uniform cases on the register accumulators before the general per-lane routine.
                const int cellc = valid ? cell : 0;
                const double2 t0 = __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const CellTabNGP *>(A.tab) + cellc));
                const double lc = (SCHEME == 2) ? (double)cell : 0.0;
                const double lw = l - lc;
                const double v_f = fma(t0.x, del_t, v);
                const double l_fw = fma((v + v_f) * del_t, t0.y, lw);
                const bool inCell = (SCHEME == 2) ? ((int)(l_fw + 1.0) == 1) : ((int)l_fw == cell);
                const bool eqSign = (v > 0.0) == (v_f > 0.0);
                bool done = valid && inCell && eqSign;
                // A particle that turns round inside its cell (slow ions in any field, electrons near their turning point) is finished too
                // unless it can still reach the face ahead first (mod_particleMover.f90:39-47: FutureAtBoundaryBool = diff_PE > -v_i_sqr;
                // otherwise, being in the cell, it keeps the unconstrained l_f, v_f).  Same operands as the general sub-step below.
                const bool turn = valid && inCell && !eqSign;
                if (__any_sync(FULL, turn)) {
                    const double two_a_dx = __ldg(reinterpret_cast<const double *>(reinterpret_cast<const CellTabNGP *>(A.tab) + cellc) + 2);
                    const int lb = cell + ((v > 0.0) ? 1 : 0);
                    const double diff_PE = __dmul_rn(two_a_dx, ((double)lb - lc) - lw);
                    done = done || (turn && !(diff_PE > -__dmul_rn(v, v)));
                }
                cont = valid && !done;
                if (!MOVE) {
                    const unsigned long long q = done ? (unsigned long long)fx_from_double(l_fw - lw) : 0ull;
                    const int k = cell - D.base;
                    if (__all_sync(FULL, !done || k == 0)) D.r0 += q;
                    else if (__all_sync(FULL, !done || (unsigned)k < 2u)) { D.r0 += (k == 0) ? q : 0ull; D.r1 += (k == 1) ? q : 0ull; }
                    else if (done) deposit_add(W, D, lane, gacc, cell, (long long)q);
                } else if (done) { A.l_out[gi] = l_fw + lc; A.vx_out[gi] = v_f; }
                nsub += done ? 1u : 0u;
                PSTAT(4, 1); PSTAT(5, done ? 1 : 0);
                // ---- entry stage (ENTRY instantiation, chosen per species by the host): in an electron batch most lanes leave their cell, and
                // almost all of those through the face ahead.  When at least `entry_min` lanes go on, that first sub-step -- the general
                // sub-step's `fwd` branch (mod_particleMover.f90:36-60) on the same operands: diff_PE, sqrt(diff_PE + v^2), (v_f - v) / a --
                // is taken here, with the whole batch, and the lane is queued sitting on the face, ready for the transit loop.  Everything
                // else (face behind, wall cells, no time left) is queued as it is.
                if (ENTRY && __popc(__ballot_sync(FULL, cont)) >= A.entry_min) {
                    const double2 t1 = __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const CellTabNGP *>(A.tab) + cellc) + 1);
                    // Every lane that goes on leaves through the face ahead (`fwd`: diff_PE > -v^2) or, decelerated to a halt before it, through the
                    // face behind (`bwd`, mod_particleMover.f90:48-52: the other face, the sign of v flipped) -- in the un-converged iterates of a
                    // step, whose fields reverse a thermal electron within one cell, a fifth of all electrons.  Same formulas, other face.
                    const int lb_f = cell + ((v > 0.0) ? 1 : 0);
                    const double v2 = __dmul_rn(v, v);
                    const double dl_f = ((double)lb_f - lc) - lw;
                    const double dpe_f = __dmul_rn(t1.x, dl_f);
                    const bool fwd = dpe_f > -v2;
                    const int lb = fwd ? lb_f : 2 * cell + 1 - lb_f;
                    const double face = (double)lb - lc;
                    const double dl = fwd ? dl_f : face - lw;
                    const double diff_PE = fwd ? dpe_f : __dmul_rn(t1.x, dl);
                    const double vf2 = copysign(lean_sqrt(__dadd_rn(diff_PE, v2)), fwd ? v : -v);
                    const bool vpos = !signbit(vf2);                       // direction of travel behind the face
                    const double tpe = fma(__dsub_rn(vf2, v), t1.y, 0.0);
                    const bool go = cont && (lb != 1) && (lb != A.NB) && ((unsigned long long)__double_as_longlong(tpe) < A.tlim_bits);
                    if (go) {
                        nsub += 1;
                        PSTAT(3, 1);
                        e_l = face + lc; e_v = vf2; e_t = tpe; e_cell = cell + (vpos ? 1 : -1);
                    }
                    // ---- exit stage: a lane that has just crossed into the next cell ends its step there more often than not (a thermal electron
                    // crosses one or two faces per step once the fields have converged): the general sub-step's first branch -- stays in the cell
                    // and keeps its direction, mod_particleMover.f90:36-47 -- on the same operands as the finish stage behind the transit loop;
                    // such a lane is done here and never sees the queue.  Everything else is queued sitting on the face.
                    bool fin2 = false;
                    if (!MOVE && A.exit_stage && __any_sync(FULL, go)) {
                        const int c2 = go ? e_cell : 0;
                        const double2 u0 = __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const CellTabNGP *>(A.tab) + c2));
                        const double lw2 = e_l - ((SCHEME == 2) ? (double)c2 : 0.0);
                        const double dtau = del_t - tpe;
                        const double v_f0 = fma(u0.x, dtau, vf2);
                        const double l_f0 = fma(__dmul_rn(__dadd_rn(vf2, v_f0), dtau), u0.y, lw2);
                        const bool inCell2 = (SCHEME == 2) ? ((int)(l_f0 + 1.0) == 1) : ((int)l_f0 == c2);
                        fin2 = go && inCell2 && (vpos == (v_f0 > 0.0));
#ifdef IPIC_PUSH_STATS
                        if (lane == 0) { atomicAdd(&g_push_stats[14], (unsigned long long)__popc(__ballot_sync(FULL, fin2))); atomicAdd(&g_push_stats[15], (unsigned long long)__popc(__ballot_sync(FULL, go))); }
                        else { __ballot_sync(FULL, fin2); __ballot_sync(FULL, go); }
#endif
                        if (fin2) {
                            nsub += 1;
                            PSTAT(3, 1);
                            deposit_add(W, D, lane, gacc, cell, fx_from_double(dl));               // (a whole cell, +-1.0, is +-FX_ONE: the run of one cell)
                            deposit_add(W, D, lane, gacc, c2, fx_from_double(l_f0 - lw2));
                            cont = false;
                        }
                    }
                    if (go && !fin2 && !MOVE) {
                        if (fabs(dl) == 1.0) e_pend = (dl > 0.0) ? cell : -cell;                 // a whole cell: the start of a run
                        else deposit_add(W, D, lane, gacc, cell, fx_from_double(dl));
                    }
                }
            } else {
                // CIC: the Picard sub-step of a particle that stays inside its cell for the whole step (every ion, bulk electrons)
                bool done = false;
                unsigned long long qa = 0ull, qb = 0ull;             // the two-point deposit: cell, cell + 1
                if (valid) {
                    const double2 *tp2 = reinterpret_cast<const double2 *>(reinterpret_cast<const CellTabCIC *>(A.tab) + cell);
                    const double2 t0 = __ldg(tp2), t1 = __ldg(tp2 + 1);
                    const double cap = __ldg(reinterpret_cast<const double *>(tp2 + 2));
                    if (cap >= del_t) {                      // otherwise the sub-step cap splits the step: general path
                        const double ls = l - (double)cell;
                        const double r = del_t * t1.x;
                        double l_prev = ls, lf = ls, d_half = ls, v_half = v;
                        bool atB = false;
                        int it;
                        for (it = 1; it <= 50; ++it) {
                            d_half = (ls + l_prev) * 0.5;
                            const double accel = fma(t0.y, d_half, t0.x);
                            v_half = fma(0.5 * accel, r, v);
                            lf = fma(v_half, r, ls);
                            atB = (lf <= 0.0) || (lf >= 1.0);
                            if (atB) break;                  // reaches a face: the general path redoes this sub-step
                            if (fabs(lf - l_prev) < 1.e-10) break;
                            l_prev = lf;
                        }
                        if (!atB) {
                            done = true;
                            const double l_f = lf + (double)cell;
                            if (MOVE) {
                                nfev += (unsigned)it;
                                A.l_out[gi] = l_f; A.vx_out[gi] = 2.0 * v_half - v;
                            } else {
                                const double Jp = l_f - l;
                                qa = (unsigned long long)fx_from_double((1.0 - d_half) * Jp);
                                qb = (unsigned long long)fx_from_double(Jp * d_half);
                            }
                        }
                    }
                }
                if (!MOVE) {
                    // sorted input keeps a sub-tile inside the warp's current cell: both points go to the register accumulators without the
                    // per-lane routine (as the NGP fresh path does)
                    const int k = cell - D.base;
                    if (__all_sync(FULL, !done || k == 0)) { D.r0 += qa; D.r1 += qb; }
                    else if (done) {
                        deposit_add(W, D, lane, gacc, cell, (long long)qa);
                        deposit_add(W, D, lane, gacc, cell + 1, (long long)qb);
                    }
                }
                if (done) nsub += 1;
                cont = valid && !done;
            }
            const unsigned cm = __ballot_sync(FULL, cont);
            if (cm) {
                if (cont) {
                    const int qq = qcount + __popc(cm & ((1u << lane) - 1u));
                    if (ENTRY) { W.ql[qq] = e_l; W.qv[qq] = e_v; W.qt[qq] = e_t; W.qcell[qq] = e_cell; W.qpend[qq] = e_pend; W.qidx[qq] = gi; }
                    else { W.ql[qq] = l; W.qv[qq] = v; W.qt[qq] = 0.0; W.qcell[qq] = cell; W.qpend[qq] = 0; W.qidx[qq] = gi; }
                }
                qcount += __popc(cm);
                __syncwarp();
            }
        }
        drain_all = !next_ok;
    }

    // ---------------- epilogue: accumulators, counters ----------------
    if (!MOVE) deposit_flush(W, D, lane, gacc);
#ifdef IPIC_PUSH_STATS
    if (!MOVE) {
        // [0] transit crossings issued per warp (x32 = lane slots)  [1] committed lane crossings  [2] general sub-steps issued per warp
        // [3] valid lanes in them  [4] fresh batches per warp  [5] lanes finished in them  [6] queue pops  [7] transit loop entries
        {   // when the warps of this launch ran out of work: [8] sum of end times, [9] latest end, [10] earliest start, [11] warps (ns, globaltimer)
            unsigned long long tnow;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow));
            if (lane == 0) {
                atomicAdd(&g_push_stats[8], tnow - st_t0_); atomicMax(&g_push_stats[9], tnow - st_t0_); atomicAdd(&g_push_stats[11], 1ull);
            }
        }
        const int warp_only[8] = {1, 0, 1, 0, 1, 0, 1, 1};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned s_ = warp_only[k] ? st_[k] : __reduce_add_sync(FULL, st_[k]);
            if (lane == 0 && s_) atomicAdd(&g_push_stats[k], (unsigned long long)s_);
        }
    }
#endif
    {
        const unsigned s0 = __reduce_add_sync(FULL, nsub);
        if (lane == 0 && s0) atomicAdd(&A.counters[MOVE ? 5 : 0], (unsigned long long)s0);
        if (MOVE) {
            const unsigned s1 = __reduce_add_sync(FULL, nfev), s2 = __reduce_add_sync(FULL, nabsL), s3 = __reduce_add_sync(FULL, nabsR),
                           s4 = __reduce_add_sync(FULL, nrefl);
            if (lane == 0) {
                if (s1) atomicAdd(&A.counters[1], (unsigned long long)s1);
                if (s2) atomicAdd(&A.counters[2], (unsigned long long)s2);
                if (s3) atomicAdd(&A.counters[3], (unsigned long long)s3);
                if (s4) atomicAdd(&A.counters[4], (unsigned long long)s4);
            }
        }
    }
    __syncwarp();
}

template <int SCHEME, bool MOVE, int TS = TILE_SUBS, bool ENTRY = false>
__global__ void __launch_bounds__(PUSH_THREADS, 1) k_push(const __grid_constant__ PushArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // a sweep enqueued ahead of the nonlinear solve's decision: the field kernel in front of it ended the solve
    if (!MOVE && A.stop != nullptr && *A.stop != 0) return;
    constexpr bool SD = ENTRY && !MOVE;                    // per-CTA copy of the difference grid behind the warps' queues (run_mark)
    int *const sd = reinterpret_cast<int *>(smem_raw + PUSH_WARPS * sizeof(WarpSmem));
    const int sdn = SD ? A.sd_cells : 0;
    if (SD && sdn > 0) {
        for (int i = threadIdx.x; i < sdn; i += blockDim.x) sd[i] = 0;
        __syncthreads();
    }
    push_body<SCHEME, MOVE, TS, ENTRY>(A, reinterpret_cast<WarpSmem *>(smem_raw)[warp], lane, (SD && sdn > 0) ? sd : nullptr, sdn);
    if (SD && sdn > 0) {
        __syncthreads();
        unsigned long long *const gd = A.dacc + (size_t)(blockIdx.x % (unsigned)A.acc_reps) * A.acc_rep_stride;
        for (int i = threadIdx.x; i < sdn; i += blockDim.x) {
            const int v = sd[i];
            if (v) atomicAdd(gd + i, (unsigned long long)((long long)v * FX_ONE));
        }
    }
}

// depositJ sweep of TWO species in one launch.  Each warp works through one species' tiles and then through the other's; the first
// `warps_b_first` warps of every CTA start with species B.  A memory-bound species (ions: one sub-step per particle, limited by load
// latency) and an issue-bound one (electrons crossing many cells: FP64 work, HBM idle) then run side by side on every SM instead of one
// after the other; whichever runs out of tiles first simply donates its warps to the other.  Results are unchanged: the species keep
// their own grids, counters and tile counters, and deposition is order-independent integer arithmetic.
// Kept as an option (`pair_species`), off by default: on B200 the ion sweep turned out to be issue-bound as well (65 % issue
// utilisation), so the pairing only adds the two instruction streams -- measured 2.34 ms per sweep against 1.88 + 0.32 ms.
template <int SCHEME>
__global__ void __launch_bounds__(PUSH_THREADS, 1) k_push_pair(const __grid_constant__ PushArgs A, const __grid_constant__ PushArgs B, const int warps_b_first) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem &W = reinterpret_cast<WarpSmem *>(smem_raw)[warp];
    if (warp < warps_b_first) {
        push_body<SCHEME, false>(B, W, lane);
        push_body<SCHEME, false>(A, W, lane);
    } else {
        push_body<SCHEME, false>(A, W, lane);
        push_body<SCHEME, false>(B, W, lane);
    }
}

template <int SCHEME>
static int launch_push_pair_v(Ctx &c, const PushArgs &A, const PushArgs &B) {
    const size_t smem = (size_t)PUSH_WARPS * sizeof(WarpSmem);
    static bool configured[64] = {};
    if (!configured[c.device & 63]) {
        IPIC_CUDA(cudaFuncSetAttribute(k_push_pair<SCHEME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        IPIC_CUDA(cudaFuncSetAttribute(k_push_pair<SCHEME>, cudaFuncAttributePreferredSharedMemoryCarveout, 52));
        configured[c.device & 63] = true;
    }
    IPIC_CUDA(cudaMemsetAsync(A.tile_counter, 0, sizeof(unsigned), c.stream));
    IPIC_CUDA(cudaMemsetAsync(B.tile_counter, 0, sizeof(unsigned), c.stream));
    c.tile_zero_mask = 0;
    const int blocks = c.push_blocks > 0 ? c.push_blocks : c.num_sms;
    const int wb = std::min(PUSH_WARPS - 1, std::max(1, c.pair_warps));
    c.launches += 1; k_push_pair<SCHEME><<<blocks, PUSH_THREADS, smem, c.stream>>>(A, B, wb);
    IPIC_CUDA(cudaGetLastError());
    return IPIC_OK;
}

template <int SCHEME, bool MOVE, int TS = TILE_SUBS, bool ENTRY = false>
static int launch_push_v(Ctx &c, const PushArgs &A) {
    if (TS == TILE_SUBS && A.tile_subs == 1) return launch_push_v<SCHEME, MOVE, 1>(c, A);      // small load: 64-particle tiles
    // a species whose particles mostly leave their cell (electrons) takes the first sub-step inside the fresh path (push_body: entry stage)
    if (!ENTRY && !MOVE && SCHEME != 1 && A.entry_min <= 32) return launch_push_v<SCHEME, MOVE, TS, (SCHEME != 1 && !MOVE)>(c, A);
    const size_t smem = (size_t)PUSH_WARPS * sizeof(WarpSmem) + ((ENTRY && !MOVE) ? (size_t)A.sd_cells * sizeof(int) : 0);
    static bool configured[64] = {};          // function attributes are per device
    if (!configured[c.device & 63]) {
        IPIC_CUDA(cudaFuncSetAttribute(k_push<SCHEME, MOVE, TS, ENTRY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)((size_t)PUSH_WARPS * sizeof(WarpSmem) + (size_t)SD_CELLS_MAX * sizeof(int))));
        // leave the rest of the 256 KB to L1: the per-cell tables are gathered through it
        IPIC_CUDA(cudaFuncSetAttribute(k_push<SCHEME, MOVE, TS, ENTRY>, cudaFuncAttributePreferredSharedMemoryCarveout, 52));
        configured[c.device & 63] = true;
    }
    {
        const int sp = (int)((A.counters - c.d_counters) / 8);
        if (c.tile_zero_mask & (1u << sp)) c.tile_zero_mask &= ~(1u << sp);        // zeroed by the half-E op in front of this sweep
        else IPIC_CUDA(cudaMemsetAsync(A.tile_counter, 0, sizeof(unsigned), c.stream));
    }
    const int blocks = c.push_blocks > 0 ? c.push_blocks : c.num_sms;
    c.launches += 1; k_push<SCHEME, MOVE, TS, ENTRY><<<blocks, PUSH_THREADS, smem, c.stream>>>(A);
    IPIC_CUDA(cudaGetLastError());
    return IPIC_OK;
}

static int make_push_args(Ctx &c, int sp, double del_t, PushArgs &A) {
    SpeciesDev &s = c.sp[sp];
    if (s.n >= (1ll << 31)) { set_error("more than 2^31-1 particles of one species on one GPU"); return IPIC_ERR_UNSUPPORTED; }
    A = PushArgs{};
    A.l = s.l; A.vx = s.vx; A.l_out = s.l; A.vx_out = s.vx;
    A.n = s.n;
    const int cells = (c.scheme == IPIC_SCHEME_CIC) ? c.Nx : c.Nh;
    const size_t tab_stride = (c.scheme == IPIC_SCHEME_CIC) ? sizeof(CellTabCIC) : sizeof(CellTabNGP);
    A.tab = (const char *)c.d_tab + (size_t)sp * (cells + 2) * tab_stride;
    const size_t NM = (size_t)std::max(c.Nx, c.Nh) + 2;
    A.acc = c.d_acc + (size_t)sp * 2 * NM;
    A.dacc = A.acc + NM;
    A.acc_reps = std::max(1, c.acc_reps); A.acc_rep_stride = c.acc_rep_stride;
    {   // stride permutation of the tile order: an odd multiplier near the golden ratio, coprime with the tile count
        // tile size: 512 particles, or 64 when that would leave resident warps without work (the default deck's 7.7e5 particles per species are
        // 1500 tiles of 512 for 4736 warps: C1 electron sweep 0.117 -> 0.066 ms with 64-particle tiles; at 5e6 particles per species the
        // large tile is still the faster one)
        const unsigned long long warps = (unsigned long long)(c.push_blocks > 0 ? c.push_blocks : c.num_sms) * PUSH_WARPS;
        int subs = TILE_SUBS;
        if (2ull * (((unsigned long long)s.n + TILE_MAX - 1) / TILE_MAX) < 3ull * warps) subs = 1;     // fewer than 1.5 tiles of 512 per resident warp
        if (c.tile_subs > 0) subs = (c.tile_subs < TILE_SUBS) ? 1 : TILE_SUBS;
        if (c.pair_species && c.sp.size() == 2) subs = TILE_SUBS;       // k_push_pair exists for the large tile only
        A.tile_subs = subs;
        const unsigned long long TILE = (unsigned long long)SUBTILE * subs;
        const unsigned long long ntiles = ((unsigned long long)s.n + TILE - 1) / TILE;
        unsigned long long st = (unsigned long long)(0.6180339887498949 * (double)ntiles) | 1ull;
        auto gcd = [](unsigned long long a, unsigned long long b) { while (b) { unsigned long long t = a % b; a = b; b = t; } return a; };
        while (ntiles > 1 && gcd(st, ntiles) != 1) st += 2;
        A.tile_stride = (unsigned)(ntiles > 1 ? st % ntiles : 1);
        if (c.tile_order == 0 || A.tile_stride == 0) A.tile_stride = 1;
        A.tile_mode = c.tile_order;
    }
    A.counters = c.d_counters + (size_t)sp * 8;
    A.tile_counter = reinterpret_cast<unsigned *>(c.d_counters + (size_t)sp * 8 + 6);
    A.Nh = c.Nh;
    A.NB = (c.scheme == IPIC_SCHEME_CIC) ? c.Nh : c.Nx;
    A.bcL = c.h_bc[0]; A.bcR = c.h_bc[c.nbc - 1];
    A.del_t = del_t;
    // NGP :79,172 ; CIC :90,208 ; NGP_Periodic :75,150 (`del_t * 1e-12` with a default-real literal)
    A.f_tol = del_t * ((c.scheme == IPIC_SCHEME_NGP) ? 1.e-15 : (c.scheme == IPIC_SCHEME_CIC) ? 1.e-10 : (double)1e-12f);
    {   // `t < t_lim` <=> `del_t - t > f_tol` for every double t: the difference is monotone in t, so the first t that fails the
        // reference's test (mod_particleMover.f90:89,183) splits the number line exactly
        // (del_t = 0 is reachable through the reference's own step halving -- a converged JFNK solve whose linear-iteration count
        // exceeds maxIter leaves remainDel_t = 0, mod_nonLinSolvers.f90:297-313 -- so it must stay a valid, empty, sweep)
        double t = -INFINITY;
        if (A.del_t > 0.0 && std::isfinite(A.del_t)) {
            t = A.del_t - A.f_tol;
            while (A.del_t - t > A.f_tol) t = std::nextafter(t, INFINITY);
            while (!(A.del_t - std::nextafter(t, -INFINITY) > A.f_tol)) t = std::nextafter(t, -INFINITY);
        }
        A.t_lim = t;
        A.tlim_bits = 0ull;
        if (t > 0.0) std::memcpy(&A.tlim_bits, &t, sizeof(double));
    }
    A.ttab = c.d_ttab + (size_t)sp * 2 * (cells + 2);
    A.stay_min = std::min(32, std::max(1, c.stay_min));
    A.transit_min = std::min(33, std::max(1, c.transit_min));   // 33 disables the transit loop
    A.exit_stage = c.exit_stage;
    A.sd_cells = c.sd_runs ? std::min(A.NB + 2, SD_CELLS_MAX) : 0;
    // entry stage: for species whose particles leave their cell in a fair share of the sweeps -- measured on the previous step (sub-steps per
    // particle per sweep above 1.1: the stage itself only runs in batches with `entry_min` lanes going on, the instantiation costs the others
    // its registers), before the first step guessed from the charge-to-mass ratio (electron-like) -- unless the option switches it off (33)
    {
        const bool many = (s.sweep_substeps_per_particle > 0.0) ? (s.sweep_substeps_per_particle > 1.1) : (std::fabs(s.q_over_m) > 1.0e10);
        A.entry_min = (c.entry_min <= 32 && many) ? std::max(1, c.entry_min) : 33;
    }
    return IPIC_OK;
}

void push_stats_dump() {
#ifdef IPIC_PUSH_STATS
    unsigned long long h[16] = {};
    if (cudaMemcpyFromSymbol(h, g_push_stats, sizeof(h)) != cudaSuccess) return;
    fprintf(stderr, "[push stats] transit: %llu warp-crossings issued (%llu lane slots), %llu committed lane crossings -> utilisation %.3f; entries %llu\n",
            h[0], h[0] * 32, h[1], h[0] ? (double)h[1] / (double)(h[0] * 32) : 0.0, h[7]);
    fprintf(stderr, "[push stats] general sub-step: %llu issued per warp, %llu valid lanes -> utilisation %.3f; queue pops %llu\n", h[2], h[3],
            h[2] ? (double)h[3] / (double)(h[2] * 32) : 0.0, h[6]);
    fprintf(stderr, "[push stats] fresh: %llu batches, %llu lanes finished (%.3f of slots)\n", h[4], h[5], h[4] ? (double)h[5] / (double)(h[4] * 32) : 0.0);
    fprintf(stderr, "[push stats] entry stage: %llu lanes taken to the face ahead, %llu of them finished in the next cell (exit stage)\n", h[15], h[14]);
    fprintf(stderr, "[push stats] lanes leaving the transit loop: %llu on a NaN (turning point / wall cell / a = 0), %llu on the clock\n", h[12], h[13]);
    fprintf(stderr, "[push stats] warps: %llu; mean busy time per warp %.1f us; the latest warp of the LONGEST launch ran %.1f us (every launch's latest >= its mean: tail)\n", h[11],
            h[11] ? (double)h[8] / (double)h[11] * 1e-3 : 0.0, (double)h[9] * 1e-3);
    unsigned long long z[16] = {};
    cudaMemcpyToSymbol(g_push_stats, z, sizeof(z));
#endif
}

int launch_deposit(Ctx &c, int sp, double del_t, const int *stop) {
    PushArgs A;
    int rc = make_push_args(c, sp, del_t, A);
    if (rc) return rc;
    if (A.n == 0) return IPIC_OK;
    A.stop = stop;
    if (c.scheme == IPIC_SCHEME_NGP) return launch_push_v<0, false>(c, A);
    if (c.scheme == IPIC_SCHEME_NGP_PERIODIC) return launch_push_v<2, false>(c, A);
    return launch_push_v<1, false>(c, A);
}

// both species of a two-species plasma in one launch (species `a` first for most warps, `b` first for c.pair_warps of them)
int launch_deposit_pair(Ctx &c, int a, int b, double del_t) {
    PushArgs A, B;
    int rc = make_push_args(c, a, del_t, A);
    if (rc) return rc;
    rc = make_push_args(c, b, del_t, B);
    if (rc) return rc;
    if (A.n == 0 || B.n == 0) {
        if (A.n) RC_(launch_deposit(c, a, del_t));
        if (B.n) RC_(launch_deposit(c, b, del_t));
        return IPIC_OK;
    }
    if (c.scheme == IPIC_SCHEME_NGP) return launch_push_pair_v<0>(c, A, B);
    if (c.scheme == IPIC_SCHEME_NGP_PERIODIC) return launch_push_pair_v<2>(c, A, B);
    return launch_push_pair_v<1>(c, A, B);
}

int launch_move(Ctx &c, int sp, double del_t) {
    PushArgs A;
    int rc = make_push_args(c, sp, del_t, A);
    if (rc) return rc;
    if (A.n == 0) return IPIC_OK;
    SpeciesDev &S = c.sp[sp];
    S.nRef = 0;
    if (A.bcL == 2 || A.bcR == 2) {          // Neumann wall: the mover records which particles it reflected (refRecordIdx)
        if (S.refl_cap < S.cap) {
            if (S.refl) cudaFree(S.refl);
            IPIC_CUDA(cudaMalloc(&S.refl, (size_t)S.cap));
            S.refl_cap = S.cap;
        }
        IPIC_CUDA(cudaMemsetAsync(S.refl, 0, (size_t)A.n, c.stream));
        A.refl = S.refl;
    }
    // arrays without slab identity and without recorded reflections are compacted by filling (compact_species): the kernel lists what it absorbs
    S.abs_listed = false;
    if (S.slab_off.empty() && !A.refl && c.fill_compaction && sp < NSP_MAX) {
        if (!c.d_abs) {
            IPIC_CUDA(cudaMalloc(&c.d_abs, (size_t)NSP_MAX * 2 * ABS_CAP * sizeof(unsigned)));
            IPIC_CUDA(cudaMalloc(&c.d_abs_count, NSP_MAX * sizeof(unsigned)));
            IPIC_CUDA(cudaMemsetAsync(c.d_abs_count, 0, NSP_MAX * sizeof(unsigned), c.stream));
        }
        A.abs_list = c.d_abs + (size_t)sp * 2 * ABS_CAP; A.abs_count = c.d_abs_count + sp; A.abs_cap = ABS_CAP;
        S.abs_listed = true;
    }
    if (c.scheme == IPIC_SCHEME_NGP) return launch_push_v<0, true>(c, A);
    if (c.scheme == IPIC_SCHEME_NGP_PERIODIC) return launch_push_v<2, true>(c, A);
    return launch_push_v<1, true>(c, A);
}

// =====================================================================================================
// rho deposition (interpolateParticleToNodes: NGP mod_Scheme.f90:50-65, CIC mod_Scheme.f90:21-66)
// =====================================================================================================
// DENS (CIC only): interpolateParticleToNodesDensity (CIC mod_Scheme.f90:68-105) -- the density diagnostic folds the weight behind a
// Dirichlet wall back into the wall cell, as at a Neumann wall, where the charge deposit subtracts it
template <int SCHEME, bool DENS = false>
__global__ void __launch_bounds__(1024, 1) k_deposit_rho(const double *__restrict__ lpos, long long n, const int *__restrict__ bc, int Nx,
                                                         unsigned long long *acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    LimbGrid G;
    const int ng = Nx + 2;
    G.g0 = reinterpret_cast<unsigned *>(smem_raw);
    G.g1 = G.g0 + ng;
    G.g2 = reinterpret_cast<int *>(G.g1 + ng);
    G.gacc = acc;
    limb_grid_zero(G, ng);
    __syncthreads();
    // four independent loads in flight per thread before the (warp-synchronous) deposits: one load per trip left the kernel latency-bound
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long nround = ((n + U * stride - 1) / (U * stride)) * (U * stride);
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < nround; i0 += U * stride) {
        double lv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const long long i = i0 + u * stride; lv[u] = (i < n) ? lpos[i] : -1.0; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool has = (lv[u] >= 0.0);
            const double l = has ? lv[u] : 0.0;
            const int lc = (int)l;
            const double d = l - (double)lc;
            if (SCHEME == 0) {
                warp_deposit2<2>(G, has, lc, fx_from_double(1.0 - d), fx_from_double(d), lane);
            } else {
                // quadratic-spline weights with wall folding
                double wc = -d * d + d + 0.5, wl = 0.5 * (1.0 - d) * (1.0 - d), wr = 0.5 * d * d;
                int il = lc - 1, ir = lc + 1;
                bool hl = has, hr = has;
                if (has) {
                    const int bv = bc[lc + 1] - bc[lc];
                    switch (bv) {
                    case 0: break;
                    case -1: case -4: wc += DENS ? wl : -wl; hl = false; break;    // Dirichlet to the left
                    case 1: case 4: wc += DENS ? wr : -wr; hr = false; break;      // Dirichlet to the right
                    case -2: wc += wl; hl = false; break;                          // Neumann to the left
                    case 2: wc += wr; hr = false; break;                           // Neumann to the right
                    case -3: il = Nx; break;                                        // periodic to the left
                    case 3: ir = 1; break;                                          // periodic to the right
                    }
                }
                warp_deposit<2>(G, has, lc, fx_from_double(wc), lane);
                warp_deposit<2>(G, hl, il, fx_from_double(wl), lane);
                warp_deposit<2>(G, hr, ir, fx_from_double(wr), lane);
            }
        }
    }
    __syncthreads();
    limb_grid_flush(G, Nx);
}

int launch_deposit_rho(Ctx &c, int sp, bool density) {
    SpeciesDev &s = c.sp[sp];
    if (s.n == 0) return IPIC_OK;
    if (c.Nx > MAX_GRID_SMEM) return IPIC_ERR_UNSUPPORTED;
    const size_t smem = (size_t)3 * (c.Nx + 2) * 4;
    unsigned long long *acc = c.d_acc + (size_t)sp * 2 * ((size_t)std::max(c.Nx, c.Nh) + 2);
    static bool cfg0s[64] = {}, cfg1s[64] = {};   // function attributes are per device
    bool &cfg0 = cfg0s[c.device & 63], &cfg1 = cfg1s[c.device & 63];
    if (c.scheme != IPIC_SCHEME_CIC) {
        if (!cfg0) { IPIC_CUDA(cudaFuncSetAttribute(k_deposit_rho<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (MAX_GRID_SMEM + 2) * 4)); cfg0 = true; }
        c.launches += 1; k_deposit_rho<0><<<c.num_sms, 1024, smem, c.stream>>>(s.l, s.n, c.d_bc, c.Nx, acc);
    } else if (!density) {
        if (!cfg1) { IPIC_CUDA(cudaFuncSetAttribute(k_deposit_rho<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (MAX_GRID_SMEM + 2) * 4)); cfg1 = true; }
        c.launches += 1; k_deposit_rho<1><<<c.num_sms, 1024, smem, c.stream>>>(s.l, s.n, c.d_bc, c.Nx, acc);
    } else {
        static bool cfg2s[64] = {};
        bool &cfg2 = cfg2s[c.device & 63];
        if (!cfg2) { IPIC_CUDA(cudaFuncSetAttribute(k_deposit_rho<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (MAX_GRID_SMEM + 2) * 4)); cfg2 = true; }
        c.launches += 1; k_deposit_rho<1, true><<<c.num_sms, 1024, smem, c.stream>>>(s.l, s.n, c.d_bc, c.Nx, acc);
    }
    IPIC_CUDA(cudaGetLastError());
    return IPIC_OK;
}

// one species' deposition grid (all CTA copies) -> doubles (sum of weights, before w_p); the grid is left zeroed
__global__ void k_acc_extract(unsigned long long *g, int reps, unsigned long long rep_stride, int n, double *out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long q = 0ull;
        for (int r = 0; r < reps; ++r) { q += g[(size_t)r * rep_stride + i + 1]; g[(size_t)r * rep_stride + i + 1] = 0ull; }
        out[i] = (double)(long long)q * FX_INV;
    }
}
int extract_species_grid(Ctx &c, int sp, int n, double *host_out) {
    unsigned long long *acc = c.d_acc + (size_t)sp * 2 * ((size_t)std::max(c.Nx, c.Nh) + 2);
    c.launches += 1; k_acc_extract<<<(n + 255) / 256, 256, 0, c.stream>>>(acc, c.acc_reps, (unsigned long long)c.acc_rep_stride, n, c.d_work);
    IPIC_CUDA(cudaGetLastError());
    IPIC_CUDA(cudaMemcpyAsync(host_out, c.d_work, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    return IPIC_OK;
}

// =====================================================================================================
// Deterministic particle reductions: fixed grid, fixed tree, fixed final order.
// =====================================================================================================
constexpr int RED_BLOCKS = 592, RED_THREADS = 256;

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *out /* [NV] per block */) {
    __shared__ double sh[NV][RED_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(FULL, x, o);
        if (lane == 0) sh[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0.0;
        for (int w = 0; w < RED_THREADS / 32; ++w) x += sh[threadIdx.x][w];
        out[threadIdx.x] = x;
    }
}

// sum(vx^2+vy^2+vz^2), sum vx, sum vy, sum vz over alive particles (mod_particle.f90:206-230)
__global__ void __launch_bounds__(RED_THREADS) k_moments(const double *__restrict__ l, const double *__restrict__ vx, const double *__restrict__ vy,
                                                         const double *__restrict__ vz, long long n, double *partial) {
    double a[4] = {0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (l[i] < 0.0) continue;
        const double x = vx[i], y = vy[i], z = vz[i];
        a[0] += x * x + y * y + z * z;
        a[1] += x; a[2] += y; a[3] += z;
    }
    block_reduce_store<4>(a, partial + (size_t)blockIdx.x * 4);
}
// fixed-order sum of per-block partials: each thread sums a strided subset, then a fixed tree over the block
template <int NV>
__global__ void __launch_bounds__(RED_THREADS) k_final_sum(const double *partial, int nblocks, double *out) {
    double a[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) a[k] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += RED_THREADS) {
#pragma unroll
        for (int k = 0; k < NV; ++k) a[k] += partial[(size_t)b * NV + k];
    }
    block_reduce_store<NV>(a, out);
}

int particle_moments(Ctx &c, int sp, double out4[4]) {
    SpeciesDev &s = c.sp[sp];
    double *partial = c.d_work;   // needs RED_BLOCKS*4 doubles: allocated with slack in ipic_create
    c.launches += 1; k_moments<<<RED_BLOCKS, RED_THREADS, 0, c.stream>>>(s.l, s.vx, s.vy, s.vz, s.n, partial);
    c.launches += 1; k_final_sum<4><<<1, RED_THREADS, 0, c.stream>>>(partial, RED_BLOCKS, c.d_scal);
    IPIC_CUDA(cudaGetLastError());
    IPIC_CUDA(cudaMemcpyAsync(c.h_scal, c.d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    for (int k = 0; k < 4; ++k) out4[k] = c.h_scal[k];
    return IPIC_OK;
}

// writeLocalTemperature (mod_particle.f90:251-283): per cell INT(l) < N_x, the sum of v^2 and the particle count.  Sorted input puts a
// warp inside one or two cells, so the lanes of a cell are combined with a match before one atomic per (warp, cell).
__global__ void __launch_bounds__(512) k_local_temperature(const double *__restrict__ l, const double *__restrict__ vx, const double *__restrict__ vy,
                                                           const double *__restrict__ vz, long long n, int Nx, int numCells, double *sumV2,
                                                           unsigned long long *count) {
    // Every CTA owns one contiguous range of the particle arrays and a private copy of the per-cell sums in shared memory: cell-sorted
    // input keeps a CTA inside a few cells, so the global atomics drop from one per (warp, cell) -- all running warps on the same few
    // addresses -- to one per (CTA, touched cell).
    extern __shared__ __align__(16) unsigned char lt_smem[];
    double *s_sum = reinterpret_cast<double *>(lt_smem);
    unsigned *s_cnt = reinterpret_cast<unsigned *>(s_sum + numCells);
    for (int k = threadIdx.x; k < numCells; k += blockDim.x) { s_sum[k] = 0.0; s_cnt[k] = 0u; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long chunk = ((n + gridDim.x - 1) / gridDim.x + 31) / 32 * 32;       // multiple of 32: whole warps stay inside one range
    const long long first = (long long)blockIdx.x * chunk, last = min(n, first + chunk);
    for (long long i = first + threadIdx.x; i - lane < last; i += blockDim.x) {
        int cell = 0;
        double e = 0.0;
        if (i < last) {
            // all four loads issued together (the velocities of a dead or out-of-range particle are read and dropped): one memory round
            // trip per trip instead of two
            const double ll = l[i], a = vx[i], b = vy[i], c = vz[i];
            if (ll >= 0.0) {
                const int idx = (int)ll;
                if (idx < Nx && idx >= 1 && idx <= numCells) { cell = idx; e = a * a + b * b + c * c; }
            }
        }
        const unsigned peers = __match_any_sync(FULL, cell);
        const int leader = __ffs(peers) - 1;
        double tot;
        if (peers == FULL) {                                // the usual case on sorted input: the whole warp in one cell, fixed xor tree
            tot = e;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL, tot, o);
        } else {
            tot = 0.0;                                      // fixed order within the group: ascending lane
            // the shuffles are synchronised over the GROUP (mask = peers): groups of different sizes run different trip counts
            for (unsigned m = peers; m; m &= m - 1) tot += __shfl_sync(peers, e, __ffs(m) - 1);
        }
        if (cell > 0 && lane == leader) { atomicAdd(s_sum + cell - 1, tot); atomicAdd(s_cnt + cell - 1, (unsigned)__popc(peers)); }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < numCells; k += blockDim.x) {
        const unsigned cn = s_cnt[k];
        if (cn) { atomicAdd(sumV2 + k, s_sum[k]); atomicAdd(count + k, (unsigned long long)cn); }
    }
}

int local_temperature(Ctx &c, int sp, int numCells, double *sumV2_host, long long *count_host) {
    SpeciesDev &s = c.sp[sp];
    if (numCells < 1 || numCells > std::max(c.Nx, c.Nh)) { set_error("local temperature: numCells out of range"); return IPIC_ERR_ARG; }
    double *d_sum = c.d_work;                                                    // numCells doubles + numCells counters
    unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(c.d_work + (std::max(c.Nx, c.Nh) + 2));
    IPIC_CUDA(cudaMemsetAsync(d_sum, 0, (size_t)numCells * sizeof(double), c.stream));
    IPIC_CUDA(cudaMemsetAsync(d_cnt, 0, (size_t)numCells * sizeof(unsigned long long), c.stream));
    if (s.n > 0) {
        const size_t smem = (size_t)numCells * (sizeof(double) + sizeof(unsigned));
        static bool cfgs[64] = {};
        if (!cfgs[c.device & 63]) {
            IPIC_CUDA(cudaFuncSetAttribute(k_local_temperature, cudaFuncAttributeMaxDynamicSharedMemorySize, (MAX_GRID_SMEM + 2) * 12));
            cfgs[c.device & 63] = true;
        }
        // a CTA's range must hold < 2^32 particles for the 32-bit shared counters: n < 2^32 per GPU is already required by the sort
        c.launches += 1;
        k_local_temperature<<<c.num_sms * (smem <= 56 * 1024 ? 4 : 2), 512, smem, c.stream>>>(s.l, s.vx, s.vy, s.vz, s.n, c.Nx, numCells, d_sum, d_cnt);
        IPIC_CUDA(cudaGetLastError());
    }
    IPIC_CUDA(cudaMemcpyAsync(sumV2_host, d_sum, (size_t)numCells * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaMemcpyAsync(count_host, d_cnt, (size_t)numCells * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    return IPIC_OK;
}

// =====================================================================================================
// Stable compaction after moveParticles (mod_particleMover.f90:252-256: survivors keep their order).
// Pass 1: per-tile survivor counts + wall-loss partial sums.  Pass 2: scan.  Pass 3: stable scatter.
// =====================================================================================================
constexpr int CT_THREADS = 256, CT_ITEMS = 16, CT_TILE = CT_THREADS * CT_ITEMS;

__global__ void __launch_bounds__(CT_THREADS) k_compact_count(const double *__restrict__ l, const double *__restrict__ vx, const double *__restrict__ vy,
                                                              const double *__restrict__ vz, long long n, unsigned *tileCount, double *lossPartial) {
    const long long base = (long long)blockIdx.x * CT_TILE;
    unsigned cnt = 0;
    double a[4] = {0, 0, 0, 0};   // energy L, energy R, momentum L, momentum R
    for (int k = 0; k < CT_ITEMS; ++k) {
        const long long i = base + (long long)k * CT_THREADS + threadIdx.x;
        if (i < n) {
            const double ll = l[i];
            if (ll >= 0.0) cnt += 1;
            else {
                // energyLoss += v_f**2 + SUM(v_y**2 + v_z**2) ; momentumLoss += v_f   (mod_particleMover.f90:206-215)
                const double vf = vx[i], y = vy[i], z = vz[i];
                const double en = vf * vf + (y * y + z * z);
                if (ll == -1.0) { a[0] += en; a[2] += vf; } else { a[1] += en; a[3] += vf; }
            }
        }
    }
    __shared__ unsigned shc[CT_THREADS / 32];
    const unsigned wc = __reduce_add_sync(FULL, cnt);
    if ((threadIdx.x & 31) == 0) shc[threadIdx.x >> 5] = wc;
    block_reduce_store<4>(a, lossPartial + (size_t)blockIdx.x * 4);   // has a __syncthreads inside
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < CT_THREADS / 32; ++w) t += shc[w];
        tileCount[blockIdx.x] = t;
    }
}

// exclusive scan of tile counts (single block), total to out[0]
__global__ void __launch_bounds__(1024) k_scan_tiles(const unsigned *tileCount, long long ntiles, long long *tileOffset, long long *total) {
    __shared__ long long sh[1024];
    const long long per = (ntiles + 1023) / 1024;
    const long long b = (long long)threadIdx.x * per, e = min(b + per, ntiles);
    long long s = 0;
    for (long long i = b; i < e; ++i) s += tileCount[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int t = 0; t < 1024; ++t) { const long long x = sh[t]; sh[t] = run; run += x; }
        *total = run;
    }
    __syncthreads();
    long long run = sh[threadIdx.x];
    for (long long i = b; i < e; ++i) { tileOffset[i] = run; run += tileCount[i]; }
}

__global__ void __launch_bounds__(CT_THREADS) k_compact_scatter(const double *__restrict__ l, const double *__restrict__ vx,
                                                                const double *__restrict__ vy, const double *__restrict__ vz, long long n,
                                                                const long long *__restrict__ tileOffset, double *lo, double *vxo, double *vyo,
                                                                double *vzo) {
    __shared__ unsigned warpBase[CT_THREADS / 32];
    __shared__ unsigned running;
    const long long base = (long long)blockIdx.x * CT_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const long long off = tileOffset[blockIdx.x];
    for (int k = 0; k < CT_ITEMS; ++k) {
        const long long i = base + (long long)k * CT_THREADS + threadIdx.x;
        double ll = -1.0;
        if (i < n) ll = l[i];
        const bool alive = ll >= 0.0;
        const unsigned m = __ballot_sync(FULL, alive);
        if (lane == 0) warpBase[warp] = __popc(m);
        __syncthreads();
        unsigned wb = 0, tot = 0;
        for (int w = 0; w < CT_THREADS / 32; ++w) { const unsigned x = warpBase[w]; if (w < warp) wb += x; tot += x; }
        const unsigned run = running;
        if (alive) {
            const long long dst = off + run + wb + __popc(m & ((1u << lane) - 1u));
            lo[dst] = ll; vxo[dst] = vx[i]; vyo[dst] = vy[i]; vzo[dst] = vz[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) running = run + tot;
        __syncthreads();
    }
}

// survivors before each slab boundary -> new slab offsets
__global__ void k_slab_offsets(const double *__restrict__ l, const long long *__restrict__ tileOffset, const long long *oldOff, long long *newOff, int nb,
                               long long n) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= nb) return;
    const long long o = oldOff[k];
    const long long tile = o / CT_TILE;
    const long long b = tile * CT_TILE;
    unsigned cnt = 0;
    for (long long i = b + lane; i < o; i += 32) if (i < n && l[i] >= 0.0) cnt += 1;
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0) {
        const long long ntiles = (n + CT_TILE - 1) / CT_TILE;
        newOff[k] = (tile < ntiles ? tileOffset[tile] : 0) + cnt;
    }
}

int ensure_capacity(Ctx &c, int sp, long long n) {
    SpeciesDev &s = c.sp[sp];
    // padded so that whole CHUNKs / tiles may be read past n
    const long long need = ((n + CT_TILE - 1) / CT_TILE + 1) * CT_TILE;
    if (need <= s.cap) return IPIC_OK;
    double *bankA = nullptr, *bankB = nullptr;
    IPIC_CUDA(cudaMalloc(&bankA, (size_t)need * 4 * sizeof(double)));
    IPIC_CUDA(cudaMalloc(&bankB, (size_t)need * 4 * sizeof(double)));
    IPIC_CUDA(cudaMemsetAsync(bankA, 0, (size_t)need * 4 * sizeof(double), c.stream));
    IPIC_CUDA(cudaMemsetAsync(bankB, 0, (size_t)need * 4 * sizeof(double), c.stream));
    if (s.l) {
        if (s.n > 0) {
            IPIC_CUDA(cudaMemcpyAsync(bankA, s.l, s.n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
            IPIC_CUDA(cudaMemcpyAsync(bankA + need, s.vx, s.n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
            IPIC_CUDA(cudaMemcpyAsync(bankA + 2 * need, s.vy, s.n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
            IPIC_CUDA(cudaMemcpyAsync(bankA + 3 * need, s.vz, s.n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
        }
        IPIC_CUDA(cudaStreamSynchronize(c.stream));
        // each bank is one allocation whose base is its l pointer
        cudaFree(s.l);
        cudaFree(s.l2);
    }
    s.l = bankA; s.vx = bankA + need; s.vy = bankA + 2 * need; s.vz = bankA + 3 * need;
    s.l2 = bankB; s.vx2 = bankB + need; s.vy2 = bankB + 2 * need; s.vz2 = bankB + 3 * need;
    s.cap = need; s.cap2 = need;
    return IPIC_OK;
}

static void swap_banks(SpeciesDev &s) {
    std::swap(s.l, s.l2); std::swap(s.vx, s.vx2); std::swap(s.vy, s.vy2); std::swap(s.vz, s.vz2);
}

// Compaction by filling: the holes among the first `total` slots take the live particles of the tail, both in ascending order (two
// stream compactions of indices, deterministic).  Touches the absorbed particles only -- a few thousand of 6e7 per step -- where the stable
// scatter rewrites all 64 B of every particle; the order of the survivors changes, so this is for arrays whose order is not the
// reference's slab order any more (device loader, after a cell sort) and without recorded reflections.
struct SelDead { const double *l; __device__ bool operator()(unsigned i) const { return l[i] < 0.0; } };
struct SelLive { const double *l; __device__ bool operator()(unsigned i) const { return l[i] >= 0.0; } };
__global__ void k_fill_holes(const unsigned *__restrict__ holes, const unsigned *__restrict__ live, const int *__restrict__ nsel, double *l, double *vx,
                             double *vy, double *vz) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nsel[0] || k >= nsel[1]) return;
    const unsigned d = holes[k], sfrom = live[k];
    l[d] = l[sfrom]; vx[d] = vx[sfrom]; vy[d] = vy[sfrom]; vz[d] = vz[sfrom];
}
static int compact_by_filling(Ctx &c, SpeciesDev &s, long long n, long long total) {
    const long long gone = n - total;
    if (c.cap_fill < 2 * gone + 64) {
        if (c.d_fill) cudaFree(c.d_fill);
        c.cap_fill = 2 * gone + gone / 2 + 4096;
        IPIC_CUDA(cudaMalloc(&c.d_fill, (size_t)c.cap_fill * sizeof(unsigned)));
    }
    unsigned *holes = c.d_fill, *live = c.d_fill + (c.cap_fill / 2);
    int *d_num = reinterpret_cast<int *>(c.d_total + 4);
    cub::CountingInputIterator<unsigned> front(0u), tail((unsigned)total);
    size_t t0 = 0, t1 = 0;
    cub::DeviceSelect::If(nullptr, t0, front, holes, d_num, (int)total, SelDead{s.l}, c.stream);
    cub::DeviceSelect::If(nullptr, t1, tail, live, d_num + 1, (int)gone, SelLive{s.l}, c.stream);
    const size_t tmp = std::max(t0, t1);
    if (tmp > c.cap_sort_tmp) {
        if (c.d_sort_tmp) cudaFree(c.d_sort_tmp);
        c.cap_sort_tmp = tmp + tmp / 8;
        IPIC_CUDA(cudaMalloc(&c.d_sort_tmp, c.cap_sort_tmp));
    }
    size_t tb = c.cap_sort_tmp;
    c.launches += 1; cub::DeviceSelect::If(c.d_sort_tmp, tb, front, holes, d_num, (int)total, SelDead{s.l}, c.stream);
    tb = c.cap_sort_tmp;
    c.launches += 1; cub::DeviceSelect::If(c.d_sort_tmp, tb, tail, live, d_num + 1, (int)gone, SelLive{s.l}, c.stream);
    c.launches += 1; k_fill_holes<<<(unsigned)((gone + 255) / 256), 256, 0, c.stream>>>(holes, live, d_num, s.l, s.vx, s.vy, s.vz);
    IPIC_CUDA(cudaGetLastError());
    s.n = total;
    return IPIC_OK;
}

// ---- compaction from the list of absorbed particles (arrays without slab identity) ----
// The move kernel appended the index of every particle it absorbed; sorted ascending, the list gives the loss sums in a fixed order and the
// holes to fill, without a pass over the arrays: the first H entries (index < total = n - gone) are the holes, the others are the dead
// slots of the tail [total, n), whose live slots -- found by rank -- move into the holes in ascending order, as compact_by_filling does.
__global__ void __launch_bounds__(RED_THREADS) k_absorbed_loss(const unsigned *__restrict__ idx, unsigned gone, const double *__restrict__ l,
                                                               const double *__restrict__ vx, const double *__restrict__ vy, const double *__restrict__ vz,
                                                               double *out4, unsigned *count_to_clear) {
    double a[4] = {0, 0, 0, 0};   // energy L, energy R, momentum L, momentum R
    const unsigned per = (gone + RED_THREADS - 1) / RED_THREADS;
    const unsigned b = threadIdx.x * per, e = min(b + per, gone);
    for (unsigned k = b; k < e; ++k) {
        const unsigned i = idx[k];
        // energyLoss += v_f**2 + SUM(v_y**2 + v_z**2) ; momentumLoss += v_f   (mod_particleMover.f90:206-215)
        const double vf = vx[i], y = vy[i], z = vz[i];
        const double en = vf * vf + (y * y + z * z);
        if (l[i] == -1.0) { a[0] += en; a[2] += vf; } else { a[1] += en; a[3] += vf; }
    }
    block_reduce_store<4>(a, out4);
    if (threadIdx.x == 0) *count_to_clear = 0u;       // ready for the next move
}
__device__ __forceinline__ unsigned upper_bound_u(const unsigned *a, unsigned n, unsigned x) {      // entries <= x
    unsigned lo = 0, hi = n;
    while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (a[mid] <= x) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void k_absorbed_fill(const unsigned *__restrict__ idx, unsigned gone, unsigned total, double *l, double *vx, double *vy, double *vz) {
    const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= gone) return;
    const unsigned H = total ? upper_bound_u(idx, gone, total - 1u) : 0u;      // holes: absorbed particles among the first `total` slots
    if (k >= H) return;
    // the (k+1)-th live slot of the tail: smallest x with (x + 1) - (dead slots of the tail up to total + x) = k + 1
    unsigned lo = 0, hi = gone - 1u;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        const unsigned live = (mid + 1u) - (upper_bound_u(idx, gone, total + mid) - H);
        if (live >= k + 1u) hi = mid; else lo = mid + 1u;
    }
    const unsigned d = idx[k], sfrom = total + lo;
    l[d] = l[sfrom]; vx[d] = vx[sfrom]; vy[d] = vy[sfrom]; vz[d] = vz[sfrom];
}
static int compact_from_list(Ctx &c, int sp, long long gone) {
    SpeciesDev &s = c.sp[sp];
    const long long n = s.n, total = n - gone;
    unsigned *raw = c.d_abs + (size_t)sp * 2 * ABS_CAP, *sorted = raw + ABS_CAP;
    size_t tmp = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp, raw, sorted, (int)gone, 0, 32, c.stream);
    if (tmp > c.cap_sort_tmp) {
        if (c.d_sort_tmp) cudaFree(c.d_sort_tmp);
        c.cap_sort_tmp = tmp + tmp / 8;
        IPIC_CUDA(cudaMalloc(&c.d_sort_tmp, c.cap_sort_tmp));
    }
    size_t tb = c.cap_sort_tmp;
    c.launches += 1; cub::DeviceRadixSort::SortKeys(c.d_sort_tmp, tb, raw, sorted, (int)gone, 0, 32, c.stream);
    c.launches += 1; k_absorbed_loss<<<1, RED_THREADS, 0, c.stream>>>(sorted, (unsigned)gone, s.l, s.vx, s.vy, s.vz, c.d_scal, c.d_abs_count + sp);
    c.launches += 1; k_absorbed_fill<<<(unsigned)((gone + 255) / 256), 256, 0, c.stream>>>(sorted, (unsigned)gone, (unsigned)total, s.l, s.vx, s.vy, s.vz);
    IPIC_CUDA(cudaGetLastError());
    IPIC_CUDA(cudaMemcpyAsync(c.h_scal, c.d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    s.stats.energyLoss[0] = c.h_scal[0]; s.stats.energyLoss[1] = c.h_scal[1];
    s.stats.momentumLoss[0] += c.h_scal[2]; s.stats.momentumLoss[1] += c.h_scal[3];
    s.n = total;
    return IPIC_OK;
}

// Runs after the move kernel.  Fills wallLoss/energyLoss/momentumLoss and removes absorbed particles.
int compact_species(Ctx &c, int sp) {
    SpeciesDev &s = c.sp[sp];
    const long long n = s.n;
    if (n == 0) return IPIC_OK;
    if (s.abs_listed) {
        // the move kernel listed what it absorbed (delIdx entries unless the list overflowed)
        s.abs_listed = false;
        const long long gone = s.stats.delIdx;
        if (gone > 0 && gone <= (long long)ABS_CAP && gone * 8 < n && s.nRef == 0 && n < (1ll << 31)) return compact_from_list(c, sp, gone);
        IPIC_CUDA(cudaMemsetAsync(c.d_abs_count + sp, 0, sizeof(unsigned), c.stream));      // not used: the array is counted below
    }
    const long long ntiles = (n + CT_TILE - 1) / CT_TILE;
    if (ntiles > c.cap_tiles) {
        if (c.d_tileCount) { cudaFree(c.d_tileCount); cudaFree(c.d_tileOffset); cudaFree(c.d_loss); }
        c.cap_tiles = ntiles + ntiles / 4 + 16;
        IPIC_CUDA(cudaMalloc(&c.d_tileCount, c.cap_tiles * sizeof(unsigned)));
        IPIC_CUDA(cudaMalloc(&c.d_tileOffset, c.cap_tiles * sizeof(long long)));
        IPIC_CUDA(cudaMalloc(&c.d_loss, c.cap_tiles * 4 * sizeof(double)));
    }
    if (!c.d_total) IPIC_CUDA(cudaMalloc(&c.d_total, 8 * sizeof(long long)));
    const int nb = (int)s.slab_off.size();
    if (nb > c.cap_slab) {
        if (c.d_slab) cudaFree(c.d_slab);
        c.cap_slab = nb + 64;
        IPIC_CUDA(cudaMalloc(&c.d_slab, (size_t)2 * c.cap_slab * sizeof(long long)));
    }
    unsigned *d_tileCount = c.d_tileCount; long long *d_tileOffset = c.d_tileOffset; double *d_loss = c.d_loss;
    long long *d_total = c.d_total; long long *d_slab = c.d_slab; const int cap_slab = c.cap_slab;
    c.launches += 1; k_compact_count<<<(unsigned)ntiles, CT_THREADS, 0, c.stream>>>(s.l, s.vx, s.vy, s.vz, n, d_tileCount, d_loss);
    c.launches += 1; k_scan_tiles<<<1, 1024, 0, c.stream>>>(d_tileCount, ntiles, d_tileOffset, d_total);
    c.launches += 1; k_final_sum<4><<<1, RED_THREADS, 0, c.stream>>>(d_loss, (int)ntiles, c.d_scal);
    IPIC_CUDA(cudaGetLastError());
    long long total = 0;
    IPIC_CUDA(cudaMemcpyAsync(&total, d_total, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaMemcpyAsync(c.h_scal, c.d_scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    s.stats.energyLoss[0] = c.h_scal[0]; s.stats.energyLoss[1] = c.h_scal[1];
    s.stats.momentumLoss[0] += c.h_scal[2]; s.stats.momentumLoss[1] += c.h_scal[3];
    if (total == n) return IPIC_OK;   // nothing absorbed: arrays already compact
    if (nb == 0 && s.nRef == 0 && c.fill_compaction && (n - total) * 8 < n && n < (1ll << 31)) return compact_by_filling(c, s, n, total);
    if (nb > 0) {
        IPIC_CUDA(cudaMemcpyAsync(d_slab, s.slab_off.data(), nb * sizeof(long long), cudaMemcpyHostToDevice, c.stream));
        c.launches += 1; k_slab_offsets<<<(nb + 7) / 8, 256, 0, c.stream>>>(s.l, d_tileOffset, d_slab, d_slab + cap_slab, nb, n);
    }
    c.launches += 1; k_compact_scatter<<<(unsigned)ntiles, CT_THREADS, 0, c.stream>>>(s.l, s.vx, s.vy, s.vz, n, d_tileOffset, s.l2, s.vx2, s.vy2, s.vz2);
    IPIC_CUDA(cudaGetLastError());
    if (nb > 0) IPIC_CUDA(cudaMemcpyAsync(s.slab_off.data(), d_slab + cap_slab, nb * sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    if (nb > 0) s.slab_off[nb - 1] = total;
    swap_banks(s);
    s.n = total;
    return IPIC_OK;
}

// refRecordIdx: the reference stores, per thread slab, the post-compaction index `i - delIdx` of every particle at its first
// reflection in the step (mod_particleMover.f90:226-231), in particle order.  Here: the move kernel flags the particle, the flagged
// (pre-compaction) indices are selected in ascending order, and after the stable compaction each is mapped to
// tileOffset[tile] + (alive particles of the tile before it); a flagged particle that was absorbed later in the same step keeps its
// entry, which then names the slot of its successor, exactly as the reference's `i - delIdx` does.
__global__ void k_map_reflected(const unsigned *__restrict__ pre, long long nref, const double *__restrict__ l_old,
                                const long long *__restrict__ tileOffset, unsigned *post, unsigned char *live) {
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;    // one warp per entry
    if (w >= nref) return;
    const unsigned i = pre[w];
    const long long tile = i / CT_TILE, base = tile * CT_TILE;
    unsigned cnt = 0;
    for (long long j = base + lane; j < (long long)i; j += 32) cnt += (l_old[j] >= 0.0) ? 1u : 0u;
    cnt = __reduce_add_sync(FULL, cnt);
    // A particle absorbed later in the same step keeps its entry: the reference stored `i - delIdx` at the reflection, which after the
    // compaction is the slot of the next surviving particle (mod_particleMover.f90:226-231 records before the particle's fate is known).
    if (lane == 0) { post[w] = (unsigned)(tileOffset[tile] + cnt); live[w] = (l_old[i] >= 0.0) ? 1 : 0; }
}

int collect_reflected(Ctx &c, int sp) {
    SpeciesDev &s = c.sp[sp];
    s.nRef = 0;
    if (!s.refl || s.n == 0) return IPIC_OK;
    if (s.refList_cap < 2 * s.cap) {
        if (s.refList) cudaFree(s.refList);
        IPIC_CUDA(cudaMalloc(&s.refList, (size_t)2 * s.cap * sizeof(unsigned)));       // [0,cap): selected, [cap,2cap): mapped
        s.refList_cap = 2 * s.cap;
    }
    if (!c.d_total) IPIC_CUDA(cudaMalloc(&c.d_total, 8 * sizeof(long long)));
    int *d_num = reinterpret_cast<int *>(c.d_total + 4);
    size_t tmp_bytes = 0;
    cub::CountingInputIterator<unsigned> it(0u);
    cub::DeviceSelect::Flagged(nullptr, tmp_bytes, it, s.refl, s.refList, d_num, (int)s.n, c.stream);
    if (tmp_bytes > c.cap_sort_tmp) {
        if (c.d_sort_tmp) cudaFree(c.d_sort_tmp);
        c.cap_sort_tmp = tmp_bytes + tmp_bytes / 8;
        IPIC_CUDA(cudaMalloc(&c.d_sort_tmp, c.cap_sort_tmp));
    }
    c.launches += 1; cub::DeviceSelect::Flagged(c.d_sort_tmp, tmp_bytes, it, s.refl, s.refList, d_num, (int)s.n, c.stream);
    int h_num = 0;
    IPIC_CUDA(cudaMemcpyAsync(&h_num, d_num, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    s.nRef = h_num;
    return IPIC_OK;
}

int map_reflected(Ctx &c, int sp, bool compacted) {
    SpeciesDev &s = c.sp[sp];
    const int nslab = std::max(1, (int)s.slab_off_prev.size() - 1);
    s.refSlabCount.assign(nslab, 0);
    if (s.nRef == 0) return IPIC_OK;
    // refIdx(T): the slab every reflection was recorded in, from the pre-compaction indices (ascending, so the list splits in slab order)
    std::vector<unsigned> hpre((size_t)s.nRef);
    IPIC_CUDA(cudaMemcpyAsync(hpre.data(), s.refList, (size_t)s.nRef * sizeof(unsigned), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    if (s.slab_off_prev.size() < 2) s.refSlabCount[0] = (int)s.nRef;
    else {
        int t = 0;
        for (unsigned i : hpre) {
            while (t + 1 < nslab && (long long)i >= s.slab_off_prev[t + 1]) ++t;
            s.refSlabCount[t] += 1;
        }
    }
    if (s.refLive_cap < s.nRef) {
        if (s.refLive) cudaFree(s.refLive);
        s.refLive_cap = s.nRef + s.nRef / 4 + 1024;
        IPIC_CUDA(cudaMalloc(&s.refLive, (size_t)s.refLive_cap));
    }
    IPIC_CUDA(cudaMemsetAsync(s.refLive, 1, (size_t)s.nRef, c.stream));
    if (!compacted) return IPIC_OK;                            // nothing absorbed: indices are already final, every entry is its own particle
    unsigned *post = s.refList + s.refList_cap / 2;
    const long long warps_per_block = 8;
    c.launches += 1;
    k_map_reflected<<<(unsigned)((s.nRef + warps_per_block - 1) / warps_per_block), 256, 0, c.stream>>>(s.refList, s.nRef, s.l2 /* pre-compaction bank */,
                                                                                                          c.d_tileOffset, post, s.refLive);
    IPIC_CUDA(cudaGetLastError());
    // every entry is kept, in order; one that names the slot behind the last survivor of its slab (the reflected particle was the last of
    // the slab and was absorbed) is what the reference holds too -- consumers skip indices >= N_p
    IPIC_CUDA(cudaMemcpyAsync(s.refList, post, (size_t)s.nRef * sizeof(unsigned), cudaMemcpyDeviceToDevice, c.stream));
    return IPIC_OK;
}

// =====================================================================================================
// AoS slab <-> SoA transposition (reference layout phaseSpace(4, finalIdx, T), mod_particle.f90:23,65)
// =====================================================================================================
__global__ void k_aos_to_soa(const double4 *__restrict__ aos, long long n, double *l, double *vx, double *vy, double *vz) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double4 p = aos[i];
        l[i] = p.x; vx[i] = p.y; vy[i] = p.z; vz[i] = p.w;
    }
}
__global__ void k_soa_to_aos(double4 *aos, long long n, const double *__restrict__ l, const double *__restrict__ vx, const double *__restrict__ vy,
                             const double *__restrict__ vz) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) aos[i] = make_double4(l[i], vx[i], vy[i], vz[i]);
}

int upload_species(Ctx &c, int sp, const double *ps, const int32_t *N_p, int T, int finalIdx) {
    SpeciesDev &s = c.sp[sp];
    long long n = 0;
    for (int t = 0; t < T; ++t) {
        if (N_p[t] < 0 || N_p[t] > finalIdx) { set_error("N_p out of range"); return IPIC_ERR_ARG; }
        n += N_p[t];
    }
    int rc = ensure_capacity(c, sp, n);
    if (rc) return rc;
    s.slab_off.assign(T + 1, 0);
    double *stage = s.l2;   // bank B as flat AoS staging (4*cap doubles)
    long long off = 0;
    for (int t = 0; t < T; ++t) {
        s.slab_off[t] = off;
        if (N_p[t] > 0)
            IPIC_CUDA(cudaMemcpyAsync(stage + 4 * off, ps + (size_t)4 * finalIdx * t, (size_t)N_p[t] * 4 * sizeof(double), cudaMemcpyHostToDevice,
                                      c.stream));
        off += N_p[t];
    }
    s.slab_off[T] = off;
    s.n = n;
    s.n_collide = n; s.coll_taken = 0; s.coll_epoch += 1;
    if (n > 0) {
        c.launches += 1; k_aos_to_soa<<<c.num_sms * 4, 256, 0, c.stream>>>(reinterpret_cast<const double4 *>(stage), n, s.l, s.vx, s.vy, s.vz);
        IPIC_CUDA(cudaGetLastError());
    }
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    return IPIC_OK;
}

int download_species(Ctx &c, int sp, double *ps, int32_t *N_p, int T, int finalIdx) {
    SpeciesDev &s = c.sp[sp];
    bool resplit = (int)s.slab_off.size() != T + 1;
    if (!resplit)   // a slab outgrew the host's finalIdx (ionization appends to the last slab): rebalance if the total fits
        for (int t = 0; t < T; ++t) if (s.slab_off[t + 1] - s.slab_off[t] > finalIdx && s.n <= (long long)T * finalIdx) resplit = true;
    if (resplit) {
        // particles did not come from slabs (device loader / sorted): split evenly, as the reference loader does
        s.slab_off.assign(T + 1, 0);
        for (int t = 0; t <= T; ++t) s.slab_off[t] = s.n * t / T;
    }
    double *stage = s.l2;
    if (s.n > 0) {
        c.launches += 1; k_soa_to_aos<<<c.num_sms * 4, 256, 0, c.stream>>>(reinterpret_cast<double4 *>(stage), s.n, s.l, s.vx, s.vy, s.vz);
        IPIC_CUDA(cudaGetLastError());
    }
    for (int t = 0; t < T; ++t) {
        const long long cnt = s.slab_off[t + 1] - s.slab_off[t];
        if (cnt > finalIdx) { set_error("slab larger than finalIdx"); return IPIC_ERR_ARG; }
        N_p[t] = (int32_t)cnt;
        if (cnt > 0)
            IPIC_CUDA(cudaMemcpyAsync(ps + (size_t)4 * finalIdx * t, stage + 4 * s.slab_off[t], (size_t)cnt * 4 * sizeof(double),
                                      cudaMemcpyDeviceToHost, c.stream));
    }
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    return IPIC_OK;
}

// =====================================================================================================
// Device synthetic loader: uniform x -> l = getLFromX(x), 3-D Maxwellian by Box-Muller
// (mod_particle.f90:84-97, 156-191) from a counter-based PCG32 stream (one short stream per particle).
// =====================================================================================================
__global__ void k_generate(double *l, double *vx, double *vy, double *vz, long long n, long long first_index, const double *__restrict__ grid,
                           const double *__restrict__ dx, int Nx, int scheme, double startX, double L_domain, double v_therm, double v_drift,
                           unsigned long long seed, double alpha, int distType) {
    const double two_pi = 6.283185307179586476925286766559;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long gid = (unsigned long long)(first_index + i);
        unsigned long long inc = ((seed ^ (gid * 0x9E3779B97F4A7C15ULL)) << 1u) | 1u;
        unsigned long long state = 0ULL;
        pcg32_next(state, inc);
        state += seed + gid;
        pcg32_next(state, inc);
        double U0 = pcg_uniform(state, inc);
        if (distType != 0) {
            // initializeRandCosine / initializeRandSine (mod_particle.f90:99-151): rejection on (1 + alpha f(2 pi R2)) / (1 + alpha)
            double R1 = U0, R2 = pcg_uniform(state, inc);
            for (int guard = 0; guard < 4096; ++guard) {
                const double tr = (distType == 1) ? cos(two_pi * R2) : sin(two_pi * R2);
                if (!(R1 > (1.0 + alpha * tr) / (1.0 + alpha))) break;
                R1 = pcg_uniform(state, inc); R2 = pcg_uniform(state, inc);
            }
            U0 = R2;
        }
        const double U1 = pcg_uniform(state, inc), U2 = pcg_uniform(state, inc), U3 = pcg_uniform(state, inc), U4 = pcg_uniform(state, inc);
        const double x = U0 * L_domain + startX;
        int lo = 1, hi = Nx;   // grid arrays here are 0-based: grid[k-1] = grid(k)
        double lpos;
        if (scheme == 0) {
            while (lo != hi - 1) { const int mid = (lo + hi) / 2; if (grid[mid - 1] <= x) lo = mid; else hi = mid; }
            lpos = lo + (x - grid[lo - 1]) / dx[lo - 1];
        } else {
            while (lo != hi - 1) { const int mid = (lo + hi) / 2; if (grid[mid - 1] - 0.5 * dx[mid - 1] <= x) lo = mid; else hi = mid; }
            if (grid[lo - 1] + 0.5 * dx[lo - 1] <= x) lo = hi;
            lpos = lo + 0.5 + (x - grid[lo - 1]) / dx[lo - 1];
        }
        l[i] = lpos;
        const double r1 = v_therm * sqrt(-2.0 * log(U1)), r3 = v_therm * sqrt(-2.0 * log(U3));
        // generate3DMaxwellian :172-181: position-dependent drift for the perturbed loads
        double v_extra = v_drift;
        if (distType == 1) v_extra = v_drift * (1.0 + alpha * cos(two_pi * (x / L_domain)));
        else if (distType == 2) v_extra = v_drift * (1.0 - alpha * sin(two_pi * (x / L_domain)));
        vx[i] = r1 * cos(two_pi * U2) + v_extra;
        vy[i] = r1 * sin(two_pi * U2);
        vz[i] = r3 * sin(two_pi * U4);
    }
}

int generate_species(Ctx &c, int sp, long long n, double T_eV, double v_drift, unsigned long long seed, double alpha, int distType) {
    SpeciesDev &s = c.sp[sp];
    int rc = ensure_capacity(c, sp, n);
    if (rc) return rc;
    double startX, L;
    if (c.scheme != IPIC_SCHEME_CIC) { startX = c.h_grid[0]; L = c.h_grid[c.Nx - 1] - c.h_grid[0]; }
    else { startX = c.h_grid[0] - 0.5 * c.h_dx[0]; L = c.h_grid[c.Nx - 1] + 0.5 * c.h_dx[c.Nx - 1] - startX; }
    const double v_therm = sqrt(T_eV * e_c / s.mass);
    const long long first = (long long)c.rank * n;   // ranks draw disjoint index ranges of one global stream
    c.launches += 1; k_generate<<<c.num_sms * 8, 256, 0, c.stream>>>(s.l, s.vx, s.vy, s.vz, n, first, c.d_grid, c.d_dx, c.Nx, c.scheme == IPIC_SCHEME_CIC ? 1 : 0, startX, L, v_therm, v_drift,
                                                     seed + 0x1000003ULL * (unsigned long long)sp, alpha, distType);
    IPIC_CUDA(cudaGetLastError());
    s.n = n;
    s.n_collide = n; s.coll_taken = 0; s.coll_epoch += 1;
    s.slab_off.clear();
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    return IPIC_OK;
}

// =====================================================================================================
// Cell sort: stable LSD radix sort of (cell key, index) + gather.  Maintenance op, not on the hot path.
// =====================================================================================================
__global__ void k_sort_keys(const double *__restrict__ l, long long n, unsigned *keys, unsigned *vals, double sub) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ll = l[i];
        keys[i] = ll >= 0.0 ? (unsigned)(int)(ll * sub) : 0xFFFFu;   // dead particles sort to the end
        vals[i] = (unsigned)i;
    }
}
__global__ void k_gather4(const unsigned *__restrict__ perm, long long n, const double *__restrict__ l, const double *__restrict__ vx,
                          const double *__restrict__ vy, const double *__restrict__ vz, double *lo, double *vxo, double *vyo, double *vzo) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned j = perm[i];
        lo[i] = l[j]; vxo[i] = vx[j]; vyo[i] = vy[j]; vzo[i] = vz[j];
    }
}

int sort_species(Ctx &c, int sp) {
    SpeciesDev &s = c.sp[sp];
    const long long n = s.n;
    if (n <= 1) return IPIC_OK;
    if (n >= (1ll << 31)) { set_error("sort: more than 2^31-1 particles of one species per GPU"); return IPIC_ERR_UNSUPPORTED; }
    // refRecordIdx (post-compaction indices of the last move) does not survive a permutation: a sort between ipic_move_particles and the
    // refluxing / injection calls that consume the list would redraw the wrong particles, so the list is dropped here (ipic1d_b200.h)
    s.nRef = 0;
    std::fill(s.refSlabCount.begin(), s.refSlabCount.end(), 0);
    if (n > c.cap_sort) {
        if (c.d_sort) cudaFree(c.d_sort);
        c.cap_sort = n + n / 8 + 4096;
        IPIC_CUDA(cudaMalloc(&c.d_sort, (size_t)c.cap_sort * 4 * sizeof(unsigned)));
    }
    unsigned *keys = c.d_sort, *keys2 = keys + c.cap_sort, *vals = keys2 + c.cap_sort, *vals2 = vals + c.cap_sort;
    size_t tmp_bytes = 0;
    // key = INT(l * sub): cells cut into `sub` slices (8 unless that needs more than 16 key bits).  The finer order costs nothing -- the
    // radix sort runs two 8-bit passes either way -- and puts neighbours in space into the same warp: C5 step 99.1 -> 97.1 ms
    const int cells = std::max(c.Nx, c.Nh) + 2;
    const int sub = std::max(1, std::min(c.sort_subbins > 0 ? c.sort_subbins : 8, 65534 / cells));
    c.launches += 1; k_sort_keys<<<c.num_sms * 8, 256, 0, c.stream>>>(s.l, n, keys, vals, (double)sub);
    const int bits = 16;                                      // keys < 65535 = the key of absorbed, not yet compacted particles
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, vals, vals2, (int)n, 0, bits, c.stream);
    if (tmp_bytes > c.cap_sort_tmp) {
        if (c.d_sort_tmp) cudaFree(c.d_sort_tmp);
        c.cap_sort_tmp = tmp_bytes + tmp_bytes / 8;
        IPIC_CUDA(cudaMalloc(&c.d_sort_tmp, c.cap_sort_tmp));
    }
    c.launches += 1; cub::DeviceRadixSort::SortPairs(c.d_sort_tmp, tmp_bytes, keys, keys2, vals, vals2, (int)n, 0, bits, c.stream);
    c.launches += 1; k_gather4<<<c.num_sms * 8, 256, 0, c.stream>>>(vals2, n, s.l, s.vx, s.vy, s.vz, s.l2, s.vx2, s.vy2, s.vz2);
    IPIC_CUDA(cudaGetLastError());
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    swap_banks(s);
    s.slab_off.clear();   // slab identity is not preserved by a sort
    s.n_collide = s.n; s.coll_taken = 0; s.coll_epoch += 1;   // nor is the prefix that excluded this step's newborn: every particle is collidable again
    return IPIC_OK;
}

}  // namespace ipic
