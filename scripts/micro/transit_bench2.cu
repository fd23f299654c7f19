#include <cstdlib>
// Micro-benchmark + numerical check of the round-2 transit loop.  lean_sqrt and transit_loop below are the device functions of
// implicitpic1d_b200/csrc/particles.cu, verbatim (scripts/micro/sync_transit.py refreshes them).
//   part 1: cycles per face-to-face crossing per SM sub-partition against resident warps (table of 4096 cells: L1 resident)
//   part 2: lean_sqrt against CUDA's correctly rounded sqrt(double), bit for bit, on random and near-square operands
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o transit_bench2 transit_bench2.cu
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>
#define FULL 0xffffffffu

#define IPIC_TRIPS_INC
// ---- from particles.cu ----
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

// ---- end of the copy ----

__global__ void __launch_bounds__(1024, 1) k_transit(const double2 *__restrict__ ttab, int ncell, int reps, unsigned long long tlim_bits, double *out, long long *cyc, int mode) {
    // mode 0: every lane of a warp starts in another 128-byte line of the table; 1: one cell per warp (sorted particles); 2: consecutive cells
    const int cell = mode == 0 ? 2 + (threadIdx.x * 37 + blockIdx.x * 11) % (ncell / 2) : mode == 1 ? 2 + ((threadIdx.x >> 5) * 37 + blockIdx.x * 11) % (ncell / 2) : 2 + (threadIdx.x + blockIdx.x * 11) % (ncell / 2);
    const double v0 = 1.0e6 + threadIdx.x;
    long long crossings = 0;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        double u = v0 + rep, t = 0.0, t_fail; bool o;
        unsigned trips = 0; crossings += transit_loop(u, t, ttab, 16u * (unsigned)cell, tlim_bits, true, 12u, o, t_fail, trips);
        acc += u + t;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = crossings; }
}

__device__ unsigned long long rng(unsigned long long &x) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; }
__global__ void k_check(unsigned long long seed, int iters, int mode, unsigned long long *bad) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long b0 = 0, b2 = 0;
    for (int i = 0; i < iters; ++i) {
        const unsigned long long r = rng(x);
        double s;
        if (mode == 0) {   // full mantissa, exponents covering v^2 of 1e2 .. 1e9 m/s
            const unsigned long long m = r & 0x000FFFFFFFFFFFFFull;
            const unsigned long long ex = 1023ull + 13ull + (rng(x) % 48ull);
            s = __longlong_as_double((long long)((ex << 52) | m));
        } else {           // u^2 + small: squares of doubles perturbed by a few ulps (near-exact-square stress)
            const double u = 1.0e5 + (double)(r >> 40);
            s = __dmul_rn(u, u) + (double)((long long)(rng(x) & 7) - 3) * 1.52587890625e-5;
        }
        const double ref = sqrt(s), a = lean_sqrt(s);
        if (a != ref) { ++b0; if (fabs(a - ref) > 1.0000001 * fabs(ref) * 2.220446049250313e-16) ++b2; }
    }
    atomicAdd(bad, b0); atomicAdd(bad + 1, b2);
}

int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int ncell = 4096, reps = 10;
    double2 *tt; double *out; long long *dc;
    cudaMalloc(&tt, ncell * sizeof(double2)); cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dc, 16);
    double2 *h = new double2[ncell];
    for (int i = 0; i < ncell; ++i) { h[i].x = 1.0e3 * (1 + (i % 7)); h[i].y = 1.0e-9 / (1 + (i % 5)); }   // small energy gain per cell, tiny time increments
    for (int i = 4000; i < ncell; ++i) h[i].x = nan("");                                                  // a wall: every lane stops there
    cudaMemcpy(tt, h, ncell * sizeof(double2), cudaMemcpyHostToDevice);
    const double tl = 1.0e30; unsigned long long tb; memcpy(&tb, &tl, 8);
    for (int threads : {128, 256, 512, 768, 1024}) {
        for (int rep = 0; rep < 2; ++rep) { k_transit<<<148, threads>>>(tt, ncell, reps, tb, out, dc, mode); cudaDeviceSynchronize(); }
        long long c[2]; cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
        const double w = threads / 128.0;
        printf("threads %4d (%.0f warps/SMSP): %.1f cycles per crossing per warp, %.1f cycles per crossing per SMSP\n", threads, w, (double)c[0] / c[1], (double)c[0] / c[1] / w);
    }
    unsigned long long *d; cudaMalloc(&d, 16);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(d, 0, 16);
        const int it2 = 8192;
        k_check<<<1184, 256>>>(12345 + mode, it2, mode, d);
        unsigned long long hb[2]; cudaMemcpy(hb, d, 16, cudaMemcpyDeviceToHost);
        printf("mode %d: %.3e operands; lean_sqrt != sqrt(): %llu, off by more than 1 ulp: %llu\n", mode, 1184.0 * 256 * it2, hb[0], hb[1]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
