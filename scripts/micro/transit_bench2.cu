// Micro-benchmark + numerical check of the round-2 transit loop (energy recurrence, signed Goldschmidt square root, integer time test,
// predicated commits into alternating register copies -- no selects).  transit_pair is the device function of particles.cu.
//   part 1: cycles per face-to-face crossing per SM sub-partition against resident warps
//   part 2: the square-root sequence against CUDA's correctly rounded sqrt(double), bit for bit, on random operands
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o transit_bench2 transit_bench2.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#define FULL 0xffffffffu

// Two crossings (state A -> B -> A).  s = v^2 carried as the energy recurrence s' = s + sgn * (2 a dx); v signed; t the time used.
// `alive` (in/out): the lane has passed every test so far; the writes of a crossing are predicated on the lane being alive when the
// crossing starts, so a lane that fails keeps its last committed copy (A after an even number of passes, B after an odd one).
// ptr walks the {2 a dx, 1/a} table by `stepb` bytes per committed crossing; e0 / e1 hold the entries of the next two cells.
__device__ __forceinline__ unsigned transit_pair(double &sA, double &vA, double &tA, double &sB, double &vB, double &tB, double2 &e0, double2 &e1,
                                                 const double2 *&ptr, long long stepb, double sgn, int ksign, int khalf, unsigned long long tlim_bits,
                                                 int &alive, int &npass) {
    unsigned mask;
    asm volatile("{\n\t"
        ".reg .pred pa, pb, pc;\n\t"
        ".reg .f64 y, ys, hs, g, ng, r, g1, ng1, h1, r1, d;\n\t"
        ".reg .b32 lo, hi, hi1, hi2;\n\t"
        ".reg .b64 tb, pn;\n\t"
        "setp.ne.s32 pa, %11, 0;\n\t"
        // ---- crossing 1: A -> B with entry e0
        "@pa fma.rn.f64 %3, %15, %6, %0;\n\t"                 // sB = sgn * dpe + sA
        "rsqrt.approx.ftz.f64 y, %3;\n\t"
        "mov.b64 {lo, hi}, y;\n\t"
        "add.s32 hi1, hi, %16;\n\t"
        "add.s32 hi2, hi, %17;\n\t"
        "mov.b64 ys, {lo, hi1};\n\t"                          // sgn / sqrt(s)
        "mov.b64 hs, {lo, hi2};\n\t"                          // sgn / (2 sqrt(s))
        "mul.rn.f64 g, %3, ys;\n\t"
        "neg.f64 ng, g;\n\t"
        "fma.rn.f64 r, ng, hs, 0d3FE0000000000000;\n\t"
        "fma.rn.f64 g1, g, r, g;\n\t"
        "fma.rn.f64 h1, hs, r, hs;\n\t"
        "neg.f64 ng1, g1;\n\t"
        "fma.rn.f64 r1, ng1, g1, %3;\n\t"
        "@pa fma.rn.f64 %4, r1, h1, g1;\n\t"                  // vB = sgn * sqrt(sB)
        "sub.rn.f64 d, %4, %1;\n\t"
        "@pa fma.rn.f64 %5, d, %7, %2;\n\t"                   // tB = (vB - vA) / a + tA
        "mov.b64 tb, %5;\n\t"
        "setp.lt.and.u64 pb, tb, %18, pa;\n\t"
        "@pb add.s64 %10, %10, %14;\n\t"
        "@pb add.s32 %12, %12, 1;\n\t"
        "add.s64 pn, %10, %14;\n\t"
        "ld.global.nc.v2.f64 {%6, %7}, [pn];\n\t"
        // ---- crossing 2: B -> A with entry e1
        "@pb fma.rn.f64 %0, %15, %8, %3;\n\t"
        "rsqrt.approx.ftz.f64 y, %0;\n\t"
        "mov.b64 {lo, hi}, y;\n\t"
        "add.s32 hi1, hi, %16;\n\t"
        "add.s32 hi2, hi, %17;\n\t"
        "mov.b64 ys, {lo, hi1};\n\t"
        "mov.b64 hs, {lo, hi2};\n\t"
        "mul.rn.f64 g, %0, ys;\n\t"
        "neg.f64 ng, g;\n\t"
        "fma.rn.f64 r, ng, hs, 0d3FE0000000000000;\n\t"
        "fma.rn.f64 g1, g, r, g;\n\t"
        "fma.rn.f64 h1, hs, r, hs;\n\t"
        "neg.f64 ng1, g1;\n\t"
        "fma.rn.f64 r1, ng1, g1, %0;\n\t"
        "@pb fma.rn.f64 %1, r1, h1, g1;\n\t"
        "sub.rn.f64 d, %1, %4;\n\t"
        "@pb fma.rn.f64 %2, d, %9, %5;\n\t"
        "mov.b64 tb, %2;\n\t"
        "setp.lt.and.u64 pc, tb, %18, pb;\n\t"
        "@pc add.s64 %10, %10, %14;\n\t"
        "@pc add.s32 %12, %12, 1;\n\t"
        "add.s64 pn, %10, %14;\n\t"
        "ld.global.nc.v2.f64 {%8, %9}, [pn];\n\t"
        "selp.s32 %11, 1, 0, pc;\n\t"
        "vote.sync.ballot.b32 %13, pc, 0xffffffff;\n\t"
        "}"
        : "+d"(sA), "+d"(vA), "+d"(tA), "+d"(sB), "+d"(vB), "+d"(tB), "+d"(e0.x), "+d"(e0.y), "+d"(e1.x), "+d"(e1.y), "+l"(ptr), "+r"(alive),
          "+r"(npass), "=r"(mask)
        : "l"(stepb), "d"(sgn), "r"(ksign), "r"(khalf), "l"(tlim_bits));
    return mask;
}

__global__ void __launch_bounds__(1024, 1) k_transit(const double2 *__restrict__ ttab, int ncell, int iters, unsigned long long tlim_bits, double *out,
                                                     long long *cyc) {
    int cell = 2 + (threadIdx.x * 37 + blockIdx.x * 11) % (ncell / 2);
    const bool right = true;
    const double sgn = right ? 1.0 : -1.0;
    const int ksign = right ? 0 : (int)0x80000000, khalf = ksign - 0x100000;
    const long long stepb = right ? 16 : -16;
    double vA = sgn * (1.0e6 + threadIdx.x), sA = vA * vA, tA = 0.0, sB = sA, vB = vA, tB = 0.0;
    const double2 *ptr = ttab + cell;
    double2 e0 = __ldg(ptr), e1 = __ldg(ptr + (right ? 1 : -1));
    int alive = 1, npass = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const unsigned m = transit_pair(sA, vA, tA, sB, vB, tB, e0, e1, ptr, stepb, sgn, ksign, khalf, tlim_bits, alive, npass);
        if (ptr >= ttab + ncell - 4) { ptr = ttab + 2; e0 = __ldg(ptr); e1 = __ldg(ptr + 1); }   // wrap (benchmark only; rare)
        if (__popc(m) < 12) break;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = sA + vA + tA + sB + vB + tB + npass + alive;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// ---- square-root check: the loop's sequence (unsigned, sgn = +1) against sqrt() ----
__device__ __forceinline__ double seq(double s, int ksign, int khalf) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
    const double ys = __hiloint2double(__double2hiint(y) + ksign, __double2loint(y));
    const double hs = __hiloint2double(__double2hiint(y) + khalf, __double2loint(y));
    const double g = __dmul_rn(s, ys);
    const double r = fma(-g, hs, 0.5);
    const double g1 = fma(g, r, g), h1 = fma(hs, r, hs);
    const double r1 = fma(-g1, g1, s);
    return fma(r1, h1, g1);
}
__device__ unsigned long long rng(unsigned long long &x) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; }
__global__ void k_check(unsigned long long seed, int iters, int mode, unsigned long long *bad) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long b0 = 0, b1 = 0, b2 = 0;
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
        const double ref = sqrt(s);
        const double a = seq(s, 0, -0x100000);
        const double b = seq(s, (int)0x80000000, (int)0x80000000 - 0x100000);
        if (a != ref) { ++b0; if (fabs(a - ref) > 1.0000001 * fabs(ref) * 2.220446049250313e-16) ++b2; }
        if (b != -ref) ++b1;
    }
    atomicAdd(bad, b0); atomicAdd(bad + 1, b1); atomicAdd(bad + 2, b2);
}

int main() {
    const int ncell = 4096, iters = 20000;
    double2 *tt; double *out; long long *dc;
    cudaMalloc(&tt, ncell * sizeof(double2)); cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    double2 *h = new double2[ncell];
    for (int i = 0; i < ncell; ++i) { h[i].x = 1.0e3 * (1 + (i % 7)); h[i].y = 1.0e-9 / (1 + (i % 5)); }   // small energy gain per cell, tiny time increments
    cudaMemcpy(tt, h, ncell * sizeof(double2), cudaMemcpyHostToDevice);
    const double tl = 1.0e30; unsigned long long tb; memcpy(&tb, &tl, 8);
    for (int threads : {128, 256, 512, 768, 1024}) {
        for (int rep = 0; rep < 2; ++rep) { k_transit<<<148, threads>>>(tt, ncell, iters, tb, out, dc); cudaDeviceSynchronize(); }
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = threads / 128.0;
        printf("threads %4d (%.0f warps/SMSP): %.1f cycles per crossing per warp, %.1f cycles per crossing per SMSP\n", threads, warps_per_smsp,
               (double)c / (2.0 * iters), (double)c / (2.0 * iters) / warps_per_smsp);
    }
    unsigned long long *d; cudaMalloc(&d, 24);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(d, 0, 24);
        const int it2 = 8192;
        k_check<<<1184, 256>>>(12345 + mode, it2, mode, d);
        unsigned long long hb[3]; cudaMemcpy(hb, d, 24, cudaMemcpyDeviceToHost);
        printf("mode %d: %.3e operands; mismatches vs sqrt(): +chain %llu, -chain %llu, more than 1 ulp %llu\n", mode, 1184.0 * 256 * it2, hb[0], hb[1], hb[2]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
