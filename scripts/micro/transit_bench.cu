// Micro-benchmark of the lean transit loop in isolation: cycles per face-to-face crossing per SM sub-partition as a function of the
// resident warps (FP64 instructions issue every 2nd cycle and do not co-issue with the rest: 13 FP64 + 14 others ~ 40 cycles).  The two device functions are the ones of particles.cu.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o transit_bench transit_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
__device__ __forceinline__ int transit_step(double u_prev, double u_cur, double t_prev, double &t_cur, double inv_a, double dpe_next, double sgn,
                                            double t_lim, bool alive, double &u_cand) {
    int ok;
    asm("{\n\t"
        ".reg .pred pa, pk;\n\t"
        ".reg .f64 uu, s, y, y1, t, nt, e, c, p, g, ng, h, r, d;\n\t"
        ".reg .b32 lo, hi;\n\t"
        "setp.ne.s32 pa, %10, 0;\n\t"
        "mul.rn.f64 uu, %4, %4;\n\t"
        "sub.rn.f64 d, %4, %3;\n\t"
        "fma.rn.f64 s, %8, %7, uu;\n\t"
        "mul.rn.f64 d, d, %6;\n\t"
        "rsqrt.approx.ftz.f64 y, s;\n\t"
        "@pa fma.rn.f64 %0, %8, d, %5;\n\t"
        "mul.rn.f64 t, y, y;\n\t"
        "neg.f64 nt, t;\n\t"
        "setp.lt.and.f64 pk, %0, %9, pa;\n\t"
        "fma.rn.f64 e, s, nt, 0d3FF0000000000000;\n\t"
        "selp.s32 %2, 1, 0, pk;\n\t"
        "mul.rn.f64 p, y, e;\n\t"
        "fma.rn.f64 y1, p, 0d3FE0000000000000, y;\n\t"
        "mul.rn.f64 g, s, y1;\n\t"
        "mov.b64 {lo, hi}, y1;\n\t"
        "add.s32 hi, hi, -1048576;\n\t"
        "mov.b64 h, {lo, hi};\n\t"
        "neg.f64 ng, g;\n\t"
        "fma.rn.f64 r, g, ng, s;\n\t"
        "fma.rn.f64 %1, r, h, g;\n\t"
        "}"
        : "+d"(t_cur), "=d"(u_cand), "=r"(ok)
        : "d"(u_prev), "d"(u_cur), "d"(t_prev), "d"(inv_a), "d"(dpe_next), "d"(sgn), "d"(t_lim), "r"((int)alive));
    return ok;
}

__device__ __forceinline__ double transit_speed(double u, double dpe, double sgn) {
    double uf;
    asm("{\n\t"
        ".reg .f64 uu, s, y, y1, t, nt, e, c, p, g, ng, h, r;\n\t"
        ".reg .b32 lo, hi;\n\t"
        "mul.rn.f64 uu, %1, %1;\n\t"
        "fma.rn.f64 s, %3, %2, uu;\n\t"
        "rsqrt.approx.ftz.f64 y, s;\n\t"
        "mul.rn.f64 t, y, y;\n\t"
        "neg.f64 nt, t;\n\t"
        "fma.rn.f64 e, s, nt, 0d3FF0000000000000;\n\t"
        "mul.rn.f64 p, y, e;\n\t"
        "fma.rn.f64 y1, p, 0d3FE0000000000000, y;\n\t"
        "mul.rn.f64 g, s, y1;\n\t"
        "mov.b64 {lo, hi}, y1;\n\t"
        "add.s32 hi, hi, -1048576;\n\t"
        "mov.b64 h, {lo, hi};\n\t"
        "neg.f64 ng, g;\n\t"
        "fma.rn.f64 r, g, ng, s;\n\t"
        "fma.rn.f64 %0, r, h, g;\n\t"
        "}"
        : "=d"(uf) : "d"(u), "d"(dpe), "d"(sgn));
    return uf;
}

__global__ void __launch_bounds__(1024, 1) k_transit(const double2 *__restrict__ ttab, int ncell, int iters, double t_lim, double *out, long long *cyc) {
    const int lane = threadIdx.x & 31;
    int cell = 2 + (threadIdx.x * 37 + blockIdx.x * 11) % (ncell - 4);
    const int step = 1, cell_end = 1 << 30;
    const double sgn = 1.0;
    bool tr = true;
    double uA = 1.0e6 + threadIdx.x, tA = 0.0, uB, tB = 0.0;
    double2 ttB = __ldg(ttab + cell), ttA = __ldg(ttab + (cell + step));
    uB = transit_speed(uA, ttB.x, sgn);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double cand;
        int ok = transit_step(uA, uB, tA, tB, ttB.y, ttA.x, sgn, t_lim, tr, cand);
        cell += ok ? step : 0; if (cell >= ncell - 2) cell = 2;
        tr = ok && (cell != cell_end);
        ttB = __ldg(ttab + (cell + step));
        uA = tr ? cand : uA;
        ok = transit_step(uB, uA, tB, tA, ttA.y, ttB.x, sgn, t_lim, tr, cand);
        cell += ok ? step : 0; if (cell >= ncell - 2) cell = 2;
        tr = ok && (cell != cell_end);
        ttA = __ldg(ttab + (cell + step));
        uB = tr ? cand : uB;
        if (__popc(__ballot_sync(FULL, tr)) < 12) break;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = uA + uB + tA + tB + cell;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    const int ncell = 4096, iters = 20000;
    double2 *tt; double *out; long long *dc;
    cudaMalloc(&tt, ncell * sizeof(double2)); cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    double2 *h = new double2[ncell];
    for (int i = 0; i < ncell; ++i) { h[i].x = 1.0e3 * (1 + (i % 7)); h[i].y = 1.0e-9 / (1 + (i % 5)); }   // small energy gain per cell, tiny time increments
    cudaMemcpy(tt, h, ncell * sizeof(double2), cudaMemcpyHostToDevice);
    for (int threads : {128, 256, 512, 768, 1024}) {
        for (int rep = 0; rep < 2; ++rep) { k_transit<<<148, threads>>>(tt, ncell, iters, 1.0e30, out, dc); cudaDeviceSynchronize(); }
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        const double warps_per_smsp = threads / 128.0;
        printf("threads %4d (%.0f warps/SMSP): %.1f cycles per crossing per warp, %.1f cycles per crossing per SMSP\n", threads, warps_per_smsp,
               (double)c / (2.0 * iters), (double)c / (2.0 * iters) / warps_per_smsp);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
