// Micro-benchmark: issue throughput (warp-instructions per cycle per SM sub-partition) of the FP64-pipe instruction types of the
// transit loop, 8 warps per sub-partition, 4 independent chains per thread.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_tput fp64_tput.cu
#include <cstdio>
#include <cuda_runtime.h>
#define BODY(NAME, ...)                                                                                           \
    __global__ void NAME(double *out, double a, double b, int iters, long long *cyc) {                           \
        double x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3; int p = 0;                            \
        long long t0 = clock64();                                                                                 \
        for (int it = 0; it < iters; ++it) {                                                                      \
            _Pragma("unroll") for (int u = 0; u < 8; ++u) { __VA_ARGS__ }                                               \
        }                                                                                                         \
        long long t1 = clock64();                                                                                 \
        out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + p;                                       \
        if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;                                                  \
    }
BODY(k_dfma, x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a);)
BODY(k_dmul, x0 = __dmul_rn(x0, b); x1 = __dmul_rn(x1, b); x2 = __dmul_rn(x2, b); x3 = __dmul_rn(x3, b);)
BODY(k_dadd, x0 = __dadd_rn(x0, b); x1 = __dadd_rn(x1, b); x2 = __dadd_rn(x2, b); x3 = __dadd_rn(x3, b);)
BODY(k_dsetp, { int q0, q1, q2, q3; asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.s32 %0, 1, 0, p;}" : "=r"(q0) : "d"(x0), "d"(b + p));
               asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.s32 %0, 1, 0, p;}" : "=r"(q1) : "d"(x1), "d"(b + p));
               asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.s32 %0, 1, 0, p;}" : "=r"(q2) : "d"(x2), "d"(b + p));
               asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.s32 %0, 1, 0, p;}" : "=r"(q3) : "d"(x3), "d"(b + p)); p += q0 + q1 + q2 + q3; })
BODY(k_rsq, { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x0)); x0 = y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x1)); x1 = y;
              asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x2)); x2 = y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x3)); x3 = y; })
BODY(k_fsel, { x0 = (p & 1) ? x1 : x0; x1 = (p & 2) ? x2 : x1; x2 = (p & 4) ? x3 : x2; x3 = (p & 8) ? x0 : x3; p += 3; })
BODY(k_mix, { double y; x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a);
              x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x0 = fma(x0, b, a);
              asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x3)); x3 = y; })
BODY(k_mix_sel, { double y; x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a);
              x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x0 = fma(x0, b, a);
              int q; asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.s32 %0, 1, 0, p;}" : "=r"(q) : "d"(x0), "d"(x1)); p += q;
              x1 = (p & 1) ? x2 : x1; x2 = (p & 2) ? x0 : x2; })
template <class K> void run(const char *name, K k, double *d, long long *dc, double b) {
    const int iters = 2048;
    for (int r = 0; r < 2; ++r) { k<<<148, 1024>>>(d, 1.5, b, iters, dc); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("%-6s: %.3f warp-instr/cycle/SMSP (%.2f cycles per warp-instruction)\n", name, 8.0 * iters * 8 * 4 / (double)c, (double)c / (8.0 * iters * 8 * 4));
}
int main() {
    double *d; long long *dc; cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    run("DFMA", k_dfma, d, dc, 0.999999); run("DMUL", k_dmul, d, dc, 0.999999); run("DADD", k_dadd, d, dc, 1e-9);
    run("DSETP", k_dsetp, d, dc, 0.5); run("RSQ64H", k_rsq, d, dc, 0.0); run("FSELx2", k_fsel, d, dc, 0.0);
    {   // 13 DFMA + 1 RSQ64H per group: 26 cycles if the MUFU overlaps the FP64 pipe, 34 if it does not
        const int iters = 2048;
        for (int r = 0; r < 2; ++r) { k_mix<<<148, 1024>>>(d, 1.5, 0.999999, iters, dc); cudaDeviceSynchronize(); }
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("13 DFMA + 1 RSQ64H: %.2f cycles per group per SMSP\n", (double)c / (8.0 * iters * 8));
        for (int r = 0; r < 2; ++r) { k_mix_sel<<<148, 1024>>>(d, 1.5, 0.999999, iters, dc); cudaDeviceSynchronize(); }
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("13 DFMA + DSETP + SELP + 4 FSEL: %.2f cycles per group per SMSP\n", (double)c / (8.0 * iters * 8));
    }
    return 0;
}
