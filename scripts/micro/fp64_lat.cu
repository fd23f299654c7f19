// Micro-benchmark: dependent-issue latency and throughput of FP64 DFMA / DMUL / MUFU.RSQ64H on this GPU.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu ; run: ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double *out, double a, double b, int iters, long long *cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], b, a);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_rsq(double *out, double a, int iters, long long *cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y; }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int threads, double *d, long long *dc) {
    const int iters = 4096;
    k_dfma<ILP><<<1, threads>>>(d, 1.0, 0.999999, iters, dc); cudaDeviceSynchronize();
    k_dfma<ILP><<<1, threads>>>(d, 1.0, 0.999999, iters, dc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 16.0);
    printf("DFMA threads %4d (warps/SMSP %.2f) ILP %d: %.2f cycles per dependent step; %.2f warp-DFMA/cycle/SMSP\n", threads, threads / 128.0, ILP, per,
           ILP * (threads / 32.0 / 4.0) / per);
}
int main() {
    double *d; long long *dc; cudaMalloc(&d, 1 << 20); cudaMalloc(&dc, 8);
    for (int th : {32, 128, 256, 512, 1024}) { run<1>(th, d, dc); run<2>(th, d, dc); run<4>(th, d, dc); }
    k_rsq<<<1, 32>>>(d, 1.5, 4096, dc); cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    printf("MUFU.RSQ64H dependent: %.2f cycles\n", (double)c / (4096 * 16.0));
    return 0;
}
