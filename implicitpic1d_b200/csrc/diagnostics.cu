// Velocity / energy histograms of the final averaging window on the device.
//
// Reference semantics restated here (never copied): OpenMP/SharedModules/src/mod_simulation.f90
//   extrema pass :437-452 (VMax = max |v_x|, EMax = max v^2, Emin = min(v^2, 0.1 eV)),
//   histogram pass :480-500 (v_x onto 2*binNumber linear bins over [-VMax, VMax], v^2 onto binNumber logarithmic bins over
//   [Emin, EMax], both with linear weights on the two neighbouring bins).
//
// One pass over (v_x, v_y, v_z), 24 B per particle (a diagnostic of the final window, not on the hot path).  Bin weights are accumulated as 64-bit fixed-point integers
// (2^-36), first in shared memory, then with one global atomic per touched bin and CTA: integer sums, so the histograms are
// bit-identical run to run and for any launch geometry.  Extrema are integer atomics on the bit patterns of non-negative doubles
// (which order like the doubles).
#include "ipic_common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace ipic {

constexpr int HIST_THREADS = 512;
constexpr int HIST_FX_BITS = 36;                       // <= 2^27 particles per call and bin before an int64 could overflow
constexpr double HIST_FX = 68719476736.0;              // 2^36
constexpr int HIST_MAX_BINS = 2048;                    // 3 * bins * 8 B of shared memory

// SUM(phaseSpace(2:4)**2) in the reference's order, without FMA contraction: the strict comparisons against Emin / EMax (:489)
// then select exactly the particles the reference selects
__device__ __forceinline__ double v2_exact(double x, double y, double z) {
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
}

__global__ void __launch_bounds__(HIST_THREADS) k_velocity_extrema(const double *__restrict__ l, const double *__restrict__ vx,
                                                                   const double *__restrict__ vy, const double *__restrict__ vz, long long n,
                                                                   unsigned long long *out /* [3]: max |vx|, max v^2, min v^2 (bit patterns) */) {
    double vmax = 0.0, emax = 0.0, emin = __longlong_as_double((long long)out[2]);
    const long long stride = (long long)gridDim.x * HIST_THREADS;
    for (long long i = (long long)blockIdx.x * HIST_THREADS + threadIdx.x; i < n; i += stride) {
        const double ll = l[i], x = vx[i], y = vy[i], z = vz[i];
        if (ll < 0.0) continue;
        const double e2 = v2_exact(x, y, z);
        vmax = fmax(vmax, fabs(x)); emax = fmax(emax, e2); emin = fmin(emin, e2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        emax = fmax(emax, __shfl_xor_sync(0xffffffffu, emax, o));
        emin = fmin(emin, __shfl_xor_sync(0xffffffffu, emin, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out + 0, (unsigned long long)__double_as_longlong(vmax));
        atomicMax(out + 1, (unsigned long long)__double_as_longlong(emax));
        atomicMin(out + 2, (unsigned long long)__double_as_longlong(emin));
    }
}

__global__ void __launch_bounds__(HIST_THREADS) k_velocity_histograms(const double *__restrict__ l, const double *__restrict__ vx,
                                                                      const double *__restrict__ vy, const double *__restrict__ vz, long long n,
                                                                      int bins, int copies, double VMax, double Emin, double EMax, double Emin_log, double diffE,
                                                                      unsigned long long *ghist /* [2*bins] VHist, [bins] EHist */) {
    // `copies` private histograms per CTA (warp w deposits into copy w % copies): the hot bins of a Maxwellian otherwise serialise every
    // warp of the CTA on the same few shared-memory words
    extern __shared__ unsigned long long sh[];
    const int hb = 3 * bins;
    for (int k = threadIdx.x; k < copies * hb; k += HIST_THREADS) sh[k] = 0ull;
    __syncthreads();
    unsigned long long *vh = sh + ((threadIdx.x >> 5) % copies) * hb, *eh = vh + 2 * bins;       // 0-based: vh[k-1] = VHist(k)
    const long long stride = (long long)gridDim.x * HIST_THREADS;
    for (long long i = (long long)blockIdx.x * HIST_THREADS + threadIdx.x; i < n; i += stride) {
        const double ll = l[i], x = vx[i], y = vy[i], z = vz[i];      // one memory round trip: the velocities of a dead particle are read and dropped
        if (ll < 0.0) continue;
        const double partV = x * (2.0 * bins - 1) / 2.0 / VMax + bins + 0.5;      // :483, in the reference's operation order
        if (partV > 1.0 && partV < (double)(bins * 2)) {
            const int ip = (int)partV;
            const double f = partV - (double)ip;
            atomicAdd(vh + ip - 1, (unsigned long long)llrint((1.0 - f) * HIST_FX));
            atomicAdd(vh + ip, (unsigned long long)llrint(f * HIST_FX));
        }
        const double partE = v2_exact(x, y, z);
        if (partE > Emin && partE < EMax) {
            double pv = (log(partE) - Emin_log) / diffE + 1.0;
            int ip = (int)pv;
            ip = min(max(ip, 1), bins - 1);                  // Emin < partE < EMax keeps 1 <= pv < binNumber up to the rounding of log
            pv = pv - (double)ip;
            atomicAdd(eh + ip - 1, (unsigned long long)llrint((1.0 - pv) * HIST_FX));
            atomicAdd(eh + ip, (unsigned long long)llrint(pv * HIST_FX));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < hb; k += HIST_THREADS) {
        unsigned long long t = 0ull;
        for (int c = 0; c < copies; ++c) t += sh[c * hb + k];
        if (t) atomicAdd(ghist + k, t);
    }
}

static int diag_scratch(Ctx &c, size_t words, unsigned long long **out) {
    const long long need = (long long)(words * sizeof(unsigned long long));
    if (need > c.cap_diag) {
        if (c.d_diag) cudaFree(c.d_diag);
        c.cap_diag = need + 1024;
        IPIC_CUDA(cudaMalloc(&c.d_diag, (size_t)c.cap_diag));
    }
    *out = c.d_diag;
    return IPIC_OK;
}
void diagnostics_free(Ctx &c) {
    if (c.d_diag) cudaFree(c.d_diag);
    c.d_diag = nullptr; c.cap_diag = 0;
}

// mod_simulation.f90:437-452 for one species of this shard: out3 = {VMax, EMax, Emin}; Emin starts at 0.1 eV (0.1 e 2 / m)
int velocity_extrema(Ctx &c, int sp, double out3[3]) {
    SpeciesDev &S = c.sp[sp];
    unsigned long long *d;
    int rc = diag_scratch(c, 4, &d); if (rc) return rc;
    const double emin0 = 0.1 * e_c * 2.0 / S.mass;
    unsigned long long h[3] = {0ull, 0ull, 0ull};
    static_assert(sizeof(double) == sizeof(unsigned long long), "bit patterns");
    std::memcpy(&h[2], &emin0, sizeof(double));
    IPIC_CUDA(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, c.stream));
    if (S.n > 0) {
        c.launches += 1; k_velocity_extrema<<<c.num_sms * 4, HIST_THREADS, 0, c.stream>>>(S.l, S.vx, S.vy, S.vz, S.n, d);
        IPIC_CUDA(cudaGetLastError());
    }
    IPIC_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    std::memcpy(out3, h, 3 * sizeof(double));
    return IPIC_OK;
}

// mod_simulation.f90:480-500 for one species of this shard: adds this call's weights to VHist[2*bins] and EHist[bins]
int velocity_histograms(Ctx &c, int sp, int bins, double VMax, double Emin, double EMax, double *VHist, double *EHist) {
    SpeciesDev &S = c.sp[sp];
    if (bins < 2 || bins > HIST_MAX_BINS) { set_error("histograms: binNumber must be 2..2048"); return IPIC_ERR_ARG; }
    if (!(VMax > 0.0) || !(Emin > 0.0) || !(EMax > Emin)) { set_error("histograms: need VMax > 0 and 0 < Emin < EMax"); return IPIC_ERR_ARG; }
    if (S.n <= 0) return IPIC_OK;
    if (S.n >= (1ll << 27)) { set_error("histograms: more than 2^27 particles of one species per GPU"); return IPIC_ERR_UNSUPPORTED; }
    unsigned long long *d;
    int rc = diag_scratch(c, (size_t)3 * bins, &d); if (rc) return rc;
    IPIC_CUDA(cudaMemsetAsync(d, 0, (size_t)3 * bins * sizeof(unsigned long long), c.stream));
    const double Emin_log = std::log(Emin);
    const double diffE = (std::log(EMax) - std::log(Emin)) / (double)(bins - 1);      // :455 (real(binNumber - 1) is single precision: exact for these integers)
    static bool cfgs[64] = {};
    if (!cfgs[c.device & 63]) {
        IPIC_CUDA(cudaFuncSetAttribute(k_velocity_histograms, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * HIST_MAX_BINS * 8));
        cfgs[c.device & 63] = true;
    }
    const int copies = std::max(1, std::min(HIST_THREADS / 32, (3 * HIST_MAX_BINS) / (3 * bins)));     // as many as fit the 48 KB of the largest histogram
    c.launches += 1;
    k_velocity_histograms<<<c.num_sms * 2, HIST_THREADS, (size_t)copies * 3 * bins * 8, c.stream>>>(S.l, S.vx, S.vy, S.vz, S.n, bins, copies, VMax, Emin, EMax, Emin_log,
                                                                                                    diffE, d);
    IPIC_CUDA(cudaGetLastError());
    std::vector<unsigned long long> h((size_t)3 * bins);
    IPIC_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    for (int k = 0; k < 2 * bins; ++k) VHist[k] += (double)(long long)h[k] / HIST_FX;
    for (int k = 0; k < bins; ++k) EHist[k] += (double)(long long)h[2 * bins + k] / HIST_FX;
    return IPIC_OK;
}

}  // namespace ipic
