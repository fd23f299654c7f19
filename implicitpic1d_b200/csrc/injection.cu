// Wall refluxing / particle injection / Maxwellian heating on the device.
//
// Reference semantics restated here (never copied): OpenMP/SharedModules/src/mod_particleInjection.f90
//   addMaxwellianLostParticles :16-40, refluxParticles :42-69, injectAtBoundary :71-130, injectUniformFlux :132-158,
//   maxwellianHeating :160-177; samplers mod_BasicFunctions.f90: getMaxwellianSample :146-160, getMaxwellianFluxSample :162-176.
//
// The reference runs these per OpenMP thread on its slab with that thread's ran2 state.  Here every new / re-drawn particle is one
// GPU thread with its own counter-based PCG32 stream keyed by (seed, rank, call, work item), so a call is bit-reproducible for a
// seed and independent of the launch geometry; the routines consume exactly what the move kernel left behind on the device
// (SUM(delIdx) of the electrons, the post-compaction refRecordIdx lists), so the particles never leave HBM.  New pairs are appended
// behind the current particles (the last thread slab on download), in work-item order.  energyAddColl is reduced per block in a
// fixed tree and over blocks in a fixed order.  Random streams differ from ran2 by design: parity with the oracle is statistical
// (moments of the drawn velocities, positions) and exact in the bookkeeping (counts, which particles change, signs, energy tally).
#include "ipic_common.cuh"
#include <cmath>

namespace ipic {

#define FULL 0xffffffffu
constexpr int INJ_THREADS = 256;
constexpr double TWO_PI = 6.283185307179586476925286766559;

// getLFromX: NGP mod_domain.f90:251-265 (binary search on the nodes), CIC mod_domain.f90 (search on the cell faces)
__device__ __forceinline__ double l_from_x(const double *__restrict__ grid, const double *__restrict__ dx, int Nx, int cic, double x) {
    int lo = 1, hi = Nx;   // 0-based arrays: grid[k-1] = grid(k)
    if (!cic) {
        while (lo != hi - 1) { const int mid = (lo + hi) / 2; if (grid[mid - 1] <= x) lo = mid; else hi = mid; }
        return lo + (x - grid[lo - 1]) / dx[lo - 1];
    }
    while (lo != hi - 1) { const int mid = (lo + hi) / 2; if (grid[mid - 1] - 0.5 * dx[mid - 1] <= x) lo = mid; else hi = mid; }
    if (grid[lo - 1] + 0.5 * dx[lo - 1] <= x) lo = hi;
    return lo + 0.5 + (x - grid[lo - 1]) / dx[lo - 1];
}

// getMaxwellianSample  mod_BasicFunctions.f90:146-160 (v_z uses sin(2 pi U4) with U3's radius, as the reference)
__device__ __forceinline__ void maxwellian_sample(double v[3], double v_therm, unsigned long long &st, unsigned long long inc) {
    const double U1 = pcg_uniform(st, inc), U2 = pcg_uniform(st, inc), U3 = pcg_uniform(st, inc), U4 = pcg_uniform(st, inc);
    const double r1 = v_therm * sqrt(-2.0 * log(U1));
    double s2, c2;
    sincos(TWO_PI * U2, &s2, &c2);
    v[0] = r1 * c2; v[1] = r1 * s2;
    v[2] = v_therm * sqrt(-2.0 * log(U3)) * sin(TWO_PI * U4);
}
// getMaxwellianFluxSample  :162-176
__device__ __forceinline__ void maxwellian_flux_sample(double v[3], double v_therm, unsigned long long &st, unsigned long long inc) {
    const double U1 = pcg_uniform(st, inc), U2 = pcg_uniform(st, inc), U3 = pcg_uniform(st, inc), U4 = pcg_uniform(st, inc);
    v[0] = v_therm * sqrt(-2.0 * log(U1)) * ((U2 - 0.5 >= 0.0) ? 1.0 : -1.0);
    const double r3 = v_therm * sqrt(-2.0 * log(U3));
    double s4, c4;
    sincos(TWO_PI * U4, &s4, &c4);
    v[1] = r3 * c4; v[2] = r3 * s4;
}
__device__ __forceinline__ double sumsq(const double v[3]) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }

// block sum in a fixed tree; thread 0 holds the result
__device__ __forceinline__ double block_sum(double x) {
    __shared__ double sh[INJ_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(FULL, x, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < INJ_THREADS / 32; ++w) r += sh[w];
    __syncthreads();
    return r;
}

struct PairArgs {
    double *el, *evx, *evy, *evz, *il, *ivx, *ivy, *ivz;    // write positions: first free slot of each species
    long long n;
    unsigned long long seed, stream0;
    const double *grid, *dx; int Nx, cic;
    double startX, endX, L;
    double vth_e, vth_i;           // thermal speeds of the electron / ion draw
    double cs;                     // ion sound speed (boundary injection)
    double del_t;
    int bcL;                       // boundaryConditions(1)
    double half_mw_e, half_mw_i;   // 0.5 * mass * w_p
    double *partial;               // [blocks] energy added
};

// MODE 0 addMaxwellianLostParticles :24-33, 1 injectAtBoundary :108-125, 2 injectUniformFlux :148-154
template <int MODE>
__global__ void __launch_bounds__(INJ_THREADS) k_inject_pairs(const PairArgs A) {
    const long long j = (long long)blockIdx.x * INJ_THREADS + threadIdx.x;
    double en = 0.0;
    if (j < A.n) {
        unsigned long long st, inc;
        pcg_seed(st, inc, A.seed, A.stream0 + (unsigned long long)j);
        double ve[3], vi[3], x;
        if (MODE == 0) {
            maxwellian_sample(ve, A.vth_e, st, inc);
            maxwellian_flux_sample(vi, A.vth_i, st, inc);
            x = pcg_uniform(st, inc) * A.L + A.startX;
        } else if (MODE == 1) {
            const double shift_rand = A.del_t * pcg_uniform(st, inc) * A.cs;
            maxwellian_flux_sample(ve, A.vth_e, st, inc);
            vi[0] = A.cs; vi[1] = 0.0; vi[2] = 0.0;
            if (A.bcL == 2) { ve[0] = fabs(ve[0]); vi[0] = fabs(vi[0]); x = A.startX + shift_rand; }
            else { ve[0] = -fabs(ve[0]); vi[0] = -fabs(vi[0]); x = A.endX - shift_rand; }
        } else {
            maxwellian_sample(ve, A.vth_e, st, inc);       // the caller passes v_therm of 20 T_e (:149)
            maxwellian_sample(vi, A.vth_i, st, inc);
            x = pcg_uniform(st, inc) * A.L + A.startX;
        }
        const double lpos = l_from_x(A.grid, A.dx, A.Nx, A.cic, x);
        A.el[j] = lpos; A.evx[j] = ve[0]; A.evy[j] = ve[1]; A.evz[j] = ve[2];
        A.il[j] = lpos; A.ivx[j] = vi[0]; A.ivy[j] = vi[1]; A.ivz[j] = vi[2];
        en = sumsq(ve) * A.half_mw_e + sumsq(vi) * A.half_mw_i;     // :29-30 (times 0.5 folded into half_mw)
    }
    const double s = block_sum(en);
    if (threadIdx.x == 0) A.partial[blockIdx.x] = s;
}

struct RedrawArgs {
    double *vx, *vy, *vz;
    const unsigned *list; long long n;     // refRecordIdx (post-compaction, 0-based)
    const unsigned char *live;             // 0: the reflected particle was absorbed later in the step -- the entry names another particle: skipped
    long long n_alive;                     // entries >= N_p name no particle (reflected, then absorbed, last of the array)
    unsigned long long seed, stream0;
    double vth, cs, half_mw;
    int bcL, mode, ion;                    // mode 0 refluxParticles :51-64, 1 injectAtBoundary :82-98
    double *partial;
};
__global__ void __launch_bounds__(INJ_THREADS) k_redraw_reflected(const RedrawArgs A) {
    const long long j = (long long)blockIdx.x * INJ_THREADS + threadIdx.x;
    double en = 0.0;
    if (j < A.n) {
        const unsigned idx = A.list[j];
        // entries are distinct particles once the absorbed ones are dropped (two threads never share an idx)
        if ((long long)idx < A.n_alive && (!A.live || A.live[j])) {
        unsigned long long st, inc;
        pcg_seed(st, inc, A.seed, A.stream0 + (unsigned long long)j);
        double v[3] = {A.vx[idx], A.vy[idx], A.vz[idx]};
        en = -sumsq(v) * A.half_mw;
        if (A.mode == 1 && A.ion) { v[0] = A.cs; v[1] = 0.0; v[2] = 0.0; }
        else maxwellian_flux_sample(v, A.vth, st, inc);
        v[0] = (A.bcL == 2) ? fabs(v[0]) : -fabs(v[0]);
        en += sumsq(v) * A.half_mw;
        A.vx[idx] = v[0]; A.vy[idx] = v[1]; A.vz[idx] = v[2];
        }
    }
    const double s = block_sum(en);
    if (threadIdx.x == 0) A.partial[blockIdx.x] = s;
}

// maxwellianHeating :160-177: every particle draws R; R < compValue -> velocity redrawn from a Maxwellian at T
__global__ void __launch_bounds__(INJ_THREADS) k_maxwellian_heating(double *vx, double *vy, double *vz, const double *__restrict__ l, long long n,
                                                                    unsigned long long seed, unsigned long long stream0, double compValue, double vth,
                                                                    unsigned long long *count) {
    const long long stride = (long long)gridDim.x * INJ_THREADS;
    unsigned mine = 0;
    for (long long i = (long long)blockIdx.x * INJ_THREADS + threadIdx.x; i < n; i += stride) {
        unsigned long long st, inc;
        pcg_seed(st, inc, seed, stream0 + (unsigned long long)i);
        const double R = pcg_uniform(st, inc);
        if (R < compValue && l[i] >= 0.0) {
            double v[3];
            maxwellian_sample(v, vth, st, inc);
            vx[i] = v[0]; vy[i] = v[1]; vz[i] = v[2];
            mine += 1;
        }
    }
    mine = __reduce_add_sync(FULL, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, (unsigned long long)mine);     // integer: order-independent
}

// out[0] = sum of partial[0..n) in a fixed order (one block)
__global__ void __launch_bounds__(INJ_THREADS) k_inj_final(const double *partial, long long n, double *out) {
    double x = 0.0;
    for (long long b = threadIdx.x; b < n; b += INJ_THREADS) x += partial[b];
    const double s = block_sum(x);
    if (threadIdx.x == 0) out[0] = s;
}

// ---- host side ------------------------------------------------------------------------------------------
static void domain_extent(const Ctx &c, double &startX, double &endX) {
    if (c.scheme != IPIC_SCHEME_CIC) { startX = c.h_grid[0]; endX = c.h_grid[c.Nx - 1]; }
    else { startX = c.h_grid[0] - 0.5 * c.h_dx[0]; endX = c.h_grid[c.Nx - 1] + 0.5 * c.h_dx[c.Nx - 1]; }
}
static int inj_scratch(Ctx &c, long long blocks, double **partial, double **out) {
    const long long need = (blocks + 8) * (long long)sizeof(double);
    if (need > c.cap_inj) {
        if (c.d_inj) cudaFree(c.d_inj);
        c.cap_inj = need + need / 4 + 4096;
        IPIC_CUDA(cudaMalloc(&c.d_inj, (size_t)c.cap_inj));
    }
    *out = c.d_inj; *partial = c.d_inj + 8;
    return IPIC_OK;
}
void injection_free(Ctx &c) {
    if (c.d_inj) cudaFree(c.d_inj);
    c.d_inj = nullptr; c.cap_inj = 0;
}
static unsigned long long inj_seed(const Ctx &c) { return c.inj_seed + 0x632BE59BD9B4E019ULL * (unsigned long long)(c.rank + 1); }
static int need_pair(Ctx &c) {
    if (c.sp.size() < 2) { set_error("particle injection: needs particleList(1) = electrons and particleList(2) = ions"); return IPIC_ERR_ARG; }
    return IPIC_OK;
}
// draws numFluxParticlesLow / High (:100-104, :143-147): one draw for the shard instead of one per OpenMP slab, same expectation
static long long draw_count(Ctx &c, double numFluxPart, unsigned long long call) {
    if (!(numFluxPart > 0.0)) return 0;
    long long k = (long long)std::floor(numFluxPart);
    unsigned long long st, inc;
    pcg_seed(st, inc, inj_seed(c), (call << 36) | 0xFFFFFFFFFULL);
    if (!(pcg_uniform(st, inc) > numFluxPart - (double)k)) k += 1;
    return k;
}

static int append_pairs(Ctx &c, int mode, long long n, double vth_e, double vth_i, double cs, double del_t, unsigned long long call, double *energyAdd) {
    if (energyAdd) *energyAdd = 0.0;
    if (n <= 0) return IPIC_OK;
    if (n >= (1ll << 31)) { set_error("particle injection: more than 2^31-1 pairs in one call"); return IPIC_ERR_UNSUPPORTED; }
    int rc = IPIC_OK;
    for (int j = 0; j < 2; ++j) {
        const long long need = c.sp[j].n + n;
        if (need + 8192 > c.sp[j].cap) rc = ensure_capacity(c, j, need + need / 8 + 8192);   // slack: injection grows the arrays every step
        if (rc) return rc;
    }
    SpeciesDev &E = c.sp[0], &I = c.sp[1];
    const long long blocks = (n + INJ_THREADS - 1) / INJ_THREADS;
    double *partial, *out;
    rc = inj_scratch(c, blocks, &partial, &out); if (rc) return rc;
    PairArgs A{};
    A.el = E.l + E.n; A.evx = E.vx + E.n; A.evy = E.vy + E.n; A.evz = E.vz + E.n;
    A.il = I.l + I.n; A.ivx = I.vx + I.n; A.ivy = I.vy + I.n; A.ivz = I.vz + I.n;
    A.n = n; A.seed = inj_seed(c); A.stream0 = call << 36;
    A.grid = c.d_grid; A.dx = c.d_dx; A.Nx = c.Nx; A.cic = (c.scheme == IPIC_SCHEME_CIC) ? 1 : 0;
    domain_extent(c, A.startX, A.endX); A.L = A.endX - A.startX;
    A.vth_e = vth_e; A.vth_i = vth_i; A.cs = cs; A.del_t = del_t; A.bcL = c.h_bc[0];
    A.half_mw_e = 0.5 * E.mass * E.w_p; A.half_mw_i = 0.5 * I.mass * I.w_p;
    A.partial = partial;
    c.launches += 2;
    if (mode == 0) k_inject_pairs<0><<<(unsigned)blocks, INJ_THREADS, 0, c.stream>>>(A);
    else if (mode == 1) k_inject_pairs<1><<<(unsigned)blocks, INJ_THREADS, 0, c.stream>>>(A);
    else k_inject_pairs<2><<<(unsigned)blocks, INJ_THREADS, 0, c.stream>>>(A);
    k_inj_final<<<1, INJ_THREADS, 0, c.stream>>>(partial, blocks, out);
    IPIC_CUDA(cudaGetLastError());
    double h = 0.0;
    IPIC_CUDA(cudaMemcpyAsync(&h, out, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    if (energyAdd) *energyAdd = h;
    E.n += n; I.n += n;                                     // N_p = N_p + delIdx / numInjected (:35-36, :126-127, :155-156)
    if (!E.slab_off.empty()) E.slab_off.back() = E.n;
    if (!I.slab_off.empty()) I.slab_off.back() = I.n;
    return IPIC_OK;
}

static int redraw_reflected(Ctx &c, int sp, int mode, double vth, double cs, unsigned long long call, double *energy) {
    SpeciesDev &S = c.sp[sp];
    *energy = 0.0;
    const long long n = S.nRef;
    if (n <= 0 || !S.refList) return IPIC_OK;
    const long long blocks = (n + INJ_THREADS - 1) / INJ_THREADS;
    double *partial, *out;
    int rc = inj_scratch(c, blocks, &partial, &out); if (rc) return rc;
    RedrawArgs A{};
    A.vx = S.vx; A.vy = S.vy; A.vz = S.vz; A.list = S.refList; A.live = S.refLive; A.n = n; A.n_alive = S.n;
    A.seed = inj_seed(c); A.stream0 = (call << 36) | ((unsigned long long)(sp + 1) << 32);
    A.vth = vth; A.cs = cs; A.half_mw = 0.5 * S.mass * S.w_p; A.bcL = c.h_bc[0]; A.mode = mode; A.ion = (sp == 1);
    A.partial = partial;
    c.launches += 2;
    k_redraw_reflected<<<(unsigned)blocks, INJ_THREADS, 0, c.stream>>>(A);
    k_inj_final<<<1, INJ_THREADS, 0, c.stream>>>(partial, blocks, out);
    IPIC_CUDA(cudaGetLastError());
    IPIC_CUDA(cudaMemcpyAsync(energy, out, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    return IPIC_OK;
}

// addMaxwellianLostParticles :16-40: one electron-ion pair per electron absorbed by the last moveParticles
int add_maxwellian_lost_particles(Ctx &c, double T_e, double T_i, double *energyAdd, long long *numAdded) {
    int rc = need_pair(c); if (rc) return rc;
    const long long n = c.sp[0].stats.delIdx;
    if (numAdded) *numAdded = n;
    const unsigned long long call = c.inj_calls++;
    return append_pairs(c, 0, n, std::sqrt(T_e * e_c / c.sp[0].mass), std::sqrt(T_i * e_c / c.sp[1].mass), 0.0, 0.0, call, energyAdd);
}
// refluxParticles :42-69
int reflux_particles(Ctx &c, double T_e, double T_i, double *energyAdd) {
    int rc = need_pair(c); if (rc) return rc;
    const unsigned long long call = c.inj_calls++;
    double sum = 0.0;
    for (int j = 0; j < 2; ++j) {
        double en = 0.0;
        rc = redraw_reflected(c, j, 0, std::sqrt((j == 0 ? T_e : T_i) * e_c / c.sp[j].mass), 0.0, call, &en);
        if (rc) return rc;
        sum += en;
    }
    if (energyAdd) *energyAdd = sum;
    return IPIC_OK;
}
// injectAtBoundary :71-130 (energyAddColl is not touched there: the tally lines are commented out in the reference)
int inject_at_boundary(Ctx &c, double T_e, double T_i, double del_t, double numFluxPart, long long *numInjected) {
    (void)T_i;
    int rc = need_pair(c); if (rc) return rc;
    const unsigned long long call = c.inj_calls++;
    const double cs = std::sqrt(T_e * e_c / c.sp[1].mass);               // ionSoundSpeed :78
    const double vth_e = std::sqrt(T_e * e_c / c.sp[0].mass);
    for (int j = 0; j < 2; ++j) {
        double en;
        rc = redraw_reflected(c, j, 1, vth_e, cs, call, &en);
        if (rc) return rc;
    }
    const long long n = draw_count(c, numFluxPart, call);
    if (numInjected) *numInjected = n;
    return append_pairs(c, 1, n, vth_e, 0.0, cs, del_t, call, nullptr);
}
// injectUniformFlux :132-158
int inject_uniform_flux(Ctx &c, double T_e, double T_i, double numFluxPart, long long *numInjected) {
    int rc = need_pair(c); if (rc) return rc;
    const unsigned long long call = c.inj_calls++;
    const long long n = draw_count(c, numFluxPart, call);
    if (numInjected) *numInjected = n;
    return append_pairs(c, 2, n, std::sqrt(20.0 * T_e * e_c / c.sp[0].mass), std::sqrt(T_i * e_c / c.sp[1].mass), 0.0, 0.0, call, nullptr);
}
// maxwellianHeating :160-177 with compValue = FracFreq * (currDel_t / del_t) / FractionFreqHeating
int maxwellian_heating(Ctx &c, int sp, double compValue, double T_eV, long long *numHeated) {
    SpeciesDev &S = c.sp[sp];
    if (numHeated) *numHeated = 0;
    if (S.n <= 0) return IPIC_OK;
    const unsigned long long call = c.inj_calls++;
    double *partial, *out;
    int rc = inj_scratch(c, 1, &partial, &out); if (rc) return rc;
    unsigned long long *cnt = reinterpret_cast<unsigned long long *>(out);
    IPIC_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), c.stream));
    c.launches += 1;
    k_maxwellian_heating<<<c.num_sms * 8, INJ_THREADS, 0, c.stream>>>(S.vx, S.vy, S.vz, S.l, S.n, inj_seed(c), (call << 36) | (1ull << 35), compValue,
                                                                       std::sqrt(T_eV * e_c / S.mass), cnt);
    IPIC_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    IPIC_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    if (numHeated) *numHeated = (long long)h;
    return IPIC_OK;
}

}  // namespace ipic
