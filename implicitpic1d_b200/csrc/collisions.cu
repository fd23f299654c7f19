// Null-collision Monte-Carlo collisions on the device.
//
// Reference semantics restated here (never copied): OpenMP/SharedModules/src/mod_NullCollision.f90
//   nullCollision_constructor :44-108, generateCollision :110-237, ionizationCollisionIsotropic :359-444,
//   elasticExcitCollisionIsotropic :511-562, scatterVector :634-661; mod_targetParticle.f90:42-56.
//
// The reference walks, per OpenMP thread, a serial loop: draw a random collidable particle, collide it, swap it behind the
// collidable range.  Here every collision candidate is one GPU thread:
//   - candidates j = 0 .. numberSelected-1 map to DISTINCT particles through a keyed bijection of [0, numToCollide)
//     (mcc_permute_hd) -- a uniform subset without replacement, no serial dependence, particles stay where they are (the
//     cell-sorted order the push kernel wants is kept);
//   - each candidate owns a counter-based PCG32 stream keyed by (seed, rank, call, j);
//   - ionization products go to a scratch record per candidate; an exclusive scan of the ionization flags gives each pair its
//     slot behind the current particles, so appends are in candidate order: bit-reproducible for a seed;
//   - per-collision tallies (count, incident energy, energy loss) are reduced per block in a fixed tree and then over blocks
//     in a fixed order.
#include "ipic_common.cuh"
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>

namespace ipic {

#define FULL 0xffffffffu
constexpr int MCC_THREADS = 256;
constexpr double m_e_c = 9.1093837015e-31;           // constants.f90:10
constexpr double k_B_c = 1.380649e-23;

struct CollArgs {
    double *l, *vx, *vy, *vz;                  // primary species
    long long N, k;                            // the step's collidable range (permutation domain), numberSelected of this set
    long long j0;                              // candidates of earlier sets of the same primary in this step: this set's are j0 .. j0+k-1
    int bits; unsigned long long key;          // candidate -> particle bijection
    unsigned long long seed, stream0;          // PCG stream of candidate j: (seed, stream0 + j)
    int nC, nE;
    const double *energy, *sigma;              // [nE], [nC][nE]
    double thr[IPIC_MAX_COLLISIONS];
    int type[IPIC_MAX_COLLISIONS];
    double sigmaVMax, reducedMass, sumMass, redMassIon, minE, maxE;
    double primary_mass, target_mass, ion_mass, v_therm;
    double *scr;                               // ionization records [7][k]: l, secondary electron v(3), ion v(3)
    int *flag;                                 // [k] 1 = candidate j ionized
    double *partial;                           // [blocks][nC][3]
};

__device__ __forceinline__ double dot3(const double a[3]) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }

// scatterVector :634-661
__device__ __forceinline__ void scatter_vector(double u[3], const double ev[3], double costheta, double phi) {
    const double sintheta = sqrt(1.0 - costheta * costheta);
    double sinphi, cosphi;
    sincos(phi, &sinphi, &cosphi);
    if (fabs(ev[1]) > fabs(ev[2])) {
        const double vv = sqrt(ev[0] * ev[0] + ev[1] * ev[1]);
        u[0] = ev[0] * costheta + (ev[1] * sinphi + ev[0] * ev[2] * cosphi) / vv * sintheta;
        u[1] = ev[1] * costheta + (-ev[0] * sinphi + ev[1] * ev[2] * cosphi) / vv * sintheta;
        u[2] = ev[2] * costheta - vv * cosphi * sintheta;
    } else {
        const double vv = sqrt(ev[0] * ev[0] + ev[2] * ev[2]);
        u[0] = ev[0] * costheta + (ev[2] * sinphi - ev[1] * ev[0] * cosphi) / vv * sintheta;
        u[1] = ev[1] * costheta + vv * cosphi * sintheta;
        u[2] = ev[2] * costheta - (ev[0] * sinphi + ev[1] * ev[2] * cosphi) / vv * sintheta;
    }
}

__global__ void __launch_bounds__(MCC_THREADS) k_collide(const CollArgs A) {
    const double two_pi = 6.283185307179586476925286766559;
    const long long j = (long long)blockIdx.x * MCC_THREADS + threadIdx.x;
    int cdone = -1;                 // collision that happened (index in the set), -1 = null collision
    double incE = 0.0, loss = 0.0;
    if (j < A.k) {
        const unsigned long long pidx = mcc_permute_hd((unsigned long long)(A.j0 + j), (unsigned long long)A.N, A.bits, A.key);
        unsigned long long st, inc;
        pcg_seed(st, inc, A.seed, A.stream0 + (unsigned long long)j);
        const double loc = A.l[pidx];
        double vin[3] = {A.vx[pidx], A.vy[pidx], A.vz[pidx]};
        // generate3DMaxwellianVelocity  mod_targetParticle.f90:42-56
        double vt[3];
        {
            const double U1 = pcg_uniform(st, inc), U2 = pcg_uniform(st, inc), U3 = pcg_uniform(st, inc), U4 = pcg_uniform(st, inc);
            const double r1 = A.v_therm * sqrt(-2.0 * log(U1)), r3 = A.v_therm * sqrt(-2.0 * log(U3));
            double s2, c2;
            sincos(two_pi * U2, &s2, &c2);
            vt[0] = r1 * c2; vt[1] = r1 * s2; vt[2] = r3 * sin(two_pi * U4);
        }
        double vcm[3] = {vin[0] - vt[0], vin[1] - vt[1], vin[2] - vt[2]};
        const double vcm2 = dot3(vcm);
        const double speedCM = sqrt(vcm2);
        const double energyCM = vcm2 * 0.5 * A.reducedMass / e_c;        // eV
        // energy-table lookup :146-169
        int lo = 0, hi = A.nE - 1;                                       // 0-based
        double d_value;
        if (energyCM < A.minE) d_value = 0.0;
        else if (energyCM > A.maxE) { lo = hi - 1; d_value = 1.0; }
        else {
            while (lo != hi - 1) {
                const int mid = (lo + hi) >> 1;
                const double em = __ldg(A.energy + mid);
                if (em < energyCM) lo = mid;
                else if (em > energyCM) hi = mid;
                else { lo = mid; hi = lo + 1; }
            }
            const double elo = __ldg(A.energy + lo);
            d_value = (energyCM - elo) / (__ldg(A.energy + hi) - elo);
        }
        const double Rand = pcg_uniform(st, inc);
        double sigma_v_low = 0.0;
        bool ionized = false;
        double v2nd[3] = {0, 0, 0};
        for (int c = 0; c < A.nC; ++c) {
            if (energyCM > A.thr[c]) {
                const double *sg = A.sigma + (size_t)c * A.nE;
                const double sigma_v = (__ldg(sg + lo) * (1.0 - d_value) + __ldg(sg + hi) * d_value) * speedCM + sigma_v_low;
                if (Rand <= sigma_v / A.sigmaVMax) {
                    cdone = c;
                    incE = dot3(vin);
                    const int ty = A.type[c];
                    if (ty == 1 || ty == 3 || ty == 2) {
                        // elasticExcitCollisionIsotropic :511-562 / ionizationCollisionIsotropic :359-444
                        const double delE = (energyCM - A.thr[c]) * e_c;
                        double P0[3], Vc[3], ev[3];
#pragma unroll
                        for (int q = 0; q < 3; ++q) { P0[q] = A.primary_mass * vin[q] + A.target_mass * vt[q]; Vc[q] = P0[q] / A.sumMass; }
                        const double cos_theta = 1.0 - 2.0 * pcg_uniform(st, inc);
                        const double phi = pcg_uniform(st, inc) * two_pi;
                        const double sin_theta = sqrt(1.0 - cos_theta * cos_theta);
                        double sin_phi, cos_phi;
                        sincos(phi, &sin_phi, &cos_phi);
                        ev[0] = cos_theta; ev[1] = sin_phi * sin_theta; ev[2] = cos_phi * sin_theta;
                        const double mu = (ty == 2) ? A.redMassIon : A.reducedMass;
                        const double speed = sqrt(2.0 * delE * mu / (A.primary_mass * A.primary_mass));
#pragma unroll
                        for (int q = 0; q < 3; ++q) vin[q] = ev[q] * speed + Vc[q];
                        if (ty == 2) {
                            // second electron at 120 degrees from the first in the centre-of-mass frame, ion takes the rest
                            double uv[3];
                            const double phi2 = pcg_uniform(st, inc) * two_pi;
                            scatter_vector(uv, ev, cos(two_pi / 3.0), phi2);
                            const double speed2 = sqrt(2.0 * delE * A.redMassIon / (m_e_c * m_e_c));
#pragma unroll
                            for (int q = 0; q < 3; ++q) v2nd[q] = uv[q] * speed2 + Vc[q];
#pragma unroll
                            for (int q = 0; q < 3; ++q) vt[q] = (P0[q] - A.primary_mass * vin[q] - m_e_c * v2nd[q]) / A.ion_mass;
                            ionized = true;
                            loss = loss - dot3(vt) * A.target_mass - dot3(v2nd) * A.primary_mass;     // :208
                        }
                    } else if (ty == 4) {
                        vin[0] = vt[0]; vin[1] = vt[1]; vin[2] = vt[2];                               // charge exchange :212-213
                    }
                    loss = loss + A.primary_mass * (incE - dot3(vin));                                 // :216
                    break;
                }
                sigma_v_low = sigma_v;
            }
        }
        if (cdone >= 0) { A.vx[pidx] = vin[0]; A.vy[pidx] = vin[1]; A.vz[pidx] = vin[2]; }
        if (A.flag) {
            A.flag[j] = ionized ? 1 : 0;
            if (ionized) {
                const size_t k = (size_t)A.k;
                A.scr[j] = loc;
                A.scr[k + j] = v2nd[0]; A.scr[2 * k + j] = v2nd[1]; A.scr[3 * k + j] = v2nd[2];
                A.scr[4 * k + j] = vt[0]; A.scr[5 * k + j] = vt[1]; A.scr[6 * k + j] = vt[2];
            }
        }
    }
    // ---- tallies: fixed tree inside the block, one partial per block per collision ----
    __shared__ double sh[MCC_THREADS / 32][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < A.nC; ++c) {
        double v[3] = {cdone == c ? 1.0 : 0.0, cdone == c ? incE : 0.0, cdone == c ? loss : 0.0};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(FULL, v[q], o);
        }
        if (lane == 0) { sh[warp][0] = v[0]; sh[warp][1] = v[1]; sh[warp][2] = v[2]; }
        __syncthreads();
        if (threadIdx.x < 3) {
            double x = 0.0;
            for (int w = 0; w < MCC_THREADS / 32; ++w) x += sh[w][threadIdx.x];
            A.partial[((size_t)blockIdx.x * A.nC + c) * 3 + threadIdx.x] = x;
        }
        __syncthreads();
    }
}

// out[v] = sum over blocks of partial[b][v], fixed order: one block per value
__global__ void __launch_bounds__(256) k_mcc_final(const double *partial, long long nblocks, int nvals, double *out) {
    __shared__ double sh[256];
    const int v = blockIdx.x;
    double x = 0.0;
    for (long long b = threadIdx.x; b < nblocks; b += 256) x += partial[(size_t)b * nvals + v];
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = sh[0];
}

// appends the ionization products in candidate order: electron to the primary species, ion to the ion species
__global__ void __launch_bounds__(256) k_mcc_append(const double *__restrict__ scr, const int *__restrict__ flag, const int *__restrict__ offs,
                                                    long long k, double *el, double *evx, double *evy, double *evz, long long n_e, double *il,
                                                    double *ivx, double *ivy, double *ivz, long long n_i, long long *total) {
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    if (j >= k) return;
    const int f = flag[j], o = offs[j];
    if (f) {
        const double loc = scr[j];
        el[n_e + o] = loc; evx[n_e + o] = scr[k + j]; evy[n_e + o] = scr[2 * k + j]; evz[n_e + o] = scr[3 * k + j];
        il[n_i + o] = loc; ivx[n_i + o] = scr[4 * k + j]; ivy[n_i + o] = scr[5 * k + j]; ivz[n_i + o] = scr[6 * k + j];
    }
    if (j == k - 1) *total = (long long)o + f;
}

// nullCollision_constructor :44-108 (collisions ordered by their maximum sigma*v, largest first)
int collision_set_create(Ctx &c, int nC, int nE, double red_mass, double sumMass, double redMassTriple, const double *energyArray,
                         const double *sigmaArray, const double *energyThreshold, const int32_t *collisionType, const int32_t *reactantsIndx,
                         const int32_t *productsIndx, unsigned long long seed, int *out) {
    if (nC < 1 || nC > IPIC_MAX_COLLISIONS || nE < 2) { set_error("null collision: numberCollisions must be 1..16, lengthArrays >= 2"); return IPIC_ERR_ARG; }
    const int prim = reactantsIndx[0] - 1, targ = reactantsIndx[1] - 1;
    if (prim < 0 || prim >= (int)c.sp.size() || targ < 0 || targ >= (int)c.targets.size()) { set_error("null collision: bad reactant index"); return IPIC_ERR_ARG; }
    CollSet S;
    S.nC = nC; S.nE = nE; S.primary = prim; S.target = targ;
    S.reducedMass = red_mass; S.sumMass = sumMass; S.redMassIon = redMassTriple; S.seed = seed;
    if (c.sp[prim].coll_seed == 0) c.sp[prim].coll_seed = seed ^ 0x2545F4914F6CDD1DULL;
    auto SIG = [&](int j, int i) { return sigmaArray[(size_t)j + (size_t)nE * (size_t)i]; };   // 0-based (lengthArrays, numberCollisions)
    std::vector<double> svmax(nC, 0.0);
    for (int i = 0; i < nC; ++i)
        for (int j = 0; j < nE; ++j) svmax[i] = std::max(svmax[i], SIG(j, i) * std::sqrt(2.0 * energyArray[j] * e_c / red_mass));
    // indexSortArray (ascending; straight insertion below 8 entries, mod_BasicFunctions.f90:336-345), then reversed (:66-70)
    std::vector<int> order(nC);
    for (int i = 0; i < nC; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return svmax[a] < svmax[b]; });
    std::reverse(order.begin(), order.end());
    std::vector<double> sig((size_t)nC * nE);
    int ionSp = -1;
    for (int i = 0; i < nC; ++i) {
        const int j = order[i];
        for (int k = 0; k < nE; ++k) sig[(size_t)i * nE + k] = SIG(k, j);
        S.thr[i] = energyThreshold[j]; S.type[i] = collisionType[j];
        S.ionSpecies[i] = -1;
        if (collisionType[j] == 2) {
            const int isp = productsIndx[1 + 3 * j] - 1;   // productsIndx(2, j): the ion (:195-207)
            if (isp < 0 || isp >= (int)c.sp.size()) { set_error("null collision: ionization product species out of range"); return IPIC_ERR_ARG; }
            if (ionSp >= 0 && ionSp != isp) { set_error("null collision: one ion species per collision set"); return IPIC_ERR_UNSUPPORTED; }
            ionSp = isp; S.ionSpecies[i] = isp; S.has_ionization = true;
        }
    }
    S.sigmaVMax = 0.0;
    for (int j = 0; j < nE; ++j) {
        double sum_sigma = 0.0;
        for (int i = 0; i < nC; ++i) sum_sigma += SIG(j, i);
        S.sigmaVMax = std::max(S.sigmaVMax, sum_sigma * std::sqrt(2.0 * energyArray[j] * e_c / red_mass));
    }
    S.minE = *std::min_element(energyArray, energyArray + nE);
    S.maxE = *std::max_element(energyArray, energyArray + nE);
    IPIC_CUDA(cudaMalloc(&S.d_energy, (size_t)nE * sizeof(double)));
    IPIC_CUDA(cudaMalloc(&S.d_sigma, (size_t)nC * nE * sizeof(double)));
    IPIC_CUDA(cudaMemcpyAsync(S.d_energy, energyArray, (size_t)nE * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    IPIC_CUDA(cudaMemcpyAsync(S.d_sigma, sig.data(), (size_t)nC * nE * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    ipic_collision_stats &st = S.stats;
    st.numberCollisions = nC; st.lengthArrays = nE;
    for (int i = 0; i < nC; ++i) { st.collisionType[i] = S.type[i]; st.energyThreshold[i] = S.thr[i]; }
    st.sigmaVMax = S.sigmaVMax; st.reducedMass = red_mass; st.sumMass = sumMass; st.reducedMassIonization = redMassTriple;
    st.minEnergy = S.minE; st.maxEnergy = S.maxE;
    c.colls.push_back(S);
    *out = (int)c.colls.size() - 1;
    return IPIC_OK;
}

void collisions_free(Ctx &c) {
    for (auto &S : c.colls) { if (S.d_energy) cudaFree(S.d_energy); if (S.d_sigma) cudaFree(S.d_sigma); }
    c.colls.clear();
    if (c.d_mcc) cudaFree(c.d_mcc);
    if (c.d_mcc_tmp) cudaFree(c.d_mcc_tmp);
    c.d_mcc = nullptr; c.d_mcc_tmp = nullptr; c.cap_mcc = 0; c.cap_mcc_tmp = 0;
}

static int grow_species(Ctx &c, int sp, long long need) {
    SpeciesDev &s = c.sp[sp];
    if (need + 8192 <= s.cap) return IPIC_OK;
    return ensure_capacity(c, sp, need + need / 8 + 8192);   // slack: ionization grows the arrays a little every step
}

// generateCollision :110-237
int generate_collision(Ctx &c, int set, double del_t) {
    CollSet &S = c.colls[set];
    SpeciesDev &P = c.sp[S.primary];
    const TargetDev &tg = c.targets[S.target];
    ipic_collision_stats &st = S.stats;
    const double P_null = 1.0 - std::exp(-S.sigmaVMax * tg.density * del_t);
    st.P_null = P_null;
    st.numberSelected = 0; st.ionizations = 0;
    if (P_null > 0.5) { set_error("P_null greater than 50%"); return IPIC_ERR_STOP; }      // :126-129
    const long long Nrange = (P.n_collide < 0) ? P.n : std::min(P.n_collide, P.n);     // the step's collidable range (permutation domain)
    const long long taken = std::min(P.coll_taken, Nrange);
    const long long N = Nrange - taken;                                                  // numToCollide as this set sees it (:221-231)
    st.totalNumberCollidableParticles += N;
    const unsigned long long seed = S.seed + 0x632BE59BD9B4E019ULL * (unsigned long long)(c.rank + 1);
    const unsigned long long call = S.calls++;
    // numberSelected :137-140 (one draw for the whole shard instead of one per OpenMP slab)
    const double numberSelectedReal = P_null * (double)N;
    long long k = (long long)numberSelectedReal;
    {
        unsigned long long stt, inc;
        pcg_seed(stt, inc, seed, (call << 36) | 0xFFFFFFFFFULL);
        if (pcg_uniform(stt, inc) < (numberSelectedReal - (double)k)) k += 1;
    }
    if (k > N) k = N;
    st.numberSelected = k;
    if (k <= 0) return IPIC_OK;
    if (k >= (1ll << 31)) { set_error("null collision: more than 2^31-1 candidates"); return IPIC_ERR_UNSUPPORTED; }
    int ionSp = -1;
    for (int i = 0; i < S.nC; ++i) if (S.ionSpecies[i] >= 0) ionSp = S.ionSpecies[i];
    if (S.has_ionization) {
        int rc = grow_species(c, S.primary, P.n + k);
        if (rc) return rc;
        rc = grow_species(c, ionSp, c.sp[ionSp].n + k);
        if (rc) return rc;
    }
    const long long blocks = (k + MCC_THREADS - 1) / MCC_THREADS;
    const int nvals = S.nC * 3;
    // scratch: [partial blocks*nvals][out nvals][total 1] doubles, then [7k] records, then [2k] ints
    const size_t n_dbl = (size_t)blocks * nvals + nvals + 2 + (S.has_ionization ? (size_t)7 * k : 0);
    const size_t bytes = n_dbl * sizeof(double) + (S.has_ionization ? (size_t)2 * k * sizeof(int) : 0) + 64;
    if ((long long)bytes > c.cap_mcc) {
        if (c.d_mcc) cudaFree(c.d_mcc);
        c.cap_mcc = (long long)(bytes + bytes / 4);
        IPIC_CUDA(cudaMalloc(&c.d_mcc, (size_t)c.cap_mcc));
    }
    double *partial = c.d_mcc, *outv = partial + (size_t)blocks * nvals;
    long long *d_total = reinterpret_cast<long long *>(outv + nvals);
    double *scr = outv + nvals + 2;
    int *flag = S.has_ionization ? reinterpret_cast<int *>(scr + (size_t)7 * k) : nullptr;
    int *offs = S.has_ionization ? flag + k : nullptr;

    CollArgs A{};
    A.l = P.l; A.vx = P.vx; A.vy = P.vy; A.vz = P.vz;
    A.N = Nrange; A.k = k; A.j0 = taken;
    A.bits = mcc_bits_hd((unsigned long long)Nrange);
    // one permutation per primary species and step: sets that share the primary walk through it one after the other
    A.key = (P.coll_seed + 0x632BE59BD9B4E019ULL * (unsigned long long)(c.rank + 1)) ^ (0xA0761D6478BD642FULL * (P.coll_epoch + 1));
    P.coll_taken = taken + k;
    A.seed = seed; A.stream0 = call << 36;
    A.nC = S.nC; A.nE = S.nE; A.energy = S.d_energy; A.sigma = S.d_sigma;
    for (int i = 0; i < S.nC; ++i) { A.thr[i] = S.thr[i]; A.type[i] = S.type[i]; }
    A.sigmaVMax = S.sigmaVMax; A.reducedMass = S.reducedMass; A.sumMass = S.sumMass; A.redMassIon = S.redMassIon; A.minE = S.minE; A.maxE = S.maxE;
    A.primary_mass = P.mass; A.target_mass = tg.mass; A.ion_mass = (ionSp >= 0) ? c.sp[ionSp].mass : 0.0; A.v_therm = tg.v_therm;
    A.scr = S.has_ionization ? scr : nullptr; A.flag = flag; A.partial = partial;
    c.launches += 1; k_collide<<<(unsigned)blocks, MCC_THREADS, 0, c.stream>>>(A);
    c.launches += 1; k_mcc_final<<<nvals, 256, 0, c.stream>>>(partial, blocks, nvals, outv);
    IPIC_CUDA(cudaGetLastError());
    long long nion = 0;
    if (S.has_ionization) {
        SpeciesDev &I = c.sp[ionSp];
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, flag, offs, (int)k, c.stream);
        if (tmp > c.cap_mcc_tmp) {
            if (c.d_mcc_tmp) cudaFree(c.d_mcc_tmp);
            c.cap_mcc_tmp = tmp + tmp / 4 + 256;
            IPIC_CUDA(cudaMalloc(&c.d_mcc_tmp, c.cap_mcc_tmp));
        }
        c.launches += 1; cub::DeviceScan::ExclusiveSum(c.d_mcc_tmp, tmp, flag, offs, (int)k, c.stream);
        c.launches += 1; k_mcc_append<<<(unsigned)((k + 255) / 256), 256, 0, c.stream>>>(scr, flag, offs, k, P.l, P.vx, P.vy, P.vz, P.n, I.l, I.vx, I.vy, I.vz, I.n,
                                                                                    d_total);
        IPIC_CUDA(cudaGetLastError());
        IPIC_CUDA(cudaMemcpyAsync(&nion, d_total, sizeof(long long), cudaMemcpyDeviceToHost, c.stream));
    }
    std::vector<double> hv(nvals);
    IPIC_CUDA(cudaMemcpyAsync(hv.data(), outv, nvals * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < S.nC; ++i) {
        st.totalAmountCollisions[i] += (int64_t)llround(hv[i * 3]);
        st.totalIncidentEnergy[i] += hv[i * 3 + 1];
        st.totalEnergyLoss[i] += hv[i * 3 + 2];
    }
    if (nion > 0) {
        SpeciesDev &I = c.sp[ionSp];
        P.n += nion; I.n += nion;                          // N_p = N_p + 1 for both (:196-197); the parents' numToCollide is untouched
        if (!P.slab_off.empty()) P.slab_off.back() = P.n;
        if (!I.slab_off.empty()) I.slab_off.back() = I.n;
        st.ionizations = nion;
    }
    return IPIC_OK;
}

}  // namespace ipic
