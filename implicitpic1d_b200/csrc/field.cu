// Grid-sized ("field") work of the implicit step as single-CTA kernels: everything between two particle
// sweeps stays on the device -- J/rho finalisation, half-time E + per-cell push tables, the tridiagonal
// solves (parallel cyclic reduction with cached level coefficients, or a Thomas sweep in the reference's
// operation order), the Anderson history / mixing and the conservation diagnostics.
//
// Reference semantics restated here (never copied): OpenMP/NGP/src/mod_potentialSolver.f90,
// OpenMP/CIC/src/mod_potentialSolver.f90, OpenMP/{NGP,CIC}/src/mod_domain.f90 (smoothField),
// OpenMP/SharedModules/src/mod_nonLinSolvers.f90:85-160, mod_BasicFunctions.f90:218-265.
#include "ipic_common.cuh"
#include <cmath>
#include <algorithm>

namespace ipic {

constexpr int FT = 1024;   // threads of the field CTA

enum : unsigned {
    OP_FINALIZE_RHO = 1u << 0,
    OP_POISSON = 1u << 1,
    OP_CONST_SOURCE = 1u << 2,
    OP_FINALIZE_J = 1u << 3,
    OP_AMPERE = 1u << 4,
    OP_AND_STORE0 = 1u << 5,   // phi_k(:,1) = phi_f
    OP_AND_FIRST = 1u << 6,    // phi_k(:,2) = phi_f ; R(:,1) = phi_k2 - phi_k1 ; norm
    OP_AND_RESID = 1u << 7,    // R(:,index) = phi_f - phi_k(:,index) ; norm
    OP_AND_MIX = 1u << 8,      // normal equations + next iterate -> phi_f
    OP_HALF_E = 1u << 9,       // EField + push tables
    OP_DIAG = 1u << 10,
    OP_RESID_F = 1u << 11,     // f = x - unknown  (funcNitsol)
    OP_AND_DECIDE = 1u << 17,  // after OP_AND_RESID: convergence / stagnation test on the device (mod_nonLinSolvers.f90:111-134); the ops that
                               // follow in the same launch are chosen by its outcome
    // NGP_Periodic (E-field form), OpenMP/NGP_Periodic/src/mod_potentialSolver.f90 / mod_Scheme.f90
    OP_P_RHO = 1u << 12,       // depositRho epilogue: raw sums, no wall fix-ups, no smoothing
    OP_P_POISSON = 1u << 13,   // solve_tridiag_Poisson + EField = -grad phi, EField_f = EField
    OP_P_FINALIZE_J = 1u << 14,
    OP_P_AMPERE = 1u << 15,    // EField_f = (<J> - J) del_t / eps_0 + EField
    OP_P_HALF_E = 1u << 16,    // EField_half + push tables
};

struct FieldArgs {
    int scheme, Nx, Nh, smooth, nsp, tridiag_mode, pcr_levels, acc_stride, tab_cells, acc_reps;
    unsigned long long acc_rep_stride;
    const double *grid, *dx, *hinv;     // hinv = 0.5 / dx_dl
    const int *bc;
    double *phi, *phi_f, *rho, *J, *E, *a, *b, *c, *cp, *m, *pcr, *work, *scal, *rho_i;
    unsigned long long *acc;
    void *tab;
    double2 *ttab;
    double qw[NSP_MAX], qm[NSP_MAX];
    double del_t, rho_const, RF_voltage, bound1, bound2, src1, src2, poisson_rf_value;
    double *Ef, *Eh;           // NGP_Periodic (E-field form): EField_f, EField_half
    double *unk; int nunk;     // unknown of the nonlinear solve: phi_f (N_x) or EField_f (N_h)
    double L_domain; int globJ;
    double *phi_k, *res_k, *norms;
    int and_i, and_m;          // loop index i and m_Anderson
    double beta;
    int diag_which;
    const double *xin; double *fout;
    // residual norm straight into pinned host memory (UVA: the pinned mirror is device-addressable): value, then a sequence number the
    // host spins on -- no D2H copy, no driver synchronisation on the critical path of a nonlinear iteration
    double *h_norm; unsigned long long *h_seq; unsigned long long seq;
    unsigned long long *counters;   // [nsp][8]: slot 6 of a species is the tile counter of its push launch, zeroed by the half-E op that precedes a sweep
    // thread-block cluster (the fused Anderson iteration): the collected integer grids [nsp][acc_stride / 2] (slice-local prefix of the difference
    // grid added), the slices' difference totals [nsp][16], the per-CTA partial sums of the normal equations [cluster size][68]
    long long *cmp, *cmp_tot; double *gram_part;
    // IPIC_TRIDIAG_SCAN
    const double *gu, *gv, *gs; double gWinv; int giL, giR;
    // OP_AND_DECIDE
    int *and_state; double eps_tol; int eform;
    volatile long long *h_and; double *h_and_norms;
};

// 1-based accessors so that every stencil reads like the reference line it restates
#define GRID(i) A.grid[(i)-1]
#define DX(i) A.dx[(i)-1]
#define BC(i) A.bc[(i)]   /* device bc table is stored 1-based */
#define PHI(i) A.phi[(i)-1]
#define PHIF(i) A.phi_f[(i)-1]
#define RHO(i) A.rho[(i)-1]
#define JJ(i) A.J[(i)-1]
#define EF(i) A.E[(i)-1]

__device__ __forceinline__ int fmodulo(int a, int p) { int r = a % p; return r < 0 ? r + p : r; }

// deterministic block sum (fixed tree); result valid in every thread
__device__ double block_sum(double x) {
    __shared__ double sh[FT / 32];
    __shared__ double tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) sh[warp] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < FT / 32; ++w) s += sh[w];
        tot = s;
    }
    __syncthreads();
    return tot;
}

// smoothField: NGP mod_domain.f90:298-324 ; CIC mod_domain.f90:289-312.  raw/out are 1-based via (i)-1.
__device__ void dev_smooth(const FieldArgs &A, const double *raw, double *out) {
    const int Nh = A.Nh, Nx = A.Nx;
#define RAW(i) raw[(i)-1]
    if (A.scheme == 0) {
        for (int i = 1 + threadIdx.x; i <= Nh; i += FT) {
            const int bv = BC(i + 1) - BC(i);
            double r = 0.0;
            switch (bv) {
            case 0: r = 0.25 * (RAW(i - 1) + 2.0 * RAW(i) + RAW(i + 1)); break;
            case -1: case -4: r = 0.25 * (3.0 * RAW(1) + RAW(2)); break;
            case 1: case 4: r = 0.25 * (3.0 * RAW(Nh) + RAW(Nh - 1)); break;
            case -2: r = 0.25 * (RAW(1) + RAW(2)); break;
            case 2: r = 0.25 * (RAW(Nh) + RAW(Nh - 1)); break;
            case -3: r = 0.25 * (RAW(Nh) + 2.0 * RAW(1) + RAW(2)); break;
            case 3: r = 0.25 * (RAW(Nh - 1) + 2.0 * RAW(Nh) + RAW(1)); break;
            default: r = out[i - 1]; break;
            }
            out[i - 1] = r;
        }
    } else {
        for (int i = 1 + threadIdx.x; i <= Nh; i += FT) {
            double r = 0.0;
            if (i == 1) {
                switch (BC(1)) {
                case 1: case 4: r = 0.25 * (2.0 * RAW(1) + 2.0 * RAW(2)); break;
                case 2: r = 0.0; break;
                case 3: r = 0.25 * (RAW(Nh - 1) + 2.0 * RAW(1) + RAW(2)); break;
                }
            } else if (i == Nh) {
                switch (BC(Nh)) {
                case 1: case 4: r = 0.25 * (2.0 * RAW(Nh) + 2.0 * RAW(Nh - 1)); break;
                case 2: r = 0.0; break;
                case 3: r = 0.25 * (RAW(Nh - 1) + 2.0 * RAW(1) + RAW(2)); break;   // newField(N_h) = newField(1)
                }
            } else {
                r = 0.25 * (RAW(i - 1) + 2.0 * RAW(i) + RAW(i + 1));
            }
            out[i - 1] = r;
        }
    }
#undef RAW
}

// in-place inclusive prefix sum of n int64 values in shared memory (fixed order; integer, so exact)
__device__ void block_scan_i64(long long *x, int n) {
    __shared__ long long wtot[FT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (n + FT - 1) / FT;
    const int b = min(threadIdx.x * per, n), e = min(b + per, n);
    long long run = 0;
    for (int i = b; i < e; ++i) { run += x[i]; x[i] = run; }
    long long inc = run;                                  // inclusive scan of the per-thread totals
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    long long off = inc - run;                            // exclusive within the warp
    for (int w = 0; w < warp; ++w) off += wtot[w];
    for (int i = b; i < e; ++i) x[i] += off;
    __syncthreads();
}

// in-place inclusive prefix sum (or, reverse = true, inclusive suffix sum) of n doubles in shared memory: every thread sums a contiguous
// chunk in order, the chunk totals are scanned over the warp and over the warps in order -- a fixed order, so the result is deterministic
__device__ void block_scan_f64(double *x, int n, bool reverse) {
    __shared__ double wtot[FT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (n + FT - 1) / FT;
    const int b = min(threadIdx.x * per, n), e = min(b + per, n);
    double run = 0.0;
    if (!reverse) { for (int i = b; i < e; ++i) { run += x[i]; x[i] = run; } }
    else { for (int i = b; i < e; ++i) { const int k = n - 1 - i; run += x[k]; x[k] = run; } }
    double inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    double off = __shfl_up_sync(0xffffffffu, inc, 1);     // exclusive within the warp
    if (lane == 0) off = 0.0;
    for (int w = 0; w < warp; ++w) off += wtot[w];
    if (!reverse) { for (int i = b; i < e; ++i) x[i] += off; }
    else { for (int i = b; i < e; ++i) x[n - 1 - i] += off; }
    __syncthreads();
}

// Sum the per-species int64 grids (species order, scaled by q*w_p) into dst[0..n-1]; zero the grids.
// with_runs: add the prefix sum of the species' difference grid (runs of fully traversed cells, see k_push).
// One CTA works through arrays that live in L2: what a phase costs is the number of dependent L2 round trips (~0.7 us each), not bytes.
// The loops below therefore issue the loads of CB cells (all copies of the grid) before the first result is used, and the zeroing
// stores after the last load.
constexpr int CB = 4;
__device__ __forceinline__ void collect_batch(unsigned long long *g, int reps, size_t rep_stride, int base, int n, unsigned long long (&q)[CB]) {
#pragma unroll
    for (int k = 0; k < CB; ++k) q[k] = 0ull;
    if (reps == 4) {
        unsigned long long v[CB][4];
#pragma unroll
        for (int k = 0; k < CB; ++k) {
            const int i = base + k * FT;
#pragma unroll
            for (int r = 0; r < 4; ++r) v[k][r] = (i < n) ? __ldcg(g + (size_t)r * rep_stride + i + 1) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < CB; ++k) {
            const int i = base + k * FT;
            q[k] = (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
            if (i < n) {
#pragma unroll
                for (int r = 0; r < 4; ++r) g[(size_t)r * rep_stride + i + 1] = 0ull;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < CB; ++k) {
            const int i = base + k * FT;
            if (i < n) {
                unsigned long long t = 0ull;
                for (int r = 0; r < reps; ++r) t += __ldcg(g + (size_t)r * rep_stride + i + 1);
                for (int r = 0; r < reps; ++r) g[(size_t)r * rep_stride + i + 1] = 0ull;
                q[k] = t;
            }
        }
    }
}

// ---- thread-block cluster helpers -------------------------------------------------------------------------------------------
// The fused Anderson iteration is launched as a cluster of CLUSTER CTAs (run_field).  CTA 0 runs the kernel as it always did; the phases
// that move the most bytes and are independent per cell -- adding up the copies of the deposition grids, the sums of the normal equations,
// the push tables -- are split over all CTAs of the cluster (one SM's load/store path is what bounds them), with a cluster barrier between
// a phase and its consumer.  Launched without a cluster the same code runs with rank 0 of 1.
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// this CTA's share of the collect: for its slice of cells and every species, the sum of the copies of the grid plus the prefix sum, inside
// the slice, of the sum of the copies of the difference grid -> compact array; the slice's difference total -> cmp_tot[species][rank];
// the copies are zeroed.  `scratch` holds one slice (shared memory).
__device__ void dev_collect_slice(const FieldArgs &A, int n, bool with_runs, unsigned crank, unsigned csize, long long *scratch) {
    const int per = (n + (int)csize - 1) / (int)csize;
    const int lo = min(n, (int)crank * per), hi = min(n, lo + per);
    const int half = A.acc_stride / 2;
    for (int s = 0; s < A.nsp; ++s) {
        unsigned long long *g = A.acc + (size_t)s * A.acc_stride, *d = g + half;
        long long *out = A.cmp + (size_t)s * half;
        if (with_runs) {
            for (int base = lo + threadIdx.x; base < hi; base += CB * FT) {
                unsigned long long q[CB];
                collect_batch(d, A.acc_reps, A.acc_rep_stride, base, hi, q);
#pragma unroll
                for (int k = 0; k < CB; ++k) { const int i = base + k * FT; if (i < hi) scratch[i - lo] = (long long)q[k]; }
            }
            if (crank == csize - 1 && threadIdx.x < A.acc_reps) d[(size_t)threadIdx.x * A.acc_rep_stride + n + 1] = 0ull;   // closing entry of a run
            __syncthreads();
            block_scan_i64(scratch, hi - lo);
            if (threadIdx.x == 0) A.cmp_tot[s * 16 + crank] = (hi > lo) ? scratch[hi - lo - 1] : 0ll;
        }
        for (int base = lo + threadIdx.x; base < hi; base += CB * FT) {
            unsigned long long q[CB];
            collect_batch(g, A.acc_reps, A.acc_rep_stride, base, hi, q);
#pragma unroll
            for (int k = 0; k < CB; ++k) { const int i = base + k * FT; if (i < hi) out[i] = (long long)q[k] + (with_runs ? scratch[i - lo] : 0ll); }
        }
        __syncthreads();
    }
}

// CTA 0 after the barrier: compact arrays + the difference totals of the slices to the left -> dst (species order, scaled by q*w_p)
__device__ void dev_collect_compact(const FieldArgs &A, double *dst, int n, bool with_runs, unsigned csize) {
    const int half = A.acc_stride / 2;
    const int per = (n + (int)csize - 1) / (int)csize;
    __shared__ long long offs[NSP_MAX][16];
    if (threadIdx.x < A.nsp * 16) {
        const int s = threadIdx.x >> 4, k = threadIdx.x & 15;
        long long o = 0;
        if (with_runs) for (int j = 0; j < k && j < (int)csize; ++j) o += __ldcg(A.cmp_tot + s * 16 + j);
        offs[s][k] = o;
    }
    __syncthreads();
    for (int base = threadIdx.x; base < n; base += CB * FT) {
        long long v[NSP_MAX > 2 ? 2 : NSP_MAX][CB];
        // two species at a time, loads together
        for (int s0_ = 0; s0_ < A.nsp; s0_ += 2) {
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                const int i = base + k * FT;
                v[0][k] = (i < n) ? __ldcg(A.cmp + (size_t)s0_ * half + i) : 0ll;
                v[1][k] = (i < n && s0_ + 1 < A.nsp) ? __ldcg(A.cmp + (size_t)(s0_ + 1) * half + i) : 0ll;
            }
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                const int i = base + k * FT;
                if (i >= n) continue;
                const int sl = i / per;
                const double x0 = (double)(v[0][k] + offs[s0_][sl]) * FX_INV;
                if (s0_ == 0) dst[i] = x0 * A.qw[0];
                else dst[i] = dst[i] + x0 * A.qw[s0_];
                if (s0_ + 1 < A.nsp) {
                    const double x1 = (double)(v[1][k] + offs[s0_ + 1][sl]) * FX_INV;
                    dst[i] = dst[i] + x1 * A.qw[s0_ + 1];
                }
            }
        }
    }
    __syncthreads();
}

__device__ void dev_collect(const FieldArgs &A, double *dst, int n, bool with_runs, long long *scratch) {
    for (int s = 0; s < A.nsp; ++s) {
        unsigned long long *g = A.acc + (size_t)s * A.acc_stride;
        if (with_runs) {
            unsigned long long *d = g + A.acc_stride / 2;
            for (int base = threadIdx.x; base < n; base += CB * FT) {
                unsigned long long q[CB];
                collect_batch(d, A.acc_reps, A.acc_rep_stride, base, n, q);
#pragma unroll
                for (int k = 0; k < CB; ++k) { const int i = base + k * FT; if (i < n) scratch[i] = (long long)q[k]; }
            }
            if (threadIdx.x < A.acc_reps) d[(size_t)threadIdx.x * A.acc_rep_stride + n + 1] = 0ull;   // closing entry of a run that ends on the last cell
            __syncthreads();
            block_scan_i64(scratch, n);
        }
        for (int base = threadIdx.x; base < n; base += CB * FT) {
            unsigned long long qq[CB];
            collect_batch(g, A.acc_reps, A.acc_rep_stride, base, n, qq);
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                const int i = base + k * FT;
                if (i < n) {
                    long long q = (long long)qq[k];
                    if (with_runs) q += scratch[i];
                    const double v = (double)q * FX_INV;
                    if (s == 0) dst[i] = v * A.qw[0];
                    else dst[i] = dst[i] + v * A.qw[s];
                }
            }
        }
        __syncthreads();
    }
}

// tridiagonal solve of (a,b,c) x = d, d in shared memory (1-based via (i)-1), result left in d.
__device__ void dev_tridiag(const FieldArgs &A, double *d, double *tmp) {
    const int n = A.Nx;
    if (A.tridiag_mode == IPIC_TRIDIAG_THOMAS) {
        // Thomas sweep in the reference's operation order (no FMA contraction): NGP :308-328 / CIC :282-299
        if (threadIdx.x == 0) {
            d[0] = __ddiv_rn(d[0], A.b[0]);
            for (int i = 2; i <= n; ++i)
                d[i - 1] = __ddiv_rn(__dsub_rn(d[i - 1], __dmul_rn(d[i - 2], A.a[i - 2])), A.m[i - 1]);
            for (int i = n - 1; i >= 1; --i) d[i - 1] = __dsub_rn(d[i - 1], __dmul_rn(A.cp[i - 1], d[i]));
        }
        __syncthreads();
        return;
    }
    if (A.tridiag_mode == IPIC_TRIDIAG_SCAN) {
        // x = T^-1 d with (S^-1)_ij = u_i v_j / W (i <= j) for the row-scaled symmetric matrix S on the unknowns [iL, iR] (Dirichlet rows are
        // x = d and go to their neighbours' right-hand sides): x_i = (v_i sum_{j<=i} u_j d_j + u_i sum_{j>i} v_j d_j) / W -- two scans.
        // tmp needs 2 n doubles (field.cu reserves them behind the two work vectors).
        const int iL = A.giL, iR = A.giR;                 // 1-based
        double *P = tmp, *Q = tmp + n;
        const double dL = d[0], dR = d[n - 1];
        for (int i = threadIdx.x; i < n; i += FT) {
            const int r = i + 1;
            double dd = 0.0;
            if (r >= iL && r <= iR) {
                dd = d[i];
                if (iL == 2 && r == 2) dd -= A.a[0] * dL;
                if (iR == n - 1 && r == n - 1) dd -= A.c[n - 2] * dR;
                dd *= A.gs[i];
            }
            P[i] = A.gu[i] * dd; Q[i] = A.gv[i] * dd;
        }
        __syncthreads();
        block_scan_f64(P, n, false);
        block_scan_f64(Q, n, true);
        for (int i = threadIdx.x; i < n; i += FT) {
            const int r = i + 1;
            if (r >= iL && r <= iR) d[i] = (A.gv[i] * P[i] + A.gu[i] * ((i + 1 < n) ? Q[i + 1] : 0.0)) * A.gWinv;
        }
        __syncthreads();
        return;
    }
    // parallel cyclic reduction; the matrix is constant, so alpha/gamma of every level are precomputed
    double *src = d, *dst = tmp;
    int stride = 1;
    for (int lev = 0; lev < A.pcr_levels; ++lev) {
        const double *al = A.pcr + (size_t)lev * 2 * n, *ga = al + n;
        for (int base = threadIdx.x; base < n; base += CB * FT) {
            double al_[CB], ga_[CB];
#pragma unroll
            for (int k = 0; k < CB; ++k) { const int i = base + k * FT; al_[k] = (i < n) ? __ldg(al + i) : 0.0; ga_[k] = (i < n) ? __ldg(ga + i) : 0.0; }
#pragma unroll
            for (int k = 0; k < CB; ++k) {
                const int i = base + k * FT;
                if (i < n) {
                    double v = src[i];
                    if (i - stride >= 0) v = fma(al_[k], src[i - stride], v);
                    if (i + stride < n) v = fma(ga_[k], src[i + stride], v);
                    dst[i] = v;
                }
            }
        }
        __syncthreads();
        double *t = src; src = dst; dst = t;
        stride <<= 1;
    }
    const double *inv = A.pcr + (size_t)A.pcr_levels * 2 * n;
    for (int i = threadIdx.x; i < n; i += FT) d[i] = src[i] * inv[i];
    __syncthreads();
}

// Normal equations of the Anderson step for m_k <= 10 history columns, register-blocked: MATMUL(A^T, A) and A^T r with columns
// fitMat(:, j+1) = r_index - r_col(j) have 55 + 10 distinct sums over the rows.  One warp per sum (the general path below) reads three
// history columns per sum, 6.4 MB through one SM at N_x = 4097; here a thread loads the 11 values of a row once, forms the 10 differences
// in registers and adds their products into the sums its group of 256 threads owns (sums p = 4 s + G of group G: 17 or 16 accumulators),
// so every history column crosses the SM once per group.  Fixed order: a thread's rows ascending, shuffle tree over the warp, the eight
// warps of the group ascending.  gram[p]: p = k (k + 1) / 2 + j for j <= k < 10, p = 55 + j for (A^T r)_j.
template <int G>
__device__ __forceinline__ void gram_group(const double *__restrict__ res_k, int NU, int r0, int r1, int index, const int *col, int m_k,
                                           double *part /* [32 warps][17] */) {
    const int t = threadIdx.x & 255, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[17];
#pragma unroll
    for (int q = 0; q < 17; ++q) acc[q] = 0.0;
    const double *Ri = res_k + (size_t)(index - 1) * NU;
    const double *Rc[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) Rc[j] = res_k + (size_t)(col[j < m_k ? j : 0] - 1) * NU;
    for (int r = r0 + t; r < r1; r += 256) {
        const double ri = __ldcg(Ri + r);
        double d[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) d[j] = __ldcg(Rc[j] + r);
#pragma unroll
        for (int j = 0; j < 10; ++j) d[j] = (j < m_k) ? ri - d[j] : 0.0;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
#pragma unroll
            for (int j = 0; j <= k; ++j) {
                const int p = k * (k + 1) / 2 + j;
                if ((p & 3) == G) acc[p >> 2] += d[j] * d[k];
            }
        }
#pragma unroll
        for (int j = 0; j < 10; ++j) {
            const int p = 55 + j;
            if ((p & 3) == G) acc[p >> 2] += d[j] * ri;
        }
    }
#pragma unroll
    for (int q = 0; q < 17; ++q) {
        double x = acc[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) part[warp * 17 + q] = x;
    }
}

#ifdef IPIC_FIELD_PHASES
// profiling aid (scripts/field_phases.sh): thread 0 stamps the clock at the end of every phase; run_field prints the differences
__device__ long long g_field_phase[16];
#define PH(k) do { if (threadIdx.x == 0) g_field_phase[k] = clock64(); } while (0)
#else
#define PH(k) do { } while (0)
#endif
__global__ void __launch_bounds__(FT, 1) k_field(const FieldArgs A, const unsigned ops_in) {
    extern __shared__ __align__(16) double sm[];
    unsigned ops = ops_in;
    const unsigned crank = cluster_rank(), csize = cluster_size();     // 0 of 1 unless run_field launched a cluster
    if (ops & OP_AND_DECIDE) {
        // an iteration enqueued ahead of the decision that ended the solve: report and leave (the sweeps in front of it did the same)
        if (A.and_state[0] != 0) {
            if (threadIdx.x == 0 && crank == 0) { __threadfence_system(); A.h_and[0] = (long long)A.seq; }
            return;
        }
    }
    const bool clustered = csize > 1;                  // run_field launches clusters for the fused Anderson iteration only
    if (crank != 0) ops &= (OP_AND_DECIDE | OP_AND_MIX | OP_HALF_E);     // helper CTAs take part in the split phases only
#ifdef IPIC_FIELD_PHASES
    if (threadIdx.x < 16) g_field_phase[threadIdx.x] = 0;
    __syncthreads();
#endif
    PH(0);
    const int Nx = A.Nx, Nh = A.Nh;
    const int NM = max(Nx, Nh) + 2;
    double *s0 = sm, *s1 = sm + NM;   // two shared work vectors

    // ---------------- depositRho epilogue: NGP :150-187 ; CIC :146-172 ----------------
    if (ops & OP_FINALIZE_RHO) {
        dev_collect(A, s0, Nx, false, reinterpret_cast<long long *>(s1));
        __syncthreads();
#define D(i) s0[(i)-1]
        if (A.scheme == 0) {
            if (threadIdx.x == 0) {
                switch (BC(1)) {
                case 1: case 4: D(1) = 0.0; break;
                case 2: D(1) = 2.0 * D(1); break;
                case 3: D(1) = D(1) + D(Nx); break;
                }
                switch (BC(Nx)) {
                case 1: case 4: D(Nx) = 0.0; break;
                case 2: D(Nx) = 2.0 * D(Nx); break;
                case 3: D(Nx) = D(1); break;
                }
            }
            __syncthreads();
            for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
                double r = D(i);
                if (A.smooth) {
                    switch (BC(i)) {
                    case 0: r = 0.25 * (D(i - 1) + 2.0 * D(i) + D(i + 1)); break;
                    case 1: case 4: r = 0.0; break;
                    case 2: r = (i == 1) ? 0.25 * (2.0 * D(1) + 2.0 * D(2)) : 0.25 * (2.0 * D(Nx) + 2.0 * D(Nh)); break;
                    case 3: r = 0.25 * (D(Nh) + 2.0 * D(1) + D(2)); break;
                    }
                }
                RHO(i) = r;
            }
        } else {
            for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
                double r = D(i);
                if (A.smooth) {
                    switch (BC(i + 1) - BC(i)) {
                    case 0: r = 0.25 * (D(i - 1) + 2.0 * D(i) + D(i + 1)); break;
                    case -1: case -4: r = 0.25 * (D(1) + D(2)); break;
                    case 1: case 4: r = 0.25 * (D(Nx) + D(Nx - 1)); break;
                    case -2: r = 0.25 * (3.0 * D(1) + D(2)); break;
                    case 2: r = 0.25 * (3.0 * D(Nx) + D(Nx - 1)); break;
                    case -3: r = 0.25 * (D(Nx) + 2.0 * D(1) + D(2)); break;
                    case 3: r = 0.25 * (D(1) + 2.0 * D(Nx) + D(Nx - 1)); break;
                    }
                }
                RHO(i) = r;
            }
        }
#undef D
        __syncthreads();
    }

    // ---------------- solve_tridiag_Poisson: NGP :190-226 ; CIC :175-218 ----------------
    if (ops & OP_POISSON) {
        for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
            double d = 0.0;
            if (A.scheme == 0) {
                switch (BC(i)) {
                case 0: case 2: d = (-RHO(i) - A.rho_const) / eps_0; break;
                case 1: case 3: d = PHI(i); break;
                case 4: d = A.poisson_rf_value; break;
                }
            } else {
                d = (-RHO(i) - A.rho_const) / eps_0;
                if (i == 1) d = d + A.src1;
                if (i == Nx) d = d + A.src2;
            }
            s0[i - 1] = d;
        }
        __syncthreads();
        dev_tridiag(A, s0, s1);
        for (int i = threadIdx.x; i < Nx; i += FT) { A.phi[i] = s0[i]; A.phi_f[i] = s0[i]; }
        __syncthreads();
    }

    // ---------------- makeConstSourceTerm: NGP :267-286 ; CIC :246-270 ----------------
    if (ops & OP_CONST_SOURCE) {
        for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
            double r = RHO(i);
            if (A.scheme == 0) {
                switch (BC(i)) {
                case 0: r = -(PHI(i) - PHI(i - 1)) / DX(i - 1) + (PHI(i + 1) - PHI(i)) / DX(i); break;
                case 1: case 3: case 4: r = PHIF(i); break;
                case 2:
                    if (i == 1) r = -2.0 * (PHI(i) - PHI(i + 1)) / DX(i);
                    else if (i == Nx) r = -2.0 * (PHI(i) - PHI(i - 1)) / DX(i - 1);
                    break;
                }
            } else {
                switch (BC(i + 1) - BC(i)) {
                case 0: r = -(PHI(i) - PHI(i + 1)) / (GRID(i + 1) - GRID(i)) + (PHI(i - 1) - PHI(i)) / (GRID(i) - GRID(i - 1)); break;
                case -1: case -4: r = -(PHI(i) - PHI(i + 1)) / (GRID(i + 1) - GRID(i)) + 2.0 * (A.bound1 - PHI(1)) / DX(1) + A.src1; break;
                case 1: case 4: r = -2.0 * (PHI(Nx) - A.bound2) / DX(Nx) + (PHI(i - 1) - PHI(i)) / (GRID(i) - GRID(i - 1)) + A.src2; break;
                case -2: r = -(PHI(i) - PHI(i + 1)) / (GRID(i + 1) - GRID(i)); break;
                case 2: r = (PHI(i - 1) - PHI(i)) / (GRID(i) - GRID(i - 1)); break;
                case -3: r = -(PHI(i) - PHI(i + 1)) / (GRID(i + 1) - GRID(i)) + 2.0 * (PHI(Nx) - PHI(1)) / (DX(1) + DX(Nx)); break;
                case 3: r = -2.0 * (PHI(Nx) - PHI(1)) / (DX(1) + DX(Nx)) + (PHI(i - 1) - PHI(i)) / (GRID(i) - GRID(i - 1)); break;
                }
            }
            RHO(i) = r;
        }
        __syncthreads();
    }

    // ---------------- depositJ epilogue: NGP mod_particleMover.f90:145-156 ; CIC mod_particleMoverPicard.f90:165-194 ----------------
    if (clustered) {
        // every CTA adds up the copies of the deposition grids for its slice of cells; CTA 0 goes on from the compact arrays
        dev_collect_slice(A, Nh, A.scheme == 0, crank, csize, reinterpret_cast<long long *>(s1));
        cluster_barrier();
    }
    if (ops & OP_FINALIZE_J) {
        if (clustered) dev_collect_compact(A, s0, Nh, A.scheme == 0, csize);
        else dev_collect(A, s0, Nh, A.scheme == 0, reinterpret_cast<long long *>(s1));
        __syncthreads();
        PH(1);
        if (A.scheme == 1 && threadIdx.x == 0) {
#define W(i) s0[(i)-1]
            switch (BC(1)) {
            case 1: case 4: W(1) = 2.0 * W(1); break;
            case 2: W(1) = 0.0; break;
            case 3: W(1) = W(1) + W(Nh); break;
            }
            switch (BC(Nh)) {
            case 1: case 4: W(Nh) = 2.0 * W(Nh); break;
            case 2: W(Nh) = 0.0; break;
            case 3: W(Nh) = W(1); break;
            }
#undef W
        }
        __syncthreads();
        if (A.smooth) {
            dev_smooth(A, s0, s1);
            __syncthreads();
            for (int i = threadIdx.x; i < Nh; i += FT) A.J[i] = s1[i] / A.del_t;
        } else {
            for (int i = threadIdx.x; i < Nh; i += FT) A.J[i] = s0[i] / A.del_t;
        }
        __syncthreads();
        PH(2);
    }

    // ---------------- solve_tridiag_Ampere: NGP :288-330 ; CIC :272-301 ----------------
    if (ops & OP_AMPERE) {
        for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
            double d = 0.0;
            if (A.scheme == 0) {
                switch (BC(i)) {
                case 0: d = (JJ(i) - JJ(i - 1)) * A.del_t / eps_0; break;
                case 1: case 3: case 4: d = 0.0; break;
                case 2:
                    if (i == 1) d = 2.0 * (A.del_t * JJ(i) / eps_0);
                    else if (i == Nx) d = 2.0 * (-A.del_t * JJ(i - 1) / eps_0);
                    break;
                }
                d = d + RHO(i);
            } else {
                d = (JJ(i + 1) - JJ(i)) * A.del_t / eps_0 + RHO(i);
            }
            s0[i - 1] = d;
        }
        __syncthreads();
        PH(3);
        dev_tridiag(A, s0, s1);
        for (int i = threadIdx.x; i < Nx; i += FT) A.phi_f[i] = s0[i];
        __syncthreads();
        PH(4);
    }


    // ================= NGP_Periodic (E-field form) =================
    // depositRho: NGP_Periodic/src/mod_Scheme.f90:57-70
    if (ops & OP_P_RHO) {
        dev_collect(A, s0, Nx, false, reinterpret_cast<long long *>(s1));
        for (int i = threadIdx.x; i < Nx; i += FT) A.rho[i] = s0[i];
        __syncthreads();
    }
    // solve_tridiag_Poisson: NGP_Periodic/src/mod_potentialSolver.f90:127-158
    if (ops & OP_P_POISSON) {
        for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
            double d = 0.0;
            switch (BC(i)) {
            case 0: case 2: d = (-RHO(i) - A.rho_const) / eps_0; break;
            case 1: case 3: d = PHI(i); break;
            case 4: d = A.poisson_rf_value; break;
            }
            s0[i - 1] = d;
        }
        __syncthreads();
        dev_tridiag(A, s0, s1);
        for (int i = threadIdx.x; i < Nx; i += FT) { A.phi[i] = s0[i]; A.phi_f[i] = s0[i]; }
        for (int i = 1 + threadIdx.x; i <= Nh; i += FT) {
            const double e = (s0[i - 1] - s0[i]) / DX(i);
            EF(i) = e; A.Ef[i - 1] = e;
        }
        __syncthreads();
    }
    // depositJ epilogue: NGP_Periodic/src/mod_particleMover.f90:131-134 (no smoothing)
    if (ops & OP_P_FINALIZE_J) {
        dev_collect(A, s0, Nh, true, reinterpret_cast<long long *>(s1));
        for (int i = threadIdx.x; i < Nh; i += FT) A.J[i] = s0[i] / A.del_t;
        __syncthreads();
    }
    // solve_tridiag_Ampere: :170-176
    if (ops & OP_P_AMPERE) {
        double aveJ = A.scal[2];
        if (!A.globJ) {
            double part = 0.0;
            for (int i = 1 + threadIdx.x; i <= Nh; i += FT) part += JJ(i) * DX(i);
            aveJ = block_sum(part) / A.L_domain;
            if (threadIdx.x == 0) A.scal[2] = aveJ;
        }
        for (int i = 1 + threadIdx.x; i <= Nh; i += FT) A.Ef[i - 1] = (aveJ - JJ(i)) * A.del_t / eps_0 + EF(i);
        __syncthreads();
    }

    // ---------------- Anderson bookkeeping: mod_nonLinSolvers.f90:98-157 ----------------
    const int M1 = A.and_m + 1;
    const int NU = A.nunk;                 // N_x (phi_f) or N_h (EField_f)
#define PK(i, c) A.phi_k[(size_t)((c)-1) * NU + (i)-1]
#define RK(i, c) A.res_k[(size_t)((c)-1) * NU + (i)-1]
#define UNK(i) A.unk[(i)-1]
    if (ops & OP_AND_STORE0) {
        for (int i = 1 + threadIdx.x; i <= NU; i += FT) PK(i, 1) = UNK(i);
        __syncthreads();
    }
    if (ops & (OP_AND_FIRST | OP_AND_RESID)) {
        const int index = (ops & OP_AND_FIRST) ? 1 : fmodulo(A.and_i, M1) + 1;
        double part = 0.0;
        for (int i = 1 + threadIdx.x; i <= NU; i += FT) {
            double r;
            if (ops & OP_AND_FIRST) { PK(i, 2) = UNK(i); r = UNK(i) - PK(i, 1); }
            else r = UNK(i) - PK(i, index);
            RK(i, index) = r;
            part += r * r;
        }
        const double nrm = sqrt(block_sum(part));
        __shared__ int s_dec;
        if (threadIdx.x == 0) {
            A.norms[index - 1] = nrm; A.scal[0] = nrm;
            if (A.h_seq) {
                *reinterpret_cast<volatile double *>(A.h_norm) = nrm;
                __threadfence_system();
                *reinterpret_cast<volatile unsigned long long *>(A.h_seq) = A.seq;
            }
            if (ops & OP_AND_DECIDE) {
                // solveDivAmpereAnderson :111-134 (NGP_Periodic :102-123): converged below eps_tol; after m iterations, stagnated when the
                // norm exceeds the mean of the last three (periodic tree: two) -- the same operations in the same order as the host loop
                const int i = A.and_i, m = A.and_m;
                int dec = 0;
                if (nrm < A.eps_tol) dec = 1;
                else if (i > m) {
                    double sumPast;
                    if (A.eform) {
                        if (m > 1) sumPast = __ddiv_rn(__dadd_rn(A.norms[fmodulo(i - 1, M1)], A.norms[fmodulo(i - 2, M1)]), 2.0);
                        else sumPast = A.norms[fmodulo(i - 1, M1)];
                    } else if (m > 3) {
                        sumPast = __dadd_rn(__dadd_rn(A.norms[fmodulo(i - 1, M1)], A.norms[fmodulo(i - 2, M1)]), A.norms[fmodulo(i - 3, M1)]);
                        sumPast = __ddiv_rn(sumPast, 3.0);
                    } else {
                        sumPast = A.norms[fmodulo(i - 1, M1)];
                    }
                    if (nrm > sumPast) dec = 2;
                }
                s_dec = dec;
                if (dec) { A.and_state[0] = dec; A.and_state[1] = i; }
                // outcome to the host (pinned, device-addressable): norm history, state, then the sequence number it waits on.  Published
                // here, before the mixing and the push tables, so that the host enqueues the next chunk while this launch finishes
                A.h_and_norms[i] = nrm;
                if (dec) { A.h_and[1] = dec; A.h_and[2] = i; }
                __threadfence_system();
                A.h_and[0] = (long long)A.seq;
            }
        }
        __syncthreads();
        if (ops & OP_AND_DECIDE) {
            const int dec = s_dec;
            if (dec == 1) ops &= ~OP_AND_MIX;                                     // converged: E-half + tables of the solved field for the mover
            else if (dec == 2) ops &= ~(OP_AND_MIX | OP_HALF_E | OP_P_HALF_E);    // stagnated: the caller halves the step and starts over
        }
        PH(5);
    }
    if (clustered) {
        // the helper CTAs waited here for the residual and the decision (and_state: 0 go on, 1 converged, 2 stagnated)
        cluster_barrier();
        if (crank != 0) {
            const int dec = __ldcg(A.and_state);
            if (dec == 1) ops &= ~OP_AND_MIX;
            else if (dec == 2) ops &= ~(OP_AND_MIX | OP_HALF_E | OP_P_HALF_E);
        }
    }
    if (ops & OP_AND_MIX) {
        const int i = A.and_i, m = A.and_m;
        const int index = fmodulo(i, M1) + 1;
        const int m_k = min(i, m);
        // normal equations A^T A alpha = A^T r with columns fitMat(:,j+1) = r_index - r_{col(i-m_k+j)}
        double *An = s0;                 // m_k x m_k column-major
        double *bn = s0 + 32 * 32;       // m_k
        double *alpha = bn + 64;         // m_k + 1
        int *col = reinterpret_cast<int *>(alpha + 64);   // history column of fitMat column j (the MODULO of :139 evaluated once)
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (threadIdx.x <= m_k) col[threadIdx.x] = fmodulo(i - m_k + (int)threadIdx.x, M1) + 1;
        __syncthreads();
        if (m_k <= 10) {
            // register-blocked sums (gram_group): four groups of 256 threads, each owning a quarter of the 65 sums
            double *part = alpha + 128;                    // [32][17]
            double *gram = part + 32 * 17;                 // [68]
            // rows of this CTA (all of them without a cluster)
            const int rper = (NU + (int)csize - 1) / (int)csize;
            const int r0 = (int)crank * rper, r1 = min(NU, r0 + rper);
            switch (threadIdx.x >> 8) {
            case 0: gram_group<0>(A.res_k, NU, r0, r1, index, col, m_k, part); break;
            case 1: gram_group<1>(A.res_k, NU, r0, r1, index, col, m_k, part); break;
            case 2: gram_group<2>(A.res_k, NU, r0, r1, index, col, m_k, part); break;
            default: gram_group<3>(A.res_k, NU, r0, r1, index, col, m_k, part); break;
            }
            __syncthreads();
            if (threadIdx.x < 65) {
                const int pp = threadIdx.x, g = pp & 3, q = pp >> 2;
                double x = 0.0;
                for (int w = 0; w < 8; ++w) x += part[(g * 8 + w) * 17 + q];
                if (clustered) A.gram_part[crank * 68 + pp] = x;
                else gram[pp] = x;
            }
            if (clustered) {
                cluster_barrier();
                if (crank == 0 && threadIdx.x < 65) {      // the CTAs' partial sums in rank order
                    double x = 0.0;
                    for (unsigned k = 0; k < csize; ++k) x += __ldcg(A.gram_part + k * 68 + threadIdx.x);
                    gram[threadIdx.x] = x;
                }
            }
            __syncthreads();
            if (threadIdx.x < m_k * m_k) {
                const int j = threadIdx.x % m_k, k = threadIdx.x / m_k;
                const int lo = min(j, k), hi = max(j, k);
                An[j + m_k * k] = gram[hi * (hi + 1) / 2 + lo];
            }
            if (threadIdx.x < m_k) bn[threadIdx.x] = gram[55 + threadIdx.x];
        } else {
        // MATMUL(A^T, A) is symmetric and each entry is the same sum of the same products either way round, so only j <= k is
        // computed (one warp per entry, lanes stride the rows, shuffle tree) and mirrored: bit-identical to computing all m_k^2
        const int ntri = m_k * (m_k + 1) / 2, npairs = ntri + m_k;
        for (int p = warp; p < npairs; p += FT / 32) {
            int j, k;
            if (p < ntri) { k = 0; int q = p; while (q > k) { q -= k + 1; ++k; } j = q; }     // p = k (k + 1) / 2 + j, j <= k
            else { j = p - ntri; k = -1; }
            const double *Ri = A.res_k + (size_t)(index - 1) * NU, *Rj = A.res_k + (size_t)(col[j] - 1) * NU;
            const double *Rk = (k >= 0) ? A.res_k + (size_t)(col[k] - 1) * NU : Ri;
            double acc = 0.0;
            if (k >= 0) {
#pragma unroll 8
                for (int r = lane; r < NU; r += 32) { const double ri = Ri[r]; acc += (ri - Rj[r]) * (ri - Rk[r]); }
            } else {
#pragma unroll 8
                for (int r = lane; r < NU; r += 32) { const double ri = Ri[r]; acc += (ri - Rj[r]) * ri; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
            if (lane == 0) { if (k >= 0) { An[j + m_k * k] = acc; An[k + m_k * j] = acc; } else bn[j] = acc; }
        }
        }
        __syncthreads();
        PH(6);
        if (crank == 0) {
        if (warp == 0) {
            // solveGaussElimination, no pivoting (mod_BasicFunctions.f90:218-254).  The elimination runs on one warp, lane = row: every
            // entry sees the reference's operations in the reference's order (its updates are independent across rows and columns
            // within a pivot step); the two substitutions stay sequential.
            const int n = m_k;
#define AA(r, c) An[((r)-1) + n * ((c)-1)]
            const int r = lane + 1;
            for (int k = 1; k <= n - 1; ++k) {
                const double k1 = AA(k, k);
                if (r > k && r <= n) AA(r, k) = AA(r, k) / k1;
                __syncwarp();
                if (r > k && r <= n) {
                    const double ark = AA(r, k);
                    for (int j = k + 1; j <= n; ++j) AA(r, j) = AA(r, j) - ark * AA(k, j);
                }
                __syncwarp();
            }
            if (lane == 0) {
                for (int rr = 1; rr <= n; ++rr) for (int j = 1; j <= rr - 1; ++j) bn[rr - 1] = bn[rr - 1] - AA(rr, j) * bn[j - 1];
                for (int rr = n; rr >= 1; --rr) {
                    for (int j = rr + 1; j <= n; ++j) bn[rr - 1] = bn[rr - 1] - AA(rr, j) * bn[j - 1];
                    bn[rr - 1] = bn[rr - 1] / AA(rr, rr);
                }
                double sa = 0.0;
                for (int j = 0; j < n; ++j) { alpha[j] = bn[j]; sa += bn[j]; }
                alpha[n] = 1.0 - sa;
            }
#undef AA
        }
        __syncthreads();
        PH(7);
        const int cn = fmodulo(i + 1, M1) + 1;
        for (int r = 1 + threadIdx.x; r <= NU; r += FT) {
            double x = alpha[0] * (A.beta * RK(r, col[0]) + PK(r, col[0]));
            for (int j = 1; j <= m_k; ++j) x = x + alpha[j] * (A.beta * RK(r, col[j]) + PK(r, col[j]));
            PK(r, cn) = x;
            UNK(r) = x;
        }
        __syncthreads();
        }
        PH(8);
    }
#undef PK
#undef RK
#undef UNK

    if (ops & OP_RESID_F) {   // funcNitsol: fcur = xcur - unknown   (mod_nonLinSolvers.f90:236 ; NGP_Periodic :219)
        for (int i = threadIdx.x; i < NU; i += FT) A.fout[i] = A.xin[i] - A.unk[i];
        __syncthreads();
    }

    // ---------------- makeHalfTimeEField (+ push tables): NGP :392-406 ; CIC :333-366, Picard :92-94 ----------------
    if (ops & OP_HALF_E) {
        if (crank == 0 && A.counters && threadIdx.x < A.nsp) A.counters[(size_t)threadIdx.x * 8 + 6] = 0ull;      // tile schedulers of the sweeps that follow
        // cells of this CTA for the push tables (all of them without a cluster)
        const int tcells = (A.scheme == 1) ? Nx : Nh;
        const int tper = (tcells + (int)csize - 1) / (int)csize;
        const int c_lo = 1 + (int)crank * tper, c_hi = min(tcells, (int)(crank + 1) * tper);
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        // cluster: the unknown (mixed by CTA 0, or solved when the iteration converged) is complete behind this barrier; every CTA forms the
        // half-time differences of the whole grid in its own shared memory (a 3-point smoothing stencil needs its neighbours) and then
        // EField and the tables of its own cells
        if (clustered) cluster_barrier();
        if (A.scheme == 0) {
            for (int i = 1 + threadIdx.x; i <= Nh; i += FT)
                s0[i - 1] = 0.5 * (__ldcg(A.phi_f + i - 1) + __ldcg(A.phi + i - 1) - __ldcg(A.phi_f + i) - __ldcg(A.phi + i));
            __syncthreads();
            const double *Esm = s0;
            if (A.smooth) {
                dev_smooth(A, s0, s1);
                __syncthreads();
                Esm = s1;
            }
            // EField and the push tables of every species in one pass over the cells (dx_dl and 0.5 / dx_dl of a batch loaded together)
            for (int base = c_lo + threadIdx.x; base <= c_hi; base += CB * FT) {
                double dx_[CB], hv_[CB];
#pragma unroll
                for (int k = 0; k < CB; ++k) {
                    const int cI = base + k * FT;
                    dx_[k] = (cI <= c_hi) ? __ldg(A.dx + cI - 1) : 1.0;
                    hv_[k] = (cI <= c_hi) ? __ldg(A.hinv + cI - 1) : 0.0;
                }
#pragma unroll
                for (int k = 0; k < CB; ++k) {
                    const int cI = base + k * FT;
                    if (cI > c_hi) continue;
                    const double dx = dx_[k];
                    const double E = Esm[cI - 1] / dx;
                    EF(cI) = E;
                    // wall cells are not transited by the lean loop: NaN makes a lane that reaches one fail its next time test.
                    // Right-movers read entry cI, left-movers the mirrored, negated copy (every lane walks upwards)
                    const bool wall = (cI == 1 || cI == Nh);
                    for (int s = 0; s < A.nsp; ++s) {
                        CellTabNGP *T = reinterpret_cast<CellTabNGP *>(A.tab) + (size_t)s * A.tab_cells;
                        double2 *TT = A.ttab + (size_t)s * 2 * A.tab_cells;
                        const double a = A.qm[s] * E;
                        CellTabNGP t;
                        t.a = a; t.hinv = hv_[k]; t.two_a_dx = 2.0 * a * dx; t.inv_a = 1.0 / a;
                        T[cI] = t;
                        TT[cI] = wall ? make_double2(qnan, 0.0) : make_double2(t.two_a_dx, t.inv_a);
                        TT[2 * A.tab_cells - 1 - cI] = wall ? make_double2(qnan, 0.0) : make_double2(-t.two_a_dx, -t.inv_a);
                    }
                }
            }
        } else {
            for (int i = 1 + threadIdx.x; i <= Nh; i += FT) {
                double w;
                if (i == 1) {
                    switch (BC(1)) {
                    case 1: w = (2.0 * A.bound1 - PHI(1) - PHIF(1)); break;
                    case 2: w = 0.0; break;
                    case 3: w = 0.5 * (PHI(Nx) + PHIF(Nx) - PHI(1) - PHIF(1)); break;
                    default: w = (A.bound1 + A.RF_voltage - PHI(1) - PHIF(1)); break;   // 4
                    }
                } else if (i == Nh) {
                    switch (BC(Nh)) {
                    case 1: w = PHI(Nx) + PHIF(Nx) - 2.0 * A.bound2; break;
                    case 2: w = 0.0; break;
                    case 3: w = 0.5 * (PHI(Nx) + PHIF(Nx) - PHI(1) - PHIF(1)); break;   // workArray(N_h) = workArray(1)
                    default: w = PHI(Nx) + PHIF(Nx) - A.bound2 - A.RF_voltage; break;   // 4
                    }
                } else {
                    w = 0.5 * (PHIF(i - 1) + PHI(i - 1) - PHIF(i) - PHI(i));
                }
                s0[i - 1] = w;
            }
            __syncthreads();
            const double *Ec = s0;
            if (A.smooth) {
                dev_smooth(A, s0, s1);
                __syncthreads();
                Ec = s1;
            }
            // EField (N_h faces): this CTA's cells' faces, the last CTA also the closing face
            for (int i = c_lo + threadIdx.x; i <= ((crank == csize - 1) ? Nh : c_hi); i += FT) A.E[i - 1] = Ec[i - 1];
            for (int s = 0; s < A.nsp; ++s) {
                CellTabCIC *T = reinterpret_cast<CellTabCIC *>(A.tab) + (size_t)s * A.tab_cells;
                for (int cI = c_lo + threadIdx.x; cI <= c_hi; cI += FT) {
                    const double dx = DX(cI), el = Ec[cI - 1], er = Ec[cI];
                    CellTabCIC t;
                    t.A0 = A.qm[s] * el; t.A1 = A.qm[s] * (er - el); t.inv_dx = 1.0 / dx; t.dx = dx;
                    t.cap = 0.1 * dx / sqrt(fabs(A.qm[s] * (el - er)));
                    t.pad[0] = t.pad[1] = t.pad[2] = 0.0;
                    T[cI] = t;
                }
            }
        }
        __syncthreads();
    }


    PH(9);
    // EField_half = 0.5 (EField + EField_f) (NGP_Periodic/src/mod_particleMover.f90:72) + push tables
    if (ops & OP_P_HALF_E) {
        if (A.counters && threadIdx.x < A.nsp) A.counters[(size_t)threadIdx.x * 8 + 6] = 0ull;      // tile schedulers of the sweeps that follow
        for (int i = threadIdx.x; i < Nh; i += FT) A.Eh[i] = 0.5 * (A.E[i] + A.Ef[i]);
        __syncthreads();
        for (int s = 0; s < A.nsp; ++s) {
            CellTabNGP *T = reinterpret_cast<CellTabNGP *>(A.tab) + (size_t)s * A.tab_cells;
            double2 *TT = A.ttab + (size_t)s * 2 * A.tab_cells;
            for (int cI = 1 + threadIdx.x; cI <= Nh; cI += FT) {
                const double a = A.qm[s] * A.Eh[cI - 1], dx = DX(cI);
                CellTabNGP t;
                t.a = a; t.hinv = A.hinv[cI - 1]; t.two_a_dx = 2.0 * a * dx; t.inv_a = 1.0 / a;
                T[cI] = t;
                const bool wall = (cI == 1 || cI == Nh);
                const double qnan = __longlong_as_double(0x7ff8000000000000ll);
                TT[cI] = wall ? make_double2(qnan, 0.0) : make_double2(t.two_a_dx, t.inv_a);
                TT[2 * A.tab_cells - 1 - cI] = wall ? make_double2(qnan, 0.0) : make_double2(-t.two_a_dx, -t.inv_a);
            }
        }
        __syncthreads();
    }

    // ---------------- conservation diagnostics ----------------
    if (ops & OP_DIAG) {
        const int wch = A.diag_which;
        double part = 0.0, res = 0.0;
        if ((wch == 0 || wch == 1) && A.scheme == 2) {   // getTotalPE: NGP_Periodic :226-236
            const double *Ev = wch ? A.Ef : A.E;
            for (int i = 1 + threadIdx.x; i <= Nh; i += FT) part += Ev[i - 1] * Ev[i - 1] * DX(i);
            res = 0.5 * eps_0 * block_sum(part);
        } else if (wch == 4 && A.scheme == 2) {          // getError_tridiag_Poisson: NGP_Periodic :159-169
            for (int i = 1 + threadIdx.x; i <= Nh; i += FT) {
                double r;
                if (i == 1) r = eps_0 * (A.Ef[0] - A.Ef[Nh - 1]) / (RHO(1) + RHO(Nx)) - 1.0;
                else r = eps_0 * (A.Ef[i - 1] - A.Ef[i - 2]) / RHO(i) - 1.0;
                part += r * r;
            }
            res = sqrt(block_sum(part) / Nh);
        } else if (wch == 5 && A.scheme == 2) {          // getChargeContinuityError: NGP_Periodic :187-225
            double cnt = 0.0;
            for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
                double dR = RHO(i) - A.rho_i[i - 1];
                if (dR != 0) {
                    switch (BC(i)) {
                    case 0: { const double v = 1.0 + A.del_t * (JJ(i) - JJ(i - 1)) / dR; part += v * v; cnt += 1.0; break; }
                    case 2: {
                        const double v = (i == 1) ? 1.0 + A.del_t * JJ(1) / dR : 1.0 - A.del_t * JJ(Nh) / dR;
                        part += v * v; cnt += 1.0; break; }
                    case 3:
                        if (i == 1) { dR = dR + RHO(Nx) - A.rho_i[Nx - 1]; const double v = 1.0 + A.del_t * (JJ(1) - JJ(Nh)) / dR; part += v * v; }
                        break;
                    }
                }
            }
            const double tot = block_sum(part);
            const double k = block_sum(cnt);
            res = sqrt(tot / k);
        } else if (wch == 3 && A.scheme == 2) {          // getEnergyFromBoundary: NGP_Periodic :237-243
            res = 0.0;
        } else if (wch == 0 || wch == 1) {   // getTotalPE(future = wch): NGP :408-420 ; CIC :369-413
            const double *ph = wch ? A.phi_f : A.phi;
            if (A.scheme == 0) {
                for (int i = 1 + threadIdx.x; i <= Nh; i += FT) { const double dphi = ph[i] - ph[i - 1]; part += dphi * dphi / DX(i); }
                res = 0.5 * eps_0 * block_sum(part);
            } else {
                for (int i = 1 + threadIdx.x; i <= Nx - 1; i += FT) { const double dphi = ph[i] - ph[i - 1]; part += dphi * dphi / (GRID(i + 1) - GRID(i)); }
                res = 0.5 * eps_0 * block_sum(part);
                const double p1 = ph[0], pN = ph[Nx - 1];
                if (wch) {
                    switch (BC(1)) {
                    case 1: res += eps_0 * (A.bound1 - p1) * (A.bound1 - p1) / DX(1); break;
                    case 3: res += eps_0 * (pN - p1) * (pN - p1) / (DX(1) + DX(Nx)); break;
                    case 4: res += eps_0 * (A.RF_voltage - p1) * (A.RF_voltage - p1) / DX(1); break;
                    }
                    switch (BC(Nh)) {
                    case 1: res += eps_0 * (A.bound2 - pN) * (A.bound2 - pN) / DX(Nx); break;
                    case 4: res += eps_0 * (A.RF_voltage - pN) * (A.RF_voltage - pN) / DX(Nx); break;
                    }
                } else {
                    switch (BC(1)) {
                    case 1: case 4: res += eps_0 * (A.bound1 - p1) * (A.bound1 - p1) / DX(1); break;
                    case 3: res += eps_0 * (pN - p1) * (pN - p1) / (DX(1) + DX(Nx)); break;
                    }
                    switch (BC(Nh)) {
                    case 1: case 4: res += eps_0 * (A.bound2 - pN) * (A.bound2 - pN) / DX(Nx); break;
                    }
                }
            }
        } else if (wch == 2) {   // getJdotE: NGP :422-430 ; CIC :415-426
            if (A.scheme == 0) {
                for (int i = 1 + threadIdx.x; i <= Nh; i += FT) part += EF(i) * DX(i) * JJ(i);
                res = block_sum(part) * A.del_t;
            } else {
                for (int i = 2 + threadIdx.x; i <= Nx; i += FT) part += JJ(i) * EF(i);
                res = block_sum(part);
                res = res + 0.5 * JJ(1) * EF(1) + JJ(Nh) * EF(Nh) * 0.5;
                res = res * A.del_t;
            }
        } else if (wch == 3) {   // getEnergyFromBoundary: NGP :432-442 ; CIC :428-456
            if (A.scheme == 0) {
                res = A.del_t * JJ(1);
                res = res + eps_0 * ((PHIF(1) - PHIF(2)) - (PHI(1) - PHI(2))) / DX(1);
                res = res * (PHIF(1) + PHI(1) - PHIF(Nx) - PHI(Nx)) * 0.5;
            } else {
                double lA = 0.0, rA = 0.0;
                switch (BC(1)) { case 1: lA = A.bound1; break; case 4: lA = 0.5 * (A.RF_voltage + A.bound1); break; }
                switch (BC(Nh)) { case 1: rA = A.bound2; break; case 4: rA = 0.5 * (A.RF_voltage + A.bound2); break; }
                res = A.del_t * JJ(2) + eps_0 * ((PHIF(1) - PHIF(2)) - (PHI(1) - PHI(2))) / (GRID(2) - GRID(1));
                res = res * (lA - rA);
            }
        } else if (wch == 4) {   // getError_tridiag_Poisson: NGP :240-265 ; CIC :232-244
            for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
                double Ax = PHIF(i) * A.b[i - 1];
                if (i > 1) Ax += PHIF(i - 1) * A.a[i - 2];
                if (i < Nx) Ax += PHIF(i + 1) * A.c[i - 1];
                if (A.scheme == 0) {
                    const int b = BC(i);
                    if (b == 0 || b == 2) {
                        if (RHO(i) + A.rho_const != 0) { const double v = Ax * eps_0 / (-RHO(i) - A.rho_const) - 1.0; part += v * v; }
                        else { const double v = Ax * eps_0; part += v * v; }
                    } else {
                        if (PHIF(i) != 0) { const double v = Ax / PHIF(i) - 1.0; part += v * v; }
                        else part += Ax * Ax;
                    }
                } else {
                    double d = (-RHO(i) - A.rho_const) / eps_0;
                    if (i == 1) d += A.src1;
                    if (i == Nx) d += A.src2;
                    const double v = 1.0 - Ax / d;
                    part += v * v;
                }
            }
            res = sqrt(block_sum(part) / Nx);
        } else if (wch == 5) {   // getChargeContinuityError: NGP :358-390 ; CIC :315-331
            double cnt = 0.0;
            for (int i = 1 + threadIdx.x; i <= Nx; i += FT) {
                const double dR = RHO(i) - A.rho_i[i - 1];
                if (dR != 0) {
                    if (A.scheme == 0) {
                        switch (BC(i)) {
                        case 0: { const double v = 1.0 + A.del_t * (JJ(i) - JJ(i - 1)) / dR; part += v * v; cnt += 1.0; break; }
                        case 2: {
                            const double v = (i == 1) ? 1.0 + 2.0 * A.del_t * JJ(1) / dR : 1.0 - 2.0 * A.del_t * JJ(Nh) / dR;
                            part += v * v; cnt += 1.0; break; }
                        case 3: if (i == 1) { const double v = 1.0 + A.del_t * (JJ(1) - JJ(Nh)) / dR; part += v * v; } break;
                        }
                    } else {
                        const double v = 1.0 + A.del_t * (JJ(i + 1) - JJ(i)) / dR; part += v * v; cnt += 1.0;
                    }
                }
            }
            const double tot = block_sum(part);
            const double k = block_sum(cnt);
            res = sqrt(tot / k);
        }
        if (threadIdx.x == 0) A.scal[1] = res;
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
static FieldArgs make_args(Ctx &c) {
    FieldArgs A{};
    A.scheme = c.scheme; A.Nx = c.Nx; A.Nh = c.Nh; A.smooth = c.smooth; A.nsp = (int)c.sp.size();
    A.tridiag_mode = c.tridiag_mode; A.pcr_levels = c.pcr_levels;
    A.acc_stride = 2 * (std::max(c.Nx, c.Nh) + 2);
    A.acc_reps = std::max(1, c.acc_reps); A.acc_rep_stride = c.acc_rep_stride;
    A.tab_cells = ((c.scheme == IPIC_SCHEME_CIC) ? c.Nx : c.Nh) + 2;
    A.grid = c.d_grid; A.dx = c.d_dx; A.hinv = c.d_hinv; A.bc = c.d_bc;
    A.phi = c.d_phi; A.phi_f = c.d_phi_f; A.rho = c.d_rho; A.J = c.d_J; A.E = c.d_E;
    A.a = c.d_a; A.b = c.d_b; A.c = c.d_c; A.cp = c.d_cp; A.m = c.d_m; A.pcr = c.d_pcr;
    A.counters = c.d_counters;
    A.work = c.d_work; A.scal = c.d_scal; A.rho_i = c.d_rho_i; A.acc = c.d_acc; A.tab = c.d_tab; A.ttab = c.d_ttab;
    for (int s = 0; s < A.nsp && s < NSP_MAX; ++s) { A.qw[s] = c.sp[s].q_times_wp; A.qm[s] = c.sp[s].q_over_m; }
    A.rho_const = c.rho_const; A.RF_voltage = c.RF_voltage;
    A.bound1 = c.boundPhi[0]; A.bound2 = c.boundPhi[1]; A.src1 = c.sourceTermVals[0]; A.src2 = c.sourceTermVals[1];
    A.Ef = c.d_Ef; A.Eh = c.d_Eh;
    A.unk = (c.scheme == IPIC_SCHEME_NGP_PERIODIC) ? c.d_Ef : c.d_phi_f;
    A.nunk = (c.scheme == IPIC_SCHEME_NGP_PERIODIC) ? c.Nh : c.Nx;
    A.L_domain = c.L_domain; A.globJ = 0;
    A.cmp = c.d_cmp; A.cmp_tot = c.d_cmp ? c.d_cmp + (size_t)NSP_MAX * ((size_t)std::max(c.Nx, c.Nh) + 2) : nullptr; A.gram_part = c.d_gram_part;
    A.gu = c.d_gu; A.gv = c.d_gv; A.gs = c.d_gs; A.gWinv = (c.gW != 0.0) ? 1.0 / c.gW : 0.0; A.giL = c.giL; A.giR = c.giR;
    A.phi_k = c.d_phi_k; A.res_k = c.d_res_k; A.norms = c.d_norms;
    A.and_m = c.m_Anderson; A.beta = c.Beta_k;
    return A;
}

static int run_field(Ctx &c, const FieldArgs &A, unsigned ops) {
    // two work vectors + scratch behind them (the scan solve uses the second vector and the scratch as 2 N_x doubles)
    const size_t smem = (size_t)2 * (std::max(c.Nx, c.Nh) + 2) * sizeof(double) + std::max<size_t>(4096, (size_t)std::max(c.Nx, c.Nh) + 2) * sizeof(double);
    static size_t configured_dev[64] = {};    // function attributes are per device
    size_t &configured = configured_dev[c.device & 63];
    if (smem > configured) {
        IPIC_CUDA(cudaFuncSetAttribute(k_field, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // the same shared-memory / L1 split as the push kernels it alternates with (particles.cu): a different carve-out makes every
        // push -> field -> push transition re-partition the SMs
        IPIC_CUDA(cudaFuncSetAttribute(k_field, cudaFuncAttributePreferredSharedMemoryCarveout, 52));
        configured = smem;
    }
    // The fused Anderson iteration runs as a thread-block cluster: CTA 0 is the kernel as ever, the others share the collect, the normal
    // equations and the push tables (k_field).  Everything else is one CTA.
    const int csz = ((ops & OP_AND_DECIDE) && c.scheme != IPIC_SCHEME_NGP_PERIODIC && c.m_Anderson <= 10 && c.d_cmp) ? std::max(1, std::min(8, c.field_cluster)) : 1;
    c.launches += 1;
    if (csz > 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csz); cfg.blockDim = dim3(FT); cfg.dynamicSmemBytes = smem; cfg.stream = c.stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        IPIC_CUDA(cudaLaunchKernelEx(&cfg, k_field, A, ops));
    } else {
        k_field<<<1, FT, smem, c.stream>>>(A, ops);
    }
    IPIC_CUDA(cudaGetLastError());
    if (ops & (OP_HALF_E | OP_P_HALF_E)) c.tile_zero_mask = ~0u;      // every species' tile counter is zero behind this launch
#ifdef IPIC_FIELD_PHASES
    {
        long long h[16];
        cudaStreamSynchronize(c.stream);
        cudaMemcpyFromSymbol(h, g_field_phase, sizeof h);
        static const char *nm[10] = {"start", "collect", "smooth+J", "rhs", "tridiag", "resid+norm", "normal-eq", "gauss", "mix", "halfE+tables"};
        fprintf(stderr, "k_field ops=%05x:", ops);
        long long prev = h[0];
        for (int k = 1; k < 10; ++k) if (h[k]) { fprintf(stderr, " %s %.1fus", nm[k], (double)(h[k] - prev) / 1965.0); prev = h[k]; }
        fprintf(stderr, "\n");
    }
#endif
    return IPIC_OK;
}

// copies 1.. of the deposition grids are added into copy 0 and zeroed (multi-GPU: copy 0 is what the all-reduce carries)
__global__ void k_fold_acc(unsigned long long *acc, size_t count, int reps, size_t rep_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long q = 0ull;
    for (int r = 1; r < reps; ++r) { unsigned long long *p = acc + (size_t)r * rep_stride + i; q += *p; *p = 0ull; }
    if (q) acc[i] += q;
}
int fold_acc(Ctx &c, size_t count) {
    if (c.acc_reps <= 1) return IPIC_OK;
    c.launches += 1; k_fold_acc<<<(unsigned)((count + 255) / 256), 256, 0, c.stream>>>(c.d_acc, count, c.acc_reps, c.acc_rep_stride);
    IPIC_CUDA(cudaGetLastError());
    return IPIC_OK;
}

// construct_diagMatrix (NGP :88-128 ; CIC :92-126) + constant-matrix factor tables
int field_setup(Ctx &c) {
    const int Nx = c.Nx;
    std::vector<double> a(Nx + 2, 0.0), b(Nx + 2, 0.0), cc(Nx + 2, 0.0);   // 1-based
    auto DXh = [&](int i) -> double { return (i >= 1 && i <= c.ndx) ? c.h_dx[i - 1] : 1.0; };
    auto BCh = [&](int i) -> int { return c.h_bc[i - 1]; };
    if (c.scheme == IPIC_SCHEME_NGP || c.scheme == IPIC_SCHEME_NGP_PERIODIC) {
        // NGP_Periodic/src/mod_potentialSolver.f90:88-126 differs only in the Neumann rows (factor 1 instead of 2)
        const double nf = (c.scheme == IPIC_SCHEME_NGP) ? 2.0 : 1.0;
        for (int i = 1; i <= Nx; ++i) {
            switch (BCh(i)) {
            case 0:
                if (i < Nx) cc[i] = 1.0 / DXh(i);
                if (i > 1) a[i - 1] = 1.0 / DXh(i - 1);
                b[i] = -(1.0 / DXh(i - 1) + 1.0 / DXh(i));
                break;
            case 1: case 4: b[i] = 1.0; break;
            case 2:
                if (i == 1) { cc[i] = nf / DXh(i); b[i] = -(nf / DXh(i)); }
                else if (i == Nx) { a[i - 1] = nf / DXh(i - 1); b[i] = -(nf / DXh(i - 1)); }
                else { set_error("Neumann boundary not on left or right most index!"); return IPIC_ERR_STOP; }
                break;
            case 3: b[i] = 1.0; break;
            default: break;
            }
        }
    } else {
        for (int i = 1; i <= Nx; ++i) {
            switch (BCh(i + 1) - BCh(i)) {
            case 0:
                b[i] = -2.0 * (1.0 / (DXh(i - 1) + DXh(i)) + 1.0 / (DXh(i + 1) + DXh(i)));
                a[i - 1] = 2.0 / (DXh(i - 1) + DXh(i));
                cc[i] = 2.0 / (DXh(i) + DXh(i + 1));
                break;
            case -1: case -4: case -3:
                b[i] = -2.0 * (1.0 / (DXh(i + 1) + DXh(i)) + 1.0 / DXh(i));
                cc[i] = 2.0 / (DXh(i + 1) + DXh(i));
                c.sourceTermVals[0] = -c.boundPhi[0] * 2.0 / DXh(1);
                break;
            case 1: case 4: case 3:
                b[i] = -2.0 * (1.0 / (DXh(i - 1) + DXh(i)) + 1.0 / DXh(i));
                a[i - 1] = 2.0 / (DXh(i - 1) + DXh(i));
                c.sourceTermVals[1] = -c.boundPhi[1] * 2.0 / DXh(Nx);
                break;
            case -2:
                b[i] = -2.0 / (DXh(i + 1) + DXh(i));
                cc[i] = 2.0 / (DXh(i + 1) + DXh(i));
                break;
            case 2:
                b[i] = -2.0 / (DXh(i - 1) + DXh(i));
                a[i - 1] = 2.0 / (DXh(i - 1) + DXh(i));
                break;
            }
        }
    }
    // Thomas: c'_i and pivots m_i exactly as the reference recomputes them on every call
    std::vector<double> cp(Nx + 2, 0.0), mm(Nx + 2, 0.0);
    cp[1] = cc[1] / b[1]; mm[1] = b[1];
    for (int i = 2; i <= Nx - 1; ++i) { mm[i] = b[i] - cp[i - 1] * a[i - 1]; cp[i] = cc[i] / mm[i]; }
    mm[Nx] = b[Nx] - cp[Nx - 1] * a[Nx - 1];
    // PCR level coefficients on the constant matrix
    int levels = 0; while ((1 << levels) < Nx) ++levels;
    std::vector<double> pcr((size_t)(levels * 2 + 1) * Nx, 0.0);
    {
        std::vector<double> la(Nx), lb(Nx), lc(Nx), na(Nx), nb(Nx), nc(Nx);   // 0-based rows
        for (int i = 0; i < Nx; ++i) { la[i] = (i > 0) ? a[i] : 0.0; lb[i] = b[i + 1]; lc[i] = (i < Nx - 1) ? cc[i + 1] : 0.0; }
        int stride = 1;
        for (int lev = 0; lev < levels; ++lev) {
            double *al = &pcr[(size_t)lev * 2 * Nx], *ga = al + Nx;
            for (int i = 0; i < Nx; ++i) {
                double alpha = 0.0, gamma = 0.0, aa = 0.0, cq = 0.0, bb = lb[i];
                if (i - stride >= 0) { alpha = -la[i] / lb[i - stride]; aa = alpha * la[i - stride]; bb += alpha * lc[i - stride]; }
                if (i + stride < Nx) { gamma = -lc[i] / lb[i + stride]; cq = gamma * lc[i + stride]; bb += gamma * la[i + stride]; }
                al[i] = alpha; ga[i] = gamma; na[i] = aa; nb[i] = bb; nc[i] = cq;
            }
            la.swap(na); lb.swap(nb); lc.swap(nc);
            stride <<= 1;
        }
        double *inv = &pcr[(size_t)levels * 2 * Nx];
        for (int i = 0; i < Nx; ++i) inv[i] = 1.0 / lb[i];
    }
    // ---- IPIC_TRIDIAG_SCAN: generators of the inverse (long double), checked against a long-double Thomas solve on a rough right-hand side ----
    c.scan_ok = false;
    if (c.scheme != IPIC_SCHEME_NGP_PERIODIC && Nx >= 8) {
        typedef long double ld;
        auto ident = [&](int i) { return b[i] == 1.0 && cc[i] == 0.0 && (i == 1 || a[i - 1] == 0.0); };
        const int iL = ident(1) ? 2 : 1, iR = ident(Nx) ? Nx - 1 : Nx;
        std::vector<ld> u(Nx + 2, 0.0L), v(Nx + 2, 0.0L), sc(Nx + 2, 0.0L);
        bool ok = (iR - iL >= 3);
        for (int i = iL; ok && i < iR; ++i) ok = (cc[i] != 0.0 && a[i] != 0.0);
        if (ok) {
            // row scales that make the matrix symmetric: s_i c_i = s_{i+1} a_i (Neumann rows of the NGP matrix carry a factor 2)
            sc[iL] = 1.0L;
            for (int i = iL; i < iR; ++i) sc[i + 1] = sc[i] * (ld)cc[i] / (ld)a[i];
            u[iL] = 1.0L;
            u[iL + 1] = -((ld)b[iL] * u[iL]) / (ld)cc[iL];
            for (int i = iL + 1; i < iR; ++i) u[i + 1] = -((ld)a[i - 1] * u[i - 1] + (ld)b[i] * u[i]) / (ld)cc[i];
            v[iR] = 1.0L;
            v[iR - 1] = -((ld)b[iR] * v[iR]) / (ld)a[iR - 1];
            for (int i = iR - 1; i > iL; --i) v[i - 1] = -((ld)cc[i] * v[i + 1] + (ld)b[i] * v[i]) / (ld)a[i - 1];
            const int j = (iL + iR) / 2;
            const ld W = sc[j] * (ld)cc[j] * (u[j] * v[j + 1] - u[j + 1] * v[j]);
            ok = std::isfinite((double)W) && W != 0.0L;
            if (ok) {
                // check: a rough right-hand side through the generators against Thomas, both in long double
                std::vector<ld> d(Nx + 2), x(Nx + 2), ref(Nx + 2), cpl(Nx + 2), ml(Nx + 2);
                unsigned long long st = 88172645463325252ull;
                for (int i = 1; i <= Nx; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; d[i] = (ld)((double)(st >> 11) / 9007199254740992.0 - 0.5) * 1.0e3L; }
                ref = d;
                cpl[1] = (ld)cc[1] / (ld)b[1]; ml[1] = (ld)b[1]; ref[1] = ref[1] / ml[1];
                for (int i = 2; i <= Nx; ++i) {
                    ml[i] = (ld)b[i] - cpl[i - 1] * (ld)a[i - 1];
                    cpl[i] = (i < Nx) ? (ld)cc[i] / ml[i] : 0.0L;
                    ref[i] = (ref[i] - ref[i - 1] * (ld)a[i - 1]) / ml[i];
                }
                for (int i = Nx - 1; i >= 1; --i) ref[i] = ref[i] - cpl[i] * ref[i + 1];
                std::vector<ld> dd(Nx + 2, 0.0L), P(Nx + 2, 0.0L), Q(Nx + 3, 0.0L);
                for (int i = iL; i <= iR; ++i) {
                    dd[i] = d[i];
                    if (iL == 2 && i == 2) dd[i] -= (ld)a[1] * d[1];
                    if (iR == Nx - 1 && i == Nx - 1) dd[i] -= (ld)cc[Nx - 1] * d[Nx];
                    dd[i] *= sc[i];
                }
                for (int i = iL; i <= iR; ++i) P[i] = P[i - 1] + u[i] * dd[i];
                for (int i = iR; i >= iL; --i) Q[i] = Q[i + 1] + v[i] * dd[i];
                ld err = 0.0L, big = 0.0L;
                for (int i = iL; i <= iR; ++i) {
                    x[i] = (v[i] * P[i] + u[i] * Q[i + 1]) / W;
                    err = std::max(err, fabsl(x[i] - ref[i])); big = std::max(big, fabsl(ref[i]));
                }
                ok = std::isfinite((double)err) && err <= 1.0e-12L * big;
            }
            if (ok) {
                std::vector<double> hu(Nx, 0.0), hv(Nx, 0.0), hs(Nx, 0.0);
                for (int i = iL; i <= iR; ++i) { hu[i - 1] = (double)u[i]; hv[i - 1] = (double)v[i]; hs[i - 1] = (double)sc[i]; }
                for (double **pp : {&c.d_gu, &c.d_gv, &c.d_gs}) { if (*pp) cudaFree(*pp); *pp = nullptr; IPIC_CUDA(cudaMalloc(pp, Nx * sizeof(double))); }
                IPIC_CUDA(cudaMemcpy(c.d_gu, hu.data(), Nx * sizeof(double), cudaMemcpyHostToDevice));
                IPIC_CUDA(cudaMemcpy(c.d_gv, hv.data(), Nx * sizeof(double), cudaMemcpyHostToDevice));
                IPIC_CUDA(cudaMemcpy(c.d_gs, hs.data(), Nx * sizeof(double), cudaMemcpyHostToDevice));
                c.gW = (double)W; c.giL = iL; c.giR = iR; c.scan_ok = true;
            }
        }
    }
    if (c.tridiag_mode == IPIC_TRIDIAG_PCR && c.scan_ok) c.tridiag_mode = IPIC_TRIDIAG_SCAN;     // the default where the matrix allows it
    if (c.tridiag_mode == IPIC_TRIDIAG_SCAN && !c.scan_ok) c.tridiag_mode = IPIC_TRIDIAG_PCR;
    c.pcr_levels = levels;
    if (c.d_pcr) cudaFree(c.d_pcr);
    IPIC_CUDA(cudaMalloc(&c.d_pcr, pcr.size() * sizeof(double)));
    IPIC_CUDA(cudaMemcpy(c.d_pcr, pcr.data(), pcr.size() * sizeof(double), cudaMemcpyHostToDevice));
    IPIC_CUDA(cudaMemcpy(c.d_a, &a[1], (Nx - 1) * sizeof(double), cudaMemcpyHostToDevice));
    IPIC_CUDA(cudaMemcpy(c.d_b, &b[1], Nx * sizeof(double), cudaMemcpyHostToDevice));
    IPIC_CUDA(cudaMemcpy(c.d_c, &cc[1], (Nx - 1) * sizeof(double), cudaMemcpyHostToDevice));
    IPIC_CUDA(cudaMemcpy(c.d_cp, &cp[1], Nx * sizeof(double), cudaMemcpyHostToDevice));
    IPIC_CUDA(cudaMemcpy(c.d_m, &mm[1], Nx * sizeof(double), cudaMemcpyHostToDevice));
    return IPIC_OK;
}

static bool eform(const Ctx &c) { return c.scheme == IPIC_SCHEME_NGP_PERIODIC; }
int field_finalize_rho(Ctx &c) { FieldArgs A = make_args(c); return run_field(c, A, eform(c) ? OP_P_RHO : OP_FINALIZE_RHO); }

int field_poisson(Ctx &c, double timeCurrent) {
    // CIC: RF wall value refreshed first (:181-189)
    if (c.scheme == IPIC_SCHEME_CIC && c.RF_bool) {
        const double v = c.RF_half_amplitude * std::sin(c.RF_rad_frequency * timeCurrent);
        if (c.RF_indx == 1) { c.boundPhi[0] = v; c.sourceTermVals[0] = -v * 2.0 / c.h_dx[0]; }
        else { c.boundPhi[1] = v; c.sourceTermVals[1] = -v * 2.0 / c.h_dx[c.Nx - 1]; }
    }
    FieldArgs A = make_args(c);
    A.poisson_rf_value = c.RF_half_amplitude * std::sin(c.RF_rad_frequency * timeCurrent);
    return run_field(c, A, eform(c) ? OP_P_POISSON : OP_POISSON);
}

int field_const_source(Ctx &c) { if (eform(c)) return IPIC_OK; FieldArgs A = make_args(c); return run_field(c, A, OP_CONST_SOURCE); }
int field_half_time_E(Ctx &c) { FieldArgs A = make_args(c); return run_field(c, A, eform(c) ? OP_P_HALF_E : OP_HALF_E); }
int field_finalize_J(Ctx &c, double del_t) { FieldArgs A = make_args(c); A.del_t = del_t; return run_field(c, A, eform(c) ? OP_P_FINALIZE_J : OP_FINALIZE_J); }
int field_ampere(Ctx &c, double del_t) { FieldArgs A = make_args(c); A.del_t = del_t; return run_field(c, A, eform(c) ? OP_P_AMPERE : OP_AMPERE); }

int field_diag(Ctx &c, int which, double del_t, double *out) {
    FieldArgs A = make_args(c);
    A.del_t = del_t; A.diag_which = which;
    int rc = run_field(c, A, OP_DIAG);
    if (rc) return rc;
    IPIC_CUDA(cudaMemcpyAsync(c.h_scal, c.d_scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    IPIC_CUDA(cudaStreamSynchronize(c.stream));
    *out = c.h_scal[1];
    return IPIC_OK;
}

int field_anderson_begin(Ctx &c) {
    FieldArgs A = make_args(c);
    if (eform(c)) return run_field(c, A, OP_AND_STORE0 | OP_P_HALF_E);
    return run_field(c, A, OP_CONST_SOURCE | OP_AND_STORE0 | OP_HALF_E);
}

// phase 0: after the first sweep   : finalize J, Ampere, column 2 + R1 + norm
// phase 1: after sweep i           : finalize J, Ampere, R_index + norm
// phase 2: mixing for iteration i  : next iterate into phi_f, then E + tables for the next sweep
// phase 3: E + tables only (converged: the mover uses the freshly solved phi_f)
int field_anderson_step(Ctx &c, int i, int phase, double del_t, int *status, double *norm) {
    FieldArgs A = make_args(c);
    A.del_t = del_t; A.and_i = i;
    unsigned ops = 0;
    const unsigned fj = eform(c) ? OP_P_FINALIZE_J : OP_FINALIZE_J, am = eform(c) ? OP_P_AMPERE : OP_AMPERE, he = eform(c) ? OP_P_HALF_E : OP_HALF_E;
    if (phase == 0) ops = fj | am | OP_AND_FIRST | he;
    else if (phase == 1) ops = fj | am | OP_AND_RESID;
    else if (phase == 2) ops = OP_AND_MIX | he;
    else ops = he;
    // h_scal[48] receives the norm, h_scal[49] the sequence number of this launch
    volatile unsigned long long *h_seq = reinterpret_cast<volatile unsigned long long *>(c.h_scal + 49);
    const bool flag = c.norm_flag && phase <= 1;
    if (flag) { A.h_norm = c.h_scal + 48; A.h_seq = reinterpret_cast<unsigned long long *>(c.h_scal + 49); A.seq = ++c.norm_seq; }
    int rc = run_field(c, A, ops);
    if (rc) return rc;
    if (flag) {
        // spin on the sequence number; the stream is polled now and then so that a failed launch cannot hang the host
        for (unsigned spins = 1;; ++spins) {
            if (*h_seq == A.seq) break;
            if ((spins & 0x3FFF) == 0) {
                const cudaError_t q = cudaStreamQuery(c.stream);
                if (q == cudaSuccess) { if (*h_seq == A.seq) break; set_error("k_field finished without publishing the residual norm"); return IPIC_ERR_CUDA; }
                if (q != cudaErrorNotReady) { set_error(std::string("k_field: ") + cudaGetErrorString(q)); return IPIC_ERR_CUDA; }
            }
        }
        *norm = *reinterpret_cast<volatile double *>(c.h_scal + 48);
    } else if (phase <= 1) {
        IPIC_CUDA(cudaMemcpyAsync(c.h_scal, c.d_scal, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        IPIC_CUDA(cudaStreamSynchronize(c.stream));
        *norm = c.h_scal[0];
    }
    if (status) *status = 0;
    return IPIC_OK;
}

// One whole Anderson iteration in one launch, decision included: finalize J, Ampere, residual + norm, convergence / stagnation test, then
// mixing + E-half + push tables (continue), E-half + tables only (converged) or nothing (stagnated).  The launch is a no-op when an
// earlier iteration ended the solve.  The host does not wait for it (solve_div_ampere_anderson reads the outcome once per chunk).
int field_anderson_iter(Ctx &c, int i, double del_t, double eps_tol, unsigned long long seq) {
    FieldArgs A = make_args(c);
    A.del_t = del_t; A.and_i = i;
    const unsigned fj = eform(c) ? OP_P_FINALIZE_J : OP_FINALIZE_J, am = eform(c) ? OP_P_AMPERE : OP_AMPERE, he = eform(c) ? OP_P_HALF_E : OP_HALF_E;
    A.and_state = c.d_and_state; A.eps_tol = eps_tol; A.eform = eform(c) ? 1 : 0;
    A.h_and = c.h_and; A.h_and_norms = c.h_and_norms; A.seq = seq;
    return run_field(c, A, fj | am | OP_AND_RESID | OP_AND_DECIDE | OP_AND_MIX | he);
}

// funcNitsol tail: phi_f already holds G(x); f = x - phi_f
int field_residual(Ctx &c, const double *d_x, double *d_f) {
    FieldArgs A = make_args(c);
    A.xin = d_x; A.fout = d_f;
    return run_field(c, A, OP_RESID_F);
}

}  // namespace ipic
