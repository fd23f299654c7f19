// Shared declarations of the B200 implicit-PIC library: context, device tables, PTX wrappers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <map>
#include "../../include/ipic1d_b200.h"

namespace ipic {

// constants.f90:7-15
constexpr double eps_0 = 8.8541878128e-12;
constexpr double e_c = 1.602176634e-19;

// ---- fixed-point deposition ---------------------------------------------------------------------
// Every deposited weight x (|x| <= 2) is rounded to a multiple of 2^-FX_BITS and accumulated as a
// 64-bit integer, so sums are order-independent: bit-identical run to run, for any particle order,
// block count or GPU count.  x + FX_MAGIC puts round(x * 2^FX_BITS) into the low mantissa bits.
constexpr int FX_BITS = 44;
constexpr double FX_MAGIC = 384.0;                                  // 1.5 * 2^(52 - FX_BITS)
constexpr double FX_INV = 1.0 / 17592186044416.0;                   // 2^-44

__device__ __forceinline__ long long fx_from_double(double x) {
    return __double_as_longlong(x + FX_MAGIC) - __double_as_longlong(FX_MAGIC);
}

// Per-cell, per-species lookup tables rebuilt from (phi, phi_f) before every sweep.
struct CellTabNGP {   // OpenMP/NGP/src/mod_particleMover.f90:21-63 operands
    double a;          // q/m * E_c
    double hinv;       // 0.5 / dx_dl(c)
    double two_a_dx;   // 2 a dx_dl(c)
    double inv_a;      // 1 / a
};
struct CellTabCIC {   // OpenMP/CIC/src/mod_particleMoverPicard.f90:17-65,92-94 operands
    double A0;         // q/m * E_left
    double A1;         // q/m * (E_right - E_left)
    double inv_dx;     // 1 / dx_dl(c)
    double dx;         // dx_dl(c)
    double cap;        // particleTimeStep(c, species)
    double pad[3];
};

// streaming geometry of the push kernels
#ifndef IPIC_PUSH_THREADS
#define IPIC_PUSH_THREADS 1024
#endif
constexpr int PUSH_THREADS = IPIC_PUSH_THREADS;
constexpr int PUSH_WARPS = PUSH_THREADS / 32;
constexpr unsigned ABS_CAP = 1u << 20;   // absorbed particles per species and move the list holds (more: the array is counted instead)
constexpr int SD_CELLS_MAX = 5632;    // 22 KB behind the 141 KB of warp queues: still the 164 KB shared-memory configuration of the SM
constexpr int MAX_GRID_SMEM = 8192;   // largest N_h the shared-memory limb grid supports (12 B per cell)

struct PushArgs {
    const double *l, *vx;      // inputs (SoA)
    double *l_out, *vx_out;    // MOVE: in-place outputs
    long long n;               // particles in the arrays (alive or marked dead)
    const void *tab;           // CellTabNGP* / CellTabCIC* for this species
    const double2 *ttab;       // NGP: transit tables of this species, 16 B per cell (the lean loop's only gather): [0, Nh+2) = {2 a dx, 1/a} by cell for
                               // right-movers, [Nh+2, 2 Nh+4) = {-2 a dx, -1/a} in reverse cell order for left-movers
    int bcL, bcR;              // boundaryConditions at the two walls (zero everywhere else)
    unsigned *tile_counter;    // dynamic tile scheduler (zeroed before the launch)
    unsigned long long *acc;   // !MOVE: int64 fixed-point grid [Nh] for this species
    unsigned long long *dacc;  // !MOVE: difference grid of the runs of fully traversed cells (prefix-summed by the field kernel)
    unsigned long long acc_rep_stride;   // !MOVE: the grids exist in acc_reps copies, this many entries apart; CTA b deposits into copy b % acc_reps
    int acc_reps;
    unsigned tile_stride;      // stride permutation of the tile order
    int tile_subs;             // 64-particle sub-tiles per scheduled tile (1, 2, 4 or 8)
    int tile_mode;             // 0 array order, 1 stride permutation, 2 from both ends of the (cell-sorted) array inwards: the wall regions, whose
                               // electrons take the most sub-steps, are claimed first and the cheap bulk fills the end of the launch
    unsigned long long *counters;  // [0] deposit substeps, [1] func evals, [2] absorbed left, [3] absorbed right, [4] first reflections, [5] move substeps
    int Nh;                    // grid entries to deposit on
    int NB;                    // wall index on the right: N_x (NGP) or N_h (CIC)
    double del_t, f_tol;
    double t_lim;              // smallest double t with !(del_t - t > f_tol): `t < t_lim` is the reference's `del_t - t > f_tol`
    unsigned long long tlim_bits;   // bit pattern of t_lim (0 when t_lim <= 0): for t >= 0, `bits(t) < tlim_bits` as unsigned integers is `t < t_lim`,
                               // and NaN, -0 and negative t compare high
    unsigned *abs_list;        // MOVE, arrays without slab identity: indices of the particles absorbed in this launch, in no order (null: not recorded)
    unsigned *abs_count;       // entries appended (may exceed abs_cap: then the list is incomplete and the caller counts the array instead)
    unsigned abs_cap;
    unsigned char *refl;       // MOVE, Neumann walls only: per-particle flag "reflected during this step" (refRecordIdx), else null
    const int *stop;           // depositJ sweeps enqueued ahead of the nonlinear solve's decision: *stop != 0 (converged / stagnated) makes the launch a no-op
    int stay_min;              // a queued batch keeps looping in registers while at least this many lanes continue
    int transit_min;           // the lean face-to-face loop runs while at least this many lanes are in transit
    int sd_cells;              // ENTRY kernels: entries of the per-CTA shared-memory copy of the difference grid (0: run ends go to the global grid)
    int exit_stage;            // ENTRY kernels: a lane the entry stage has taken into the next cell is finished there when it stays in that cell
    int entry_min;             // ENTRY kernels: a fresh batch takes its first sub-step to the face ahead in place when at least this many lanes leave their cell; 33 = never
};

// ---- device context -------------------------------------------------------------------------------
struct SpeciesDev {
    double mass = 0, q = 0, w_p = 1, q_over_m = 0, q_times_wp = 0;
    long long n = 0, cap = 0;
    double *l = nullptr, *vx = nullptr, *vy = nullptr, *vz = nullptr;
    double *l2 = nullptr, *vx2 = nullptr, *vy2 = nullptr, *vz2 = nullptr;
    long long cap2 = 0;
    std::vector<long long> slab_off;       // numThread+1 offsets of the reference's thread slabs
    ipic_move_stats stats{};
    long long depositSubsteps = 0;
    bool abs_listed = false;               // the last move recorded the indices of the particles it absorbed (Ctx::d_abs)
    double sweep_substeps_per_particle = 0.0;   // of the last implicit step's depositJ sweeps (0 = not measured yet): picks the push kernel variant
    long long n_collide = -1;              // numToCollide (mod_particle.f90:17); < 0 = all
    // null-collision sets that share this primary within one step draw their candidates from ONE keyed permutation of the collidable
    // range, one after the other: a particle is selected at most once per step, as in the reference, whose serial loop swaps every
    // selected particle behind the range and stores the reduced numToCollide (mod_NullCollision.f90:221-231)
    long long coll_taken = 0;              // candidates already drawn since numToCollide was last set
    unsigned long long coll_epoch = 0;     // bumped whenever numToCollide is set: keys the permutation of the step
    unsigned long long coll_seed = 0;      // seed of the first collision set created with this primary
    // refRecordIdx (mod_particle.f90:19, mod_particleMover.f90:226-231): particles first reflected at a Neumann wall in the last move
    unsigned char *refl = nullptr; long long refl_cap = 0;      // per-particle flags written by the move kernel
    unsigned *refList = nullptr; long long refList_cap = 0;     // flagged indices, ascending; after the move: post-compaction indices
    long long nRef = 0;
    // refLive[j] = 1: entry j names the reflected particle itself; 0: that particle was absorbed later in the same step, so the entry
    // (kept for parity with the reference's list) names its successor's slot -- the refluxing kernels skip it
    unsigned char *refLive = nullptr; long long refLive_cap = 0;
    std::vector<long long> slab_off_prev;  // thread-slab offsets before the last compaction: the slab a reflection was recorded in
    std::vector<int> refSlabCount;         // refIdx(T): entries of refList per thread slab, in list order
};

// ---- null-collision MCC (collisions.cu) -------------------------------------------------------------
struct TargetDev { double mass = 0, density = 0, temperature = 0, v_therm = 0; };
struct CollSet {
    int nC = 0, nE = 0, primary = 0, target = 0;
    int type[IPIC_MAX_COLLISIONS] = {}; int ionSpecies[IPIC_MAX_COLLISIONS] = {};
    double thr[IPIC_MAX_COLLISIONS] = {};
    double sigmaVMax = 0, reducedMass = 0, sumMass = 0, redMassIon = 0, minE = 0, maxE = 0;
    double *d_energy = nullptr, *d_sigma = nullptr;      // [nE], [nC][nE]
    unsigned long long seed = 0, calls = 0;
    bool has_ionization = false;
    ipic_collision_stats stats{};
};

struct Timing { double ms = 0; long long launches = 0; };

struct Ctx {
    int scheme = 0, Nx = 0, Nh = 0, nbc = 0, ndx = 0, smooth = 0, device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    long long launches = 0;                                       // kernels launched by this context
    // device tables
    double *d_grid = nullptr, *d_dx = nullptr, *d_hinv = nullptr;   // d_hinv = 0.5 / dx_dl
    int *d_bc = nullptr;                                          // boundaryConditions, 1-based (d_bc[k] = bc(k))
    std::vector<double> h_grid, h_dx;
    std::vector<int> h_bc;
    // field vectors (device)
    double *d_phi = nullptr, *d_phi_f = nullptr, *d_rho = nullptr, *d_J = nullptr, *d_E = nullptr;
    double *d_Ef = nullptr, *d_Eh = nullptr;                      // NGP_Periodic: EField_f, EField_half
    double L_domain = 0.0;
    double *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;       // tridiagonal
    double *d_cp = nullptr, *d_m = nullptr;                       // Thomas: c' and pivots m_i (constant matrix)
    double *d_pcr = nullptr; int pcr_levels = 0;                  // PCR level coefficients [levels][2][Nx] + inv diag [Nx]
    // IPIC_TRIDIAG_SCAN: generators of the inverse of the (row-scaled, symmetric) tridiagonal matrix on the unknowns [iL, iR] -- u solves the
    // homogeneous recurrence from the left end, v from the right end, gs are the row scales, gW = s_j c_j (u_j v_j+1 - u_j+1 v_j) -- built
    // in long double by field_setup and checked against the Thomas solve; scan_ok = false (periodic / singular matrices) keeps PCR
    double *d_gu = nullptr, *d_gv = nullptr, *d_gs = nullptr; double gW = 0.0; int giL = 1, giR = 0; bool scan_ok = false;
    double *d_work = nullptr;                                     // scratch (4 * max(Nx,Nh))
    double *d_rho_i = nullptr;
    // [acc_reps][nsp_max][2][max(Nx,Nh)+2] int64 accumulators: grid, difference grid.  The copies spread the global RED traffic of the
    // push kernel over more addresses (and L2 slices): on a 256-cell grid every CTA otherwise hammers the same 2 KB.  Integer sums, so
    // the split changes nothing in the result; the field kernel adds the copies when it collects.
    unsigned long long *d_acc = nullptr;
    int acc_reps = 1; size_t acc_rep_stride = 0;
    double2 *d_ttab = nullptr;                                    // [nsp_max][2][cells] NGP transit tables (by direction of travel)
    void *d_tab = nullptr;                                        // [nsp_max][cells] CellTab*
    unsigned long long *d_counters = nullptr;                     // [nsp_max][8]
    double *d_scal = nullptr;                                     // small device scalar outputs (64 doubles)
    double *h_scal = nullptr;                                     // pinned mirror
    // Anderson
    double *d_phi_k = nullptr, *d_res_k = nullptr;                // [m+1][Nx]
    double *d_norms = nullptr;
    // Anderson loop with the decision on the device (field.cu OP_AND_DECIDE): the host enqueues `and_chunk` iterations ahead -- sweeps, the
    // J all-reduce, the fused field kernel -- and reads the outcome once per chunk.  d_and_state[0]: 0 running, 1 converged, 2 stagnated;
    // [1]: the iteration that ended the solve.  h_and (pinned): [0] sequence number of the last field kernel that reported, [1] state,
    // [2] iteration; h_and_norms[i]: residual norm of iteration i
    int *d_and_state = nullptr;
    volatile long long *h_and = nullptr; double *h_and_norms = nullptr; int cap_and_norms = 0;
    int and_chunk = 6;
    // thread-block cluster of the fused iteration's field kernel (field.cu): compact copies of the collected grids, partial sums of the
    // normal equations per CTA; field_cluster = CTAs in the cluster (1 = single CTA)
    long long *d_cmp = nullptr; double *d_gram_part = nullptr; int field_cluster = 8;
    long long noop_sweeps = 0;                                    // sweeps enqueued ahead that turned out to be no-ops (timing counts exclude them)
    int tridiag_mode = IPIC_TRIDIAG_PCR;
    // solver scalars (host mirror; NGP mod_potentialSolver.f90:17-22, CIC :17-23)
    double rho_const = 0, RF_rad_frequency = 0, RF_half_amplitude = 0, RF_voltage = 0;
    double boundPhi[2] = {0, 0}, sourceTermVals[2] = {0, 0};
    int RF_indx = 0; bool RF_bool = false;
    bool solver_ready = false;
    // nonlinear solver (mod_nonLinSolvers.f90:29-35)
    int iterNumPicard = 0, iterNumAdaptiveSteps = 0, amountTimeSplits = 0;
    int maxIter = 100, solverType = 0, m_Anderson = 10;
    double Beta_k = 0.5, eps_r = 0, eps_a = 1e-8;
    long long sweeps = 0, substeps = 0;
    std::vector<double> normHist;
    void *nitsol_state = nullptr;                                 // nitsol::Solver (solverType 1)
    int nitsol_info[7] = {0, 0, 0, 0, 0, 0, 0};                   // nfe, njve, nrpre, nli, nni, nbt, iterm of the last JFNK solve
    double lastNorm = 0;
    std::vector<SpeciesDev> sp;
    std::vector<TargetDev> targets;
    std::vector<CollSet> colls;
    // MCC scratch
    double *d_mcc = nullptr; long long cap_mcc = 0; void *d_mcc_tmp = nullptr; size_t cap_mcc_tmp = 0;
    // particle injection / refluxing / heating (injection.cu)
    double *d_inj = nullptr; long long cap_inj = 0;
    unsigned long long inj_seed = 0x1A2B3C4D5E6F7081ULL, inj_calls = 0;
    // histogram / extrema scratch (diagnostics.cu)
    unsigned long long *d_diag = nullptr; long long cap_diag = 0;
    // options
    int stay_min = 20;
    int transit_min = 12;
    int entry_min = 8;         // 33 = no entry stage for any species
    int sd_runs = 1;           // entry-stage kernels count the ends of runs of traversed cells in shared memory (PushArgs::sd_cells)
    int exit_stage = 1;        // the entry stage also finishes the lanes that end their step in the cell they enter
    int sort_subbins = 0;      // cell sort key = INT(l * sub); 0 = automatic (8, or what 16 key bits allow)
    int pair_species = 0;      // 1 = depositJ of a two-species plasma in one launch (k_push_pair).  Off: measured slower (C5: 2.34 ms per
                               // sweep against 1.88 + 0.32 ms) -- both species' sweeps are issue-bound, so running them side by side frees nothing
    int pair_warps = 4;        // warps per CTA (of 32) that start with the second species
    int tile_order = 1;           // 1 = stride-permuted tile order, 0 = in array order, 2 = from both ends inwards
    int tile_subs = 0;            // 0 = automatic (8 sub-tiles of 64 particles, fewer for small loads)
    int push_blocks = 0;          // 0 = one CTA per SM
    int lazy_compaction = 0;
    int fill_compaction = 1;      // arrays without slab identity (device loader, sorted): absorbed particles' slots are filled from the tail instead of a stable scatter
    unsigned *d_fill = nullptr; long long cap_fill = 0;
    // indices of the particles the last move absorbed, appended by the move kernel itself (arrays without slab identity): the compaction
    // then touches the absorbed particles only instead of reading every l.  [NSP_MAX][2][ABS_CAP] (as appended, sorted) + [NSP_MAX] counts
    unsigned *d_abs = nullptr; unsigned *d_abs_count = nullptr;
    int norm_flag = 1;            // 1 = k_field publishes the residual norm into pinned host memory and the host spins on a sequence number;
                                  // 0 = D2H copy + cudaStreamSynchronize per nonlinear iteration
    unsigned long long norm_seq = 0;
    unsigned tile_zero_mask = 0;  // bit s set: species s' tile counter was zeroed on the stream by the last half-E op and not consumed yet, so the
                                  // push launch skips its own memset (one stream operation less between two kernels)
    // multi-GPU
    void *nccl_comm = nullptr; int nranks = 1, rank = 0;
    // peer-memory all-reduce of the deposition grids (peer.cu): this rank's window (cudaMalloc, exported through CUDA IPC) and the mapped
    // windows of all ranks of the node
    void *peer_window = nullptr; std::vector<void *> peer_ptr; bool peer_ready = false; int peer_reduce = 1;
    unsigned long long peer_seq = 0;
    // timing
    std::map<std::string, Timing> timing;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing_on = true;
    int num_sms = 148;
    // compaction scratch
    unsigned *d_tileCount = nullptr; long long *d_tileOffset = nullptr; double *d_loss = nullptr;
    long long cap_tiles = 0; long long *d_total = nullptr; long long *d_slab = nullptr; int cap_slab = 0;
    // cell-sort scratch (kept between sorts: cudaMalloc/cudaFree of GB-sized buffers costs more than the sort)
    unsigned *d_sort = nullptr; long long cap_sort = 0; void *d_sort_tmp = nullptr; size_t cap_sort_tmp = 0;
};

static constexpr int NSP_MAX = 8;

void set_error(const std::string &s);
#define IPIC_CUDA(call)                                                                             \
    do {                                                                                            \
        cudaError_t _e = (call);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::ipic::set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
            return IPIC_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

// ---- launchers implemented in particles.cu ---------------------------------------------------------
int launch_deposit(Ctx &c, int sp, double del_t, const int *stop = nullptr);
void push_stats_dump();                                 // measurement builds (-DIPIC_PUSH_STATS) only: prints and clears the lane-utilisation counters
int launch_move(Ctx &c, int sp, double del_t);
int launch_deposit_pair(Ctx &c, int a, int b, double del_t);
int launch_deposit_rho(Ctx &c, int sp, bool density = false);   // density: the diagnostic's weights (differs from rho at CIC Dirichlet walls)
int extract_species_grid(Ctx &c, int sp, int n, double *host_out);   // int64 grid of one species -> doubles on the host, grid zeroed
int compact_species(Ctx &c, int sp);
int local_temperature(Ctx &c, int sp, int numCells, double *sumV2_host, long long *count_host);
int fold_acc(Ctx &c, size_t count);                    // CTA copies of the deposition grids -> copy 0
int collect_reflected(Ctx &c, int sp);                 // flagged pre-compaction indices -> refList
int map_reflected(Ctx &c, int sp, bool compacted);     // -> post-compaction indices (absorbed ones dropped)
int particle_moments(Ctx &c, int sp, double out4[4]);      // sum v^2, sum vx, vy, vz
int upload_species(Ctx &c, int sp, const double *ps, const int32_t *N_p, int T, int finalIdx);
int download_species(Ctx &c, int sp, double *ps, int32_t *N_p, int T, int finalIdx);
int generate_species(Ctx &c, int sp, long long n, double T_eV, double v_drift, unsigned long long seed, double alpha, int distType);
int sort_species(Ctx &c, int sp);
int ensure_capacity(Ctx &c, int sp, long long n);

// ---- peer.cu: all-reduce of the deposition grids through peer memory (NVLink), fused with the fold of the CTA copies ----
int peer_export(Ctx &c, void *handle64);
int peer_import(Ctx &c, int nranks, int rank, const void *handles);
int peer_allreduce_acc(Ctx &c, size_t count);
int peer_check(Ctx &c);
void peer_free(Ctx &c);

// ---- collisions.cu -----------------------------------------------------------------------------------
int collision_set_create(Ctx &c, int nC, int nE, double red_mass, double sumMass, double redMassTriple, const double *energyArray, const double *sigmaArray,
                         const double *energyThreshold, const int32_t *collisionType, const int32_t *reactantsIndx, const int32_t *productsIndx,
                         unsigned long long seed, int *out);
int generate_collision(Ctx &c, int set, double del_t);
void collisions_free(Ctx &c);

// ---- injection.cu (mod_particleInjection.f90) -----------------------------------------------------------
int add_maxwellian_lost_particles(Ctx &c, double T_e, double T_i, double *energyAdd, long long *numAdded);
int reflux_particles(Ctx &c, double T_e, double T_i, double *energyAdd);
int inject_at_boundary(Ctx &c, double T_e, double T_i, double del_t, double numFluxPart, long long *numInjected);
int inject_uniform_flux(Ctx &c, double T_e, double T_i, double numFluxPart, long long *numInjected);
int maxwellian_heating(Ctx &c, int sp, double compValue, double T_eV, long long *numHeated);
void injection_free(Ctx &c);

// ---- diagnostics.cu (mod_simulation.f90:437-500) -------------------------------------------------------------
int velocity_extrema(Ctx &c, int sp, double out3[3]);
int velocity_histograms(Ctx &c, int sp, int bins, double VMax, double Emin, double EMax, double *VHist, double *EHist);
void diagnostics_free(Ctx &c);

// ---- launchers implemented in field.cu --------------------------------------------------------------
int field_setup(Ctx &c);                                   // construct_diagMatrix + Thomas/PCR factor tables
int field_finalize_rho(Ctx &c);
int field_poisson(Ctx &c, double timeCurrent);
int field_const_source(Ctx &c);
int field_half_time_E(Ctx &c);                             // EField + per-species cell tables
int field_finalize_J(Ctx &c, double del_t);
int field_ampere(Ctx &c, double del_t);
int field_diag(Ctx &c, int which, double del_t, double *out);
int field_anderson_begin(Ctx &c);
int field_anderson_step(Ctx &c, int i, int phase, double eps_tol, int *status, double *norm);
int field_anderson_iter(Ctx &c, int i, double del_t, double eps_tol, unsigned long long seq);   // fused iteration, decision on the device
int field_residual(Ctx &c, const double *d_x, double *d_f);

// ---- counter-based PCG32 (one short stream per work item) ------------------------------------------------
__host__ __device__ __forceinline__ unsigned pcg32_next(unsigned long long &state, unsigned long long inc) {
    const unsigned long long old = state;
    state = old * 6364136223846793005ULL + inc;
    const unsigned xorshifted = (unsigned)(((old >> 18u) ^ old) >> 27u);
    const unsigned rot = (unsigned)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
}
__host__ __device__ __forceinline__ double pcg_uniform(unsigned long long &state, unsigned long long inc) {
    // (res + 1) / (2^32 + 1): in (0,1), the mapping of PIC1D/SharedModules/src/PCG_func.c:9-20
    return ((double)pcg32_next(state, inc) + 1.0) / 4294967297.0;
}
// stream `gid` of generator `seed`
__host__ __device__ __forceinline__ void pcg_seed(unsigned long long &state, unsigned long long &inc, unsigned long long seed, unsigned long long gid) {
    inc = ((seed ^ (gid * 0x9E3779B97F4A7C15ULL)) << 1u) | 1u;
    state = 0ULL;
    pcg32_next(state, inc);
    state += seed + gid;
    pcg32_next(state, inc);
}

// ---- keyed bijection of [0, n): how the MCC draws its collision candidates without replacement -------------
// Four rounds of (add round key, xorshift, odd multiply) on `bits`-bit integers -- each step is invertible modulo 2^bits, so
// the composition is a permutation of [0, 2^bits); values >= n are walked through the permutation again (cycle walking,
// < 2 trips on average since 2^bits < 2n).
__host__ __device__ __forceinline__ int mcc_bits_hd(unsigned long long n) {
    int b = 1;
    while (b < 63 && (1ull << b) < n) ++b;
    return b;
}
__host__ __device__ __forceinline__ unsigned long long mcc_permute_hd(unsigned long long x, unsigned long long n, int bits, unsigned long long key) {
    const unsigned long long mask = (1ull << bits) - 1ull;
    const int s1 = (bits + 1) / 2, s2 = (bits + 2) / 3;
    do {
        unsigned long long z = key;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            z += 0x9E3779B97F4A7C15ULL;                                  // splitmix64 round keys
            unsigned long long k = z;
            k = (k ^ (k >> 30)) * 0xBF58476D1CE4E5B9ULL;
            k = (k ^ (k >> 27)) * 0x94D049BB133111EBULL;
            k ^= k >> 31;
            x = (x + k) & mask;
            x ^= x >> s1;
            x = (x * 0xD6E8FEB86659FD93ULL) & mask;
            x ^= x >> s2;
            x = (x * 0xCA5A826395121157ULL) & mask;
            x ^= x >> s1;
        }
    } while (x >= n);
    return x;
}

// ---- PTX wrappers: mbarrier + bulk async copy (TMA 1-D) ----------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

}  // namespace ipic
