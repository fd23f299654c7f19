// =============================================================================
// ipic_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// A plain C++17/OpenMP restatement of the implicit-PIC inner loop of
// nsavard01/ImplicitPic1D (trees OpenMP/NGP and OpenMP/CIC).  It exists only
// to check the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
// cpu_baseline / --impl reference leg).  Nothing under implicitpic1d_b200/
// may import, link or call it.
//
// PARITY STATUS: "parity unpinned" against the Fortran binary -- the reference
// ships no golden vectors for this path, cannot be compiled here (no Fortran
// compiler, Intel-only modules, MKL) and seeds its RNG from the OS.  The
// oracle is pinned instead on the reference's own acceptance invariants
// (main.f90:93-114: energy error, Gauss error, charge-continuity error) and on
// analytic known-answer tests (tests/test_oracle_*.py).  The only outputs of
// the real binary available -- the start-up lines of the run logs it ships
// (OpenMP/{NGP,CIC,NGP_Periodic}/src/output.txt: time step, grid spacing,
// species constants, null-collision sets) -- ARE pinned, to the digits printed
// (tests/golden/reference_run_logs.json, tests/test_reference_log_pins.py);
// the mover / deposition / solver results stay unpinned.
//
// Conventions: arrays are 1-based like the Fortran (index 0 is padding) so
// that every loop bound and stencil reads like the file:line it restates.
// Thread-slab layout phaseSpace(4, finalIdx, T), per-thread private grid
// copies and the thread-ordered / species-ordered reduction are kept, so the
// summation order is the reference's.  Compiled WITHOUT fast-math.
// =============================================================================
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// constants.f90:7-15
constexpr double eps_0 = 8.8541878128e-12;
constexpr double m_amu = 1.66053906660e-27;
constexpr double m_e   = 9.1093837015e-31;
constexpr double e_c   = 1.602176634e-19;
constexpr double k_B   = 1.380649e-23;
const double pi = 4.0 * std::atan(1.0);

inline double fsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }  // Fortran SIGN(a,b)

typedef std::vector<double> dvec;
typedef std::vector<int> ivec;

// ---- mod_BasicFunctions.f90:45-58  ran2 (Park-Miller / Schrage) -------------
inline double ran2(int32_t &irand) {
    const int32_t ia = 16807, im = 2147483647, iq = 127773, ir = 2836;
    const double am = 1.0 / im;
    int32_t k = irand / iq;
    irand = ia * (irand - k * iq) - ir * k;
    if (irand < 0) irand += im;
    return am * irand;
}

// ---- mod_BasicFunctions.f90:76-96  randPCG ----------------------------------
inline double randPCG(uint64_t &state) {
    uint64_t x = state;
    state = x * 6364136223846793005ULL + 1442695040888963407ULL;
    x ^= x >> 18;
    int32_t x_shifted = (int32_t)(uint32_t)(x >> 27);
    int32_t count = (int32_t)(x >> 59);
    uint32_t xs = (uint32_t)x_shifted;
    uint32_t rn = (xs >> count) | (xs << ((-count) & 31));
    double res = (double)(int32_t)rn + 2147483648.0;
    return res / 4294967296.0;
}

struct Domain {
    int scheme = 0;           // 0 NGP, 1 CIC
    int Nx = 0, Nh = 0;       // NumberXNodes, NumberXHalfNodes
    int T = 1;                // numThread
    dvec grid, dx_dl, nodeVol, centerDiff;
    ivec bc;                  // boundaryConditions: NGP size Nx, CIC size Nh
    double L_domain = 0, startX = 0, endX = 0;
    bool smooth = false;
    int nTN = 0, nTH = 0;
    ivec tnL, tnR, thL, thR;  // threadNodeIndx / threadHalfNodeIndx (1-based thread)
};

// NGP mod_domain.f90:62-100 / CIC mod_domain.f90:59-98: node slices per thread
static void make_slices(int N, int T, int &num, ivec &L, ivec &R) {
    num = (T < N) ? T : N;
    L.assign(num + 1, 0); R.assign(num + 1, 0);
    int spacing = N / num - 1, modT = N % num, k = 1;
    for (int i = 1; i <= num; ++i) {
        L[i] = k;
        if (i <= modT) k += spacing + 1; else k += spacing;
        R[i] = k;
        k += 1;
    }
}

// domain_constructor: NGP mod_domain.f90:46-101, CIC mod_domain.f90:43-100
static void domain_init(Domain &w, int scheme, int Nx, int bcL, int bcR, int T) {
    w.scheme = scheme; w.Nx = Nx; w.T = T;
    w.Nh = (scheme == 0) ? Nx - 1 : Nx + 1;      // NGP :370 / CIC :355
    int nb = (scheme == 0) ? Nx : w.Nh;
    w.grid.assign(Nx + 2, 0.0);
    w.dx_dl.assign(std::max(Nx, w.Nh) + 2, 1.0);
    w.nodeVol.assign(Nx + 2, 1.0);
    w.centerDiff.assign(Nx + 2, 0.0);
    w.bc.assign(nb + 2, 0);
    w.bc[1] = bcL; w.bc[nb] = bcR;
    if (bcL == 3 || bcR == 3) { w.bc[1] = 3; w.bc[nb] = 3; }
    make_slices(Nx, T, w.nTN, w.tnL, w.tnR);
    make_slices(w.Nh, T, w.nTH, w.thL, w.thR);
}

// Grid constructors.  NGP mod_domain.f90:104-265; CIC mod_domain.f90:103-252.
// Returns 0 ok, nonzero = the reference would `stop`.
static int domain_grid(Domain &w, double del_x, double L, int gridType) {
    const int Nx = w.Nx, Nh = w.Nh;
    if (w.scheme == 0) {
        w.grid[1] = 0.0; w.grid[Nx] = L;
        switch (gridType) {
        case 0:  // constructUniformGrid :216-225
            for (int i = 2; i <= Nh; ++i) w.grid[i] = (i - 1) * L / Nh;
            break;
        case 1:  // constructSineGrid :127-144
            if (del_x / L >= 1.0 / ((double)Nx - 1.0)) return 1;
            for (int i = 2; i <= Nh; ++i)
                w.grid[i] = L * (((double)i - 1.0) / ((double)Nx - 1.0) - (1.0 / ((double)Nx - 1.0) - del_x / L)
                                 * std::sin(2 * pi * (i - 1) / (Nx - 1)) / std::sin(2 * pi / (Nx - 1)));
            break;
        case 2:  // constructHalfSineGrid :166-186
            if (del_x / L >= 1.0 / ((double)Nx - 1.0)) return 1;
            if (w.bc[1] == 3) return 2;
            w.grid[Nx] = L / 2.0;
            for (int i = 2; i <= Nh; ++i)
                w.grid[i] = L * (((double)i - 1.0) / ((double)Nx - 1.0) / 2.0 - (1.0 / ((double)Nx - 1.0) / 2.0 - del_x / L)
                                 * std::sin(pi * ((double)i - 1) / ((double)Nx - 1)) / std::sin(pi / ((double)Nx - 1.0)));
            break;
        case 3: {  // constructExpHalfGrid :188-214
            double F = 100.0 * del_x;
            if (F / L > 0.5) return 1;
            if (w.bc[1] == 3) return 2;
            double k = 2.0 * std::log(L / F - 1.0) / (double)(float)Nh;
            double A = L / (std::exp(k * (double)Nh) - 1.0);
            for (int i = 2; i <= Nh; ++i) w.grid[i] = A * (std::exp(k * (double)(i - 1)) - 1.0);
            break; }
        case 5:  // constructInvSineGrid :146-164
            if (del_x / L >= 1.0 / ((double)Nx - 1.0)) return 1;
            for (int i = 2; i <= Nh; ++i) {
                double phase = (double)((float)(i - 1) / (float)Nh);   // real(i-1)/real(N_h): single precision
                w.grid[i] = L * (phase + (1.0 - (double)Nx * del_x / L) * std::sin(2.0 * pi * phase) / 2.0 / pi);
            }
            break;
        default: return 3;
        }
        // makeArraysFromGrid :251-265
        for (int i = 1; i <= Nh; ++i) w.dx_dl[i] = w.grid[i + 1] - w.grid[i];
        for (int i = 2; i <= Nh; ++i) w.nodeVol[i] = (w.dx_dl[i - 1] + w.dx_dl[i]) / 2.0;
        w.nodeVol[1] = 0.5 * w.dx_dl[1];
        w.nodeVol[Nx] = 0.5 * w.dx_dl[Nh];
        w.startX = w.grid[1]; w.endX = w.grid[Nx];
        w.L_domain = w.endX - w.startX;
    } else {
        const int NN = Nx + 1;   // number of field faces
        dvec gf(NN + 2, 0.0);
        gf[1] = 0.0; gf[NN] = L;
        switch (gridType) {
        case 0:  // CIC constructUniformGrid :236-252
            for (int i = 2; i <= NN - 1; ++i) gf[i] = (i - 1) * L / (NN - 1);
            break;
        case 1:  // CIC constructSineGrid :128-151
            if (del_x / L >= 1.0 / ((double)NN - 1.0)) return 1;
            for (int i = 2; i <= NN - 1; ++i)
                gf[i] = L * (((double)i - 1.0) / ((double)NN - 1.0) - (1.0 / ((double)NN - 1.0) - del_x / L)
                             * std::sin(2 * pi * (i - 1) / (NN - 1)) / std::sin(2 * pi / (NN - 1)));
            break;
        case 2:  // CIC constructHalfSineGrid :179-201
            if (del_x / L >= 1.0 / ((double)NN - 1.0)) return 1;
            gf[NN] = L / 2.0;
            for (int i = 2; i <= NN - 1; ++i)
                gf[i] = L * (((double)i - 1.0) / ((double)NN - 1.0) / 2.0 - (1.0 / ((double)NN - 1.0) / 2.0 - del_x / L)
                             * std::sin(pi * ((double)i - 1) / ((double)NN - 1)) / std::sin(pi / ((double)NN - 1.0)));
            break;
        case 5:  // CIC constructInvSineGrid :154-177
            if (del_x / L >= 1.0 / ((double)Nh - 1.0)) return 1;
            for (int i = 2; i <= Nx; ++i) {
                double phase = (double)((float)(i - 1) / (float)Nx);
                gf[i] = L * (phase + (1.0 - (double)Nh * del_x / L) * std::sin(2.0 * pi * phase) / 2.0 / pi);
            }
            break;
        default: return 3;   // CIC exp grid (type 3) reads unset memory in the reference (:226) -> unsupported
        }
        for (int i = 1; i <= NN - 1; ++i) w.dx_dl[i] = gf[i + 1] - gf[i];
        for (int i = 1; i <= Nx; ++i) w.grid[i] = 0.5 * (gf[i] + gf[i + 1]);
        for (int i = 1; i <= Nx - 1; ++i) w.centerDiff[i] = w.grid[i + 1] - w.grid[i];
        // constructGrid :123-125
        w.startX = w.grid[1] - 0.5 * w.dx_dl[1];
        w.endX = w.grid[Nx] + 0.5 * w.dx_dl[Nx];
        w.L_domain = w.endX - w.startX;
    }
    return 0;
}

// getLFromX: NGP mod_domain.f90:275-296 ; CIC mod_domain.f90:254-279
static double getLFromX(const Domain &w, double x) {
    int lo = 1, hi = w.Nx;
    if (w.scheme == 0) {
        while (lo != hi - 1) {
            int mid = (lo + hi) / 2;
            if (w.grid[mid] <= x) lo = mid; else hi = mid;
        }
        return lo + (x - w.grid[lo]) / w.dx_dl[lo];
    }
    while (lo != hi - 1) {
        int mid = (lo + hi) / 2;
        if (w.grid[mid] - 0.5 * w.dx_dl[mid] <= x) lo = mid; else hi = mid;
    }
    if (w.grid[lo] + 0.5 * w.dx_dl[lo] <= x) lo = hi;
    return lo + 0.5 + (x - w.grid[lo]) / w.dx_dl[lo];
}

// getXFromL: NGP :267-273 ; CIC :281-287
static double getXFromL(const Domain &w, double l) {
    int il = (int)l;
    if (w.scheme == 0) return w.grid[il] + w.dx_dl[il] * (l - il);
    return w.grid[il] + w.dx_dl[il] * (l - il - 0.5);
}

// smoothField: NGP mod_domain.f90:298-324 ; CIC mod_domain.f90:289-312
static void smoothField(const Domain &w, const double *raw, double *out) {
    const int Nh = w.Nh, Nx = w.Nx;
    if (w.scheme == 0) {
        for (int i = 1; i <= Nh; ++i) {
            int bv = w.bc[i + 1] - w.bc[i];
            switch (bv) {
            case 0: out[i] = 0.25 * (raw[i - 1] + 2.0 * raw[i] + raw[i + 1]); break;
            case -1: case -4: out[i] = 0.25 * (3.0 * raw[1] + raw[2]); break;
            case 1: case 4: out[i] = 0.25 * (3.0 * raw[Nh] + raw[Nh - 1]); break;
            case -2: out[i] = 0.25 * (raw[1] + raw[2]); break;
            case 2: out[i] = 0.25 * (raw[Nh] + raw[Nh - 1]); break;
            case -3: out[i] = 0.25 * (raw[Nh] + 2.0 * raw[1] + raw[2]); break;
            case 3: out[i] = 0.25 * (raw[Nh - 1] + 2.0 * raw[Nh] + raw[1]); break;
            default: break;
            }
        }
    } else {
        switch (w.bc[1]) {
        case 1: case 4: out[1] = 0.25 * (2.0 * raw[1] + 2.0 * raw[2]); break;
        case 2: out[1] = 0.0; break;
        case 3: out[1] = 0.25 * (raw[Nh - 1] + 2.0 * raw[1] + raw[2]); break;
        }
        switch (w.bc[Nh]) {
        case 1: case 4: out[Nh] = 0.25 * (2.0 * raw[Nh] + 2.0 * raw[Nh - 1]); break;
        case 2: out[Nh] = 0.0; break;
        case 3: out[Nh] = out[1]; break;
        }
        for (int i = 2; i <= Nx; ++i) out[i] = 0.25 * (raw[i - 1] + 2.0 * raw[i] + raw[i + 1]);
    }
}

// ---- mod_particle.f90:15-74  type Particle ----------------------------------
struct Species {
    double mass = 0, q = 0, w_p = 1, q_over_m = 0, q_times_wp = 0;
    int finalIdx = 0, T = 1, maxNodes = 0, Nx = 0;
    dvec ps;                    // phaseSpace(4, finalIdx, T)
    ivec N_p, delIdx, refIdx, numToCollide, refRecordIdx;
    int refRecLen = 0;
    dvec wallLoss, energyLoss, momentumLoss;   // (2,T)
    dvec workSpace, densities;  // (maxNodes,T), (Nx,T), 1-based first index
    double numFuncEvalAve = 0, numSubStepsAve = 0;
    double accumEnergyLoss[2] = {0, 0};
    int accumWallLoss[2] = {0, 0};
    long long substepsDeposit = 0;   // oracle-only instrumentation: do-while trips in the last depositJ
    long long substepsMove = 0, funcEvalsMove = 0;
    inline double &P(int k, int i, int t) { return ps[(size_t)(k - 1) + 4 * ((size_t)(i - 1) + (size_t)finalIdx * (size_t)(t - 1))]; }
    inline double &WS(int i, int t) { return workSpace[(size_t)i + (size_t)(maxNodes + 2) * (size_t)(t - 1)]; }
    inline double &DN(int i, int t) { return densities[(size_t)i + (size_t)(Nx + 2) * (size_t)(t - 1)]; }
    inline double &WL(int s, int t) { return wallLoss[(s - 1) + 2 * (t - 1)]; }
    inline double &EL(int s, int t) { return energyLoss[(s - 1) + 2 * (t - 1)]; }
    inline double &ML(int s, int t) { return momentumLoss[(s - 1) + 2 * (t - 1)]; }
    long long sumNp() const { long long s = 0; for (int t = 1; t <= T; ++t) s += N_p[t]; return s; }
    long long sumDel() const { long long s = 0; for (int t = 1; t <= T; ++t) s += delIdx[t]; return s; }
};

struct Solver {   // NGP mod_potentialSolver.f90:17-22 ; CIC :17-23
    dvec phi, phi_f, J, rho, EField, a_tri, b_tri, c_tri;
    double rho_const = 0, RF_rad_frequency = 0, RF_half_amplitude = 0;
    double RF_voltage = 0, boundPhi[3] = {0, 0, 0}, sourceTermVals[3] = {0, 0, 0};  // CIC only, 1-based
    int RF_indx = 0;
    bool RF_bool = false;
    // NGP_Periodic (E-field form): OpenMP/NGP_Periodic/src/mod_potentialSolver.f90:10-13
    dvec EField_f, EField_half;
    double aveJ = 0.0;
    bool eform = false, globJBool = false;
};

struct NonLin {   // mod_nonLinSolvers.f90:29-35
    int iterNumPicard = 0, iterNumAdaptiveSteps = 0, amountTimeSplits = 0;
    int maxIter = 100, solverType = 0, m_Anderson = 10;
    double Beta_k = 0.5, eps_r = 0, eps_a = 1e-8;
    dvec Residual_k, phi_k, fitMat;   // (Nx, m+1), (Nx, m+1), (Nx, m)
    double moverSeconds = 0, potSeconds = 0;
    long long sweeps = 0;         // number of depositJ calls in the last solvePotential
    long long substeps = 0;       // total do-while trips (depositJ sweeps + final moveParticles)
    dvec normHist;                // residual norms of the last solveDivAmpereAnderson
};

// ---- mod_targetParticle.f90:13-37  type targetParticle -------------------------
struct TargetParticle { double mass = 0, density = 0, temperature = 0, v_therm = 0; };

// ---- mod_NullCollision.f90:18-33  type nullCollision ----------------------------
struct NullCollision {
    int numberCollisions = 0, lengthArrays = 0;
    int primary = 0, target = 0;       // reactantsIndx(1), reactantsIndx(2), 1-based
    dvec energyArray;                  // (lengthArrays), 1-based
    dvec sigmaArray;                   // (lengthArrays, numberCollisions), 1-based via SA()
    dvec energyThreshold;              // 1-based
    ivec collisionType, numberProducts, productsIndx;   // productsIndx(3, numberCollisions) 1-based species numbers
    double sigmaVMax = 0, reducedMass = 0, minEnergy = 0, maxEnergy = 0, sumMass = 0, reducedMassIonization = 0;
    std::vector<long long> totalAmountCollisions;
    dvec totalIncidentEnergy, totalEnergyLoss;
    long long totalNumberCollidableParticles = 0;
    double P_null_last = 0;
    inline double &SA(int j, int i) { return sigmaArray[(size_t)j + (size_t)(lengthArrays + 1) * (size_t)(i - 1)]; }
    inline int &PI(int k, int i) { return productsIndx[(size_t)k + 4 * (size_t)(i - 1)]; }
};

struct Oracle {
    Domain world;
    Solver solver;
    std::vector<Species> plist;
    NonLin nl;
    dvec particleTimeStep;        // CIC (Nx, nSpecies)
    std::vector<TargetParticle> targets;   // targetParticleList
    std::vector<NullCollision> colls;       // nullCollisionList
    int err = 0;
};

// particle_constructor mod_particle.f90:46-74
static void species_init(Species &s, const Domain &w, double mass, double q, double w_p, int N_p, int finalIdx) {
    s.mass = mass; s.q = q; s.w_p = w_p; s.q_over_m = q / mass; s.q_times_wp = q * w_p;
    s.finalIdx = finalIdx; s.T = w.T; s.Nx = w.Nx;
    s.maxNodes = std::max(w.Nx, w.Nh);
    s.ps.assign((size_t)4 * finalIdx * w.T, 0.0);
    s.refRecLen = finalIdx / 10;
    s.refRecordIdx.assign((size_t)(s.refRecLen + 1) * (w.T + 1), 0);
    s.N_p.assign(w.T + 1, N_p); s.delIdx.assign(w.T + 1, 0); s.refIdx.assign(w.T + 1, 0);
    s.numToCollide.assign(w.T + 1, 0);
    s.wallLoss.assign(2 * w.T, 0.0); s.energyLoss.assign(2 * w.T, 0.0); s.momentumLoss.assign(2 * w.T, 0.0);
    s.workSpace.assign((size_t)(s.maxNodes + 2) * w.T, 0.0);
    s.densities.assign((size_t)(w.Nx + 2) * w.T, 0.0);
}

// initializeRandUniform mod_particle.f90:84-97
static void initializeRandUniform(Species &s, const Domain &w, int32_t *irand) {
#pragma omp parallel for schedule(static, 1)
    for (int iThread = 1; iThread <= w.T; ++iThread) {
        for (int i = 1; i <= s.N_p[iThread]; ++i) {
            double x_pos = ran2(irand[iThread]) * w.L_domain + w.startX;
            s.P(1, i, iThread) = getLFromX(w, x_pos);
        }
    }
}

// initializeRandCosine / initializeRandSine mod_particle.f90:99-151 (sine=true -> :126-151)
static void initializeRandTrig(Species &s, const Domain &w, int32_t *irand, double alpha, bool sine) {
#pragma omp parallel for schedule(static, 1)
    for (int iThread = 1; iThread <= w.T; ++iThread) {
        int32_t randNum = irand[iThread];
        for (int i = 1; i <= s.N_p[iThread]; ++i) {
            double R1 = ran2(randNum), R2 = ran2(randNum);
            double val = (1.0 + alpha * (sine ? std::sin(2 * pi * R2) : std::cos(2 * pi * R2))) / (1.0 + alpha);
            while (R1 > val) {
                R1 = ran2(randNum); R2 = ran2(randNum);
                val = (1.0 + alpha * (sine ? std::sin(2 * pi * R2) : std::cos(2 * pi * R2))) / (1.0 + alpha);
            }
            double x_pos = R2 * w.L_domain + w.startX;
            s.P(1, i, iThread) = getLFromX(w, x_pos);
        }
        irand[iThread] = randNum;
    }
}

// generate3DMaxwellian mod_particle.f90:156-191
static void generate3DMaxwellian(Species &s, const Domain &w, double T, int32_t *irand, double alpha, int distType, double v_drift) {
    double v_therm = std::sqrt(T * e_c / s.mass);
#pragma omp parallel for schedule(static, 1)
    for (int iThread = 1; iThread <= w.T; ++iThread) {
        int32_t ir = irand[iThread];
        for (int i = 1; i <= s.N_p[iThread]; ++i) {
            double U1 = ran2(ir), U2 = ran2(ir), U3 = ran2(ir), U4 = ran2(ir);
            double v_extra = 0.0, x_pos;
            switch (distType) {
            case 0: v_extra = v_drift; break;
            case 1: x_pos = getXFromL(w, s.P(1, i, iThread)) / w.L_domain; v_extra = v_drift * (1.0 + alpha * std::cos(2 * pi * x_pos)); break;
            case 2: x_pos = getXFromL(w, s.P(1, i, iThread)) / w.L_domain; v_extra = v_drift * (1.0 - alpha * std::sin(2 * pi * x_pos)); break;
            }
            s.P(2, i, iThread) = v_therm * std::sqrt(-2.0 * std::log(U1)) * std::cos(2.0 * pi * U2) + v_extra;
            s.P(3, i, iThread) = v_therm * std::sqrt(-2.0 * std::log(U1)) * std::sin(2.0 * pi * U2);
            s.P(4, i, iThread) = v_therm * std::sqrt(-2.0 * std::log(U3)) * std::sin(2.0 * pi * U4);
        }
        irand[iThread] = ir;
    }
}

// getTotalKE mod_particle.f90:218-230 (thread partial sums added in thread order)
static double getSumV2(Species &s) {
    dvec part(s.T + 1, 0.0);
#pragma omp parallel for schedule(static, 1)
    for (int t = 1; t <= s.T; ++t) {
        double r = 0.0;
        for (int i = 1; i <= s.N_p[t]; ++i)
            for (int k = 2; k <= 4; ++k) r += s.P(k, i, t) * s.P(k, i, t);
        part[t] = r;
    }
    double res = 0.0;
    for (int t = 1; t <= s.T; ++t) res += part[t];
    return res;
}
static double getTotalKE(Species &s) { return getSumV2(s) * s.mass * 0.5 * s.w_p; }
// getKEAve :193-204
static double getKEAve(Species &s) { return getSumV2(s) * s.mass * 0.5 / e_c / (double)s.sumNp(); }
// writeLocalTemperature :251-283 (the per-cell average it writes, not the file): EHist(j) = SUM(v^2) m / count / 3 / e over the particles with
// INT(l) == j, for INT(l) < NumberXNodes; per-thread partial sums, then summed over threads in thread order
static void localTemperature(Species &s, int Nx, int numCells, double *EHist /*0-based numCells*/) {
    dvec temp((size_t)(numCells + 1) * (s.T + 1), 0.0);
    std::vector<long long> counter((size_t)(numCells + 1) * (s.T + 1), 0);
    for (int t = 1; t <= s.T; ++t)
        for (int j = 1; j <= s.N_p[t]; ++j) {
            const int index = (int)s.P(1, j, t);
            if (index < Nx && index >= 1 && index <= numCells) {
                temp[(size_t)index + (size_t)(numCells + 1) * t] += s.P(2, j, t) * s.P(2, j, t) + s.P(3, j, t) * s.P(3, j, t) + s.P(4, j, t) * s.P(4, j, t);
                counter[(size_t)index + (size_t)(numCells + 1) * t] += 1;
            }
        }
    for (int j = 1; j <= numCells; ++j) {
        double sm = 0.0; long long cn = 0;
        for (int t = 1; t <= s.T; ++t) { sm += temp[(size_t)j + (size_t)(numCells + 1) * t]; cn += counter[(size_t)j + (size_t)(numCells + 1) * t]; }
        EHist[j - 1] = cn > 0 ? sm * s.mass / (double)cn / 3.0 / e_c : 0.0;
    }
}
// getTotalMomentum :206-216
static void getTotalMomentum(Species &s, double res[3]) {
    double tmp[3] = {0, 0, 0};
    for (int t = 1; t <= s.T; ++t) {
        double a[3] = {0, 0, 0};
        for (int i = 1; i <= s.N_p[t]; ++i)
            for (int k = 0; k < 3; ++k) a[k] += s.P(k + 2, i, t);
        for (int k = 0; k < 3; ++k) tmp[k] += a[k];
    }
    for (int k = 0; k < 3; ++k) res[k] = tmp[k] * s.w_p * s.mass;
}

// ---- potentialSolver constructor: NGP mod_potentialSolver.f90:50-86 ; CIC :49-90
static void construct_diagMatrix(Oracle &o);
static int solver_init(Oracle &o, double leftV, double rightV, double RF_frequency, double timeCurrent) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    s.J.assign(Nh + 2, 0.0); s.rho.assign(Nx + 2, 0.0); s.phi.assign(Nx + 2, 0.0); s.phi_f.assign(Nx + 2, 0.0);
    s.EField.assign(Nh + 2, 0.0);
    s.a_tri.assign(std::max(Nx, Nh) + 2, 0.0); s.b_tri.assign(Nx + 2, 0.0); s.c_tri.assign(std::max(Nx, Nh) + 2, 0.0);
    s.rho_const = 0.0; s.RF_half_amplitude = 0.0;
    s.RF_rad_frequency = 2.0 * pi * RF_frequency;
    if (w.scheme == 0) {
        construct_diagMatrix(o);
        if (w.bc[1] == 1) s.phi[1] = leftV;
        if (w.bc[1] == 4) { s.RF_half_amplitude = leftV; s.phi[1] = s.RF_half_amplitude * std::sin(timeCurrent * s.RF_rad_frequency); }
        if (w.bc[Nx] == 1) s.phi[Nx] = rightV;
        if (w.bc[Nx] == 4) {
            if (s.RF_half_amplitude != 0) return 1;   // two RF boundaries -> stop
            s.RF_half_amplitude = rightV;
            s.phi[Nx] = s.RF_half_amplitude * std::sin(timeCurrent * s.RF_rad_frequency);
        }
        if (w.bc[1] == 3) { s.phi[1] = leftV; s.phi[Nx] = leftV; }
        s.RF_bool = s.RF_rad_frequency > 0 && s.RF_half_amplitude > 0 && (w.bc[1] == 4 || w.bc[Nx] == 4);
        s.phi_f = s.phi;
    } else {
        s.boundPhi[1] = s.boundPhi[2] = 0.0; s.sourceTermVals[1] = s.sourceTermVals[2] = 0.0;
        s.RF_voltage = 0.0;   // (uninitialised in the reference when no RF wall; never read then)
        if (w.bc[1] == 1) { s.boundPhi[1] = leftV; s.sourceTermVals[1] = -s.boundPhi[1] * 2.0 / w.dx_dl[1]; }
        else if (w.bc[1] == 4) { s.RF_half_amplitude = leftV; s.RF_voltage = s.RF_half_amplitude * std::sin(timeCurrent * s.RF_rad_frequency); s.RF_indx = 1; }
        if (w.bc[Nh] == 1) { s.boundPhi[2] = rightV; s.sourceTermVals[2] = -s.boundPhi[2] * 2.0 / w.dx_dl[Nx]; }
        else if (w.bc[Nh] == 4) { s.RF_half_amplitude = rightV; s.RF_voltage = s.RF_half_amplitude * std::sin(timeCurrent * s.RF_rad_frequency); s.RF_indx = 2; }
        s.RF_bool = s.RF_rad_frequency > 0 && s.RF_half_amplitude > 0 && (w.bc[1] == 4 || w.bc[Nh] == 4);
        construct_diagMatrix(o);
    }
    return 0;
}

// construct_diagMatrix: NGP :88-128 ; CIC :92-126
static void construct_diagMatrix(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx;
    std::fill(s.a_tri.begin(), s.a_tri.end(), 0.0);
    std::fill(s.b_tri.begin(), s.b_tri.end(), 0.0);
    std::fill(s.c_tri.begin(), s.c_tri.end(), 0.0);
    if (w.scheme == 0) {
        for (int i = 1; i <= Nx; ++i) {
            switch (w.bc[i]) {
            case 0:
                if (i < Nx) s.c_tri[i] = 1.0 / w.dx_dl[i];
                if (i > 1) s.a_tri[i - 1] = 1.0 / w.dx_dl[i - 1];
                s.b_tri[i] = -(1.0 / w.dx_dl[i - 1] + 1.0 / w.dx_dl[i]);
                break;
            case 1: case 4: s.b_tri[i] = 1.0; break;
            case 2:
                if (i == 1) { s.c_tri[i] = 2.0 / w.dx_dl[i]; s.b_tri[i] = -(2.0 / w.dx_dl[i]); }
                else if (i == Nx) { s.a_tri[i - 1] = 2.0 / w.dx_dl[i - 1]; s.b_tri[i] = -(2.0 / w.dx_dl[i - 1]); }
                else o.err = 10;
                break;
            case 3: s.b_tri[i] = 1.0; break;
            default: break;
            }
        }
    } else {
        for (int i = 1; i <= Nx; ++i) {
            switch (w.bc[i + 1] - w.bc[i]) {
            case 0:
                s.b_tri[i] = -2.0 * (1.0 / (w.dx_dl[i - 1] + w.dx_dl[i]) + 1.0 / (w.dx_dl[i + 1] + w.dx_dl[i]));
                s.a_tri[i - 1] = 2.0 / (w.dx_dl[i - 1] + w.dx_dl[i]);
                s.c_tri[i] = 2.0 / (w.dx_dl[i] + w.dx_dl[i + 1]);
                break;
            case -1: case -4: case -3:
                s.b_tri[i] = -2.0 * (1.0 / (w.dx_dl[i + 1] + w.dx_dl[i]) + 1.0 / w.dx_dl[i]);
                s.c_tri[i] = 2.0 / (w.dx_dl[i + 1] + w.dx_dl[i]);
                s.sourceTermVals[1] = -s.boundPhi[1] * 2.0 / w.dx_dl[1];
                break;
            case 1: case 4: case 3:
                s.b_tri[i] = -2.0 * (1.0 / (w.dx_dl[i - 1] + w.dx_dl[i]) + 1.0 / w.dx_dl[i]);
                s.a_tri[i - 1] = 2.0 / (w.dx_dl[i - 1] + w.dx_dl[i]);
                s.sourceTermVals[2] = -s.boundPhi[2] * 2.0 / w.dx_dl[Nx];
                break;
            case -2:
                s.b_tri[i] = -2.0 / (w.dx_dl[i + 1] + w.dx_dl[i]);
                s.c_tri[i] = 2.0 / (w.dx_dl[i + 1] + w.dx_dl[i]);
                break;
            case 2:
                s.b_tri[i] = -2.0 / (w.dx_dl[i - 1] + w.dx_dl[i]);
                s.a_tri[i - 1] = 2.0 / (w.dx_dl[i - 1] + w.dx_dl[i]);
                break;
            }
        }
    }
}

// interpolateParticleToNodes: NGP mod_Scheme.f90:50-65 ; CIC mod_Scheme.f90:21-66
static void interpolateParticleToNodes(Species &p, const Domain &w, int t) {
    if (w.scheme == 0) {
        for (int j = 1; j <= p.N_p[t]; ++j) {
            double l = p.P(1, j, t);
            int l_left = (int)l, l_right = l_left + 1;
            double d = l - (double)l_left;
            p.DN(l_left, t) += (1.0 - d);
            p.DN(l_right, t) += d;
        }
    } else {
        const int Nx = w.Nx;
        for (int j = 1; j <= p.N_p[t]; ++j) {
            double l = p.P(1, j, t);
            int lc = (int)l; double d = l - (double)lc;
            p.DN(lc, t) += (-d * d + d + 0.5);
            int lr = lc + 1, ll = lc - 1;
            switch (w.bc[lr] - w.bc[lc]) {
            case 0: p.DN(ll, t) += 0.5 * (1.0 - d) * (1.0 - d); p.DN(lr, t) += 0.5 * d * d; break;
            case -1: case -4: p.DN(lc, t) -= 0.5 * (1.0 - d) * (1.0 - d); p.DN(lr, t) += 0.5 * d * d; break;
            case 1: case 4: p.DN(lc, t) -= 0.5 * d * d; p.DN(ll, t) += 0.5 * (1.0 - d) * (1.0 - d); break;
            case -2: p.DN(lc, t) += 0.5 * (1.0 - d) * (1.0 - d); p.DN(lr, t) += 0.5 * d * d; break;
            case 2: p.DN(lc, t) += 0.5 * d * d; p.DN(ll, t) += 0.5 * (1.0 - d) * (1.0 - d); break;
            case -3: p.DN(Nx, t) += 0.5 * (1.0 - d) * (1.0 - d); p.DN(lr, t) += 0.5 * d * d; break;
            case 3: p.DN(ll, t) += 0.5 * (1.0 - d) * (1.0 - d); p.DN(1, t) += 0.5 * d * d; break;
            }
        }
    }
}

// cross-thread sliced reduction of a per-thread grid copy, species-ordered scaling
// (NGP mod_particleMover.f90:145-149, mod_potentialSolver.f90:145-149; CIC twins)
template <class GET>
static void reduce_slices(Oracle &o, int nSlices, const ivec &L, const ivec &R, GET get) {
    const int T = o.world.T, nS = (int)o.plist.size();
#pragma omp parallel for schedule(static, 1)
    for (int it = 1; it <= T; ++it) {
        if (it <= nSlices) {
            for (int i = L[it]; i <= R[it]; ++i) {
                double s = 0.0;
                for (int t = 1; t <= T; ++t) s += get(o.plist[0], i, t);
                double acc = s * o.plist[0].q_times_wp;
                for (int j = 1; j < nS; ++j) {
                    double sj = 0.0;
                    for (int t = 1; t <= T; ++t) sj += get(o.plist[j], i, t);
                    acc = acc + sj * o.plist[j].q_times_wp;
                }
                get(o.plist[0], i, 1) = acc;
            }
        }
    }
}

// depositRho: NGP mod_potentialSolver.f90:130-188 ; CIC :128-173
static void depositRho(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh, T = w.T;
#pragma omp parallel for schedule(static, 1)
    for (int it = 1; it <= T; ++it) {
        for (auto &p : o.plist) {
            for (int i = 0; i <= Nx + 1; ++i) p.DN(i, it) = 0.0;
            interpolateParticleToNodes(p, w, it);
        }
    }
    reduce_slices(o, w.nTN, w.tnL, w.tnR, [](Species &p, int i, int t) -> double & { return p.DN(i, t); });
    Species &p1 = o.plist[0];
    auto D = [&](int i) -> double & { return p1.DN(i, 1); };
    if (w.scheme == 0) {
        switch (w.bc[1]) {
        case 1: case 4: D(1) = 0.0; break;
        case 2: D(1) = 2.0 * D(1); break;
        case 3: D(1) = D(1) + D(Nx); break;
        }
        switch (w.bc[Nx]) {
        case 1: case 4: D(Nx) = 0.0; break;
        case 2: D(Nx) = 2.0 * D(Nx); break;
        case 3: D(Nx) = D(1); break;
        }
        if (w.smooth) {
            for (int i = 1; i <= Nx; ++i) {
                switch (w.bc[i]) {
                case 0: s.rho[i] = 0.25 * (D(i - 1) + 2.0 * D(i) + D(i + 1)); break;
                case 1: case 4: s.rho[i] = 0.0; break;
                case 2:
                    if (i == 1) s.rho[i] = 0.25 * (2.0 * D(1) + 2.0 * D(2));
                    else s.rho[i] = 0.25 * (2.0 * D(Nx) + 2.0 * D(Nh));
                    break;
                case 3: s.rho[i] = 0.25 * (D(Nh) + 2.0 * D(1) + D(2)); break;
                }
            }
        } else {
            for (int i = 1; i <= Nx; ++i) s.rho[i] = D(i);
        }
    } else {
        if (w.smooth) {
            for (int i = 1; i <= Nx; ++i) {
                switch (w.bc[i + 1] - w.bc[i]) {
                case 0: s.rho[i] = 0.25 * (D(i - 1) + 2.0 * D(i) + D(i + 1)); break;
                case -1: case -4: s.rho[i] = 0.25 * (D(1) + D(2)); break;
                case 1: case 4: s.rho[i] = 0.25 * (D(Nx) + D(Nx - 1)); break;
                case -2: s.rho[i] = 0.25 * (3.0 * D(1) + D(2)); break;
                case 2: s.rho[i] = 0.25 * (3.0 * D(Nx) + D(Nx - 1)); break;
                case -3: s.rho[i] = 0.25 * (D(Nx) + 2.0 * D(1) + D(2)); break;
                case 3: s.rho[i] = 0.25 * (D(1) + 2.0 * D(Nx) + D(Nx - 1)); break;
                }
            }
        } else {
            for (int i = 1; i <= Nx; ++i) s.rho[i] = D(i);
        }
    }
}

// Thomas sweep shared by Poisson/Ampere (NGP :203-221, :308-328 ; CIC :192-210, :282-299)
static void thomas(const Solver &s, int Nx, const dvec &d, dvec &x) {
    dvec cp(Nx + 2), dp(Nx + 2);
    cp[1] = s.c_tri[1] / s.b_tri[1];
    dp[1] = d[1] / s.b_tri[1];
    for (int i = 2; i <= Nx - 1; ++i) {
        double m = s.b_tri[i] - cp[i - 1] * s.a_tri[i - 1];
        cp[i] = s.c_tri[i] / m;
        dp[i] = (d[i] - dp[i - 1] * s.a_tri[i - 1]) / m;
    }
    double m = s.b_tri[Nx] - cp[Nx - 1] * s.a_tri[Nx - 1];
    dp[Nx] = (d[Nx] - dp[Nx - 1] * s.a_tri[Nx - 1]) / m;
    x[Nx] = dp[Nx];
    for (int i = Nx - 1; i >= 1; --i) x[i] = dp[i] - cp[i] * x[i + 1];
}

// solve_tridiag_Poisson: NGP :190-226 ; CIC :175-218
static void solve_tridiag_Poisson(Oracle &o, double timeCurrent) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx;
    dvec d(Nx + 2, 0.0);
    if (w.scheme == 0) {
        for (int i = 1; i <= Nx; ++i) {
            switch (w.bc[i]) {
            case 0: case 2: d[i] = (-s.rho[i] - s.rho_const) / eps_0; break;
            case 1: case 3: d[i] = s.phi[i]; break;
            case 4: d[i] = s.RF_half_amplitude * std::sin(s.RF_rad_frequency * timeCurrent); break;
            }
        }
    } else {
        if (s.RF_bool) {
            if (s.RF_indx == 1) { s.boundPhi[1] = s.RF_half_amplitude * std::sin(s.RF_rad_frequency * timeCurrent); s.sourceTermVals[1] = -s.boundPhi[1] * 2.0 / w.dx_dl[1]; }
            else { s.boundPhi[2] = s.RF_half_amplitude * std::sin(s.RF_rad_frequency * timeCurrent); s.sourceTermVals[2] = -s.boundPhi[2] * 2.0 / w.dx_dl[Nx]; }
        }
        for (int i = 1; i <= Nx; ++i) d[i] = (-s.rho[i] - s.rho_const) / eps_0;
        d[1] = d[1] + s.sourceTermVals[1];
        d[Nx] = d[Nx] + s.sourceTermVals[2];
    }
    thomas(s, Nx, d, s.phi);
    s.phi_f = s.phi;
}

// triMul mod_BasicFunctions.f90:205-216
static void triMul(int n, const dvec &lo, const dvec &up, const dvec &di, const dvec &x, dvec &res) {
    res[1] = x[1] * di[1] + x[2] * up[1];
    for (int i = 2; i <= n - 1; ++i) res[i] = x[i] * di[i] + x[i - 1] * lo[i - 1] + x[i + 1] * up[i];
    res[n] = x[n] * di[n] + x[n - 1] * lo[n - 1];
}

// getError_tridiag_Poisson: NGP :240-265 ; CIC :232-244
static double getError_tridiag_Poisson(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx;
    dvec Ax(Nx + 2, 0.0);
    triMul(Nx, s.a_tri, s.c_tri, s.b_tri, s.phi_f, Ax);
    double res = 0.0;
    if (w.scheme == 0) {
        for (int i = 1; i <= Nx; ++i) {
            switch (w.bc[i]) {
            case 0: case 2:
                if (s.rho[i] + s.rho_const != 0) { double v = Ax[i] * eps_0 / (-s.rho[i] - s.rho_const) - 1.0; res += v * v; }
                else { double v = Ax[i] * eps_0; res += v * v; }
                break;
            case 1: case 3: case 4:
                if (s.phi_f[i] != 0) { double v = Ax[i] / s.phi_f[i] - 1.0; res += v * v; }
                else res += Ax[i] * Ax[i];
                break;
            }
        }
        return std::sqrt(res / Nx);
    }
    dvec d(Nx + 2);
    for (int i = 1; i <= Nx; ++i) d[i] = (-s.rho[i] - s.rho_const) / eps_0;
    d[1] += s.sourceTermVals[1]; d[Nx] += s.sourceTermVals[2];
    for (int i = 1; i <= Nx; ++i) { double v = 1.0 - Ax[i] / d[i]; res += v * v; }
    return std::sqrt(res / Nx);
}

// makeConstSourceTerm: NGP :267-286 ; CIC :246-270
static void makeConstSourceTerm(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx;
    if (w.scheme == 0) {
        for (int i = 1; i <= Nx; ++i) {
            switch (w.bc[i]) {
            case 0: s.rho[i] = -(s.phi[i] - s.phi[i - 1]) / w.dx_dl[i - 1] + (s.phi[i + 1] - s.phi[i]) / w.dx_dl[i]; break;
            case 1: case 3: case 4: s.rho[i] = s.phi_f[i]; break;
            case 2:
                if (i == 1) s.rho[i] = -2.0 * (s.phi[i] - s.phi[i + 1]) / w.dx_dl[i];
                else if (i == Nx) s.rho[i] = -2.0 * (s.phi[i] - s.phi[i - 1]) / w.dx_dl[i - 1];
                break;
            }
        }
    } else {
        for (int i = 1; i <= Nx; ++i) {
            switch (w.bc[i + 1] - w.bc[i]) {
            case 0: s.rho[i] = -(s.phi[i] - s.phi[i + 1]) / (w.grid[i + 1] - w.grid[i]) + (s.phi[i - 1] - s.phi[i]) / (w.grid[i] - w.grid[i - 1]); break;
            case -1: case -4: s.rho[i] = -(s.phi[i] - s.phi[i + 1]) / (w.grid[i + 1] - w.grid[i]) + 2.0 * (s.boundPhi[1] - s.phi[1]) / w.dx_dl[1] + s.sourceTermVals[1]; break;
            case 1: case 4: s.rho[i] = -2.0 * (s.phi[Nx] - s.boundPhi[2]) / w.dx_dl[Nx] + (s.phi[i - 1] - s.phi[i]) / (w.grid[i] - w.grid[i - 1]) + s.sourceTermVals[2]; break;
            case -2: s.rho[i] = -(s.phi[i] - s.phi[i + 1]) / (w.grid[i + 1] - w.grid[i]); break;
            case 2: s.rho[i] = (s.phi[i - 1] - s.phi[i]) / (w.grid[i] - w.grid[i - 1]); break;
            case -3: s.rho[i] = -(s.phi[i] - s.phi[i + 1]) / (w.grid[i + 1] - w.grid[i]) + 2.0 * (s.phi[Nx] - s.phi[1]) / (w.dx_dl[1] + w.dx_dl[Nx]); break;
            case 3: s.rho[i] = -2.0 * (s.phi[Nx] - s.phi[1]) / (w.dx_dl[1] + w.dx_dl[Nx]) + (s.phi[i - 1] - s.phi[i]) / (w.grid[i] - w.grid[i - 1]); break;
            }
        }
    }
}

// solve_tridiag_Ampere: NGP :288-330 ; CIC :272-301
static void solve_tridiag_Ampere(Oracle &o, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx;
    dvec d(Nx + 2, 0.0);
    if (w.scheme == 0) {
        for (int i = 1; i <= Nx; ++i) {
            switch (w.bc[i]) {
            case 0: d[i] = (s.J[i] - s.J[i - 1]) * del_t / eps_0; break;
            case 1: case 3: case 4: d[i] = 0.0; break;
            case 2:
                if (i == 1) d[i] = 2.0 * (del_t * s.J[i] / eps_0);
                else if (i == Nx) d[i] = 2.0 * (-del_t * s.J[i - 1] / eps_0);
                break;
            }
            d[i] = d[i] + s.rho[i];
        }
    } else {
        for (int i = 1; i <= Nx; ++i) d[i] = (s.J[i + 1] - s.J[i]) * del_t / eps_0 + s.rho[i];
    }
    thomas(s, Nx, d, s.phi_f);
}

// getChargeContinuityError: NGP :358-390 ; CIC :315-331
static double getChargeContinuityError(Oracle &o, const double *rho_i /*1-based*/, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    double ce = 0.0; int k = 0;
    for (int i = 1; i <= Nx; ++i) {
        double dR = s.rho[i] - rho_i[i];
        if (dR != 0) {
            if (w.scheme == 0) {
                switch (w.bc[i]) {
                case 0: { double v = 1.0 + del_t * (s.J[i] - s.J[i - 1]) / dR; ce += v * v; k++; break; }
                case 1: case 4: break;
                case 2:
                    if (i == 1) { double v = 1.0 + 2.0 * del_t * s.J[1] / dR; ce += v * v; }
                    else { double v = 1.0 - 2.0 * del_t * s.J[Nh] / dR; ce += v * v; }
                    k++; break;
                case 3:
                    if (i == 1) { double v = 1.0 + del_t * (s.J[1] - s.J[Nh]) / dR; ce += v * v; }
                    break;
                }
            } else {
                double v = 1.0 + del_t * (s.J[i + 1] - s.J[i]) / dR; ce += v * v; k++;
            }
        }
    }
    return std::sqrt(ce / k);
}

// makeHalfTimeEField: NGP :392-406 ; CIC :333-366
static void makeHalfTimeEField(Oracle &o, double *work /*1-based, Nh*/) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    if (w.scheme == 0) {
        for (int i = 1; i <= Nh; ++i) work[i] = 0.5 * (s.phi_f[i] + s.phi[i] - s.phi_f[i + 1] - s.phi[i + 1]);
        if (w.smooth) {
            smoothField(w, work, s.EField.data());
            for (int i = 1; i <= Nh; ++i) s.EField[i] = s.EField[i] / w.dx_dl[i];
        } else {
            for (int i = 1; i <= Nh; ++i) s.EField[i] = work[i] / w.dx_dl[i];
        }
    } else {
        switch (w.bc[1]) {
        case 1: work[1] = (2.0 * s.boundPhi[1] - s.phi[1] - s.phi_f[1]); break;
        case 2: work[1] = 0.0; break;
        case 3: work[1] = 0.5 * (s.phi[Nx] + s.phi_f[Nx] - s.phi[1] - s.phi_f[1]); break;
        case 4: work[1] = (s.boundPhi[1] + s.RF_voltage - s.phi[1] - s.phi_f[1]); break;
        }
        switch (w.bc[Nh]) {
        case 1: work[Nh] = s.phi[Nx] + s.phi_f[Nx] - 2.0 * s.boundPhi[2]; break;
        case 2: work[Nh] = 0.0; break;
        case 3: work[Nh] = work[1]; break;
        case 4: work[Nh] = s.phi[Nx] + s.phi_f[Nx] - s.boundPhi[2] - s.RF_voltage; break;
        }
        for (int i = 2; i <= Nx; ++i) work[i] = 0.5 * (s.phi_f[i - 1] + s.phi[i - 1] - s.phi_f[i] - s.phi[i]);
        if (w.smooth) smoothField(w, work, s.EField.data());
        else for (int i = 1; i <= Nh; ++i) s.EField[i] = work[i];
    }
}

// getTotalPE: NGP :408-420 ; CIC :369-413
static double getTotalPE(Oracle &o, bool future) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    const dvec &ph = future ? s.phi_f : s.phi;
    double sum = 0.0;
    if (w.scheme == 0) {
        for (int i = 1; i <= Nh; ++i) { double dphi = ph[i + 1] - ph[i]; sum += dphi * dphi / w.dx_dl[i]; }
        return 0.5f * eps_0 * sum;   // literal 0.5 is default-real in the Fortran; exactly representable
    }
    for (int i = 1; i <= Nx - 1; ++i) { double dphi = ph[i + 1] - ph[i]; sum += dphi * dphi / (w.grid[i + 1] - w.grid[i]); }
    double res = 0.5f * eps_0 * sum;
    if (future) {
        switch (w.bc[1]) {
        case 1: res += eps_0 * (s.boundPhi[1] - ph[1]) * (s.boundPhi[1] - ph[1]) / w.dx_dl[1]; break;
        case 3: res += eps_0 * (ph[Nx] - ph[1]) * (ph[Nx] - ph[1]) / (w.dx_dl[1] + w.dx_dl[Nx]); break;
        case 4: res += eps_0 * (s.RF_voltage - ph[1]) * (s.RF_voltage - ph[1]) / w.dx_dl[1]; break;
        }
        switch (w.bc[Nh]) {
        case 1: res += eps_0 * (s.boundPhi[2] - ph[Nx]) * (s.boundPhi[2] - ph[Nx]) / w.dx_dl[Nx]; break;
        case 4: res += eps_0 * (s.RF_voltage - ph[Nx]) * (s.RF_voltage - ph[Nx]) / w.dx_dl[Nx]; break;
        }
    } else {
        switch (w.bc[1]) {
        case 1: case 4: res += eps_0 * (s.boundPhi[1] - ph[1]) * (s.boundPhi[1] - ph[1]) / w.dx_dl[1]; break;
        case 3: res += eps_0 * (ph[Nx] - ph[1]) * (ph[Nx] - ph[1]) / (w.dx_dl[1] + w.dx_dl[Nx]); break;
        }
        switch (w.bc[Nh]) {
        case 1: case 4: res += eps_0 * (s.boundPhi[2] - ph[Nx]) * (s.boundPhi[2] - ph[Nx]) / w.dx_dl[Nx]; break;
        }
    }
    return res;
}

// getJdotE: NGP :422-430 ; CIC :415-426
static double getJdotE(Oracle &o, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    double res = 0.0;
    if (w.scheme == 0) {
        for (int i = 1; i <= w.Nh; ++i) res += s.EField[i] * w.dx_dl[i] * s.J[i];
        return res * del_t;
    }
    for (int i = 2; i <= w.Nx; ++i) res += s.J[i] * s.EField[i];
    res = res + 0.5 * s.J[1] * s.EField[1] + s.J[w.Nh] * s.EField[w.Nh] * 0.5;
    return res * del_t;
}

// getEnergyFromBoundary: NGP :432-442 ; CIC :428-456
static double getEnergyFromBoundary(Oracle &o, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    if (w.scheme == 0) {
        double res = del_t * s.J[1];
        res = res + eps_0 * ((s.phi_f[1] - s.phi_f[2]) - (s.phi[1] - s.phi[2])) / w.dx_dl[1];
        res = res * (s.phi_f[1] + s.phi[1] - s.phi_f[Nx] - s.phi[Nx]) * 0.5;
        return res;
    }
    double lA = 0, rA = 0;
    switch (w.bc[1]) { case 1: lA = s.boundPhi[1]; break; case 2: case 3: lA = 0; break; case 4: lA = 0.5 * (s.RF_voltage + s.boundPhi[1]); break; }
    switch (w.bc[Nh]) { case 1: rA = s.boundPhi[2]; break; case 2: case 3: rA = 0; break; case 4: rA = 0.5 * (s.RF_voltage + s.boundPhi[2]); break; }
    double res = del_t * s.J[2] + eps_0 * ((s.phi_f[1] - s.phi_f[2]) - (s.phi[1] - s.phi[2])) / (w.grid[2] - w.grid[1]);
    return res * (lA - rA);
}

// setRFVoltage: NGP :444-453 ; CIC :458-470
static void setRFVoltage(Oracle &o, double timeFuture) {
    Domain &w = o.world; Solver &s = o.solver;
    if (w.scheme == 0) {
        if (w.bc[1] == 4) s.phi_f[1] = s.RF_half_amplitude * std::sin(s.RF_rad_frequency * timeFuture);
        else s.phi_f[w.Nx] = s.RF_half_amplitude * std::sin(s.RF_rad_frequency * timeFuture);
    } else {
        s.RF_voltage = s.RF_half_amplitude * std::sin(s.RF_rad_frequency * timeFuture);
        if (s.RF_indx == 1) s.sourceTermVals[1] = -s.RF_voltage * 2.0 / w.dx_dl[1];
        else s.sourceTermVals[2] = -s.RF_voltage * 2.0 / w.dx_dl[w.Nx];
    }
}

// resetVoltage: NGP :455-459 ; CIC :472-478
static void resetVoltage(Oracle &o) {
    Solver &s = o.solver;
    s.phi = s.phi_f;
    if (o.world.scheme == 1 && s.RF_bool) s.boundPhi[s.RF_indx] = s.RF_voltage;
}

// ---- NGP particleSubStepNoBField  OpenMP/NGP/src/mod_particleMover.f90:21-63 --
static inline void particleSubStepNoBField(double l_sub, double v_sub, double &l_f, double &v_f, double &del_tau, double E_x,
                                           double q_over_m, double dx_dl, int l_cell, bool &FutureAtBoundaryBool, int &l_boundary) {
    double a = q_over_m * E_x;
    v_f = v_sub + a * del_tau;
    l_f = l_sub + 0.5 * (v_sub + v_f) * del_tau / dx_dl;
    bool inCellBool = ((int)l_f == l_cell);
    bool equalVSignBool = ((v_sub > 0) == (v_f > 0));
    FutureAtBoundaryBool = (!inCellBool) || (!equalVSignBool);
    if (FutureAtBoundaryBool) {
        if (v_sub > 0) l_boundary = l_cell + 1; else l_boundary = l_cell;
        double diff_PE = 2.0 * a * ((double)l_boundary - l_sub) * dx_dl;
        double v_i_sqr = v_sub * v_sub;
        FutureAtBoundaryBool = (diff_PE > -v_i_sqr);
        if (FutureAtBoundaryBool) {
            l_f = (double)l_boundary;
            v_f = fsign(1.0, v_sub) * std::sqrt(diff_PE + v_i_sqr);
            del_tau = (v_f - v_sub) / a;
        } else if (!inCellBool) {
            FutureAtBoundaryBool = true;
            l_boundary = 2 * l_cell + 1 - l_boundary;
            l_f = (double)l_boundary;
            diff_PE = 2.0 * a * (l_f - l_sub) * dx_dl;
            v_f = -fsign(1.0, v_sub) * std::sqrt(diff_PE + v_i_sqr);
            del_tau = (v_f - v_sub) / a;
        }
    }
}

// ---- CIC particleSubStepPicard  OpenMP/CIC/src/mod_particleMoverPicard.f90:17-65
static inline void particleSubStepPicard(double l_sub, double v_sub, double &l_f, double &v_f, double &d_half, double &del_tau,
                                         double E_left, double E_right, double dx_dl, double q_over_m, int l_cell,
                                         bool &FutureAtBoundaryBool, int &l_boundary, int &numIter) {
    const int maxPartIter = 50;
    double accel = 0, v_half = 0, l_f_prev = l_sub;
    int i;
    for (i = 1; i <= maxPartIter; ++i) {
        d_half = (l_sub + l_f_prev) * 0.5;
        accel = q_over_m * (E_left * (1.0 - d_half) + E_right * d_half);
        v_half = v_sub + 0.5 * accel * del_tau / dx_dl;
        l_f = l_sub + v_half * del_tau / dx_dl;
        FutureAtBoundaryBool = ((l_f <= 0) || (l_f >= 1));
        if (FutureAtBoundaryBool) { if (v_half < 0) l_f = 0.0; else l_f = 1.0; }
        if (std::fabs(l_f - l_f_prev) < 1.e-10) break;
        l_f_prev = l_f;
    }
    numIter = i;   // Fortran: loop variable is maxPartIter+1 after an exhausted loop
    if (!FutureAtBoundaryBool) {
        v_f = 2.0 * v_half - v_sub;
    } else {
        if (l_f != l_sub) {
            v_f = fsign(1.0, v_half) * std::sqrt(2.0 * accel * (l_f - l_sub) + v_sub * v_sub);
            del_tau = (l_f - l_sub) * dx_dl / v_half;
            numIter = numIter + 1;
        } else {
            del_tau = -2.0 * v_sub * dx_dl / accel;
            v_f = -v_sub;
        }
        l_boundary = (int)l_f + l_cell;
    }
    l_f = l_f + (double)l_cell;
}

// CIC per-cell sub-step cap table  mod_particleMoverPicard.f90:92-94, 212-214
static void makeParticleTimeStep(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, nS = (int)o.plist.size();
    o.particleTimeStep.resize((size_t)(Nx + 2) * nS);
    for (int j = 0; j < nS; ++j)
        for (int i = 1; i <= Nx; ++i)
            o.particleTimeStep[(size_t)i + (size_t)(Nx + 2) * j] =
                0.1 * w.dx_dl[i] / std::sqrt(std::fabs(o.plist[j].q_over_m * (s.EField[i] - s.EField[i + 1])));
}

// depositJ: NGP mod_particleMover.f90:67-157 ; CIC mod_particleMoverPicard.f90:77-195
static void depositJ(Oracle &o, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh, T = w.T, nS = (int)o.plist.size();
    const bool cic = (w.scheme == 1);
    const double f_tol = del_t * (cic ? 1.e-10 : 1.e-15);
    const int NB = cic ? Nh : Nx;   // wall index on the right
    makeHalfTimeEField(o, &o.plist[0].WS(0, 1));
    if (cic) makeParticleTimeStep(o);
    for (auto &p : o.plist) p.substepsDeposit = 0;
#pragma omp parallel for schedule(static, 1)
    for (int iThread = 1; iThread <= T; ++iThread) {
        for (int j = 0; j < nS; ++j) {
            Species &p = o.plist[j];
            long long trips = 0;
            for (int i = 1; i <= Nh; ++i) p.WS(i, iThread) = 0.0;
            const double *pts = cic ? &o.particleTimeStep[(size_t)(Nx + 2) * j] : nullptr;
            for (int i = 1; i <= p.N_p[iThread]; ++i) {
                double v_sub = p.P(2, i, iThread), l_sub = p.P(1, i, iThread);
                double timePassed = 0.0, l_f = 0, v_f = 0, d_half = 0;
                int l_cell = (int)(l_sub + fsign(1.e-12, v_sub));
                double del_tau = del_t;
                bool atB = false; int l_boundary = 0, numIter = 0;
                while (del_tau > f_tol) {
                    ++trips;
                    double dx = w.dx_dl[l_cell];
                    if (!cic) {
                        double E_x = s.EField[l_cell];
                        particleSubStepNoBField(l_sub, v_sub, l_f, v_f, del_tau, E_x, p.q_over_m, dx, l_cell, atB, l_boundary);
                        p.WS(l_cell, iThread) = p.WS(l_cell, iThread) + l_f - l_sub;
                    } else {
                        del_tau = std::min(del_tau, pts[l_cell]);
                        particleSubStepPicard(l_sub - (double)l_cell, v_sub, l_f, v_f, d_half, del_tau, s.EField[l_cell], s.EField[l_cell + 1],
                                              dx, p.q_over_m, l_cell, atB, l_boundary, numIter);
                        double J_part = l_f - l_sub;
                        p.WS(l_cell, iThread) = p.WS(l_cell, iThread) + (1.0 - d_half) * J_part;
                        p.WS(l_cell + 1, iThread) = p.WS(l_cell + 1, iThread) + J_part * d_half;
                    }
                    if (atB) {
                        int b = w.bc[l_boundary];
                        if (b == 0) l_cell = l_cell + (int)fsign(1.0, v_f);
                        else if (b == 1 || b == 4) break;
                        else if (b == 2) {
                            if (l_boundary == NB) { if (v_f > 0) v_f = -v_f; }
                            else { if (v_f < 0) v_f = -v_f; }
                        } else if (b == 3) {
                            l_f = (double)std::abs(l_boundary - NB - 1);
                            l_cell = std::abs(l_cell - NB);
                        }
                    }
                    timePassed = timePassed + del_tau;
                    del_tau = del_t - timePassed;
                    l_sub = l_f; v_sub = v_f;
                }
            }
#pragma omp atomic
            p.substepsDeposit += trips;
        }
    }
    reduce_slices(o, w.nTH, w.thL, w.thR, [](Species &p, int i, int t) -> double & { return p.WS(i, t); });
    Species &p1 = o.plist[0];
    if (cic) {   // wall fix-ups :173-188
        switch (w.bc[1]) {
        case 1: case 4: p1.WS(1, 1) = 2.0 * p1.WS(1, 1); break;
        case 2: p1.WS(1, 1) = 0.0; break;
        case 3: p1.WS(1, 1) = p1.WS(1, 1) + p1.WS(Nh, 1); break;
        }
        switch (w.bc[Nh]) {
        case 1: case 4: p1.WS(Nh, 1) = 2.0 * p1.WS(Nh, 1); break;
        case 2: p1.WS(Nh, 1) = 0.0; break;
        case 3: p1.WS(Nh, 1) = p1.WS(1, 1); break;
        }
    }
    if (w.smooth) {
        smoothField(w, &p1.WS(0, 1), s.J.data());
        for (int i = 1; i <= Nh; ++i) s.J[i] = s.J[i] / del_t;
    } else {
        for (int i = 1; i <= Nh; ++i) s.J[i] = p1.WS(i, 1) / del_t;
    }
    o.nl.sweeps += 1;
    for (auto &p : o.plist) o.nl.substeps += p.substepsDeposit;
}

// moveParticles: NGP mod_particleMover.f90:162-270 ; CIC mod_particleMoverPicard.f90:198-317
static void moveParticles(Oracle &o, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh, T = w.T, nS = (int)o.plist.size();
    const bool cic = (w.scheme == 1);
    const double f_tol = del_t * (cic ? 1.e-10 : 1.e-15);
    const int NB = cic ? Nh : Nx;
    std::vector<long long> numSubStepAve(nS, 0), funcEvalCounter(nS, 0);
    makeHalfTimeEField(o, &o.plist[0].WS(0, 1));
    if (cic) makeParticleTimeStep(o);
#pragma omp parallel for schedule(static, 1)
    for (int iThread = 1; iThread <= T; ++iThread) {
        for (int j = 0; j < nS; ++j) {
            Species &p = o.plist[j];
            long long trips = 0, fevals = 0;
            int delIdx = 0, refIdx = 0;
            p.WL(1, iThread) = 0; p.WL(2, iThread) = 0;
            p.EL(1, iThread) = 0.0; p.EL(2, iThread) = 0.0;
            const int N_p = p.N_p[iThread];
            const double *pts = cic ? &o.particleTimeStep[(size_t)(Nx + 2) * j] : nullptr;
            for (int i = 1; i <= N_p; ++i) {
                double v_sub = p.P(2, i, iThread), l_sub = p.P(1, i, iThread);
                double timePassed = 0.0, l_f = 0, v_f = 0, d_half = 0;
                bool refluxedBool = false;
                int l_cell = (int)(l_sub + fsign(1.e-12, v_sub));
                double del_tau = del_t;
                bool convergeTimeBool = true, atB = false;
                int l_boundary = 0, numIter = 0;
                while (convergeTimeBool) {
                    ++trips;
                    double dx = w.dx_dl[l_cell];
                    if (!cic) {
                        double E_x = s.EField[l_cell];
                        particleSubStepNoBField(l_sub, v_sub, l_f, v_f, del_tau, E_x, p.q_over_m, dx, l_cell, atB, l_boundary);
                    } else {
                        del_tau = std::min(del_tau, pts[l_cell]);
                        particleSubStepPicard(l_sub - (double)l_cell, v_sub, l_f, v_f, d_half, del_tau, s.EField[l_cell], s.EField[l_cell + 1],
                                              dx, p.q_over_m, l_cell, atB, l_boundary, numIter);
                        fevals += numIter;
                    }
                    if (atB) {
                        int b = w.bc[l_boundary];
                        if (b == 0) l_cell = l_cell + (int)fsign(1.0, v_f);
                        else if (b == 1 || b == 4) {
                            delIdx = delIdx + 1;
                            if (l_boundary == 1) {
                                p.EL(1, iThread) = p.EL(1, iThread) + v_f * v_f + (p.P(3, i, iThread) * p.P(3, i, iThread) + p.P(4, i, iThread) * p.P(4, i, iThread));
                                p.WL(1, iThread) = p.WL(1, iThread) + 1;
                                p.ML(1, iThread) = p.ML(1, iThread) + v_f;
                            } else if (l_boundary == NB) {
                                p.EL(2, iThread) = p.EL(2, iThread) + v_f * v_f + (p.P(3, i, iThread) * p.P(3, i, iThread) + p.P(4, i, iThread) * p.P(4, i, iThread));
                                p.WL(2, iThread) = p.WL(2, iThread) + 1;
                                p.ML(2, iThread) = p.ML(2, iThread) + v_f;
                            }
                            break;
                        } else if (b == 2) {
                            if (l_boundary == NB) { if (v_f > 0) v_f = -v_f; }
                            else { if (v_f < 0) v_f = -v_f; }
                            if (!refluxedBool) {
                                refIdx = refIdx + 1;
                                if (refIdx <= p.refRecLen) p.refRecordIdx[(size_t)refIdx + (size_t)(p.refRecLen + 1) * iThread] = i - delIdx;
                                refluxedBool = true;
                            }
                        } else if (b == 3) {
                            l_f = (double)std::abs(l_boundary - NB - 1);
                            l_cell = std::abs(l_cell - NB);
                        }
                    }
                    timePassed = timePassed + del_tau;
                    del_tau = del_t - timePassed;
                    convergeTimeBool = del_tau > f_tol;
                    l_sub = l_f; v_sub = v_f;
                }
                if (!convergeTimeBool) {
                    double vy = p.P(3, i, iThread), vz = p.P(4, i, iThread);
                    p.P(1, i - delIdx, iThread) = l_f;
                    p.P(2, i - delIdx, iThread) = v_f;
                    p.P(3, i - delIdx, iThread) = vy;
                    p.P(4, i - delIdx, iThread) = vz;
                }
            }
            p.N_p[iThread] = N_p - delIdx;
            p.delIdx[iThread] = delIdx;
            p.refIdx[iThread] = refIdx;
#pragma omp critical
            { numSubStepAve[j] += trips; funcEvalCounter[j] += fevals; }
        }
    }
    for (int j = 0; j < nS; ++j) {
        Species &p = o.plist[j];
        for (int t = 1; t <= T; ++t) p.numToCollide[t] = p.N_p[t];
        // real(int)/real(int): default-real (single precision) division in the Fortran (:265-266)
        float denom = (float)(int)(p.sumNp() + p.sumDel());
        p.numSubStepsAve = (double)((float)(int)numSubStepAve[j] / denom);
        p.numFuncEvalAve = (double)((float)(int)funcEvalCounter[j] / denom);
        for (int sd = 1; sd <= 2; ++sd) {
            double se = 0.0, sw = 0.0;
            for (int t = 1; t <= T; ++t) { se += p.EL(sd, t); sw += p.WL(sd, t); }
            p.accumEnergyLoss[sd - 1] += se;
            p.accumWallLoss[sd - 1] += (int)sw;
        }
        p.substepsMove = numSubStepAve[j]; p.funcEvalsMove = funcEvalCounter[j];
        o.nl.substeps += numSubStepAve[j];
    }
}

// solveGaussElimination + solveNormalEquation  mod_BasicFunctions.f90:218-265
// A is (m x n) column-major 1-based via lambda; returns alpha[1..n]
static void solveNormalEquation(const double *A /*col-major m x n, 0-based storage*/, const double *b, int m, int n, double *out /*0-based n*/) {
    std::vector<double> An((size_t)n * n), bn(n);
    for (int j = 0; j < n; ++j) { double s = 0.0; for (int i = 0; i < m; ++i) s += A[(size_t)i + (size_t)m * j] * b[i]; bn[j] = s; }
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j) { double s = 0.0; for (int i = 0; i < m; ++i) s += A[(size_t)i + (size_t)m * j] * A[(size_t)i + (size_t)m * k]; An[(size_t)j + (size_t)n * k] = s; }
    auto a = [&](int i, int j) -> double & { return An[(size_t)(i - 1) + (size_t)n * (j - 1)]; };
    for (int k = 1; k <= n - 1; ++k) {
        double k1 = a(k, k);
        for (int i = k + 1; i <= n; ++i) a(i, k) = a(i, k) / k1;
        for (int j = k + 1; j <= n; ++j) {
            double k2 = a(k, j);
            for (int i = k + 1; i <= n; ++i) a(i, j) = a(i, j) - a(i, k) * k2;
        }
    }
    for (int i = 1; i <= n; ++i) for (int j = 1; j <= i - 1; ++j) bn[i - 1] = bn[i - 1] - a(i, j) * bn[j - 1];
    for (int i = n; i >= 1; --i) {
        for (int j = i + 1; j <= n; ++j) bn[i - 1] = bn[i - 1] - a(i, j) * bn[j - 1];
        bn[i - 1] = bn[i - 1] / a(i, i);
    }
    for (int i = 0; i < n; ++i) out[i] = bn[i];
}

static inline int fmodulo(int a, int p) { int r = a % p; return (r < 0) ? r + p : r; }

static double wtime() {
#ifdef _OPENMP
    return omp_get_wtime();
#else
    return 0.0;
#endif
}

// The Anderson(m) fixed-point loop of solveDivAmpereAnderson (mod_nonLinSolvers.f90:96-158; NGP_Periodic twin :93-140), written once
// over the map x <- G(x): `G()` overwrites x[1..N] with G(x), `onConverged()` is what the reference does when the norm test passes
// (moveParticles).  The two trees differ only in the stagnation window (`nAvg` past norms, used when m > mMin) and in the value
// stored in iterNumPicard on convergence (`iterAdd`).  Also driven stand-alone (orc_anderson_fixed_point) on the reference's own
// Anderson test problem (Serial/TestAndersonAcceleration/main.f90).
template <class GFun, class DoneFun>
static void andersonLoop(NonLin &nl, int N, double *x /* 1-based */, int nAvg, int mMin, int iterAdd, GFun G, DoneFun onConverged) {
    const int m = nl.m_Anderson, M1 = m + 1;
    if ((int)nl.phi_k.size() < N * M1) { nl.phi_k.assign((size_t)N * M1, 0.0); nl.Residual_k.assign((size_t)N * M1, 0.0); nl.fitMat.assign((size_t)N * m, 0.0); }
    auto PK = [&](int i, int c) -> double & { return nl.phi_k[(size_t)(i - 1) + (size_t)N * (c - 1)]; };
    auto RK = [&](int i, int c) -> double & { return nl.Residual_k[(size_t)(i - 1) + (size_t)N * (c - 1)]; };
    dvec normResidual(M1 + 1, 0.0), alpha(M1 + 1, 0.0);
    nl.normHist.clear();
    for (int i = 1; i <= N; ++i) PK(i, 1) = x[i];
    G();
    for (int i = 1; i <= N; ++i) { PK(i, 2) = x[i]; RK(i, 1) = PK(i, 2) - PK(i, 1); }
    { double sm = 0.0; for (int i = 1; i <= N; ++i) sm += RK(i, 1) * RK(i, 1); normResidual[1] = std::sqrt(sm); }
    nl.normHist.push_back(normResidual[1]);
    const double eps_tol = nl.eps_r * normResidual[1] + nl.eps_a * (double)std::sqrt((float)N)   /* SQRT(real(N)) is single precision */;
    for (int i = 1; i <= nl.maxIter; ++i) {
        const int index = fmodulo(i, M1) + 1, m_k = std::min(i, m);
        G();
        for (int k = 1; k <= N; ++k) RK(k, index) = x[k] - PK(k, index);
        { double sm = 0.0; for (int k = 1; k <= N; ++k) sm += RK(k, index) * RK(k, index); normResidual[index] = std::sqrt(sm); }
        nl.normHist.push_back(normResidual[index]);
        if (normResidual[index] < eps_tol) {
            onConverged();
            nl.iterNumPicard = i + iterAdd;
            return;
        }
        if (i > m) {
            double sumPast;
            if (m > mMin) {
                sumPast = 0.0;
                for (int k = 1; k <= nAvg; ++k) sumPast += normResidual[fmodulo(i - k, M1) + 1];   // same left-to-right order as the source
                sumPast = sumPast / (double)nAvg;
            } else sumPast = normResidual[fmodulo(i - 1, M1) + 1];
            if (normResidual[index] > sumPast) { nl.iterNumPicard = nl.maxIter; return; }
        }
        for (int j = 0; j <= m_k - 1; ++j) {
            const int cj = fmodulo(i - m_k + j, M1) + 1;
            for (int k = 1; k <= N; ++k) nl.fitMat[(size_t)(k - 1) + (size_t)N * j] = RK(k, index) - RK(k, cj);
        }
        solveNormalEquation(nl.fitMat.data(), &RK(1, index), N, m_k, &alpha[1]);
        { double sa = 0.0; for (int j = 1; j <= m_k; ++j) sa += alpha[j]; alpha[m_k + 1] = 1.0 - sa; }
        const int cn = fmodulo(i + 1, M1) + 1, c0 = fmodulo(i - m_k, M1) + 1;
        for (int k = 1; k <= N; ++k) PK(k, cn) = alpha[1] * (nl.Beta_k * RK(k, c0) + PK(k, c0));
        for (int j = 1; j <= m_k; ++j) {
            const int cj = fmodulo(i - m_k + j, M1) + 1;
            for (int k = 1; k <= N; ++k) PK(k, cn) = PK(k, cn) + alpha[j + 1] * (nl.Beta_k * RK(k, cj) + PK(k, cj));
        }
        for (int k = 1; k <= N; ++k) x[k] = PK(k, cn);
    }
    // loop exhausted: the reference leaves iterNumPicard untouched (:112-158)
}

// solveDivAmpereAnderson  mod_nonLinSolvers.f90:85-160 (stagnation: mean of the last three norms when m > 3, :120-126)
static void solveDivAmpereAnderson(Oracle &o, double del_t) {
    Solver &s = o.solver; NonLin &nl = o.nl;
    makeConstSourceTerm(o);
    andersonLoop(nl, o.world.Nx, s.phi_f.data(), 3, 3, 1,
        [&]() {
            double t0 = wtime(); depositJ(o, del_t); nl.moverSeconds += wtime() - t0;
            t0 = wtime(); solve_tridiag_Ampere(o, del_t); nl.potSeconds += wtime() - t0;
        },
        [&]() { double t0 = wtime(); moveParticles(o, del_t); nl.moverSeconds += wtime() - t0; });
}

// adaptiveSolveDivAmpereAnderson  mod_nonLinSolvers.f90:162-196 ; returns nonzero where the reference `stop`s
static int adaptiveSolveDivAmpereAnderson(Oracle &o, double del_t, double &remainDel_t, double &currDel_t, double timeCurrent) {
    NonLin &nl = o.nl;
    currDel_t = remainDel_t;
    if (o.solver.RF_bool) setRFVoltage(o, timeCurrent + remainDel_t);
    solveDivAmpereAnderson(o, remainDel_t);
    if (nl.iterNumPicard < nl.maxIter) {
        remainDel_t = del_t;
    } else {
        nl.amountTimeSplits += 1;
        nl.iterNumAdaptiveSteps = 0;
        while (nl.iterNumPicard == nl.maxIter) {
            currDel_t = currDel_t / 2.0;
            nl.iterNumAdaptiveSteps += 1;
            if (nl.iterNumAdaptiveSteps > 4) return 1;
            if (o.solver.RF_bool) setRFVoltage(o, timeCurrent + currDel_t);
            solveDivAmpereAnderson(o, currDel_t);
        }
        remainDel_t = remainDel_t - currDel_t;
    }
    return 0;
}

// solvePotential  mod_nonLinSolvers.f90:198-215 (Anderson branch)
static int solvePotential(Oracle &o, double del_t, double &remainDel_t, double &currDel_t, double timeCurrent) {
    o.nl.sweeps = 0; o.nl.substeps = 0;
    resetVoltage(o);
    if (o.nl.solverType == 0) return adaptiveSolveDivAmpereAnderson(o, del_t, remainDel_t, currDel_t, timeCurrent);
    return 2;   // JFNK: not restated in this file
}


// =============================================================================
// NGP_Periodic tree (E-field form): the unknown of the nonlinear solve is EField_f (N_h), the field update is the
// explicit Ampere law, positions are advanced in cell-local coordinates.
//   OpenMP/NGP_Periodic/src/mod_particleMover.f90, mod_potentialSolver.f90, mod_nonLinSolvers.f90, mod_Scheme.f90
// The grid (mod_domain.f90) is the NGP grid without smoothing; world.scheme stays 0.
// =============================================================================

// construct_diagMatrix  NGP_Periodic/src/mod_potentialSolver.f90:88-126 (Neumann rows carry factor 1, not 2)
static int P_construct_diagMatrix(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx;
    std::fill(s.a_tri.begin(), s.a_tri.end(), 0.0); std::fill(s.b_tri.begin(), s.b_tri.end(), 0.0); std::fill(s.c_tri.begin(), s.c_tri.end(), 0.0);
    for (int i = 1; i <= Nx; ++i) {
        switch (w.bc[i]) {
        case 0:
            if (i < Nx) s.c_tri[i] = 1.0 / w.dx_dl[i];
            if (i > 1) s.a_tri[i - 1] = 1.0 / w.dx_dl[i - 1];
            s.b_tri[i] = -(1.0 / w.dx_dl[i - 1] + 1.0 / w.dx_dl[i]);
            break;
        case 1: case 4: s.b_tri[i] = 1.0; break;
        case 2:
            if (i == 1) { s.c_tri[i] = 1.0 / w.dx_dl[i]; s.b_tri[i] = -(1.0 / w.dx_dl[i]); }
            else if (i == Nx) { s.a_tri[i - 1] = 1.0 / w.dx_dl[i - 1]; s.b_tri[i] = -(1.0 / w.dx_dl[i - 1]); }
            else return 1;
            break;
        case 3: s.b_tri[i] = 1.0; break;
        default: break;
        }
    }
    return 0;
}

// potentialSolver_constructor  :33-86
static int P_solver_init(Oracle &o, double leftV, double rightV, double RF_frequency) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    s.eform = true;
    s.phi.assign(Nx + 2, 0.0); s.phi_f.assign(Nx + 2, 0.0); s.rho.assign(Nx + 2, 0.0); s.J.assign(Nh + 2, 0.0);
    s.EField.assign(Nh + 2, 0.0); s.EField_f.assign(Nh + 2, 0.0); s.EField_half.assign(Nh + 2, 0.0);
    s.a_tri.assign(Nx + 2, 0.0); s.b_tri.assign(Nx + 2, 0.0); s.c_tri.assign(Nx + 2, 0.0);
    if (P_construct_diagMatrix(o)) return 1;
    s.rho_const = 0.0; s.RF_half_amplitude = 0.0;
    if (w.bc[1] == 1) s.phi[1] = leftV;
    if (w.bc[1] == 4) s.RF_half_amplitude = leftV;
    if (w.bc[Nx] == 1) s.phi[Nx] = rightV;
    if (w.bc[Nx] == 4) { if (s.RF_half_amplitude != 0) return 2; s.RF_half_amplitude = rightV; }
    if (w.bc[1] == 3) { s.phi[1] = leftV; s.phi[Nx] = leftV; }
    s.RF_rad_frequency = 2.0 * pi * RF_frequency;
    s.RF_bool = s.RF_rad_frequency > 0 && s.RF_half_amplitude > 0 && (w.bc[1] == 4 || w.bc[Nx] == 4);
    s.aveJ = 0.0; s.globJBool = false;
    return 0;
}

// addThreadedDomainArray  NGP_Periodic/src/mod_domain.f90:261-273: array(slice) += SUM_t x(slice,t) * const, species by species
template <class GET>
static void P_add_threaded(Oracle &o, double *arr /*1-based*/, int nSlices, const ivec &L, const ivec &R, GET get) {
    const int T = o.world.T;
    for (auto &p : o.plist) {
#pragma omp parallel for schedule(static, 1)
        for (int it = 1; it <= T; ++it) {
            if (it <= nSlices) {
                for (int i = L[it]; i <= R[it]; ++i) {
                    double sm = 0.0;
                    for (int t = 1; t <= T; ++t) sm += get(p, i, t);
                    arr[i] = arr[i] + sm * p.q_times_wp;
                }
            }
        }
    }
}

// depositRho  NGP_Periodic/src/mod_Scheme.f90:57-70 (interpolateParticleToNodes :25-37): raw linear weights, no wall fix-ups
static void P_depositRho(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, T = w.T;
    for (int i = 1; i <= Nx; ++i) s.rho[i] = 0.0;
#pragma omp parallel for schedule(static, 1)
    for (int it = 1; it <= T; ++it) {
        for (auto &p : o.plist) {
            for (int i = 0; i <= Nx + 1; ++i) p.DN(i, it) = 0.0;
            for (int j = 1; j <= p.N_p[it]; ++j) {
                const int l_left = (int)p.P(1, j, it);
                const double d = p.P(1, j, it) - (double)l_left;
                p.DN(l_left, it) = p.DN(l_left, it) + (1.0 - d);
                p.DN(l_left + 1, it) = p.DN(l_left + 1, it) + d;
            }
        }
    }
    P_add_threaded(o, s.rho.data(), w.nTN, w.tnL, w.tnR, [](Species &p, int i, int t) -> double & { return p.DN(i, t); });
}

// solve_tridiag_Poisson  :127-158 (also sets EField and EField_f)
static void P_solve_tridiag_Poisson(Oracle &o, double timeCurrent) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    dvec d(Nx + 2, 0.0), cp(Nh + 2, 0.0), dp(Nx + 2, 0.0);
    for (int i = 1; i <= Nx; ++i) {
        switch (w.bc[i]) {
        case 0: case 2: d[i] = (-s.rho[i] - s.rho_const) / eps_0; break;
        case 1: case 3: d[i] = s.phi[i]; break;
        case 4: d[i] = s.RF_half_amplitude * std::sin(s.RF_rad_frequency * timeCurrent); break;
        }
    }
    cp[1] = s.c_tri[1] / s.b_tri[1];
    dp[1] = d[1] / s.b_tri[1];
    for (int i = 2; i <= Nh; ++i) {
        const double m = s.b_tri[i] - cp[i - 1] * s.a_tri[i - 1];
        cp[i] = s.c_tri[i] / m;
        dp[i] = (d[i] - dp[i - 1] * s.a_tri[i - 1]) / m;
    }
    {
        const double m = s.b_tri[Nx] - cp[Nh] * s.a_tri[Nh];
        dp[Nx] = (d[Nx] - dp[Nh] * s.a_tri[Nh]) / m;
    }
    for (int i = 1; i <= Nx; ++i) s.phi[i] = dp[i];
    for (int i = Nh; i >= 1; --i) s.phi[i] = dp[i] - cp[i] * s.phi[i + 1];
    for (int i = 1; i <= Nh; ++i) { s.EField[i] = (s.phi[i] - s.phi[i + 1]) / w.dx_dl[i]; s.EField_f[i] = s.EField[i]; }
    for (int i = 1; i <= Nx; ++i) s.phi_f[i] = s.phi[i];
}

// solve_tridiag_Ampere  :170-176: explicit Ampere update of EField_f
static void P_solve_tridiag_Ampere(Oracle &o, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nh = w.Nh;
    if (!s.globJBool) {
        double sm = 0.0;
        for (int i = 1; i <= Nh; ++i) sm += s.J[i] * w.dx_dl[i];
        s.aveJ = sm / w.L_domain;
    }
    for (int i = 1; i <= Nh; ++i) s.EField_f[i] = (s.aveJ - s.J[i]) * del_t / eps_0 + s.EField[i];
}

// particleSubStepNoBField  NGP_Periodic/src/mod_particleMover.f90:18-59 (cell-local l in [0,1])
static inline void P_subStep(double l_sub, double v_sub, double &l_f, double &v_f, double &del_tau, double E_x, double q_over_m, double dx_dl,
                             bool &atB, int &l_boundary) {
    const double a = q_over_m * E_x;
    v_f = v_sub + a * del_tau;
    l_f = l_sub + 0.5 * (v_sub + v_f) * del_tau / dx_dl;
    const bool inCell = ((int)(l_f + 1.0) == 1);
    const bool eqSign = ((v_sub > 0) == (v_f > 0));
    atB = (!inCell) || (!eqSign);
    if (atB) {
        l_boundary = (v_sub > 0) ? 1 : 0;
        double diff_PE = 2.0 * a * ((double)l_boundary - l_sub) * dx_dl;
        const double v_i_sqr = v_sub * v_sub;
        atB = (diff_PE > -v_i_sqr);
        if (atB) {
            l_f = (double)l_boundary;
            v_f = fsign(1.0, v_sub) * std::sqrt(diff_PE + v_i_sqr);
            del_tau = (v_f - v_sub) / a;
        } else if (!inCell) {
            atB = true;
            l_boundary = 1 - l_boundary;
            l_f = (double)l_boundary;
            diff_PE = 2.0 * a * (l_f - l_sub) * dx_dl;
            v_f = -fsign(1.0, v_sub) * std::sqrt(diff_PE + v_i_sqr);
            del_tau = (v_f - v_sub) / a;
        }
    }
}

// depositJ :60-135 / moveParticles :136-238 share this sweep (MOVE selects the bookkeeping)
template <bool MOVE>
static void P_sweep(Oracle &o, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh, T = w.T, nS = (int)o.plist.size();
    for (int i = 1; i <= Nh; ++i) s.EField_half[i] = 0.5 * (s.EField[i] + s.EField_f[i]);
    if (!MOVE) for (int i = 1; i <= Nh; ++i) s.J[i] = 0.0;
    const double f_tol = del_t * (double)1e-12f;                  // `del_t * 1e-12`: default-real literal (:75,:150)
    std::vector<long long> numSubStepAve(nS, 0);
    if (!MOVE) for (auto &p : o.plist) p.substepsDeposit = 0;
#pragma omp parallel for schedule(static, 1)
    for (int iThread = 1; iThread <= T; ++iThread) {
        for (int j = 0; j < nS; ++j) {
            Species &p = o.plist[j];
            long long trips = 0;
            int delIdx = 0, refIdx = 0;
            if (MOVE) { p.WL(1, iThread) = 0; p.WL(2, iThread) = 0; p.EL(1, iThread) = 0.0; p.EL(2, iThread) = 0.0; }
            else for (int i = 1; i <= Nh; ++i) p.WS(i, iThread) = 0.0;
            const int N_p = p.N_p[iThread];
            for (int i = 1; i <= N_p; ++i) {
                double v_sub = p.P(2, i, iThread), l_sub = p.P(1, i, iThread);
                double timePassed = 0.0, l_f = 0, v_f = 0, del_tau;
                bool refluxedBool = false, futB = false, absorbed = false;
                bool atBoundary = (std::fmod(l_sub, 1.0) == 0.0);
                bool timeNotConverged = true;
                int l_cell, l_boundary = 0;
                while (timeNotConverged) {
                    ++trips;
                    if (!atBoundary) l_cell = (int)l_sub;
                    else l_cell = (int)l_sub + ((int)fsign(1.0, v_sub) - 1) / 2;
                    const double E_x = s.EField_half[l_cell], dx = w.dx_dl[l_cell];
                    del_tau = del_t - timePassed;
                    l_sub = l_sub - (double)l_cell;
                    P_subStep(l_sub, v_sub, l_f, v_f, del_tau, E_x, p.q_over_m, dx, futB, l_boundary);
                    if (!MOVE) p.WS(l_cell, iThread) = p.WS(l_cell, iThread) + l_f - l_sub;
                    l_f = l_f + (double)l_cell;
                    if (futB) {
                        l_boundary = l_boundary + l_cell;
                        const int b = w.bc[l_boundary];
                        if (b == 1 || b == 4) {
                            if (MOVE) {
                                delIdx = delIdx + 1;
                                const double e_perp = p.P(3, i, iThread) * p.P(3, i, iThread) + p.P(4, i, iThread) * p.P(4, i, iThread);
                                if (l_boundary == 1) { p.EL(1, iThread) = p.EL(1, iThread) + v_f * v_f + e_perp; p.WL(1, iThread) = p.WL(1, iThread) + 1; }
                                else if (l_boundary == Nx) { p.EL(2, iThread) = p.EL(2, iThread) + v_f * v_f + e_perp; p.WL(2, iThread) = p.WL(2, iThread) + 1; }
                            }
                            absorbed = true;
                            break;
                        } else if (b == 2) {
                            if (l_boundary == Nx) { if (v_f > 0) v_f = -v_f; }
                            else { if (v_f < 0) v_f = -v_f; }
                            if (MOVE && !refluxedBool) {
                                refIdx = refIdx + 1;
                                if (refIdx <= p.refRecLen) p.refRecordIdx[(size_t)refIdx + (size_t)(p.refRecLen + 1) * iThread] = i - delIdx;
                                refluxedBool = true;
                            }
                        } else if (b == 3) {
                            l_f = (double)std::abs(l_boundary - Nx - 1);
                        }
                    }
                    timePassed = timePassed + del_tau;
                    timeNotConverged = (del_t - timePassed > f_tol);
                    l_sub = l_f; v_sub = v_f;
                    atBoundary = futB;
                }
                if (MOVE && !absorbed) {
                    const double vy = p.P(3, i, iThread), vz = p.P(4, i, iThread);
                    p.P(1, i - delIdx, iThread) = l_f; p.P(2, i - delIdx, iThread) = v_f;
                    p.P(3, i - delIdx, iThread) = vy; p.P(4, i - delIdx, iThread) = vz;
                }
            }
            if (MOVE) { p.N_p[iThread] = N_p - delIdx; p.delIdx[iThread] = delIdx; p.refIdx[iThread] = refIdx; }
#pragma omp critical
            { if (MOVE) numSubStepAve[j] += trips; else p.substepsDeposit += trips; }
        }
    }
    if (!MOVE) {
        P_add_threaded(o, s.J.data(), w.nTH, w.thL, w.thR, [](Species &p, int i, int t) -> double & { return p.WS(i, t); });
        for (int i = 1; i <= Nh; ++i) s.J[i] = s.J[i] / del_t;
        o.nl.sweeps += 1;
        for (auto &p : o.plist) o.nl.substeps += p.substepsDeposit;
    } else {
        for (int j = 0; j < nS; ++j) {
            Species &p = o.plist[j];
            for (int t = 1; t <= T; ++t) p.numToCollide[t] = p.N_p[t];
            const float denom = (float)(int)(p.sumNp() + p.sumDel());
            p.numSubStepsAve = (double)((float)(int)numSubStepAve[j] / denom);
            p.numFuncEvalAve = 0.0;
            for (int sd = 1; sd <= 2; ++sd) {
                double se = 0.0, sw = 0.0;
                for (int t = 1; t <= T; ++t) { se += p.EL(sd, t); sw += p.WL(sd, t); }
                p.accumEnergyLoss[sd - 1] += se;
                p.accumWallLoss[sd - 1] += (int)sw;
            }
            p.substepsMove = numSubStepAve[j]; p.funcEvalsMove = 0;
            o.nl.substeps += numSubStepAve[j];
        }
    }
}

// solveDivAmpereAnderson  NGP_Periodic/src/mod_nonLinSolvers.f90:85-141: Anderson on EField_f (N_h); iterNumPicard = i;
// stagnation: mean of the last two norms when m > 1
static void P_solveDivAmpereAnderson(Oracle &o, double del_t) {
    Solver &s = o.solver; NonLin &nl = o.nl;
    andersonLoop(nl, o.world.Nh, s.EField_f.data(), 2, 1, 0,
        [&]() { P_sweep<false>(o, del_t); P_solve_tridiag_Ampere(o, del_t); },
        [&]() { P_sweep<true>(o, del_t); });
}

// adaptiveSolveDivAmpereAnderson :142-172 (setRFVoltage is an empty routine in this tree) ; solvePotential :173-189
static int P_solvePotential(Oracle &o, double del_t, double &remainDel_t, double &currDel_t) {
    NonLin &nl = o.nl; Solver &s = o.solver;
    nl.sweeps = 0; nl.substeps = 0;
    for (int i = 1; i <= o.world.Nh; ++i) s.EField[i] = s.EField_f[i];      // resetVoltage :283-286
    if (nl.solverType != 0) return 2;
    currDel_t = remainDel_t;
    P_solveDivAmpereAnderson(o, remainDel_t);
    if (nl.iterNumPicard < nl.maxIter) { remainDel_t = del_t; return 0; }
    nl.amountTimeSplits += 1;
    nl.iterNumAdaptiveSteps = 0;
    while (nl.iterNumPicard == nl.maxIter) {
        currDel_t = currDel_t / 2.0;
        nl.iterNumAdaptiveSteps += 1;
        if (nl.iterNumAdaptiveSteps > 4) return 1;
        P_solveDivAmpereAnderson(o, currDel_t);
    }
    remainDel_t = remainDel_t - currDel_t;
    return 0;
}

// getTotalPE :226-236 ; getError_tridiag_Poisson :159-169 ; getChargeContinuityError :187-225
static double P_getTotalPE(Oracle &o, bool future) {
    Domain &w = o.world; Solver &s = o.solver;
    const dvec &E = future ? s.EField_f : s.EField;
    double sm = 0.0;
    for (int i = 1; i <= w.Nh; ++i) sm += E[i] * E[i] * w.dx_dl[i];
    return 0.5 * eps_0 * sm;
}
static double P_gaussError(Oracle &o) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nh = w.Nh, Nx = w.Nx;
    double r = eps_0 * (s.EField_f[1] - s.EField_f[Nh]) / (s.rho[1] + s.rho[Nx]) - 1.0;
    double res = r * r;
    for (int i = 2; i <= Nh; ++i) { r = eps_0 * (s.EField_f[i] - s.EField_f[i - 1]) / s.rho[i] - 1.0; res = res + r * r; }
    return std::sqrt(res / Nh);
}
static double P_chargeError(Oracle &o, const double *rho_i /*1-based*/, double del_t) {
    Domain &w = o.world; Solver &s = o.solver;
    const int Nx = w.Nx, Nh = w.Nh;
    double err = 0.0; int k = 0;
    for (int i = 1; i <= Nx; ++i) {
        double dR = s.rho[i] - rho_i[i];
        if (dR != 0) {
            switch (w.bc[i]) {
            case 0: { const double v = 1.0 + del_t * (s.J[i] - s.J[i - 1]) / dR; err += v * v; k += 1; break; }
            case 2: {
                const double v = (i == 1) ? 1.0 + del_t * s.J[1] / dR : 1.0 - del_t * s.J[Nh] / dR;
                err += v * v; k += 1; break; }
            case 3:
                if (i == 1) { dR = dR + s.rho[Nx] - rho_i[Nx]; const double v = 1.0 + del_t * (s.J[1] - s.J[Nh]) / dR; err += v * v; }
                break;
            default: break;
            }
        }
    }
    return std::sqrt(err / k);
}


// =============================================================================
// Null-collision Monte-Carlo collisions
//   OpenMP/SharedModules/src/mod_NullCollision.f90 (constructor :44-108, generateCollision :110-237,
//   ionizationCollisionIsotropic :359-444, elasticExcitCollisionIsotropic :511-562, scatterVector :634-661)
//   OpenMP/SharedModules/src/mod_targetParticle.f90:42-56 (generate3DMaxwellianVelocity)
// Pinned on the reference's own known-answer test OpenMP-BField/TestMCC/MCCExample.f90:13,97-98 (electron swarm in He at
// 25 Td against the BOLSIG collision frequency) in tests/test_mcc_cpu.py.
// =============================================================================

// indexSortArray  mod_BasicFunctions.f90:316-403 (Numerical Recipes indexx).  For n < 7 the routine is the straight insertion
// of :336-345 (stable, ascending); larger inputs take its quicksort, whose result differs from a stable sort only in the order
// of equal keys -- which no caller depends on (equal energies are removed, equal sigma_v maxima do not occur).
static void indexSortArray(int n, const double *arr /*1-based*/, int *indx /*1-based*/) {
    for (int j = 1; j <= n; ++j) indx[j] = j;
    if (n < 8) {
        for (int j = 2; j <= n; ++j) {
            const int indxt = indx[j];
            const double a = arr[indxt];
            int i = j - 1;
            for (; i >= 1; --i) {
                if (arr[indx[i]] <= a) break;
                indx[i + 1] = indx[i];
            }
            indx[i + 1] = indxt;
        }
    } else {
        std::stable_sort(indx + 1, indx + n + 1, [&](int a, int b) { return arr[a] < arr[b]; });
    }
}

// nullCollision_constructor :44-108.  sigma (lengthArrays, numberCollisions) column-major 0-based storage at this interface.
static void nullCollision_init(NullCollision &nc, int numberCollisions, int lengthArrays, double red_mass, double sumMass,
                               double redMassTripleProducts, const double *energyArray, const double *sigmaArray, const double *energyThreshold,
                               const int *collisionType, int primary, int target, const int *numberProducts, const int *productsIndx /*(3,nc)*/) {
    nc.numberCollisions = numberCollisions; nc.lengthArrays = lengthArrays;
    nc.sumMass = sumMass; nc.reducedMass = red_mass; nc.reducedMassIonization = redMassTripleProducts;
    nc.primary = primary; nc.target = target;
    auto SIG = [&](int j, int i) { return sigmaArray[(size_t)(j - 1) + (size_t)lengthArrays * (size_t)(i - 1)]; };
    // order collisions by maximum sigma_v in each collision (:55-70)
    dvec sigma_v_max(numberCollisions + 1, 0.0);
    ivec indxOrder(numberCollisions + 1, 0);
    for (int i = 1; i <= numberCollisions; ++i) {
        sigma_v_max[i] = 0;
        for (int j = 1; j <= lengthArrays; ++j) {
            const double v_r = std::sqrt(2.0 * energyArray[j - 1] * e_c / nc.reducedMass);
            sigma_v_max[i] = std::max(sigma_v_max[i], SIG(j, i) * v_r);
        }
    }
    indexSortArray(numberCollisions, sigma_v_max.data(), indxOrder.data());
    for (int i = 1; i <= numberCollisions / 2; ++i) std::swap(indxOrder[i], indxOrder[numberCollisions - i + 1]);
    nc.energyArray.assign(lengthArrays + 1, 0.0);
    nc.sigmaArray.assign((size_t)(lengthArrays + 1) * numberCollisions + 1, 0.0);
    nc.energyThreshold.assign(numberCollisions + 1, 0.0);
    nc.collisionType.assign(numberCollisions + 1, 0); nc.numberProducts.assign(numberCollisions + 1, 0);
    nc.productsIndx.assign((size_t)4 * numberCollisions + 4, 0);
    for (int j = 1; j <= lengthArrays; ++j) nc.energyArray[j] = energyArray[j - 1];
    for (int i = 1; i <= numberCollisions; ++i) {
        const int j = indxOrder[i];
        for (int k = 1; k <= lengthArrays; ++k) nc.SA(k, i) = SIG(k, j);
        nc.energyThreshold[i] = energyThreshold[j - 1];
        nc.collisionType[i] = collisionType[j - 1];
        nc.numberProducts[i] = numberProducts[j - 1];
        for (int k = 1; k <= 3; ++k) nc.PI(k, i) = productsIndx[(k - 1) + 3 * (j - 1)];
    }
    // sigma_v_max total along the energy array (:88-98)
    nc.sigmaVMax = 0.0;
    for (int i = 1; i <= lengthArrays; ++i) {
        double sum_sigma = 0;
        const double v_r = std::sqrt(2.0 * energyArray[i - 1] * e_c / nc.reducedMass);
        for (int j = 1; j <= numberCollisions; ++j) sum_sigma = sum_sigma + SIG(i, j);
        nc.sigmaVMax = std::max(nc.sigmaVMax, sum_sigma * v_r);
    }
    nc.totalAmountCollisions.assign(numberCollisions + 1, 0);
    nc.totalIncidentEnergy.assign(numberCollisions + 1, 0.0);
    nc.totalEnergyLoss.assign(numberCollisions + 1, 0.0);
    nc.totalNumberCollidableParticles = 0;
    nc.minEnergy = *std::min_element(energyArray, energyArray + lengthArrays);
    nc.maxEnergy = *std::max_element(energyArray, energyArray + lengthArrays);
}

// generate3DMaxwellianVelocity  mod_targetParticle.f90:42-56
static inline void targetMaxwellian(const TargetParticle &tp, int32_t &irand, double res[3]) {
    const double U1 = ran2(irand), U2 = ran2(irand), U3 = ran2(irand), U4 = ran2(irand);
    res[0] = tp.v_therm * std::sqrt(-2 * std::log(U1)) * std::cos(2 * pi * U2);
    res[1] = tp.v_therm * std::sqrt(-2 * std::log(U1)) * std::sin(2 * pi * U2);
    res[2] = tp.v_therm * std::sqrt(-2 * std::log(U3)) * std::sin(2 * pi * U4);
}

// scatterVector :634-661
static inline void scatterVector(double u[3], const double ev[3], double costheta, double phi) {
    const double sintheta = std::sqrt(1.0 - costheta * costheta), sinphi = std::sin(phi), cosphi = std::cos(phi);
    if (std::fabs(ev[1]) > std::fabs(ev[2])) {
        const double vv = std::sqrt(ev[0] * ev[0] + ev[1] * ev[1]);
        u[0] = ev[0] * costheta + (ev[1] * sinphi + ev[0] * ev[2] * cosphi) / vv * sintheta;
        u[1] = ev[1] * costheta + (-ev[0] * sinphi + ev[1] * ev[2] * cosphi) / vv * sintheta;
        u[2] = ev[2] * costheta - vv * cosphi * sintheta;
    } else {
        const double vv = std::sqrt(ev[0] * ev[0] + ev[2] * ev[2]);
        u[0] = ev[0] * costheta + (ev[2] * sinphi - ev[1] * ev[0] * cosphi) / vv * sintheta;
        u[1] = ev[1] * costheta + vv * cosphi * sintheta;
        u[2] = ev[2] * costheta - (ev[0] * sinphi + ev[1] * ev[2] * cosphi) / vv * sintheta;
    }
}

// elasticExcitCollisionIsotropic :511-562
static inline void elasticExcitCollisionIsotropic(const NullCollision &nc, double primary_mass, double target_mass, int32_t &irand, double energyCM,
                                                  double E_thres, double incidentVelocity[3], double targetVelocity[3]) {
    const double delE = (energyCM - E_thres) * e_c;
    double P_beginning[3], V_cm[3], e_vector[3];
    for (int k = 0; k < 3; ++k) { P_beginning[k] = primary_mass * incidentVelocity[k] + target_mass * targetVelocity[k]; V_cm[k] = P_beginning[k] / nc.sumMass; }
    const double cos_theta = 1.0 - 2.0 * ran2(irand);
    const double phi = ran2(irand) * 2.0 * pi;
    const double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    const double cos_phi = std::cos(phi), sin_phi = std::sin(phi);
    e_vector[0] = cos_theta; e_vector[1] = sin_phi * sin_theta; e_vector[2] = cos_phi * sin_theta;
    const double speedPerParticle = std::sqrt(2.0 * delE * nc.reducedMass / (primary_mass * primary_mass));
    for (int k = 0; k < 3; ++k) incidentVelocity[k] = e_vector[k] * speedPerParticle + V_cm[k];
    for (int k = 0; k < 3; ++k) targetVelocity[k] = (P_beginning[k] - primary_mass * incidentVelocity[k]) / target_mass;
}

// ionizationCollisionIsotropic :359-444 (active statements only)
static inline void ionizationCollisionIsotropic(const NullCollision &nc, double primary_mass, double ion_mass, double target_mass, int32_t &irand,
                                                double energyCM, double E_thres, double incidentVelocity[3], double targetVelocity[3],
                                                double velocityCM[3]) {
    const double delE = (energyCM - E_thres) * e_c;
    double P_beginning[3], V_cm[3], e_vector[3], u_vector[3];
    for (int k = 0; k < 3; ++k) { P_beginning[k] = primary_mass * incidentVelocity[k] + target_mass * targetVelocity[k]; V_cm[k] = P_beginning[k] / nc.sumMass; }
    double cos_theta = 1.0 - 2.0 * ran2(irand);
    double phi = ran2(irand) * 2.0 * pi;
    const double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
    const double cos_phi = std::cos(phi), sin_phi = std::sin(phi);
    e_vector[0] = cos_theta; e_vector[1] = sin_phi * sin_theta; e_vector[2] = cos_phi * sin_theta;
    double speedPerParticle = std::sqrt(2.0 * delE * nc.reducedMassIonization / (primary_mass * primary_mass));
    for (int k = 0; k < 3; ++k) incidentVelocity[k] = e_vector[k] * speedPerParticle + V_cm[k];
    // second electron: 120 degrees from the first in the centre-of-mass frame
    cos_theta = std::cos(2.0 * pi / 3.0);
    phi = ran2(irand) * 2.0 * pi;
    scatterVector(u_vector, e_vector, cos_theta, phi);
    speedPerParticle = std::sqrt(2.0 * delE * nc.reducedMassIonization / (m_e * m_e));
    for (int k = 0; k < 3; ++k) velocityCM[k] = u_vector[k] * speedPerParticle + V_cm[k];
    for (int k = 0; k < 3; ++k) targetVelocity[k] = (P_beginning[k] - primary_mass * incidentVelocity[k] - m_e * velocityCM[k]) / ion_mass;
}

// generateCollision :110-237.  irand: T seeds (1-based).  Returns nonzero where the reference `stop`s (P_null > 0.5) or a slab would
// overflow finalIdx (the reference would write out of bounds).
static int generateCollision(Oracle &o, NullCollision &nc, int32_t *irand, double del_t) {
    Species &prim = o.plist[nc.primary - 1];
    const TargetParticle &tg = o.targets[nc.target - 1];
    const int T = o.world.T, nC = nc.numberCollisions;
    const double P_null = 1.0 - std::exp(-nc.sigmaVMax * tg.density * del_t);
    nc.P_null_last = P_null;
    if (P_null > 0.5) return 1;
    const double primary_mass = prim.mass, target_mass = tg.mass;
    { long long s = 0; for (int t = 1; t <= T; ++t) s += prim.numToCollide[t]; nc.totalNumberCollidableParticles += s; }
    // OpenMP reduction(+) arrays: per-thread copies added in thread order here
    std::vector<dvec> energyLossT(T + 1, dvec(nC + 1, 0.0)), totIncT(T + 1, dvec(nC + 1, 0.0));
    std::vector<std::vector<long long>> totalCollT(T + 1, std::vector<long long>(nC + 1, 0));
    int overflow = 0;
#pragma omp parallel for schedule(static, 1) reduction(| : overflow)
    for (int iThread = 1; iThread <= T; ++iThread) {
        int32_t irand_thread = irand[iThread];
        int numberTotalParticles = prim.numToCollide[iThread];
        const double numberSelectedReal = P_null * (double)numberTotalParticles;   // real(...) of an int32 is exact in single only below 2^24; the product is real64
        int numberSelected = (int)numberSelectedReal;
        double Rand = ran2(irand_thread);
        if (Rand < (numberSelectedReal - numberSelected)) numberSelected = numberSelected + 1;
        dvec &energyLoss = energyLossT[iThread], &totIncidentEnergy = totIncT[iThread];
        std::vector<long long> &totalCollisions = totalCollT[iThread];
        for (int i = 1; i <= numberSelected; ++i) {
            Rand = ran2(irand_thread);
            const int particleIndx = (int)((numberTotalParticles)*Rand) + 1;
            const double particleLocation = prim.P(1, particleIndx, iThread);
            double incidentVelocity[3] = {prim.P(2, particleIndx, iThread), prim.P(3, particleIndx, iThread), prim.P(4, particleIndx, iThread)};
            double targetVelocity[3], velocity_CM[3];
            targetMaxwellian(tg, irand_thread, targetVelocity);
            for (int k = 0; k < 3; ++k) velocity_CM[k] = incidentVelocity[k] - targetVelocity[k];
            const double vcm2 = velocity_CM[0] * velocity_CM[0] + velocity_CM[1] * velocity_CM[1] + velocity_CM[2] * velocity_CM[2];
            const double speedCM = std::sqrt(vcm2);
            const double energyCM = vcm2 * 0.5 * nc.reducedMass / e_c;   // eV
            int indxLow = 1, indxHigh = nc.lengthArrays;
            double d_value;
            if (energyCM < nc.minEnergy) {
                d_value = 0.0;
            } else if (energyCM > nc.maxEnergy) {
                indxLow = indxHigh - 1;
                d_value = 1.0;
            } else {
                while (indxLow != indxHigh - 1) {
                    const int indxMiddle = (indxLow + indxHigh) / 2;
                    if (nc.energyArray[indxMiddle] < energyCM) indxLow = indxMiddle;
                    else if (nc.energyArray[indxMiddle] > energyCM) indxHigh = indxMiddle;
                    else { indxLow = indxMiddle; indxHigh = indxLow + 1; }
                }
                d_value = (energyCM - nc.energyArray[indxLow]) / (nc.energyArray[indxHigh] - nc.energyArray[indxLow]);
            }
            Rand = ran2(irand_thread);
            double sigma_v_low = 0.0;
            for (int collIdx = 1; collIdx <= nC; ++collIdx) {
                if (energyCM > nc.energyThreshold[collIdx]) {
                    const double sigma_v = (nc.SA(indxLow, collIdx) * (1.0 - d_value) + nc.SA(indxHigh, collIdx) * d_value) * speedCM + sigma_v_low;
                    const bool collisionBool = (Rand <= sigma_v / nc.sigmaVMax);
                    if (collisionBool) {
                        totalCollisions[collIdx] += 1;
                        const double IncidentEnergy = incidentVelocity[0] * incidentVelocity[0] + incidentVelocity[1] * incidentVelocity[1] +
                                                      incidentVelocity[2] * incidentVelocity[2];
                        totIncidentEnergy[collIdx] += IncidentEnergy;
                        switch (nc.collisionType[collIdx]) {
                        case 1: case 3:
                            elasticExcitCollisionIsotropic(nc, primary_mass, target_mass, irand_thread, energyCM, nc.energyThreshold[collIdx],
                                                           incidentVelocity, targetVelocity);
                            break;
                        case 2: {
                            Species &ion = o.plist[nc.PI(2, collIdx) - 1];
                            ionizationCollisionIsotropic(nc, primary_mass, ion.mass, target_mass, irand_thread, energyCM, nc.energyThreshold[collIdx],
                                                         incidentVelocity, targetVelocity, velocity_CM);
                            if (prim.N_p[iThread] + 1 > prim.finalIdx || ion.N_p[iThread] + 1 > ion.finalIdx) { overflow |= 1; break; }
                            prim.N_p[iThread] += 1;
                            ion.N_p[iThread] += 1;
                            const int ne = prim.N_p[iThread], ni = ion.N_p[iThread];
                            for (int k = 0; k < 3; ++k) prim.P(2 + k, ne, iThread) = velocity_CM[k];
                            prim.P(1, ne, iThread) = particleLocation;
                            for (int k = 0; k < 3; ++k) ion.P(2 + k, ni, iThread) = targetVelocity[k];
                            ion.P(1, ni, iThread) = particleLocation;
                            energyLoss[collIdx] = energyLoss[collIdx] -
                                                  (targetVelocity[0] * targetVelocity[0] + targetVelocity[1] * targetVelocity[1] + targetVelocity[2] * targetVelocity[2]) * target_mass -
                                                  (velocity_CM[0] * velocity_CM[0] + velocity_CM[1] * velocity_CM[1] + velocity_CM[2] * velocity_CM[2]) * primary_mass;
                            break; }
                        case 4:
                            for (int k = 0; k < 3; ++k) incidentVelocity[k] = targetVelocity[k];
                            break;
                        default: break;   // 5: dissociation, nothing happens
                        }
                        energyLoss[collIdx] = energyLoss[collIdx] + primary_mass * (IncidentEnergy - (incidentVelocity[0] * incidentVelocity[0] +
                                              incidentVelocity[1] * incidentVelocity[1] + incidentVelocity[2] * incidentVelocity[2]));
                        break;
                    }
                    sigma_v_low = sigma_v;
                }
            }
            // swap with the last collidable particle so that there are no repeats (:222-227)
            for (int k = 1; k <= 4; ++k) prim.P(k, particleIndx, iThread) = prim.P(k, numberTotalParticles, iThread);
            for (int k = 0; k < 3; ++k) prim.P(2 + k, numberTotalParticles, iThread) = incidentVelocity[k];
            prim.P(1, numberTotalParticles, iThread) = particleLocation;
            numberTotalParticles = numberTotalParticles - 1;
        }
        irand[iThread] = irand_thread;
        prim.numToCollide[iThread] = numberTotalParticles;
    }
    for (int t = 1; t <= T; ++t)
        for (int c = 1; c <= nC; ++c) {
            nc.totalAmountCollisions[c] += totalCollT[t][c];
            nc.totalEnergyLoss[c] += energyLossT[t][c];
            nc.totalIncidentEnergy[c] += totIncT[t][c];
        }
    return overflow ? 2 : 0;
}

}  // namespace

// =============================================================================
// C ABI for the test harness (ctypes).  0-based plain arrays at this boundary.
// =============================================================================
// =====================================================================================================
// density diagnostic: loadParticleDensity + WriteParticleDensity (NGP mod_Scheme.f90:68-158, CIC mod_Scheme.f90:68-185)
// =====================================================================================================
// interpolateParticleToNodesDensity  CIC mod_Scheme.f90:68-105 (NGP uses interpolateParticleToNodes, mod_Scheme.f90:79)
static void interpolateParticleToNodesDensity(Species &p, const Domain &w, int t) {
    if (w.scheme != 1) { interpolateParticleToNodes(p, w, t); return; }
    const int Nx = w.Nx;
    for (int j = 1; j <= p.N_p[t]; ++j) {
        const double l = p.P(1, j, t);
        const int lc = (int)l; const double d = l - (double)lc;
        p.DN(lc, t) += (-d * d + d + 0.5);
        const int lr = lc + 1, ll = lc - 1;
        const double wl = 0.5 * (1.0 - d) * (1.0 - d), wr = 0.5 * d * d;
        switch (w.bc[lr] - w.bc[lc]) {
        case 0: p.DN(ll, t) += wl; p.DN(lr, t) += wr; break;
        case -2: case -1: case -4: p.DN(lc, t) += wl; p.DN(lr, t) += wr; break;
        case 2: case 1: case 4: p.DN(lc, t) += wr; p.DN(ll, t) += wl; break;
        case -3: p.DN(Nx, t) += wl; p.DN(lr, t) += wr; break;
        case 3: p.DN(ll, t) += wl; p.DN(1, t) += wr; break;
        }
    }
}
static void loadParticleDensity(Oracle &o, bool reset) {
    const int T = o.world.T;
#pragma omp parallel for schedule(static, 1)
    for (int it = 1; it <= T; ++it)
        for (auto &p : o.plist) {
            if (reset) for (int i = 0; i <= o.world.Nx + 1; ++i) p.DN(i, it) = 0.0;
            interpolateParticleToNodesDensity(p, o.world, it);
        }
}
// the array WriteParticleDensity writes for species sp (out[0..Nx-1]); like the reference it leaves the reduced grid in densities(:,1)
static void particleDensityArray(Oracle &o, int sp, double *out) {
    Domain &w = o.world; Species &p = o.plist[sp];
    const int Nx = w.Nx, Nh = w.Nh, T = w.T;
    for (int i = 1; i <= Nx; ++i) {                       // SUM(densities(i, :), DIM=2) * w_p
        double s = 0.0;
        for (int t = 1; t <= T; ++t) s += p.DN(i, t);
        p.DN(i, 1) = s * p.w_p;
    }
    auto D = [&](int i) -> double & { return p.DN(i, 1); };
    dvec dens(Nx + 2, 0.0);
    if (w.scheme == 0) {
        if (w.bc[1] == 3) { D(1) = D(1) + D(Nx); D(Nx) = D(1); }
        if (w.smooth) {
            for (int j = 1; j <= Nx; ++j) {
                switch (w.bc[j]) {
                case 0: dens[j] = dens[j] + 0.25 * (D(j - 1) + 2.0 * D(j) + D(j + 1)); break;
                case 1: case 4: case 2:
                    if (j == 1) { dens[j] = dens[j] + 0.25 * (2.0 * D(1) + D(2)); dens[j + 1] = dens[j + 1] + 0.25 * D(1); }
                    else { dens[j] = dens[j] + 0.25 * (2.0 * D(Nx) + D(Nh)); dens[j - 1] = dens[j - 1] + 0.25 * D(Nx); }
                    break;
                case 3: dens[j] = dens[j] + 0.25 * (D(Nh) + 2.0 * D(1) + D(2)); break;
                }
            }
        } else {
            for (int j = 1; j <= Nx; ++j) dens[j] = D(j);
        }
        switch (w.bc[1]) {
        case 1: case 2: case 4: dens[1] = 2.0 * dens[1] / w.dx_dl[1]; break;
        case 3: dens[1] = 2.0 * dens[1] / (w.dx_dl[1] + w.dx_dl[Nh]); break;
        }
        for (int j = 2; j <= Nh; ++j) dens[j] = 2.0 * dens[j] / (w.dx_dl[j - 1] + w.dx_dl[j]);
        switch (w.bc[Nx]) {
        case 1: case 2: case 4: dens[Nx] = 2.0 * dens[Nx] / w.dx_dl[Nh]; break;
        case 3: dens[Nx] = dens[1]; break;
        }
    } else {
        if (w.smooth) {
            for (int j = 1; j <= Nx; ++j) {
                switch (w.bc[j + 1] - w.bc[j]) {
                case 0: dens[j] = 0.25 * (D(j - 1) + 2.0 * D(j) + D(j + 1)); break;
                case -2: case -1: case -4: dens[j] = 0.25 * (D(2) + 3.0 * D(1)); break;
                case 2: case 1: case 4: dens[j] = 0.25 * (D(Nx - 1) + 3.0 * D(Nx)); break;
                case -3: dens[j] = 0.25 * (D(Nx) + 2.0 * D(1) + D(2)); break;
                case 3: dens[j] = 0.25 * (D(1) + 2.0 * D(Nx) + D(Nx - 1)); break;
                }
            }
            for (int j = 1; j <= Nx; ++j) D(j) = dens[j];
        }
        for (int j = 1; j <= Nx; ++j) dens[j] = D(j) / w.dx_dl[j];
    }
    for (int j = 1; j <= Nx; ++j) out[j - 1] = dens[j];
}

// =====================================================================================================
// EEDF / velocity histograms of solveSimulationFinalAverage  mod_simulation.f90:433-500 (one species)
// =====================================================================================================
// extrema pass :437-452 ; out3 = {VMax, EMax, Emin}
static void velocityExtrema(Oracle &o, int sp, double out3[3]) {
    Species &p = o.plist[sp];
    double VMax = 0.0, EMax = 0.0, Emin = 0.1 * e_c * 2.0 / p.mass;
    for (int t = 1; t <= o.world.T; ++t)
        for (int i = 1; i <= p.N_p[t]; ++i) {
            double partV = std::fabs(p.P(2, i, t));
            VMax = std::max(VMax, partV);
            partV = p.P(2, i, t) * p.P(2, i, t) + p.P(3, i, t) * p.P(3, i, t) + p.P(4, i, t) * p.P(4, i, t);
            EMax = std::max(EMax, partV);
            Emin = std::min(Emin, partV);
        }
    out3[0] = VMax; out3[1] = EMax; out3[2] = Emin;
}
// histogram pass :480-500 ; VHist(2*binNumber), EHist(binNumber) 1-based in the reference, 0-based arrays here; E_grid :453-461
static void velocityHistograms(Oracle &o, int sp, int binNumber, double VMax, double Emin, double EMax, double *VHist, double *EHist, double *E_grid) {
    Species &p = o.plist[sp];
    const double Emin_log = std::log(Emin);
    const double diffE = (std::log(EMax) - std::log(Emin)) / (double)(float)(binNumber - 1);
    if (E_grid) {
        double partV = std::log(Emin);
        for (int i = 1; i <= binNumber; ++i) { E_grid[i - 1] = std::exp(partV); partV = partV + diffE; }
    }
    for (int t = 1; t <= o.world.T; ++t)
        for (int k = 1; k <= p.N_p[t]; ++k) {
            double partV = p.P(2, k, t) * (2.0 * binNumber - 1) / 2.0 / VMax + binNumber + 0.5;
            if (partV > 1 && partV < binNumber * 2) {
                const int intPartV = (int)partV;
                VHist[intPartV - 1] += (1.0 - (partV - intPartV));
                VHist[intPartV] += partV - intPartV;
            }
            const double partE = p.P(2, k, t) * p.P(2, k, t) + p.P(3, k, t) * p.P(3, k, t) + p.P(4, k, t) * p.P(4, k, t);
            if (partE > Emin && partE < EMax) {
                partV = (std::log(partE) - Emin_log) / diffE + 1.0;
                const int intPartV = (int)partV;
                partV = partV - intPartV;
                EHist[intPartV - 1] += (1.0 - partV);
                EHist[intPartV] += partV;
            }
        }
}

// =====================================================================================================
// mod_particleInjection.f90 (the routines that consume the mover's delIdx / refIdx / refRecordIdx) + the samplers they use
// =====================================================================================================
// getMaxwellianSample  mod_BasicFunctions.f90:146-160
static inline void getMaxwellianSample(double v[3], double M, double T_gas, int32_t &irand) {
    const double v_therm = std::sqrt(T_gas * e_c / M);
    const double U1 = ran2(irand), U2 = ran2(irand), U3 = ran2(irand), U4 = ran2(irand);
    v[0] = v_therm * std::sqrt(-2.0 * std::log(U1)) * std::cos(2.0 * pi * U2);
    v[1] = v_therm * std::sqrt(-2.0 * std::log(U1)) * std::sin(2.0 * pi * U2);
    v[2] = v_therm * std::sqrt(-2.0 * std::log(U3)) * std::sin(2.0 * pi * U4);
}
// getMaxwellianFluxSample  mod_BasicFunctions.f90:162-176
static inline void getMaxwellianFluxSample(double v[3], double M, double T_gas, int32_t &irand) {
    const double v_therm = std::sqrt(T_gas * e_c / M);
    const double U1 = ran2(irand), U2 = ran2(irand), U3 = ran2(irand), U4 = ran2(irand);
    v[0] = v_therm * std::sqrt(-2 * std::log(U1)) * ((U2 - 0.5 >= 0.0) ? 1.0 : -1.0);       // SIGN(1.0d0, U(2) - 0.5d0)
    v[1] = v_therm * std::sqrt(-2 * std::log(U3)) * std::cos(2 * pi * U4);
    v[2] = v_therm * std::sqrt(-2 * std::log(U3)) * std::sin(2 * pi * U4);
}
static inline double sumsq3(Species &s, int i, int t) { return s.P(2, i, t) * s.P(2, i, t) + s.P(3, i, t) * s.P(3, i, t) + s.P(4, i, t) * s.P(4, i, t); }

// addMaxwellianLostParticles  mod_particleInjection.f90:16-40 ; returns nonzero where a slab would overflow
static int addMaxwellianLostParticles(Oracle &o, double T_e, double T_i, int32_t *irand /*1-based*/, double *energyAddColl /*1-based*/) {
    Species &el = o.plist[0], &io = o.plist[1];
    for (int t = 1; t <= o.world.T; ++t) {
        int32_t ir = irand[t];
        if (el.N_p[t] + el.delIdx[t] > el.finalIdx || io.N_p[t] + el.delIdx[t] > io.finalIdx) return 1;
        for (int i = 1; i <= el.delIdx[t]; ++i) {
            double v[3];
            getMaxwellianSample(v, el.mass, T_e, ir);
            for (int k = 0; k < 3; ++k) el.P(k + 2, el.N_p[t] + i, t) = v[k];
            getMaxwellianFluxSample(v, io.mass, T_i, ir);
            for (int k = 0; k < 3; ++k) io.P(k + 2, io.N_p[t] + i, t) = v[k];
            energyAddColl[t] += (sumsq3(el, el.N_p[t] + i, t) * el.mass * el.w_p + sumsq3(io, io.N_p[t] + i, t) * io.mass * io.w_p) * 0.5;
            const double x_random = ran2(ir) * o.world.L_domain + o.world.startX;
            el.P(1, el.N_p[t] + i, t) = getLFromX(o.world, x_random);
            io.P(1, io.N_p[t] + i, t) = el.P(1, el.N_p[t] + i, t);
        }
        io.N_p[t] += el.delIdx[t];
        el.N_p[t] += el.delIdx[t];
        irand[t] = ir;
    }
    return 0;
}
// refluxParticles  :42-69
static void refluxParticles(Oracle &o, double T_e, double T_i, int32_t *irand, double *energyAddColl) {
    for (int j = 0; j < 2; ++j) {
        Species &p = o.plist[j];
        for (int t = 1; t <= o.world.T; ++t) {
            int32_t ir = irand[t];
            for (int i = 1; i <= std::min(p.refIdx[t], p.refRecLen); ++i) {     // refRecordIdx(finalIdx/10, T): entries past the end were never stored
                const int idx = p.refRecordIdx[(size_t)i + (size_t)(p.refRecLen + 1) * t];
                energyAddColl[t] -= sumsq3(p, idx, t) * p.mass * p.w_p * 0.5;
                double v[3];
                getMaxwellianFluxSample(v, p.mass, j == 0 ? T_e : T_i, ir);
                for (int k = 0; k < 3; ++k) p.P(k + 2, idx, t) = v[k];
                p.P(2, idx, t) = (o.world.bc[1] == 2) ? std::fabs(p.P(2, idx, t)) : -std::fabs(p.P(2, idx, t));
                energyAddColl[t] += sumsq3(p, idx, t) * p.mass * p.w_p * 0.5;
            }
            irand[t] = ir;
        }
    }
}
// injectAtBoundary  :71-130 ; numFluxParticlesLow/High, injectionR from readInjectionInputs :232-238
static int injectAtBoundary(Oracle &o, double T_e, double del_t, int numLow, int numHigh, double injectionR, int32_t *irand) {
    Species &el = o.plist[0], &io = o.plist[1];
    const double ionSoundSpeed = std::sqrt(T_e * e_c / io.mass);
    const bool leftNeumann = (o.world.bc[1] == 2);
    for (int j = 0; j < 2; ++j) {
        Species &p = o.plist[j];
        for (int t = 1; t <= o.world.T; ++t)
            for (int i = 1; i <= std::min(p.refIdx[t], p.refRecLen); ++i) {
                const int idx = p.refRecordIdx[(size_t)i + (size_t)(p.refRecLen + 1) * t];
                if (j == 0) {
                    double v[3];
                    getMaxwellianFluxSample(v, p.mass, T_e, irand[t]);
                    for (int k = 0; k < 3; ++k) p.P(k + 2, idx, t) = v[k];
                } else { p.P(2, idx, t) = ionSoundSpeed; p.P(3, idx, t) = 0.0; p.P(4, idx, t) = 0.0; }
                p.P(2, idx, t) = leftNeumann ? std::fabs(p.P(2, idx, t)) : -std::fabs(p.P(2, idx, t));
            }
    }
    const double endX = o.world.startX + o.world.L_domain;
    for (int t = 1; t <= o.world.T; ++t) {
        const int numInjected = (ran2(irand[t]) > injectionR) ? numLow : numHigh;
        if (el.N_p[t] + numInjected > el.finalIdx || io.N_p[t] + numInjected > io.finalIdx) return 1;
        for (int i = 1; i <= numInjected; ++i) {
            const double shift_rand = del_t * ran2(irand[t]) * ionSoundSpeed;
            double v[3];
            getMaxwellianFluxSample(v, el.mass, T_e, irand[t]);
            const int ie = el.N_p[t] + i, ii = io.N_p[t] + i;
            for (int k = 0; k < 3; ++k) el.P(k + 2, ie, t) = v[k];
            io.P(2, ii, t) = ionSoundSpeed; io.P(3, ii, t) = 0.0; io.P(4, ii, t) = 0.0;
            if (leftNeumann) {
                el.P(2, ie, t) = std::fabs(el.P(2, ie, t)); io.P(2, ii, t) = std::fabs(io.P(2, ii, t));
                el.P(1, ie, t) = getLFromX(o.world, o.world.startX + shift_rand);
            } else {
                el.P(2, ie, t) = -std::fabs(el.P(2, ie, t)); io.P(2, ii, t) = -std::fabs(io.P(2, ii, t));
                el.P(1, ie, t) = getLFromX(o.world, endX - shift_rand);
            }
            io.P(1, ii, t) = el.P(1, ie, t);
        }
        el.N_p[t] += numInjected; io.N_p[t] += numInjected;
    }
    return 0;
}
// injectUniformFlux  :132-158
static int injectUniformFlux(Oracle &o, double T_e, double T_i, int numLow, int numHigh, double injectionR, int32_t *irand) {
    Species &el = o.plist[0], &io = o.plist[1];
    for (int t = 1; t <= o.world.T; ++t) {
        const int numInjected = (ran2(irand[t]) > injectionR) ? numLow : numHigh;
        if (el.N_p[t] + numInjected > el.finalIdx || io.N_p[t] + numInjected > io.finalIdx) return 1;
        for (int i = 1; i <= numInjected; ++i) {
            double v[3];
            const int ie = el.N_p[t] + i, ii = io.N_p[t] + i;
            getMaxwellianSample(v, el.mass, 20.0 * T_e, irand[t]);
            for (int k = 0; k < 3; ++k) el.P(k + 2, ie, t) = v[k];
            getMaxwellianSample(v, io.mass, T_i, irand[t]);
            for (int k = 0; k < 3; ++k) io.P(k + 2, ii, t) = v[k];
            const double x_random = ran2(irand[t]) * o.world.L_domain + o.world.startX;
            el.P(1, ie, t) = getLFromX(o.world, x_random);
            io.P(1, ii, t) = el.P(1, ie, t);
        }
        el.N_p[t] += numInjected; io.N_p[t] += numInjected;
    }
    return 0;
}
// maxwellianHeating  :160-177 with compValue = FracFreq * (currDel_t / del_t) / FractionFreqHeating
static void maxwellianHeating(Oracle &o, int sp, double compValue, double T, int32_t *irand) {
    Species &p = o.plist[sp];
    for (int t = 1; t <= o.world.T; ++t)
        for (int i = 1; i <= p.N_p[t]; ++i) {
            const double R = ran2(irand[t]);
            if (R < compValue) { double v[3]; getMaxwellianSample(v, p.mass, T, irand[t]); for (int k = 0; k < 3; ++k) p.P(k + 2, i, t) = v[k]; }
        }
}

extern "C" {

void *orc_create(int scheme, int Nx, double L_domain, int gridType, double del_x, int bcL, int bcR, int smooth, int T) {
    Oracle *o = new Oracle();
    if (bcL == 3 || bcR == 3) { bcL = 3; bcR = 3; }
    domain_init(o->world, scheme, Nx, bcL, bcR, T);
    int rc = domain_grid(o->world, del_x, L_domain, gridType);
    if (rc) { delete o; return nullptr; }
    o->world.smooth = smooth != 0;
    return o;
}
void orc_destroy(void *h) { delete (Oracle *)h; }
int orc_Nx(void *h) { return ((Oracle *)h)->world.Nx; }
int orc_Nh(void *h) { return ((Oracle *)h)->world.Nh; }
int orc_T(void *h) { return ((Oracle *)h)->world.T; }
double orc_L_domain(void *h) { return ((Oracle *)h)->world.L_domain; }
void orc_get_grid(void *h, double *grid, double *dx_dl, int *bc) {
    Oracle *o = (Oracle *)h; Domain &w = o->world;
    for (int i = 1; i <= w.Nx; ++i) grid[i - 1] = w.grid[i];
    int nd = (w.scheme == 0) ? w.Nh : w.Nx;
    for (int i = 1; i <= nd; ++i) dx_dl[i - 1] = w.dx_dl[i];
    int nb = (w.scheme == 0) ? w.Nx : w.Nh;
    for (int i = 1; i <= nb; ++i) bc[i - 1] = w.bc[i];
}
void orc_get_slices(void *h, int *half_lr /*2*nTH*/, int *node_lr /*2*nTN*/, int *nTH, int *nTN) {
    Domain &w = ((Oracle *)h)->world;
    *nTH = w.nTH; *nTN = w.nTN;
    for (int i = 1; i <= w.nTH; ++i) { half_lr[2 * (i - 1)] = w.thL[i]; half_lr[2 * (i - 1) + 1] = w.thR[i]; }
    for (int i = 1; i <= w.nTN; ++i) { node_lr[2 * (i - 1)] = w.tnL[i]; node_lr[2 * (i - 1) + 1] = w.tnR[i]; }
}
double orc_getLFromX(void *h, double x) { return getLFromX(((Oracle *)h)->world, x); }
double orc_getXFromL(void *h, double l) { return getXFromL(((Oracle *)h)->world, l); }

int orc_init_solver(void *h, double leftV, double rightV, double RF_frequency, double timeCurrent) {
    return solver_init(*(Oracle *)h, leftV, rightV, RF_frequency, timeCurrent);
}
// species: mass in kg, charge in C.  Returns species index.
int orc_add_species(void *h, double mass, double q, int N_p_per_thread, int finalIdx) {
    Oracle *o = (Oracle *)h;
    o->plist.emplace_back();
    species_init(o->plist.back(), o->world, mass, q, 1.0, N_p_per_thread, finalIdx);
    return (int)o->plist.size() - 1;
}
// initialize_n_ave  mod_particle.f90:76-82
void orc_initialize_n_ave(void *h, int sp, double n_ave) {
    Oracle *o = (Oracle *)h; Species &s = o->plist[sp];
    s.w_p = n_ave * o->world.L_domain / (double)s.sumNp();
    s.q_times_wp = s.q * s.w_p;
}
void orc_set_weight(void *h, int sp, double w_p) { Species &s = ((Oracle *)h)->plist[sp]; s.w_p = w_p; s.q_times_wp = s.q * w_p; }
double orc_get_weight(void *h, int sp) { return ((Oracle *)h)->plist[sp].w_p; }
// loader with per-slab ran2 seeds (irand has T entries, updated in place) -- mod_particle.f90:352-369
void orc_load_particles(void *h, int sp, int32_t *irand, double T_eV, double alpha, int distType, double v_drift) {
    Oracle *o = (Oracle *)h; Species &s = o->plist[sp];
    std::vector<int32_t> ir(o->world.T + 1);
    for (int t = 1; t <= o->world.T; ++t) ir[t] = irand[t - 1];
    switch (distType) {
    case 0: initializeRandUniform(s, o->world, ir.data()); break;
    case 1: initializeRandTrig(s, o->world, ir.data(), alpha, false); break;
    case 2: initializeRandTrig(s, o->world, ir.data(), alpha, true); break;
    }
    generate3DMaxwellian(s, o->world, T_eV, ir.data(), (sp == 0) ? alpha : 1.236, distType, v_drift);
    for (int t = 1; t <= o->world.T; ++t) irand[t - 1] = ir[t];
}
double *orc_phase_space(void *h, int sp) { return ((Oracle *)h)->plist[sp].ps.data(); }   // (4, finalIdx, T) Fortran order
int orc_final_idx(void *h, int sp) { return ((Oracle *)h)->plist[sp].finalIdx; }
void orc_get_Np(void *h, int sp, int *N_p) { Species &s = ((Oracle *)h)->plist[sp]; for (int t = 1; t <= s.T; ++t) N_p[t - 1] = s.N_p[t]; }
void orc_set_Np(void *h, int sp, const int *N_p) { Species &s = ((Oracle *)h)->plist[sp]; for (int t = 1; t <= s.T; ++t) s.N_p[t] = N_p[t - 1]; }
void orc_get_counters(void *h, int sp, int *delIdx, int *refIdx, double *wallLoss, double *energyLoss, double *momentumLoss) {
    Species &s = ((Oracle *)h)->plist[sp];
    for (int t = 1; t <= s.T; ++t) {
        delIdx[t - 1] = s.delIdx[t]; refIdx[t - 1] = s.refIdx[t];
        for (int sd = 1; sd <= 2; ++sd) {
            wallLoss[(sd - 1) + 2 * (t - 1)] = s.WL(sd, t);
            energyLoss[(sd - 1) + 2 * (t - 1)] = s.EL(sd, t);
            momentumLoss[(sd - 1) + 2 * (t - 1)] = s.ML(sd, t);
        }
    }
}
void orc_get_ref_record(void *h, int sp, int t /*1-based*/, int *out, int n) {
    Species &s = ((Oracle *)h)->plist[sp];
    for (int k = 1; k <= n && k <= s.refRecLen; ++k) out[k - 1] = s.refRecordIdx[(size_t)k + (size_t)(s.refRecLen + 1) * t];
}
void orc_get_move_stats(void *h, int sp, double *numSubStepsAve, double *numFuncEvalAve, long long *substepsMove, long long *funcEvalsMove,
                        long long *substepsDeposit) {
    Species &s = ((Oracle *)h)->plist[sp];
    *numSubStepsAve = s.numSubStepsAve; *numFuncEvalAve = s.numFuncEvalAve;
    *substepsMove = s.substepsMove; *funcEvalsMove = s.funcEvalsMove; *substepsDeposit = s.substepsDeposit;
}
void orc_config_solver(void *h, double eps_a, double eps_r, int solverType, int m_Anderson, double Beta_k, int maxIter) {
    Oracle *o = (Oracle *)h; NonLin &nl = o->nl;
    nl.eps_a = eps_a; nl.eps_r = eps_r; nl.solverType = solverType; nl.m_Anderson = m_Anderson; nl.Beta_k = Beta_k; nl.maxIter = maxIter;
    const int Nx = o->world.Nx;
    nl.Residual_k.assign((size_t)Nx * (m_Anderson + 1), 0.0);
    nl.phi_k.assign((size_t)Nx * (m_Anderson + 1), 0.0);
    nl.fitMat.assign((size_t)Nx * std::max(1, m_Anderson), 0.0);
}

void orc_deposit_rho(void *h) { depositRho(*(Oracle *)h); }
void orc_solve_tridiag_poisson(void *h, double t) { solve_tridiag_Poisson(*(Oracle *)h, t); }
void orc_solve_initial_potential(void *h, double t) { depositRho(*(Oracle *)h); solve_tridiag_Poisson(*(Oracle *)h, t); }
void orc_make_const_source_term(void *h) { makeConstSourceTerm(*(Oracle *)h); }
void orc_make_half_time_efield(void *h) { Oracle *o = (Oracle *)h; makeHalfTimeEField(*o, &o->plist[0].WS(0, 1)); }
void orc_deposit_J(void *h, double del_t) { depositJ(*(Oracle *)h, del_t); }
void orc_solve_tridiag_ampere(void *h, double del_t) { solve_tridiag_Ampere(*(Oracle *)h, del_t); }
void orc_move_particles(void *h, double del_t) { moveParticles(*(Oracle *)h, del_t); }
void orc_reset_voltage(void *h) { resetVoltage(*(Oracle *)h); }
void orc_set_rf_voltage(void *h, double t) { setRFVoltage(*(Oracle *)h, t); }
int orc_solve_potential(void *h, double del_t, double *remainDel_t, double *currDel_t, double timeCurrent) {
    return solvePotential(*(Oracle *)h, del_t, *remainDel_t, *currDel_t, timeCurrent);
}
void orc_solve_div_ampere_anderson(void *h, double del_t) { solveDivAmpereAnderson(*(Oracle *)h, del_t); }

// field getters/setters (0-based out arrays). which: 0 phi, 1 phi_f, 2 J, 3 rho, 4 EField, 5 a_tri, 6 b_tri, 7 c_tri
static dvec *field_ptr(Oracle *o, int which, int *n) {
    Solver &s = o->solver; const int Nx = o->world.Nx, Nh = o->world.Nh;
    switch (which) {
    case 0: *n = Nx; return &s.phi;
    case 1: *n = Nx; return &s.phi_f;
    case 2: *n = Nh; return &s.J;
    case 3: *n = Nx; return &s.rho;
    case 4: *n = Nh; return &s.EField;
    case 5: *n = Nx - 1; return &s.a_tri;
    case 6: *n = Nx; return &s.b_tri;
    case 7: *n = Nx - 1; return &s.c_tri;
    case 8: *n = Nh; return &s.EField_f;
    case 9: *n = Nh; return &s.EField_half;
    }
    *n = 0; return nullptr;
}
int orc_get_field(void *h, int which, double *out) {
    int n; dvec *v = field_ptr((Oracle *)h, which, &n);
    if (!v) return 0;
    for (int i = 1; i <= n; ++i) out[i - 1] = (*v)[i];
    return n;
}
int orc_set_field(void *h, int which, const double *in) {
    int n; dvec *v = field_ptr((Oracle *)h, which, &n);
    if (!v) return 0;
    for (int i = 1; i <= n; ++i) (*v)[i] = in[i - 1];
    return n;
}
// CIC wall scalars: boundPhi(2), sourceTermVals(2), RF_voltage
void orc_get_cic_walls(void *h, double *out5) {
    Solver &s = ((Oracle *)h)->solver;
    out5[0] = s.boundPhi[1]; out5[1] = s.boundPhi[2]; out5[2] = s.sourceTermVals[1]; out5[3] = s.sourceTermVals[2]; out5[4] = s.RF_voltage;
}
void orc_get_rf(void *h, double *rad_freq, double *half_amp, int *rf_bool, int *rf_indx) {
    Solver &s = ((Oracle *)h)->solver;
    *rad_freq = s.RF_rad_frequency; *half_amp = s.RF_half_amplitude; *rf_bool = s.RF_bool ? 1 : 0; *rf_indx = s.RF_indx;
}
void orc_get_particle_time_step(void *h, int sp, double *out /*Nx*/) {
    Oracle *o = (Oracle *)h; const int Nx = o->world.Nx;
    for (int i = 1; i <= Nx; ++i) out[i - 1] = o->particleTimeStep[(size_t)i + (size_t)(Nx + 2) * sp];
}

// diagnostics (main.f90:58-114)
double orc_total_PE(void *h, int future) { return getTotalPE(*(Oracle *)h, future != 0); }
double orc_total_KE(void *h, int sp) { return getTotalKE(((Oracle *)h)->plist[sp]); }
double orc_KE_ave(void *h, int sp) { return getKEAve(((Oracle *)h)->plist[sp]); }
void orc_total_momentum(void *h, int sp, double *out3) { getTotalMomentum(((Oracle *)h)->plist[sp], out3); }
double orc_JdotE(void *h, double del_t) { return getJdotE(*(Oracle *)h, del_t); }
double orc_energy_from_boundary(void *h, double del_t) { return getEnergyFromBoundary(*(Oracle *)h, del_t); }
double orc_gauss_error(void *h) { return getError_tridiag_Poisson(*(Oracle *)h); }
double orc_charge_error(void *h, const double *rho_i /*0-based Nx*/, double del_t) {
    Oracle *o = (Oracle *)h; dvec r(o->world.Nx + 2, 0.0);
    for (int i = 1; i <= o->world.Nx; ++i) r[i] = rho_i[i - 1];
    return getChargeContinuityError(*o, r.data(), del_t);
}
double orc_lost_energy_sum(void *h, int sp) {   // SUM(energyLoss) as used in main.f90:83
    Species &s = ((Oracle *)h)->plist[sp]; double r = 0.0;
    for (int t = 1; t <= s.T; ++t) for (int sd = 1; sd <= 2; ++sd) r += s.EL(sd, t);
    return r;
}
void orc_get_solver_stats(void *h, int *iterNumPicard, int *iterNumAdaptiveSteps, int *amountTimeSplits, long long *sweeps, long long *substeps,
                          double *moverSeconds, double *potSeconds) {
    NonLin &nl = ((Oracle *)h)->nl;
    *iterNumPicard = nl.iterNumPicard; *iterNumAdaptiveSteps = nl.iterNumAdaptiveSteps; *amountTimeSplits = nl.amountTimeSplits;
    *sweeps = nl.sweeps; *substeps = nl.substeps; *moverSeconds = nl.moverSeconds; *potSeconds = nl.potSeconds;
}
int orc_get_norm_history(void *h, double *out, int nmax) {
    NonLin &nl = ((Oracle *)h)->nl; int n = (int)nl.normHist.size();
    for (int i = 0; i < n && i < nmax; ++i) out[i] = nl.normHist[i];
    return n;
}
void orc_solve_normal_equation(const double *A, const double *b, int m, int n, double *out) { solveNormalEquation(A, b, m, n, out); }
// ---- null-collision MCC ------------------------------------------------------------------------------------------
// targetParticle_constructor mod_targetParticle.f90:27-37 (mass in kg)
int orc_add_target(void *h, double mass, double density, double temperature) {
    Oracle *o = (Oracle *)h;
    TargetParticle t; t.mass = mass; t.density = density; t.temperature = temperature;
    t.v_therm = std::sqrt(t.temperature * k_B / t.mass);
    o->targets.push_back(t);
    return (int)o->targets.size() - 1;
}
// one reactant pair = one nullCollision (readNullCollisionInputs :821-905 tail: masses -> red_mass, sumMass; constructor).
// sp/target are 0-based here; productsIndx (3, nc) holds 1-based species numbers as the reference (0 = unused / neutral slot).
int orc_add_null_collision(void *h, int sp, int target, int numberCollisions, int lengthArrays, const double *energyArray, const double *sigmaArray,
                           const double *energyThreshold, const int *collisionType, const int *numberProducts, const int *productsIndx,
                           double redMassTripleProducts) {
    Oracle *o = (Oracle *)h;
    const double m1 = o->plist[sp].mass, m2 = o->targets[target].mass;
    const double red_mass = m1 * m2 / (m1 + m2), sumMass = (m1 + m2);
    o->colls.emplace_back();
    nullCollision_init(o->colls.back(), numberCollisions, lengthArrays, red_mass, sumMass, redMassTripleProducts, energyArray, sigmaArray, energyThreshold,
                       collisionType, sp + 1, target + 1, numberProducts, productsIndx);
    return (int)o->colls.size() - 1;
}
int orc_generate_collision(void *h, int set, int32_t *irand, double del_t) {
    Oracle *o = (Oracle *)h;
    std::vector<int32_t> ir(o->world.T + 1);
    for (int t = 1; t <= o->world.T; ++t) ir[t] = irand[t - 1];
    const int rc = generateCollision(*o, o->colls[set], ir.data(), del_t);
    for (int t = 1; t <= o->world.T; ++t) irand[t - 1] = ir[t];
    return rc;
}
// scalars7 = sigmaVMax, reducedMass, sumMass, reducedMassIonization, minEnergy, maxEnergy, P_null of the last call
void orc_get_collision_scalars(void *h, int set, double *scalars7) {
    NullCollision &nc = ((Oracle *)h)->colls[set];
    scalars7[0] = nc.sigmaVMax; scalars7[1] = nc.reducedMass; scalars7[2] = nc.sumMass; scalars7[3] = nc.reducedMassIonization;
    scalars7[4] = nc.minEnergy; scalars7[5] = nc.maxEnergy; scalars7[6] = nc.P_null_last;
}
// the collision tables in the constructor's (sorted) order; sigma as (lengthArrays, numberCollisions) column-major
void orc_get_collision_tables(void *h, int set, double *energyThreshold, int *collisionType, double *sigmaArray) {
    NullCollision &nc = ((Oracle *)h)->colls[set];
    for (int i = 1; i <= nc.numberCollisions; ++i) {
        energyThreshold[i - 1] = nc.energyThreshold[i]; collisionType[i - 1] = nc.collisionType[i];
        for (int j = 1; j <= nc.lengthArrays; ++j) sigmaArray[(size_t)(j - 1) + (size_t)nc.lengthArrays * (i - 1)] = nc.SA(j, i);
    }
}
void orc_get_collision_totals(void *h, int set, long long *totalAmountCollisions, double *totalEnergyLoss, double *totalIncidentEnergy,
                              long long *totalNumberCollidableParticles) {
    NullCollision &nc = ((Oracle *)h)->colls[set];
    for (int i = 1; i <= nc.numberCollisions; ++i) {
        totalAmountCollisions[i - 1] = nc.totalAmountCollisions[i]; totalEnergyLoss[i - 1] = nc.totalEnergyLoss[i];
        totalIncidentEnergy[i - 1] = nc.totalIncidentEnergy[i];
    }
    *totalNumberCollidableParticles = nc.totalNumberCollidableParticles;
}
void orc_reset_collision_totals(void *h, int set) {
    NullCollision &nc = ((Oracle *)h)->colls[set];
    std::fill(nc.totalAmountCollisions.begin(), nc.totalAmountCollisions.end(), 0LL);
    std::fill(nc.totalEnergyLoss.begin(), nc.totalEnergyLoss.end(), 0.0);
    std::fill(nc.totalIncidentEnergy.begin(), nc.totalIncidentEnergy.end(), 0.0);
    nc.totalNumberCollidableParticles = 0;
}
// numToCollide = N_p, the epilogue of moveParticles (mod_particleMover.f90:263), for callers that collide without moving
void orc_set_num_to_collide(void *h, int sp) { Species &s = ((Oracle *)h)->plist[sp]; for (int t = 1; t <= s.T; ++t) s.numToCollide[t] = s.N_p[t]; }
void orc_index_sort(int n, const double *arr, int *indx) {   // 0-based in/out wrapper of indexSortArray
    dvec a(n + 1); ivec ix(n + 1);
    for (int i = 0; i < n; ++i) a[i + 1] = arr[i];
    indexSortArray(n, a.data(), ix.data());
    for (int i = 0; i < n; ++i) indx[i] = ix[i + 1] - 1;
}
double orc_ran2(int32_t *irand) { return ran2(*irand); }
double orc_randPCG(uint64_t *state) { return randPCG(*state); }
void orc_species_scalars(void *h, int sp, double *out5) {
    Species &s = ((Oracle *)h)->plist[sp];
    out5[0] = s.mass; out5[1] = s.q; out5[2] = s.w_p; out5[3] = s.q_over_m; out5[4] = s.q_times_wp;
}
int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}


// ---- NGP_Periodic (E-field form) ----
int orc_p_init_solver(void *h, double leftV, double rightV, double RF_frequency) { return P_solver_init(*(Oracle *)h, leftV, rightV, RF_frequency); }
void orc_p_deposit_rho(void *h) { P_depositRho(*(Oracle *)h); }
void orc_p_solve_tridiag_poisson(void *h, double t) { P_solve_tridiag_Poisson(*(Oracle *)h, t); }
void orc_p_deposit_J(void *h, double del_t) { P_sweep<false>(*(Oracle *)h, del_t); }
void orc_p_move_particles(void *h, double del_t) { P_sweep<true>(*(Oracle *)h, del_t); }
void orc_p_solve_tridiag_ampere(void *h, double del_t) { P_solve_tridiag_Ampere(*(Oracle *)h, del_t); }
void orc_p_reset_voltage(void *h) { Oracle *o = (Oracle *)h; for (int i = 1; i <= o->world.Nh; ++i) o->solver.EField[i] = o->solver.EField_f[i]; }
int orc_p_solve_potential(void *h, double del_t, double *remainDel_t, double *currDel_t) { return P_solvePotential(*(Oracle *)h, del_t, *remainDel_t, *currDel_t); }
void orc_p_solve_div_ampere_anderson(void *h, double del_t) { P_solveDivAmpereAnderson(*(Oracle *)h, del_t); }
double orc_p_total_PE(void *h, int future) { return P_getTotalPE(*(Oracle *)h, future != 0); }
double orc_p_gauss_error(void *h) { return P_gaussError(*(Oracle *)h); }
double orc_p_charge_error(void *h, const double *rho_i, double del_t) {
    Oracle *o = (Oracle *)h;
    dvec r(o->world.Nx + 2, 0.0);
    for (int i = 1; i <= o->world.Nx; ++i) r[i] = rho_i[i - 1];
    return P_chargeError(*o, r.data(), del_t);
}
double orc_p_aveJ(void *h) { return ((Oracle *)h)->solver.aveJ; }
void orc_reset_sweep_counters(void *h) { Oracle *o = (Oracle *)h; o->nl.sweeps = 0; o->nl.substeps = 0; }

// The Anderson loop of solveDivAmpereAnderson on a caller-supplied map x <- G(x) (known-answer tests: the reference's own
// Serial/TestAndersonAcceleration/main.f90 problem).  Returns iterNumPicard as the reference would store it (maxIter on stagnation,
// 0 when the loop is exhausted); the residual norms go to normHist.
int orc_anderson_fixed_point(void (*G)(int n, const double *x, double *gx, void *user), void *user, int n, double *x, int m_Anderson, double Beta_k,
                             double eps_a, double eps_r, int maxIter, int periodicTree, double *normHist, int histCap, int *nHist) {
    NonLin nl;
    nl.m_Anderson = m_Anderson; nl.Beta_k = Beta_k; nl.eps_a = eps_a; nl.eps_r = eps_r; nl.maxIter = maxIter; nl.iterNumPicard = 0;
    dvec xv(n + 1, 0.0), gx(n);
    for (int i = 0; i < n; ++i) xv[i + 1] = x[i];
    auto map = [&]() { G(n, xv.data() + 1, gx.data(), user); for (int i = 0; i < n; ++i) xv[i + 1] = gx[i]; };
    if (periodicTree) andersonLoop(nl, n, xv.data(), 2, 1, 0, map, []() {});
    else andersonLoop(nl, n, xv.data(), 3, 3, 1, map, []() {});
    for (int i = 0; i < n; ++i) x[i] = xv[i + 1];
    if (nHist) *nHist = (int)nl.normHist.size();
    for (int i = 0; i < (int)nl.normHist.size() && i < histCap; ++i) normHist[i] = nl.normHist[i];
    return nl.iterNumPicard;
}
void orc_local_temperature(void *h, int sp, int numCells, double *EHist) {
    Oracle *o = (Oracle *)h;
    localTemperature(o->plist[sp], o->world.Nx, numCells, EHist);
}
// mod_particleInjection.f90: irand is the caller's per-thread ran2 state (T entries), energyAdd gets SUM(energyAddColl) of this call
int orc_add_maxwellian_lost_particles(void *h, double T_e, double T_i, int32_t *irand, double *energyAdd) {
    Oracle *o = (Oracle *)h; const int T = o->world.T;
    std::vector<int32_t> ir(T + 1); dvec en(T + 1, 0.0);
    for (int t = 1; t <= T; ++t) ir[t] = irand[t - 1];
    const int rc = addMaxwellianLostParticles(*o, T_e, T_i, ir.data(), en.data());
    double sm = 0.0; for (int t = 1; t <= T; ++t) { irand[t - 1] = ir[t]; sm += en[t]; }
    *energyAdd = sm;
    return rc;
}
void orc_reflux_particles(void *h, double T_e, double T_i, int32_t *irand, double *energyAdd) {
    Oracle *o = (Oracle *)h; const int T = o->world.T;
    std::vector<int32_t> ir(T + 1); dvec en(T + 1, 0.0);
    for (int t = 1; t <= T; ++t) ir[t] = irand[t - 1];
    refluxParticles(*o, T_e, T_i, ir.data(), en.data());
    double sm = 0.0; for (int t = 1; t <= T; ++t) { irand[t - 1] = ir[t]; sm += en[t]; }
    *energyAdd = sm;
}
void orc_velocity_extrema(void *h, int sp, double *out3) { velocityExtrema(*(Oracle *)h, sp, out3); }
void orc_velocity_histograms(void *h, int sp, int binNumber, double VMax, double Emin, double EMax, double *VHist, double *EHist, double *E_grid) {
    velocityHistograms(*(Oracle *)h, sp, binNumber, VMax, Emin, EMax, VHist, EHist, E_grid);
}
void orc_load_particle_density(void *h, int reset) { loadParticleDensity(*(Oracle *)h, reset != 0); }
void orc_particle_density_array(void *h, int sp, double *out) { particleDensityArray(*(Oracle *)h, sp, out); }
int orc_inject_at_boundary(void *h, double T_e, double del_t, int numLow, int numHigh, double injectionR, int32_t *irand) {
    Oracle *o = (Oracle *)h; const int T = o->world.T;
    std::vector<int32_t> ir(T + 1);
    for (int t = 1; t <= T; ++t) ir[t] = irand[t - 1];
    const int rc = injectAtBoundary(*o, T_e, del_t, numLow, numHigh, injectionR, ir.data());
    for (int t = 1; t <= T; ++t) irand[t - 1] = ir[t];
    return rc;
}
int orc_inject_uniform_flux(void *h, double T_e, double T_i, int numLow, int numHigh, double injectionR, int32_t *irand) {
    Oracle *o = (Oracle *)h; const int T = o->world.T;
    std::vector<int32_t> ir(T + 1);
    for (int t = 1; t <= T; ++t) ir[t] = irand[t - 1];
    const int rc = injectUniformFlux(*o, T_e, T_i, numLow, numHigh, injectionR, ir.data());
    for (int t = 1; t <= T; ++t) irand[t - 1] = ir[t];
    return rc;
}
void orc_maxwellian_heating(void *h, int sp, double compValue, double T_eV, int32_t *irand) {
    Oracle *o = (Oracle *)h; const int T = o->world.T;
    std::vector<int32_t> ir(T + 1);
    for (int t = 1; t <= T; ++t) ir[t] = irand[t - 1];
    maxwellianHeating(*o, sp, compValue, T_eV, ir.data());
    for (int t = 1; t <= T; ++t) irand[t - 1] = ir[t];
}
}  // extern "C"
