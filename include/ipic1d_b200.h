/* =============================================================================
 * ipic1d_b200.h -- C ABI of the B200-native implicit-PIC inner loop
 *
 * Drop-in boundary for the hot path of nsavard01/ImplicitPic1D (trees
 * OpenMP/NGP and OpenMP/CIC): particle mover, J/rho deposition, field solve
 * and the nonlinear (Anderson / JFNK-residual) iteration.  The reference has
 * no FFI boundary of its own: callers reach the path through Fortran module
 * procedures.  Every entry point below names the module procedure it replaces
 * (file:line under /root/reference/OpenMP); fortran/ipic_b200_shim.f90 shows
 * the iso_c_binding interface a maintainer adds.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every call returns an int
 *     status (0 = ok); no exceptions cross the boundary.  The Fortran shim maps
 *     nonzero to `print *, ...; stop`, the reference's only error behaviour.
 *   - arrays are 0-based C arrays holding the reference's 1-based arrays in
 *     the same order (element k here = Fortran element k+1).
 *   - all floating point is IEEE fp64, indices are int32, as in the reference.
 *   - one host thread per context; all kernels of a context run on one CUDA
 *     stream; calls that return host data are synchronous.
 *   - particles live on the device between calls (per-species SoA l, v_x, v_y,
 *     v_z).  Host copies move only through ipic_upload_particles /
 *     ipic_download_particles.
 * ========================================================================== */
#ifndef IPIC1D_B200_H
#define IPIC1D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ipic_ctx *ipic_handle;

/* scheme = which reference tree's semantics are used */
#define IPIC_SCHEME_NGP 0 /* OpenMP/NGP : phi on N_x nodes, J/E on N_h = N_x-1 cells */
#define IPIC_SCHEME_CIC 1 /* OpenMP/CIC : phi on N_x centres, J/E on N_h = N_x+1 faces */
#define IPIC_SCHEME_NGP_PERIODIC 2 /* OpenMP/NGP_Periodic : E-field form; the nonlinear unknown is EField_f (N_h), explicit Ampere update,
                                      cell-local mover arithmetic (mod_particleMover.f90:18-238, mod_potentialSolver.f90:127-176) */

/* status codes */
#define IPIC_OK 0
#define IPIC_ERR_ARG 1         /* bad argument */
#define IPIC_ERR_CUDA 2        /* CUDA runtime error (see ipic_last_error) */
#define IPIC_ERR_STOP 3        /* a condition on which the reference executes `stop` */
#define IPIC_ERR_NOT_CONVERGED 4
#define IPIC_ERR_UNSUPPORTED 5
#define IPIC_ERR_NCCL 6

/* field selectors for ipic_get_field / ipic_set_field */
#define IPIC_F_PHI 0    /* solver%phi    (N_x) */
#define IPIC_F_PHI_F 1  /* solver%phi_f  (N_x) */
#define IPIC_F_J 2      /* solver%J      (N_h) */
#define IPIC_F_RHO 3    /* solver%rho    (N_x) */
#define IPIC_F_EFIELD 4 /* solver%EField (N_h) */
#define IPIC_F_A_TRI 5  /* solver%a_tri  (N_x-1) */
#define IPIC_F_B_TRI 6  /* solver%b_tri  (N_x) */
#define IPIC_F_C_TRI 7  /* solver%c_tri  (N_x-1) */
#define IPIC_F_EFIELD_F 8    /* solver%EField_f    (N_h)  NGP_Periodic */
#define IPIC_F_EFIELD_HALF 9 /* solver%EField_half (N_h)  NGP_Periodic */

/* linear-solve modes for the tridiagonal Ampere/Poisson systems */
#define IPIC_TRIDIAG_PCR 0    /* single-block parallel cyclic reduction, cached level coefficients */
#define IPIC_TRIDIAG_THOMAS 1 /* sequential Thomas sweep in the reference's operation order (parity mode) */
#define IPIC_TRIDIAG_SCAN 2   /* x = T^-1 d through the generators of the constant matrix's inverse, (T^-1)_ij = u_i v_j / W for i <= j: two
                               * block scans instead of log2(N_x) reduction levels.  Default where the matrix allows it (bounded decks) */

const char *ipic_version(void);
const char *ipic_last_error(void);
int ipic_device_count(int *n);

/* ---- world + solver construction -------------------------------------------
 * replaces: type(Domain) contents built by readWorld        NGP/src/mod_domain.f90:347-391, CIC/src/mod_domain.f90:332-376
 *           potentialSolver_constructor + construct_diagMatrix
 *                                                           NGP/src/mod_potentialSolver.f90:50-128, CIC/src/mod_potentialSolver.f90:49-126
 * grid: N_x doubles; dx_dl: N_h (NGP) or N_x (CIC) doubles; boundaryConditions: N_x (NGP) or N_h (CIC) int32.
 * The caller (Fortran host) keeps owning its Domain; this copies the tables to the device. */
int ipic_create(ipic_handle *out, int scheme, int32_t NumberXNodes, const double *grid, const double *dx_dl,
                const int32_t *boundaryConditions, int gridSmoothBool, int device);
int ipic_destroy(ipic_handle h);
int ipic_init_solver(ipic_handle h, double leftVoltage, double rightVoltage, double RF_frequency, double timeCurrent);
int ipic_set_tridiag_mode(ipic_handle h, int mode);

/* ---- particle container ----------------------------------------------------
 * replaces: type(Particle)                                   SharedModules/src/mod_particle.f90:15-74
 * mass [kg], q [C], w_p [1/m^2].  Returns the species index through *species. */
int ipic_add_species(ipic_handle h, double mass, double q, double w_p, int32_t *species);
int ipic_set_weight(ipic_handle h, int32_t species, double w_p);
/* phaseSpace(4, finalIdx, numThread) real64 AoS slabs + N_p(numThread), exactly the reference's arrays
 * (mod_particle.f90:23,65); transposed to device SoA, slab order preserved. */
int ipic_upload_particles(ipic_handle h, int32_t species, const double *phaseSpace, const int32_t *N_p, int32_t numThread,
                          int32_t finalIdx);
int ipic_download_particles(ipic_handle h, int32_t species, double *phaseSpace, int32_t *N_p, int32_t numThread, int32_t finalIdx);
int ipic_particle_count(ipic_handle h, int32_t species, int64_t *n);
/* zero-copy view of the device-resident SoA arrays (valid until the next call that moves, sorts, compacts, uploads or grows
 * the species); for hosts that run their own kernels on the particles (e.g. through __cuda_array_interface__). */
int ipic_particle_device_pointers(ipic_handle h, int32_t species, double **l, double **v_x, double **v_y, double **v_z, int64_t *n);
/* device-side synthetic loader for loads too large to stage through the host (bench): uniform positions,
 * Maxwellian velocities (mod_particle.f90:84-97,156-191 rules) from a counter-based PCG32 stream. */
int ipic_generate_particles(ipic_handle h, int32_t species, int64_t n, double T_eV, double v_drift, uint64_t seed);
/* same with the perturbed loads of readChargedParticleInputs (mod_particle.f90:99-151,172-181): distType 0 uniform, 1 cosine, 2 sine
 * density perturbation of amplitude alpha (rejection sampling) and the matching position-dependent drift. */
int ipic_generate_particles_dist(ipic_handle h, int32_t species, int64_t n, double T_eV, double v_drift, double alpha, int distType,
                                 uint64_t seed);
/* periodic cell sort (stable radix sort on INT(l * subbins) + gather); does not change any deposited quantity.
 * Ordering contract: call it at a step boundary, i.e. BEFORE ipic_move_particles / ipic_solve_potential of the step, never between the
 * move and the calls that consume the move's per-particle records.  A sort permutes the arrays, so it drops refRecordIdx (the list
 * ipic_reflux_particles / ipic_inject_at_boundary / ipic_get_ref_record read: empty afterwards) and resets numToCollide to N_p (the
 * prefix that excluded particles born during the step no longer exists).  At most 2^31 - 1 particles of one species per GPU.
 * A sort also ends the reference's thread-slab identity of the arrays (ipic_upload_particles sets it, ipic_download_particles then returns
 * one slab): from then on ipic_move_particles removes absorbed particles by moving live particles of the tail into their slots (option
 * "fill_compaction") instead of the stable, slab-wise compaction of mod_particleMover.f90:252-256 -- the same particles, another order. */
int ipic_sort_particles(ipic_handle h, int32_t species);

/* ---- field state -----------------------------------------------------------*/
int ipic_get_field(ipic_handle h, int which, double *out, int32_t n);
int ipic_set_field(ipic_handle h, int which, const double *in, int32_t n);

/* ---- the path ---------------------------------------------------------------
 * depositRho            NGP/src/mod_potentialSolver.f90:130-188   CIC/src/mod_potentialSolver.f90:128-173
 * solve_tridiag_Poisson NGP :190-226                              CIC :175-218
 * solveInitialPotential NGP :228-238                              CIC :220-230
 * makeConstSourceTerm   NGP :267-286                              CIC :246-270
 * makeHalfTimeEField    NGP :392-406                              CIC :333-366
 * solve_tridiag_Ampere  NGP :288-330                              CIC :272-301
 * setRFVoltage          NGP :444-453                              CIC :458-470
 * resetVoltage          NGP :455-459                              CIC :472-478
 * depositJ              NGP/src/mod_particleMover.f90:67-157      CIC/src/mod_particleMoverPicard.f90:77-195
 * moveParticles         NGP/src/mod_particleMover.f90:162-270     CIC/src/mod_particleMoverPicard.f90:198-317 */
int ipic_deposit_rho(ipic_handle h);
/* loadParticleDensity (NGP mod_Scheme.f90:68-83, CIC mod_Scheme.f90:107-121) for one species: densities[N_x] receives the sum of the
 * interpolation weights of this call over all ranks (the reference's SUM(part%densities, DIM=2) of a reset load, before w_p); the host
 * adds calls up for time averages and applies WriteParticleDensity's scaling / smoothing / normalisation (O(N_x)). */
int ipic_load_particle_density(ipic_handle h, int32_t species, double *densities);
int ipic_solve_tridiag_poisson(ipic_handle h, double timeCurrent);
int ipic_solve_initial_potential(ipic_handle h, double timeCurrent);
int ipic_make_const_source_term(ipic_handle h);
int ipic_make_half_time_efield(ipic_handle h);
int ipic_solve_tridiag_ampere(ipic_handle h, double del_t);
int ipic_set_rf_voltage(ipic_handle h, double timeFuture);
int ipic_reset_voltage(ipic_handle h);
/* CIC only: the wall state the reference keeps outside phi -- boundPhi(2), RF_voltage, sourceTermVals(2)
 * (CIC mod_potentialSolver.f90:19,69-85; written by setRFVoltage :458-470 and resetVoltage :472-478) -- for a host that owns the solver object */
int ipic_set_cic_walls(ipic_handle h, const double *boundPhi2, double RF_voltage, const double *sourceTermVals2);
int ipic_deposit_J(ipic_handle h, double del_t);
int ipic_move_particles(ipic_handle h, double del_t);

/* per-species outputs of moveParticles (mod_particle.f90:18-27): totals over the reference's thread slabs.
 * wallLoss/energyLoss/momentumLoss are [left,right]; energyLoss is the raw sum of v^2 as in the reference. */
typedef struct {
    int64_t N_p;            /* SUM(N_p) after the move */
    int64_t delIdx;         /* SUM(delIdx): absorbed at Dirichlet/RF walls */
    int64_t refIdx;         /* SUM(refIdx): first reflections at Neumann walls */
    int64_t numSubSteps;    /* executed do-while trips (numerator of numSubStepsAve) */
    int64_t numFuncEvals;   /* CIC Picard iterations (numerator of numFuncEvalAve) */
    double numSubStepsAve;  /* mod_particleMover.f90:265 (single-precision division, as the reference) */
    double numFuncEvalAve;  /* :266 */
    double wallLoss[2];
    double energyLoss[2];
    double momentumLoss[2]; /* accumulates across steps, as the reference (never reset) */
    double accumEnergyLoss[2];
    int64_t accumWallLoss[2];
} ipic_move_stats;
int ipic_get_move_stats(ipic_handle h, int32_t species, ipic_move_stats *out);
/* refRecordIdx (mod_particle.f90:19; written by moveParticles, NGP mod_particleMover.f90:226-231, CIC mod_particleMoverPicard.f90:276-281):
 * the particles first reflected at a Neumann wall during the last ipic_move_particles / converged ipic_solve_potential, as ascending
 * 0-based indices into the species' particle order after compaction (the order ipic_download_particles writes, slab after slab).
 * *n receives the count; at most cap indices are written.  A particle reflected and then absorbed in the same step is listed with the slot
 * of the next surviving particle, as the reference's `i - delIdx` (recorded at the reflection) comes out after its compaction. */
int ipic_get_ref_record(ipic_handle h, int32_t species, int32_t *idx, int64_t cap, int64_t *n);
/* refIdx(numThread) of the same move: how many consecutive entries of the list above belong to each thread slab (the slab the particle
 * was in when it was reflected).  Entry k of slab t is the 0-based global index; its refRecordIdx value is index - (first index of slab t
 * after the compaction) + 1, which is N_p(t) + 1 when the reflected particle was the last of its slab and was absorbed. */
int ipic_get_ref_record_counts(ipic_handle h, int32_t species, int32_t *refIdx, int32_t numThread);
/* executed sub-steps of a species in the last depositJ sweep, or, after ipic_solve_potential, in all depositJ sweeps of that step
 * (instrumentation for the throughput metric) */
int ipic_get_deposit_substeps(ipic_handle h, int32_t species, int64_t *n);

/* ---- nonlinear solvers -------------------------------------------------------
 * initializeSolver               SharedModules/src/mod_nonLinSolvers.f90:41-81
 * solveDivAmpereAnderson         :85-160
 * adaptiveSolveDivAmpereAnderson :162-196
 * solvePotential                 :198-215
 * funcNitsol                     :218-239  (residual callback F(x) = x - G(x) for NITSOL's JFNK) */
int ipic_initialize_solver(ipic_handle h, double eps_a, double eps_r, int solverType, int m_Anderson, double Beta_k, int maxIter);
int ipic_solve_div_ampere_anderson(ipic_handle h, double del_t);
int ipic_solve_potential(ipic_handle h, double del_t, double *remainDel_t, double *currDel_t, double timeCurrent);
int ipic_func_nitsol(ipic_handle h, int32_t n, const double *xcur, double *fcur, double del_t);
/* solveJFNK :255-282 (solverType 1; ipic_solve_potential adds the adaptive step halving of :284-314).
 * info7 = nfe, njve, nrpre, nli, nni, nbt (mod_nitsol.f90:3430-3435) and iterm of the last solve. */
int ipic_solve_jfnk(ipic_handle h, double del_t);
int ipic_get_nitsol_info(ipic_handle h, int32_t *info7);
/* nitsol (SharedModules/src/mod_nitsol.f90:2964-3443) with the settings of initializeNitsol (:128-140) on a HOST residual
 * callback f(n, xcur, fcur, user) -> itrmf: the driver the JFNK path uses, exposed for hosts that want to keep their own
 * residual (and for CPU tests of the solver logic).  maxIter -> input(1), kdmax -> input(4), input(6) = 2*kdmax. */
typedef int (*ipic_residual_fn)(int32_t n, const double *xcur, double *fcur, void *user);
int ipic_nitsol_host(int32_t n, double *x, ipic_residual_fn f, void *user, double abstol, double reltol, double stptol, int32_t maxIter,
                     int32_t kdmax, int32_t *info6, int32_t *iterm);

typedef struct {
    int32_t iterNumPicard, iterNumAdaptiveSteps, amountTimeSplits; /* mod_nonLinSolvers.f90:29 */
    int64_t sweeps;        /* depositJ calls in the last solvePotential */
    int64_t substeps;      /* executed sub-steps in the last solvePotential (all sweeps + final move, all species) */
    double moverSeconds;   /* device time in depositJ/moveParticles kernels (moverTimer analogue) */
    double potSeconds;     /* device time in the field kernels (potTimer analogue) */
    double lastResidualNorm;
} ipic_solver_stats;
int ipic_get_solver_stats(ipic_handle h, ipic_solver_stats *out);
int ipic_get_norm_history(ipic_handle h, double *out, int32_t nmax, int32_t *n);

/* ---- conservation diagnostics (main.f90:58-114) -------------------------------
 * getTotalPE NGP :408-420 CIC :369-413 | getJdotE NGP :422-430 CIC :415-426 | getEnergyFromBoundary NGP :432-442 CIC :428-456
 * getError_tridiag_Poisson NGP :240-265 CIC :232-244 | getChargeContinuityError NGP :358-390 CIC :315-331
 * getTotalKE / getTotalMomentum mod_particle.f90:206-230 */
int ipic_get_total_PE(ipic_handle h, int future, double *out);
int ipic_get_JdotE(ipic_handle h, double del_t, double *out);
int ipic_get_energy_from_boundary(ipic_handle h, double del_t, double *out);
int ipic_get_error_tridiag_poisson(ipic_handle h, double *out);
int ipic_get_charge_continuity_error(ipic_handle h, const double *rho_i, double del_t, double *out);
int ipic_get_total_KE(ipic_handle h, int32_t species, double *out);
/* writeLocalTemperature (mod_particle.f90:251-283) without the file: EHist[j-1] = SUM(v^2) mass / count / 3 / e over the particles with
 * INT(l) == j (j < N_x, j <= numCells), in eV; 0 for an empty cell.  The Python mirror writes the reference's unformatted record. */
int ipic_local_temperature(ipic_handle h, int32_t species, int32_t numCells, double *EHist);
int ipic_get_total_momentum(ipic_handle h, int32_t species, double *out3);
/* EEDF / velocity histograms of solveSimulationFinalAverage (SharedModules/src/mod_simulation.f90), per species, for THIS process's
 * particles (a multi-GPU host combines ranks with max / min / sum, as the reference's OpenMP reduction clauses do over threads):
 *   ipic_velocity_extrema     :437-452  out3 = { VMax = max |v_x|, EMax = max v^2, Emin = min(v^2, 0.1 e 2 / m) }
 *   ipic_velocity_histograms  :480-500  ADDS one pass over the particles to VHist[2 * binNumber] (v_x on linear bins over [-VMax, VMax])
 *                             and EHist[binNumber] (v^2 on logarithmic bins E_grid(i) = Emin exp((i-1) diffE), diffE = (ln EMax - ln Emin) /
 *                             (binNumber - 1)), linear weights on the two neighbouring bins; binNumber 2..2048 (the reference uses 100) */
int ipic_velocity_extrema(ipic_handle h, int32_t species, double *out3);
int ipic_velocity_histograms(ipic_handle h, int32_t species, int32_t binNumber, double VMax, double Emin, double EMax, double *VHist,
                             double *EHist);

/* ---- null-collision Monte-Carlo collisions ---------------------------------------------------------
 * replaces: type(targetParticle) + constructor          SharedModules/src/mod_targetParticle.f90:13-37
 *           type(nullCollision) + constructor           SharedModules/src/mod_NullCollision.f90:18-108
 *           generateCollision                           :110-237  (called once per set per step, mod_simulation.f90:166-168)
 *           elasticExcitCollisionIsotropic :511-562, ionizationCollisionIsotropic :359-444, scatterVector :634-661
 * The arguments of ipic_null_collision_create are those of nullCollision_constructor (:44), arrays in the reference's
 * order: sigmaArray(lengthArrays, numberCollisions) column-major; reactantsIndx = (charged species, target particle) and
 * productsIndx(3, numberCollisions) hold the reference's 1-based numbers (species k here = ipic_add_species index k-1).
 * collisionType: 1 elastic, 2 ionization, 3 excitation, 4 charge exchange, 5 dissociation (:783-792).
 * Random numbers: a counter-based PCG32 stream per (seed, rank, call, collision candidate) instead of the reference's
 * per-thread ran2 states, and the `numberSelected` collision candidates are drawn without replacement through a keyed
 * bijection of [0, numToCollide) instead of the swap-to-end loop (:222-227) -- the same uniform subset statistically,
 * with no serial dependence.  Collision sets that share a primary species draw from ONE such bijection per step, one after the other
 * (the second set's candidates start where the first set's ended, and its numToCollide is the reduced one, :221-231), so a particle is
 * selected at most once per step however many sets name its species; a "step" starts whenever numToCollide is set (by the mover, by
 * ipic_set_num_to_collide, by an upload / load / sort).  Particles keep their array positions; products of ionization are appended in
 * candidate order, so a run is bit-reproducible for a given seed.  Results agree with the reference statistically (moments), not bit-wise. */
#define IPIC_MAX_COLLISIONS 16
int ipic_add_target_particle(ipic_handle h, double mass, double density, double temperature, int32_t *target);
int ipic_null_collision_create(ipic_handle h, int32_t numberCollisions, int32_t lengthArrays, double red_mass, double sumMass,
                               double redMassTripleProducts, const double *energyArray, const double *sigmaArray, const double *energyThreshold,
                               const int32_t *collisionType, const int32_t *reactantsIndx, const int32_t *numberProducts,
                               const int32_t *productsIndx, uint64_t seed, int32_t *collision_set);
int ipic_generate_collision(ipic_handle h, int32_t collision_set, double del_t);
typedef struct {
    int32_t numberCollisions, lengthArrays;
    int32_t collisionType[IPIC_MAX_COLLISIONS];     /* in the constructor's order (descending max sigma*v, :55-86) */
    double energyThreshold[IPIC_MAX_COLLISIONS];
    double sigmaVMax, reducedMass, sumMass, reducedMassIonization, minEnergy, maxEnergy;
    double P_null;                                  /* of the last call */
    int64_t numberSelected;                         /* collision candidates of the last call */
    int64_t ionizations;                            /* particle pairs appended by the last call */
    int64_t totalAmountCollisions[IPIC_MAX_COLLISIONS];
    double totalEnergyLoss[IPIC_MAX_COLLISIONS];    /* raw sums in the reference's units (mass * v^2), :216 */
    double totalIncidentEnergy[IPIC_MAX_COLLISIONS];
    int64_t totalNumberCollidableParticles;
} ipic_collision_stats;
int ipic_get_collision_stats(ipic_handle h, int32_t collision_set, ipic_collision_stats *out);
int ipic_reset_collision_stats(ipic_handle h, int32_t collision_set);
/* numToCollide (mod_particle.f90:17): set to N_p by moveParticles (mod_particleMover.f90:263) and by upload; this sets it
 * explicitly (n < 0: all particles). */
int ipic_set_num_to_collide(ipic_handle h, int32_t species, int64_t n);
/* the keyed bijection used to draw collision candidates: out[k] = image of j0 + k in [0, n).  Host function (no GPU needed). */
int ipic_mcc_permutation(uint64_t n, uint64_t key, uint64_t j0, uint64_t count, uint64_t *out);

/* ---- wall refluxing, particle injection, Maxwellian heating ---------------------------------------------------
 * replaces: SharedModules/src/mod_particleInjection.f90 -- the routines mod_simulation.f90:158-176 calls right after
 * moveParticles, on the particles where the mover left them (device resident).  particleList(1) = species 0 (electrons),
 * particleList(2) = species 1 (ions), as the reference assumes.  Temperatures in eV.  Random numbers come from counter-based
 * PCG32 streams keyed by (seed, rank, call, work item) instead of the per-thread ran2 states: bit-reproducible for a seed.
 *   ipic_add_maxwellian_lost_particles  addMaxwellianLostParticles :16-40: appends SUM(electron delIdx of the last move) electron-ion
 *                                       pairs at one uniform position each; *energyAdd = what the call adds to SUM(energyAddColl)
 *   ipic_reflux_particles               refluxParticles :42-69: re-draws the particles listed in refRecordIdx (see ipic_get_ref_record)
 *                                       from a flux Maxwellian pointing away from the Neumann wall; *energyAdd = new - old energy
 *   ipic_inject_at_boundary             injectAtBoundary :71-130: reflected electrons get a flux Maxwellian, reflected ions the ion sound
 *                                       speed; then floor(numFluxPart) or floor(numFluxPart) + 1 pairs enter within del_t * c_s of the
 *                                       Neumann wall.  numFluxPart = injectionFlux * del_t / w_p for this process's share of the wall
 *                                       (readInjectionInputs :232-236 divides the same number by numThread, one draw per slab; here one
 *                                       draw per call, the same expectation)
 *   ipic_inject_uniform_flux            injectUniformFlux :132-158 (electrons at 20 T_e, as the reference)
 *   ipic_maxwellian_heating             maxwellianHeating :160-177 with compValue = FracFreq * (currDel_t / del_t) / FractionFreqHeating */
int ipic_set_injection_seed(ipic_handle h, uint64_t seed);
int ipic_add_maxwellian_lost_particles(ipic_handle h, double T_e, double T_i, double *energyAdd, int64_t *numAdded);
int ipic_reflux_particles(ipic_handle h, double T_e, double T_i, double *energyAdd);
int ipic_inject_at_boundary(ipic_handle h, double T_e, double T_i, double del_t, double numFluxPart, int64_t *numInjected);
int ipic_inject_uniform_flux(ipic_handle h, double T_e, double T_i, double numFluxPart, int64_t *numInjected);
int ipic_maxwellian_heating(ipic_handle h, int32_t species, double compValue, double T_eV, int64_t *numHeated);

/* ---- multi-GPU: particles sharded over ranks, one all-reduce of the fixed-point J/rho grid per deposit ------
 * (pattern of PIC1D/ExplicitPIC_Cpp/src/potential_solver.cpp:142).  unique_id is the 128-byte ncclUniqueId. */
int ipic_nccl_unique_id(void *out128);
int ipic_comm_init(ipic_handle h, int32_t nranks, int32_t rank, const void *unique_id128);
/* Ranks on ONE node (NVLink / NVSwitch): the all-reduce of the deposition grids runs through peer memory instead of NCCL -- every rank
 * publishes its folded integer grids in a window, raises a flag on all peers, and adds up all ranks' windows itself (integer sums: the
 * same bits on every rank whatever the order; two launches per residual evaluation).  ipic_peer_export returns the 64-byte CUDA IPC
 * handle of this rank's window; the host exchanges the handles (torch.distributed all_gather, MPI_Allgather in a Fortran host) and
 * hands all of them, in rank order, to ipic_peer_import.  Optional: without it (or with option "peer_reduce" = 0) the all-reduce is
 * NCCL's.  A peer that never shows up makes the waiting kernel give up after ~3 s and the next ipic_move_particles return IPIC_ERR_NCCL. */
int ipic_peer_export(ipic_handle h, void *handle64);
int ipic_peer_import(ipic_handle h, int32_t nranks, int32_t rank, const void *handles64_by_rank);

/* ---- tuning / introspection ----------------------------------------------------*/
/* Scheduling knobs of the kernels; none of them changes a result (J, fields, particles and counters are the same bit for bit -- the
 * deposits are integers -- except where noted).  Unknown names return IPIC_ERR_ARG.
 *   stay_min, transit_min, entry_min   lane-count thresholds of the push kernel's register loop / face-to-face loop / entry stage (33 = off)
 *   exit_stage (1)      the entry stage also finishes the lanes that end their step in the cell they enter
 *   sd_runs (1)         ends of runs of traversed cells counted in shared memory per CTA instead of global atomics
 *   tile_order (1), tile_subs (0 = automatic), push_blocks (0 = one CTA per SM)   tile schedule / size / grid of the push kernels
 *   pair_species (0), pair_warps      both species of a two-species plasma in one launch
 *   sort_subbins (0 = automatic)      slices of a cell in the key of ipic_sort_particles
 *   fill_compaction (1) arrays without thread-slab identity (device loader, after a sort): absorbed particles' slots are filled from the tail
 *                       instead of a stable scatter -- the same survivors and loss sums, another particle ORDER
 *   anderson_chunk (6)  nonlinear iterations enqueued ahead of the device-side convergence decision
 *   field_cluster (8)   CTAs of the fused field kernel's thread-block cluster (sums of the normal equations added in another order: last bits)
 *   peer_reduce (1)     J all-reduce through peer memory (0: NCCL);  norm_flag (1)  residual norm through pinned memory
 *   timing (1)          CUDA-event brackets of ipic_get_timing */
int ipic_set_option(ipic_handle h, const char *name, double value);
/* run all kernels of this context on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. the
 * host program's / torch's current stream; NULL restores the context's own stream. */
int ipic_set_stream(ipic_handle h, void *cuda_stream);
/* kernels launched by this context since creation (cub sort passes count as one each) */
int ipic_get_launch_count(ipic_handle h, int64_t *n);
/* page-locked host buffers for phaseSpace so that upload/download run at full PCIe rate; a Fortran host
 * maps them with c_f_pointer.  Plain (pageable) memory works too, only slower. */
int ipic_host_alloc(void **out, uint64_t bytes);
int ipic_host_free(void *p);
/* device time and bracket count of a named section since the last reset (CUDA events on the library's stream; "depositJ_kernels" is the
 * sum of the per-species sweep entries).  The call waits for the stream.  Brackets are folded into the totals as their events complete,
 * so a host that never asks keeps a bounded list (name "__pending": launches = brackets not folded in yet, ms = events parked for
 * re-use; no wait). */
int ipic_get_timing(ipic_handle h, const char *name, double *ms, int64_t *launches);
int ipic_reset_timing(ipic_handle h);
int ipic_synchronize(ipic_handle h);

#ifdef __cplusplus
}
#endif
#endif /* IPIC1D_B200_H */
