! =============================================================================
! mod_nonLinSolvers_b200.f90 -- drop-in replacement of OpenMP/SharedModules/src/mod_nonLinSolvers.f90 (level-2 integration, INTEGRATION.md 3)
!
! Same module name, same public entities (:26-30) and procedures -- initializeSolver (:41-81), solveDivAmpereAnderson (:85-160),
! adaptiveSolveDivAmpereAnderson (:162-196), solvePotential (:198-215), funcNitsol / jacNitsol (:218-253), solveJFNK (:255-282),
! adaptiveSolveDivAmpereJFNK (:284-314) -- so main.f90 and mod_simulation.f90 keep compiling unchanged.  One implicit step is ONE
! library call, ipic_solve_potential: fields, Anderson history / NITSOL state and particles stay on the device for the whole step;
! the host gets back what it reads after a step (phi, phi_f, J, EField, the moved particles, counters and loss sums).
! Compile together with ipic_b200_c.f90 and mod_particleMover_b200.f90 (which owns the device handle), same -DIPIC_SCHEME.
! =============================================================================
module mod_nonLinSolvers
    use iso_fortran_env, only: int32, int64, real64
    use iso_c_binding
    use constants
    use mod_BasicFunctions
    use mod_particle
    use mod_domain
    use mod_potentialSolver
    use mod_particleMover
    use mod_Scheme
    use ipic_b200_c
    implicit none

    ! the reference's module state (:26-30): NITSOL reaches the program objects through these globals
    type(Domain) :: globalWorld
    type(Particle), allocatable :: globalParticleList(:)
    type(potentialSolver) :: globalSolver
    integer(int32), protected :: iterNumPicard, iterNumParticle, iterNumAdaptiveSteps, amountTimeSplits
    integer(int64), protected :: potTimer, moverTimer
    integer(int32), private :: maxIter, solverType, m_Anderson
    real(real64), private :: Beta_k, eps_r, eps_a
    logical, private :: solverOnDevice = .false.

contains

    subroutine initializeSolver()
        ! reads SolverState.inp exactly as the reference (:41-81); the history arrays / NITSOL work space are allocated on the device
        integer(int32) :: io
        print *, "Reading non-linear solver inputs:"
        if (.not. restartBool) then
            open(10,file='../InputData/SolverState.inp', IOSTAT=io)
        else
            open(10,file=restartDirectory//'/InputData/SolverState.inp', IOSTAT=io)
        end if
        read(10, *, IOSTAT = io) eps_a
        read(10, *, IOSTAT = io) eps_r
        read(10, *, IOSTAT = io) solverType
        read(10, *, IOSTAT = io) m_Anderson
        read(10, *, IOSTAT = io) Beta_k
        read(10, *, IOSTAT = io) maxIter
        close(10)
        print *, 'Absolute error:', eps_a
        print *, "Relative error:", eps_r
        select case (solverType)
        case(0)
            print *, "Solver type is: Anderson Acceleration"
        case(1)
            print *, "Solver type is: JFNK"
        case default
            print *, "Solver type incorrect!"
            stop
        end select
        print *, "Maximum iteration number:", maxIter
        print *, "Anderson number m is:", m_Anderson
        print *, "Relaxation parameter is:", Beta_k
        print *, ""
        potTimer = 0
        moverTimer = 0
        solverOnDevice = .false.
    end subroutine initializeSolver

    subroutine ipic_solver_attach(solver, particleList, world)
        ! first use: create the device context (shared with mod_particleMover) and hand it the SolverState.inp values
        type(potentialSolver), intent(in) :: solver
        type(Particle), intent(in) :: particleList(:)
        type(Domain), intent(in) :: world
        call ipic_attach(solver, particleList, world)
        if (.not. solverOnDevice) then
            call ipic_check(ipic_initialize_solver(ipic_h, eps_a, eps_r, int(solverType, c_int), int(m_Anderson, c_int), Beta_k, &
                int(maxIter, c_int)), 'ipic_initialize_solver')
            solverOnDevice = .true.
        end if
    end subroutine ipic_solver_attach

    subroutine ipic_pull_step(solver, particleList)
        ! what the host reads after a step: main.f90:75-114 (energies, charge / Gauss errors), mod_simulation.f90:157-285
        type(potentialSolver), intent(in out) :: solver
        type(Particle), intent(in out) :: particleList(:)
        type(ipic_solver_stats) :: st
        call ipic_check(ipic_get_field(ipic_h, IPIC_F_PHI, solver%phi, NumberXNodes), 'get phi')
        call ipic_check(ipic_get_field(ipic_h, IPIC_F_PHI_F, solver%phi_f, NumberXNodes), 'get phi_f')
        call ipic_check(ipic_get_field(ipic_h, IPIC_F_J, solver%J, NumberXHalfNodes), 'get J')
        call ipic_check(ipic_get_field(ipic_h, IPIC_F_EFIELD, solver%EField, NumberXHalfNodes), 'get EField')
        call ipic_pull_particles(particleList)
        particlesOnDevice = particlesStayOnDevice      ! level 2: collisions / injection run on the host before the next step
        call ipic_check(ipic_get_solver_stats(ipic_h, st), 'ipic_get_solver_stats')
        iterNumPicard = st%iterNumPicard
        iterNumAdaptiveSteps = st%iterNumAdaptiveSteps
        amountTimeSplits = st%amountTimeSplits
        ! the reference accumulates system_clock ticks (:102-131); here: device seconds in the same unit
        moverTimer = int(st%moverSeconds * 1.0d9, int64)
        potTimer = int(st%potSeconds * 1.0d9, int64)
    end subroutine ipic_pull_step

    subroutine solvePotential(solver, particleList, world, del_t, remainDel_t, currDel_t, timeCurrent)
        ! :198-215 -- resetVoltage, then the adaptive Anderson / JFNK solve with step halving, then moveParticles: all inside the library
        type(potentialSolver), intent(in out) :: solver
        type(Particle), intent(in out) :: particleList(numberChargedParticles)
        type(Domain), intent(in) :: world
        real(real64), intent(in) :: del_t, timeCurrent
        real(real64), intent(in out) :: remainDel_t, currDel_t
        call ipic_solver_attach(solver, particleList, world)
        call ipic_push_state(solver, particleList)
        call ipic_check(ipic_solve_potential(ipic_h, del_t, remainDel_t, currDel_t, timeCurrent), 'ipic_solve_potential')
        call ipic_pull_step(solver, particleList)
#if IPIC_SCHEME == 1
        ! keep the host copy of the CIC wall state in step with the device (setRFVoltage :458-470, resetVoltage :472-478)
        if (solver%RF_bool) then
            solver%boundPhi(solver%RF_indx) = solver%RF_voltage
            call solver%setRFVoltage(world, timeCurrent + currDel_t)
        end if
#endif
    end subroutine solvePotential

    subroutine solveDivAmpereAnderson(solver, particleList, world, del_t)
        ! :85-160 -- one Anderson solve at a fixed del_t, no step halving (iterNumPicard = maxIter reports a failed solve, as in the reference)
        type(potentialSolver), intent(in out) :: solver
        type(Particle), intent(in out) :: particleList(numberChargedParticles)
        type(Domain), intent(in) :: world
        real(real64), intent(in) :: del_t
        call ipic_solver_attach(solver, particleList, world)
        call ipic_push_state(solver, particleList)
        call ipic_check(ipic_solve_div_ampere_anderson(ipic_h, del_t), 'ipic_solve_div_ampere_anderson')
        call ipic_pull_step(solver, particleList)
    end subroutine solveDivAmpereAnderson

    subroutine adaptiveSolveDivAmpereAnderson(solver, particleList, world, del_t, remainDel_t, currDel_t, timeCurrent)
        ! :162-196 -- the library's entry also copies phi_f to phi first (resetVoltage); the host has already done so, which makes it a no-op
        type(potentialSolver), intent(in out) :: solver
        type(Particle), intent(in out) :: particleList(numberChargedParticles)
        type(Domain), intent(in) :: world
        real(real64), intent(in) :: del_t, timeCurrent
        real(real64), intent(in out) :: remainDel_t, currDel_t
        call solvePotential(solver, particleList, world, del_t, remainDel_t, currDel_t, timeCurrent)
    end subroutine adaptiveSolveDivAmpereAnderson

    ! ---- JFNK: NITSOL callbacks with the reference's abstract interfaces (mod_nitsol.f90:66-80) -------------------------------------
    subroutine funcNitsol(n, xcur, fcur, rpar, ipar, itrmf)
        ! :218-239 -- fcur = xcur - G(xcur): one particle sweep + field solve on the device
        integer, intent(in) :: n
        integer, intent(in out) :: itrmf, ipar(*)
        real(real64), intent(in) :: xcur(n)
        real(real64), intent(in out) :: rpar(*), fcur(n)
        call ipic_check(ipic_func_nitsol(ipic_h, int(n, c_int32_t), xcur, fcur, rpar(1)), 'ipic_func_nitsol')
        itrmf = 0
    end subroutine funcNitsol

    subroutine jacNitsol(n, xcur, fcur, ijob, v, z, rpar, ipar, itrmjv)
        ! :241-253 -- never called: NITSOL runs with the finite-difference J.v (input(3) = 0)
        integer, intent(in) :: ijob, n
        integer, intent(in out) :: itrmjv, ipar(*)
        real(real64), intent(in) :: fcur(n), v(n), xcur(n)
        real(real64), intent(in out) :: rpar(*), z(n)
        itrmjv = 0
    end subroutine jacNitsol

    subroutine solveJFNK(del_t)
        ! :255-282 -- the library runs its own NITSOL (csrc/nitsol.hpp) on the device residual, then moves the particles on success
        real(real64), intent(in) :: del_t
        call ipic_solver_attach(globalSolver, globalParticleList, globalWorld)
        call ipic_push_state(globalSolver, globalParticleList)
        call ipic_check(ipic_solve_jfnk(ipic_h, del_t), 'ipic_solve_jfnk')
        call ipic_pull_step(globalSolver, globalParticleList)
    end subroutine solveJFNK

    subroutine adaptiveSolveDivAmpereJFNK(del_t, remainDel_t, currDel_t, timeCurrent)
        ! :284-314 -- solverType = 1 makes ipic_solve_potential take this path
        real(real64), intent(in) :: del_t, timeCurrent
        real(real64), intent(in out) :: remainDel_t, currDel_t
        call solvePotential(globalSolver, globalParticleList, globalWorld, del_t, remainDel_t, currDel_t, timeCurrent)
    end subroutine adaptiveSolveDivAmpereJFNK

end module mod_nonLinSolvers
