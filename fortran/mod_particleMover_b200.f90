! =============================================================================
! mod_particleMover_b200.f90 -- drop-in replacement of
!     OpenMP/NGP/src/mod_particleMover.f90           (scheme 0)
!     OpenMP/CIC/src/mod_particleMoverPicard.f90     (scheme 1; the file the CIC Makefile compiles as mod_particleMover.o)
! Same module name, same public procedures and argument lists (NGP :17,67-74,162-167 ; CIC :67-70,77-84,198-203), so
! mod_nonLinSolvers / mod_simulation / main keep compiling unchanged.  The bodies call the C ABI (include/ipic1d_b200.h).
!
! Residency protocol (level-1 integration, see INTEGRATION.md): the host keeps owning particleList and solver.  Particles are
! uploaded at the first depositJ after the host may have touched them (start of a step: collisions / injection happen on the
! host between steps) and downloaded by moveParticles; phi, phi_f and rho (8 B x N_x each) go up and J (and phi_f) come back on
! every call.  Build with -DIPIC_SCHEME=0 (NGP tree) or 1 (CIC tree).
! =============================================================================
module mod_particleMover
    use iso_fortran_env, only: int32, real64
    use iso_c_binding
    use constants
    use mod_particle
    use mod_domain
    use mod_potentialSolver
    use ipic_b200_c
    implicit none
#ifndef IPIC_SCHEME
#define IPIC_SCHEME 0
#endif
    ! public: the level-2 shim (mod_nonLinSolvers_b200.f90) drives the same handle
    type(c_ptr) :: ipic_h = c_null_ptr
    logical :: particlesOnDevice = .false.
    ! level 3 (INTEGRATION.md section 3b): injection (mod_particleInjection_b200.f90) and MCC (ipic_generate_collision) also run on the
    ! device, so nothing on the host touches phaseSpace between steps: the per-step download / upload is skipped and the host copy is
    ! refreshed on demand with ipic_sync_host (before diagnostics dumps, writePhaseSpace, restart files)
    logical :: particlesStayOnDevice = .false.

contains

    subroutine allocateParticleMoverData()
        ! NGP: no-op (:17-19).  CIC allocated particleTimeStep here (:67-70); the per-cell cap table now lives on the device.
    end subroutine allocateParticleMoverData

    subroutine ipic_attach(solver, particleList, world)
        class(potentialSolver), intent(in) :: solver
        type(Domain), intent(in) :: world
        type(Particle), intent(in) :: particleList(:)
        integer(c_int32_t) :: sp
        integer :: j
        if (c_associated(ipic_h)) return
        call ipic_check(ipic_create(ipic_h, IPIC_SCHEME, NumberXNodes, world%grid, world%dx_dl, world%boundaryConditions, &
            merge(1_c_int, 0_c_int, world%gridSmoothBool), 0_c_int), 'ipic_create')
        ! wall voltages as the constructor received them (NGP mod_potentialSolver.f90:54-84, CIC :56-90): a Dirichlet wall keeps its
        ! voltage in phi (NGP) / boundPhi (CIC), the RF wall in RF_half_amplitude
        call ipic_check(ipic_init_solver(ipic_h, ipic_wall_voltage(solver, world, 1), ipic_wall_voltage(solver, world, 2), &
            solver%RF_rad_frequency / (2.0d0 * pi), 0.0_c_double), 'ipic_init_solver')
        do j = 1, size(particleList)
            call ipic_check(ipic_add_species(ipic_h, particleList(j)%mass, particleList(j)%q, particleList(j)%w_p, sp), 'ipic_add_species')
        end do
    end subroutine ipic_attach

    function ipic_wall_voltage(solver, world, side) result(v)
        class(potentialSolver), intent(in) :: solver
        type(Domain), intent(in) :: world
        integer, intent(in) :: side
        real(c_double) :: v
        integer :: bc
        bc = world%boundaryConditions(merge(1, size(world%boundaryConditions), side == 1))
        if (bc == 4) then
            v = solver%RF_half_amplitude
        else if (bc == 1) then
#if IPIC_SCHEME == 1
            v = solver%boundPhi(side)
#else
            v = solver%phi(merge(1, NumberXNodes, side == 1))
#endif
        else
            v = 0.0d0
        end if
    end function ipic_wall_voltage

    subroutine ipic_push_state(solver, particleList)
        class(potentialSolver), intent(in) :: solver
        type(Particle), intent(in) :: particleList(:)
        integer :: j
        call ipic_check(ipic_set_field(ipic_h, IPIC_F_PHI, solver%phi, NumberXNodes), 'set phi')
        call ipic_check(ipic_set_field(ipic_h, IPIC_F_PHI_F, solver%phi_f, NumberXNodes), 'set phi_f')
#if IPIC_SCHEME == 1
        ! CIC keeps its walls outside phi; the host's setRFVoltage / resetVoltage have just updated them
        call ipic_check(ipic_set_cic_walls(ipic_h, solver%boundPhi, solver%RF_voltage, solver%sourceTermVals), 'ipic_set_cic_walls')
#endif
        if (.not. particlesOnDevice) then
            do j = 1, size(particleList)
                call ipic_check(ipic_set_weight(ipic_h, j - 1, particleList(j)%w_p), 'ipic_set_weight')
                call ipic_check(ipic_upload_particles(ipic_h, j - 1, particleList(j)%phaseSpace, particleList(j)%N_p, numThread, &
                    particleList(j)%finalIdx), 'ipic_upload_particles')
            end do
            particlesOnDevice = .true.
        end if
    end subroutine ipic_push_state

    subroutine depositJ(solver, particleList, world, del_t)
        class(potentialSolver), intent(in out), target :: solver
        type(Domain), intent(in) :: world
        type(Particle), intent(in out) :: particleList(:)
        real(real64), intent(in) :: del_t
        call ipic_attach(solver, particleList, world)
        call ipic_push_state(solver, particleList)
        call ipic_check(ipic_deposit_J(ipic_h, del_t), 'ipic_deposit_J')
        call ipic_check(ipic_get_field(ipic_h, IPIC_F_J, solver%J, NumberXHalfNodes), 'get J')
        call ipic_check(ipic_get_field(ipic_h, IPIC_F_EFIELD, solver%EField, NumberXHalfNodes), 'get EField')
    end subroutine depositJ

    subroutine ipic_pull_particles(particleList)
        ! the mover's outputs (mod_particleMover.f90:162-270): moved and compacted phaseSpace, N_p, loss sums, reflux bookkeeping, counters
        type(Particle), intent(in out) :: particleList(:)
        type(ipic_move_stats) :: st
        integer :: i, j, k, iThread
        integer(c_int64_t) :: nRef, slabStart
        integer(c_int32_t) :: refCount(numThread)
        integer(c_int32_t), allocatable :: refGlobal(:)
        integer(c_int64_t) :: nTotal
        do j = 1, size(particleList)
            if (particlesStayOnDevice) then
                ! only the count comes back; it is booked on the last slab, where appended particles land on the next download
                call ipic_check(ipic_particle_count(ipic_h, j - 1, nTotal), 'ipic_particle_count')
                particleList(j)%N_p = 0
                particleList(j)%N_p(numThread) = int(nTotal, int32)
            else
                call ipic_check(ipic_download_particles(ipic_h, j - 1, particleList(j)%phaseSpace, particleList(j)%N_p, numThread, &
                    particleList(j)%finalIdx), 'ipic_download_particles')
            end if
            call ipic_check(ipic_get_move_stats(ipic_h, j - 1, st), 'ipic_get_move_stats')
            ! totals are reported in slab 1; callers only ever use SUM(...) over the thread dimension
            ! (main.f90:79-84, mod_simulation.f90:171-200)
            particleList(j)%delIdx = 0
            particleList(j)%delIdx(1) = int(st%delIdx, int32)
            ! refIdx(T), refRecordIdx(:, T) (mod_particleMover.f90:226-231): the library returns the reflected particles as ascending
            ! 0-based indices into the downloaded order (slab after slab); split them back into 1-based per-slab indices
            particleList(j)%refIdx = 0
            if (st%refIdx > 0) then
                allocate(refGlobal(st%refIdx))
                call ipic_check(ipic_get_ref_record(ipic_h, j - 1, refGlobal, int(st%refIdx, c_int64_t), nRef), 'ipic_get_ref_record')
                if (particlesStayOnDevice) then
                    ! the device-side refluxing reads the list where it is; refIdx keeps the total
                    particleList(j)%refIdx(1) = int(st%refIdx, int32)
                else
                ! refIdx(T) comes from the library (the slab each reflection was recorded in); the list splits in that order, and an
                ! entry is turned into the slab's 1-based index by subtracting the slab's first global index
                call ipic_check(ipic_get_ref_record_counts(ipic_h, j - 1, refCount, numThread), 'ipic_get_ref_record_counts')
                k = 0
                slabStart = 0
                do iThread = 1, numThread
                    particleList(j)%refIdx(iThread) = refCount(iThread)
                    do i = 1, refCount(iThread)
                        k = k + 1
                        if (i <= size(particleList(j)%refRecordIdx, 1) .and. k <= nRef) &
                            particleList(j)%refRecordIdx(i, iThread) = int(refGlobal(k) - slabStart + 1, int32)
                    end do
                    slabStart = slabStart + particleList(j)%N_p(iThread)
                end do
                end if
                deallocate(refGlobal)
            end if
            particleList(j)%wallLoss = 0
            particleList(j)%wallLoss(:, 1) = st%wallLoss
            particleList(j)%energyLoss = 0.0d0
            particleList(j)%energyLoss(:, 1) = st%energyLoss
            particleList(j)%momentumLoss(:, 1) = st%momentumLoss
            particleList(j)%numToCollide = particleList(j)%N_p
            particleList(j)%numSubStepsAve = st%numSubStepsAve
            particleList(j)%numFuncEvalAve = st%numFuncEvalAve
            particleList(j)%accumEnergyLoss = particleList(j)%accumEnergyLoss + st%energyLoss
            particleList(j)%accumWallLoss = particleList(j)%accumWallLoss + int(st%wallLoss, int32)
        end do
    end subroutine ipic_pull_particles

    subroutine moveParticles(solver, particleList, world, del_t)
        class(potentialSolver), intent(in out) :: solver
        type(Domain), intent(in) :: world
        type(Particle), intent(in out) :: particleList(:)
        real(real64), intent(in) :: del_t
        call ipic_attach(solver, particleList, world)
        call ipic_push_state(solver, particleList)
        call ipic_check(ipic_move_particles(ipic_h, del_t), 'ipic_move_particles')
        call ipic_pull_particles(particleList)
        particlesOnDevice = particlesStayOnDevice     ! level 1/2: collisions / injection run on the host before the next step
    end subroutine moveParticles

    subroutine ipic_sync_host(particleList)
        ! level 3: bring the host copy of phaseSpace / N_p up to date (diagnostics dumps, writePhaseSpace); the slab split of the
        ! last upload is kept, particles appended since sit behind the last slab
        type(Particle), intent(in out) :: particleList(:)
        integer :: j
        do j = 1, size(particleList)
            call ipic_check(ipic_download_particles(ipic_h, j - 1, particleList(j)%phaseSpace, particleList(j)%N_p, numThread, &
                particleList(j)%finalIdx), 'ipic_download_particles')
        end do
    end subroutine ipic_sync_host

end module mod_particleMover
