! =============================================================================
! ipic_b200_c.f90 -- ISO_C_BINDING interface to libipic1d_b200.so (include/ipic1d_b200.h)
!
! Reference-side binding a maintainer of nsavard01/ImplicitPic1D adds to
! OpenMP/SharedModules/src/ (listed before mod_particleMover.o in the Makefile) and links with
!     -L<repo>/implicitpic1d_b200/lib -lipic1d_b200 -Wl,-rpath,<repo>/implicitpic1d_b200/lib
! Precedent in the reference for Fortran -> C calls: pcg32_random_r, bind(c)
! (PIC1D/SharedModules/src/mod_BasicFunctions.f90:13-19).
!
! NOTE: this container has no Fortran compiler (SURVEY.md section 0), so this file is exercised
! only through its C twin: tests/test_c_abi_host.py drives the same entry points in the same
! order from C (tests/c_host/host_driver.c).
! =============================================================================
module ipic_b200_c
    use iso_c_binding
    implicit none
    public

    integer(c_int), parameter :: IPIC_SCHEME_NGP = 0, IPIC_SCHEME_CIC = 1, IPIC_SCHEME_NGP_PERIODIC = 2
    integer(c_int), parameter :: IPIC_F_PHI = 0, IPIC_F_PHI_F = 1, IPIC_F_J = 2, IPIC_F_RHO = 3, IPIC_F_EFIELD = 4
    integer(c_int), parameter :: IPIC_TRIDIAG_PCR = 0, IPIC_TRIDIAG_THOMAS = 1

    type, bind(c) :: ipic_move_stats
        integer(c_int64_t) :: N_p, delIdx, refIdx, numSubSteps, numFuncEvals
        real(c_double) :: numSubStepsAve, numFuncEvalAve
        real(c_double) :: wallLoss(2), energyLoss(2), momentumLoss(2), accumEnergyLoss(2)
        integer(c_int64_t) :: accumWallLoss(2)
    end type ipic_move_stats

    type, bind(c) :: ipic_solver_stats
        integer(c_int32_t) :: iterNumPicard, iterNumAdaptiveSteps, amountTimeSplits
        integer(c_int64_t) :: sweeps, substeps
        real(c_double) :: moverSeconds, potSeconds, lastResidualNorm
    end type ipic_solver_stats

    integer(c_int), parameter :: IPIC_MAX_COLLISIONS = 16
    type, bind(c) :: ipic_collision_stats
        integer(c_int32_t) :: numberCollisions, lengthArrays
        integer(c_int32_t) :: collisionType(IPIC_MAX_COLLISIONS)
        real(c_double) :: energyThreshold(IPIC_MAX_COLLISIONS)
        real(c_double) :: sigmaVMax, reducedMass, sumMass, reducedMassIonization, minEnergy, maxEnergy
        real(c_double) :: P_null
        integer(c_int64_t) :: numberSelected, ionizations
        integer(c_int64_t) :: totalAmountCollisions(IPIC_MAX_COLLISIONS)
        real(c_double) :: totalEnergyLoss(IPIC_MAX_COLLISIONS), totalIncidentEnergy(IPIC_MAX_COLLISIONS)
        integer(c_int64_t) :: totalNumberCollidableParticles
    end type ipic_collision_stats

    interface
        function ipic_last_error() bind(c, name="ipic_last_error") result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function
        function ipic_create(handle, scheme, NumberXNodes, grid, dx_dl, boundaryConditions, gridSmoothBool, device) &
                bind(c, name="ipic_create") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), intent(out) :: handle
            integer(c_int), value :: scheme, gridSmoothBool, device
            integer(c_int32_t), value :: NumberXNodes
            real(c_double), intent(in) :: grid(*), dx_dl(*)
            integer(c_int32_t), intent(in) :: boundaryConditions(*)
            integer(c_int) :: rc
        end function
        function ipic_destroy(handle) bind(c, name="ipic_destroy") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: handle
            integer(c_int) :: rc
        end function
        function ipic_init_solver(handle, leftVoltage, rightVoltage, RF_frequency, timeCurrent) bind(c, name="ipic_init_solver") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: leftVoltage, rightVoltage, RF_frequency, timeCurrent
            integer(c_int) :: rc
        end function
        function ipic_set_tridiag_mode(handle, mode) bind(c, name="ipic_set_tridiag_mode") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: handle
            integer(c_int), value :: mode
            integer(c_int) :: rc
        end function
        function ipic_add_species(handle, mass, q, w_p, species) bind(c, name="ipic_add_species") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: mass, q, w_p
            integer(c_int32_t), intent(out) :: species
            integer(c_int) :: rc
        end function
        function ipic_set_weight(handle, species, w_p) bind(c, name="ipic_set_weight") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            real(c_double), value :: w_p
            integer(c_int) :: rc
        end function
        function ipic_upload_particles(handle, species, phaseSpace, N_p, numThread, finalIdx) bind(c, name="ipic_upload_particles") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species, numThread, finalIdx
            real(c_double), intent(in) :: phaseSpace(4, finalIdx, numThread)
            integer(c_int32_t), intent(in) :: N_p(numThread)
            integer(c_int) :: rc
        end function
        function ipic_download_particles(handle, species, phaseSpace, N_p, numThread, finalIdx) bind(c, name="ipic_download_particles") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species, numThread, finalIdx
            real(c_double), intent(in out) :: phaseSpace(4, finalIdx, numThread)
            integer(c_int32_t), intent(out) :: N_p(numThread)
            integer(c_int) :: rc
        end function
        function ipic_get_field(handle, which, out, n) bind(c, name="ipic_get_field") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int), value :: which
            integer(c_int32_t), value :: n
            real(c_double), intent(out) :: out(*)
            integer(c_int) :: rc
        end function
        function ipic_set_field(handle, which, vals, n) bind(c, name="ipic_set_field") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int), value :: which
            integer(c_int32_t), value :: n
            real(c_double), intent(in) :: vals(*)
            integer(c_int) :: rc
        end function
        function ipic_deposit_rho(handle) bind(c, name="ipic_deposit_rho") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: handle
            integer(c_int) :: rc
        end function
        function ipic_solve_tridiag_poisson(handle, timeCurrent) bind(c, name="ipic_solve_tridiag_poisson") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: timeCurrent
            integer(c_int) :: rc
        end function
        function ipic_make_const_source_term(handle) bind(c, name="ipic_make_const_source_term") result(rc)
            import :: c_ptr, c_int
            type(c_ptr), value :: handle
            integer(c_int) :: rc
        end function
        function ipic_solve_tridiag_ampere(handle, del_t) bind(c, name="ipic_solve_tridiag_ampere") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: del_t
            integer(c_int) :: rc
        end function
        function ipic_set_rf_voltage(handle, timeFuture) bind(c, name="ipic_set_rf_voltage") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: timeFuture
            integer(c_int) :: rc
        end function
        function ipic_set_cic_walls(handle, boundPhi, RF_voltage, sourceTermVals) bind(c, name="ipic_set_cic_walls") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), intent(in) :: boundPhi(2), sourceTermVals(2)
            real(c_double), value :: RF_voltage
            integer(c_int) :: rc
        end function
        function ipic_deposit_J(handle, del_t) bind(c, name="ipic_deposit_J") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: del_t
            integer(c_int) :: rc
        end function
        function ipic_move_particles(handle, del_t) bind(c, name="ipic_move_particles") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: del_t
            integer(c_int) :: rc
        end function
        function ipic_get_move_stats(handle, species, stats) bind(c, name="ipic_get_move_stats") result(rc)
            import :: c_ptr, c_int, c_int32_t, ipic_move_stats
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            type(ipic_move_stats), intent(out) :: stats
            integer(c_int) :: rc
        end function
        function ipic_get_ref_record(handle, species, idx, cap, n) bind(c, name="ipic_get_ref_record") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            integer(c_int32_t), intent(out) :: idx(*)
            integer(c_int64_t), value :: cap
            integer(c_int64_t), intent(out) :: n
            integer(c_int) :: rc
        end function
        function ipic_get_ref_record_counts(handle, species, refIdx, numThread) bind(c, name="ipic_get_ref_record_counts") result(rc)
            import :: c_ptr, c_int, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species, numThread
            integer(c_int32_t), intent(out) :: refIdx(*)
            integer(c_int) :: rc
        end function
        function ipic_initialize_solver(handle, eps_a, eps_r, solverType, m_Anderson, Beta_k, maxIter) &
                bind(c, name="ipic_initialize_solver") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: eps_a, eps_r, Beta_k
            integer(c_int), value :: solverType, m_Anderson, maxIter
            integer(c_int) :: rc
        end function
        function ipic_solve_potential(handle, del_t, remainDel_t, currDel_t, timeCurrent) bind(c, name="ipic_solve_potential") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: del_t, timeCurrent
            real(c_double), intent(in out) :: remainDel_t, currDel_t
            integer(c_int) :: rc
        end function
        function ipic_solve_div_ampere_anderson(handle, del_t) bind(c, name="ipic_solve_div_ampere_anderson") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: del_t
            integer(c_int) :: rc
        end function
        function ipic_solve_jfnk(handle, del_t) bind(c, name="ipic_solve_jfnk") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: del_t
            integer(c_int) :: rc
        end function
        function ipic_func_nitsol(handle, n, xcur, fcur, del_t) bind(c, name="ipic_func_nitsol") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: n
            real(c_double), intent(in) :: xcur(n)
            real(c_double), intent(out) :: fcur(n)
            real(c_double), value :: del_t
            integer(c_int) :: rc
        end function
        function ipic_get_solver_stats(handle, stats) bind(c, name="ipic_get_solver_stats") result(rc)
            import :: c_ptr, c_int, ipic_solver_stats
            type(c_ptr), value :: handle
            type(ipic_solver_stats), intent(out) :: stats
            integer(c_int) :: rc
        end function
        function ipic_particle_count(handle, species, n) bind(c, name="ipic_particle_count") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            integer(c_int64_t), intent(out) :: n
            integer(c_int) :: rc
        end function
        function ipic_load_particle_density(handle, species, densities) bind(c, name="ipic_load_particle_density") result(rc)
            ! loadParticleDensity (mod_Scheme.f90) for one species: SUM(part%densities, DIM=2) of a reset load, before w_p
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            real(c_double), intent(out) :: densities(*)
            integer(c_int) :: rc
        end function
        function ipic_velocity_extrema(handle, species, out3) bind(c, name="ipic_velocity_extrema") result(rc)
            ! mod_simulation.f90:437-452: out3 = (VMax, EMax, Emin) of this process's particles
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            real(c_double), intent(out) :: out3(3)
            integer(c_int) :: rc
        end function
        function ipic_velocity_histograms(handle, species, binNumber, VMax, Emin, EMax, VHist, EHist) &
                bind(c, name="ipic_velocity_histograms") result(rc)
            ! mod_simulation.f90:480-500: adds one pass to VHist(2*binNumber), EHist(binNumber)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species, binNumber
            real(c_double), value :: VMax, Emin, EMax
            real(c_double), intent(in out) :: VHist(*), EHist(*)
            integer(c_int) :: rc
        end function
        function ipic_local_temperature(handle, species, numCells, EHist) bind(c, name="ipic_local_temperature") result(rc)
            ! writeLocalTemperature (mod_particle.f90:251-283): the array it writes
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species, numCells
            real(c_double), intent(out) :: EHist(*)
            integer(c_int) :: rc
        end function
        function ipic_get_total_KE(handle, species, res) bind(c, name="ipic_get_total_KE") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            real(c_double), intent(out) :: res
            integer(c_int) :: rc
        end function
        function ipic_get_total_momentum(handle, species, res3) bind(c, name="ipic_get_total_momentum") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            real(c_double), intent(out) :: res3(3)
            integer(c_int) :: rc
        end function
        function ipic_add_target_particle(handle, mass, density, temperature, target) bind(c, name="ipic_add_target_particle") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: mass, density, temperature
            integer(c_int32_t), intent(out) :: target
            integer(c_int) :: rc
        end function
        function ipic_null_collision_create(handle, numberCollisions, lengthArrays, red_mass, sumMass, redMassTripleProducts, energyArray, &
                sigmaArray, energyThreshold, collisionType, reactantsIndx, numberProducts, productsIndx, seed, collision_set) &
                bind(c, name="ipic_null_collision_create") result(rc)
            ! the arguments of nullCollision_constructor (mod_NullCollision.f90:44), arrays in the reference's own (column-major) layout
            import :: c_ptr, c_int, c_int32_t, c_int64_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: numberCollisions, lengthArrays
            real(c_double), value :: red_mass, sumMass, redMassTripleProducts
            real(c_double), intent(in) :: energyArray(*), sigmaArray(*), energyThreshold(*)
            integer(c_int32_t), intent(in) :: collisionType(*), reactantsIndx(*), numberProducts(*), productsIndx(*)
            integer(c_int64_t), value :: seed
            integer(c_int32_t), intent(out) :: collision_set
            integer(c_int) :: rc
        end function
        function ipic_get_collision_stats(handle, collision_set, stats) bind(c, name="ipic_get_collision_stats") result(rc)
            import :: c_ptr, c_int, c_int32_t, ipic_collision_stats
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: collision_set
            type(ipic_collision_stats), intent(out) :: stats
            integer(c_int) :: rc
        end function
        function ipic_reset_collision_stats(handle, collision_set) bind(c, name="ipic_reset_collision_stats") result(rc)
            import :: c_ptr, c_int, c_int32_t
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: collision_set
            integer(c_int) :: rc
        end function
        function ipic_generate_collision(handle, collision_set, del_t) bind(c, name="ipic_generate_collision") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: collision_set
            real(c_double), value :: del_t
            integer(c_int) :: rc
        end function
        ! ---- mod_particleInjection.f90 on the device (include/ipic1d_b200.h, "wall refluxing, particle injection, Maxwellian heating")
        function ipic_set_injection_seed(handle, seed) bind(c, name="ipic_set_injection_seed") result(rc)
            import :: c_ptr, c_int, c_int64_t
            type(c_ptr), value :: handle
            integer(c_int64_t), value :: seed
            integer(c_int) :: rc
        end function
        function ipic_add_maxwellian_lost_particles(handle, T_e, T_i, energyAdd, numAdded) &
                bind(c, name="ipic_add_maxwellian_lost_particles") result(rc)
            import :: c_ptr, c_int, c_int64_t, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: T_e, T_i
            real(c_double), intent(out) :: energyAdd
            integer(c_int64_t), intent(out) :: numAdded
            integer(c_int) :: rc
        end function
        function ipic_reflux_particles(handle, T_e, T_i, energyAdd) bind(c, name="ipic_reflux_particles") result(rc)
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: T_e, T_i
            real(c_double), intent(out) :: energyAdd
            integer(c_int) :: rc
        end function
        function ipic_inject_at_boundary(handle, T_e, T_i, del_t, numFluxPart, numInjected) bind(c, name="ipic_inject_at_boundary") result(rc)
            import :: c_ptr, c_int, c_int64_t, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: T_e, T_i, del_t, numFluxPart
            integer(c_int64_t), intent(out) :: numInjected
            integer(c_int) :: rc
        end function
        function ipic_inject_uniform_flux(handle, T_e, T_i, numFluxPart, numInjected) bind(c, name="ipic_inject_uniform_flux") result(rc)
            import :: c_ptr, c_int, c_int64_t, c_double
            type(c_ptr), value :: handle
            real(c_double), value :: T_e, T_i, numFluxPart
            integer(c_int64_t), intent(out) :: numInjected
            integer(c_int) :: rc
        end function
        function ipic_maxwellian_heating(handle, species, compValue, T_eV, numHeated) bind(c, name="ipic_maxwellian_heating") result(rc)
            import :: c_ptr, c_int, c_int32_t, c_int64_t, c_double
            type(c_ptr), value :: handle
            integer(c_int32_t), value :: species
            real(c_double), value :: compValue, T_eV
            integer(c_int64_t), intent(out) :: numHeated
            integer(c_int) :: rc
        end function
    end interface

contains

    subroutine ipic_check(rc, where)
        ! the reference's only error behaviour is `print *, ...; stop` (e.g. mod_nonLinSolvers.f90:186)
        integer(c_int), intent(in) :: rc
        character(len=*), intent(in) :: where
        character(kind=c_char), pointer :: cmsg(:)
        integer :: k
        if (rc /= 0) then
            print *, 'ipic1d_b200 error ', rc, ' in ', where
            call c_f_pointer(ipic_last_error(), cmsg, [512])
            do k = 1, 512
                if (cmsg(k) == c_null_char) exit
            end do
            print *, cmsg(1:k-1)
            stop
        end if
    end subroutine ipic_check

end module ipic_b200_c
