! =============================================================================
! mod_particleInjection_b200.f90 -- drop-in replacement of OpenMP/SharedModules/src/mod_particleInjection.f90 for the level-2
! integration (particles stay on the device between steps, see INTEGRATION.md).  Same module name, public variables and
! procedure signatures (:9-12, :16, :42, :71, :132, :160, :181, :254), so mod_simulation.f90:158-165 keeps compiling unchanged;
! the bodies call the C ABI (csrc/injection.cu).  `irand` is accepted and left untouched: the device draws from counter-based
! PCG32 streams (ipic_set_injection_seed).  energyAddColl(1) receives the whole tally (the reference only uses SUM(energyAddColl),
! mod_simulation.f90:190,426).  After a call that appends pairs, N_p(numThread) of both species grows by the number appended
! (new pairs land behind the last thread slab, where ipic_download_particles writes them).
! =============================================================================
module mod_particleInjection
    use iso_fortran_env, only: int32, real64
    use iso_c_binding
    use constants
    use mod_BasicFunctions
    use mod_domain
    use mod_particle
    use ipic_b200_c
    use mod_particleMover, only: ipic_h
    implicit none
    real(real64), allocatable, public :: energyAddColl(:)
    logical, public, protected :: addLostPartBool, refluxPartBool, injectionBool, uniformInjectionBool, heatingBool
    integer(int32), private :: numFluxParticlesHigh, numFluxParticlesLow
    real(real64), private :: injectionFlux, injectionR, FractionFreqHeating
    real(real64), private :: numFluxPartTotal      ! expected pairs per call for the whole process = numThread x the per-thread figure

contains

    subroutine addMaxwellianLostParticles(particleList, T_e, T_i, irand, world)
        type(Particle), intent(in out) :: particleList(2)
        type(Domain), intent(in) :: world
        integer(int32), intent(in out) :: irand(numThread)
        real(real64), intent(in) :: T_e, T_i
        real(c_double) :: en
        integer(c_int64_t) :: n
        call ipic_check(ipic_add_maxwellian_lost_particles(ipic_h, T_e, T_i, en, n), 'addMaxwellianLostParticles')
        energyAddColl(1) = energyAddColl(1) + en
        particleList(1)%N_p(numThread) = particleList(1)%N_p(numThread) + int(n, int32)
        particleList(2)%N_p(numThread) = particleList(2)%N_p(numThread) + int(n, int32)
    end subroutine addMaxwellianLostParticles

    subroutine refluxParticles(particleList, T_e, T_i, irand, world)
        type(Particle), intent(in out) :: particleList(2)
        type(Domain), intent(in) :: world
        integer(int32), intent(in out) :: irand(numThread)
        real(real64), intent(in) :: T_e, T_i
        real(c_double) :: en
        call ipic_check(ipic_reflux_particles(ipic_h, T_e, T_i, en), 'refluxParticles')
        energyAddColl(1) = energyAddColl(1) + en
    end subroutine refluxParticles

    subroutine injectAtBoundary(particleList, T_e, T_i, irand, world, del_t)
        type(Particle), intent(in out) :: particleList(2)
        type(Domain), intent(in) :: world
        integer(int32), intent(in out) :: irand(numThread)
        real(real64), intent(in) :: T_e, T_i, del_t
        integer(c_int64_t) :: n
        call ipic_check(ipic_inject_at_boundary(ipic_h, T_e, T_i, del_t, numFluxPartTotal, n), 'injectAtBoundary')
        particleList(1)%N_p(numThread) = particleList(1)%N_p(numThread) + int(n, int32)
        particleList(2)%N_p(numThread) = particleList(2)%N_p(numThread) + int(n, int32)
    end subroutine injectAtBoundary

    subroutine injectUniformFlux(particleList, T_e, T_i, irand, world)
        type(Particle), intent(in out) :: particleList(2)
        type(Domain), intent(in) :: world
        integer(int32), intent(in out) :: irand(numThread)
        real(real64), intent(in) :: T_e, T_i
        integer(c_int64_t) :: n
        call ipic_check(ipic_inject_uniform_flux(ipic_h, T_e, T_i, numFluxPartTotal, n), 'injectUniformFlux')
        particleList(1)%N_p(numThread) = particleList(1)%N_p(numThread) + int(n, int32)
        particleList(2)%N_p(numThread) = particleList(2)%N_p(numThread) + int(n, int32)
    end subroutine injectUniformFlux

    subroutine maxwellianHeating(species, irand, FracFreq, T, currDel_t, del_t)
        ! the reference passes particleList(1): species 0 of the device context
        type(Particle), intent(in out) :: species
        integer(int32), intent(in out) :: irand(numThread)
        real(real64), intent(in) :: FracFreq, T, currDel_t, del_t
        integer(c_int64_t) :: n
        call ipic_check(ipic_maxwellian_heating(ipic_h, 0_c_int32_t, FracFreq * (currDel_t/del_t)/FractionFreqHeating, T, n), 'maxwellianHeating')
    end subroutine maxwellianHeating

    subroutine readInjectionInputs(InjFilename, w_p)
        ! Same file, same meaning as the reference reader (:181-252): five lines `flag [, value]`; the uniform-injection line
        ! is only consumed when boundary injection is off.  What the device needs from it is the expected number of pairs per
        ! call for the whole process (numThread x the reference's per-thread numFluxPart, :232).
        real(real64), intent(in) :: w_p
        character(len=*), intent(in) :: InjFilename
        integer(int32) :: flag(5), unitNum, io
        real(real64) :: val(5), perThread
        character(len=:), allocatable :: path
        path = '../InputData/'//InjFilename
        if (restartBool) path = restartDirectory//"/InputData/"//InjFilename
        flag = 0
        val = 0.0d0
        open(newunit=unitNum, file=path, IOSTAT=io)
        read(unitNum, *, IOSTAT=io) flag(1)
        read(unitNum, *, IOSTAT=io) flag(2)
        read(unitNum, *, IOSTAT=io) flag(3), val(3)
        if (flag(3) /= 1) read(unitNum, *, IOSTAT=io) flag(4), val(4)
        read(unitNum, *, IOSTAT=io) flag(5), val(5)
        close(unitNum)
        addLostPartBool = flag(1) == 1
        refluxPartBool = flag(2) == 1
        injectionBool = flag(3) == 1
        uniformInjectionBool = flag(4) == 1
        heatingBool = flag(5) == 1
        injectionFlux = merge(val(3), val(4), injectionBool)
        FractionFreqHeating = val(5)
        numFluxPartTotal = 0.0d0
        if (injectionBool .or. uniformInjectionBool) then
            perThread = injectionFlux * del_t / w_p / real(numThread, real64)
            numFluxParticlesLow = floor(perThread)
            numFluxParticlesHigh = numFluxParticlesLow + 1
            injectionR = perThread - real(numFluxParticlesLow, real64)
            numFluxPartTotal = perThread * real(numThread, real64)
        end if
        print *, "Particle injection switches (lost re-injected, reflux, boundary injection, uniform injection, heating):", &
            addLostPartBool, refluxPartBool, injectionBool, uniformInjectionBool, heatingBool
        if (numFluxPartTotal > 0.0d0) print *, "flux", injectionFlux, "-> pairs per call", numFluxPartTotal
        if (heatingBool) print *, "FractionFreqHeating", FractionFreqHeating
    end subroutine readInjectionInputs

    subroutine writeParticleInjectionInputs(dirName)
        ! ParticleInjectionInput.dat in the layout Analysis/ reads (:254-261): a header line, then five logicals and two reals
        character(*), intent(in) :: dirName
        integer :: u
        open(newunit=u, file=dirName//'/ParticleInjectionInput.dat')
        write(u, '(a)') "AddLostPartBool, refluxPartBool, injectionBool, uniformInjectionBool, heatingBool, injectionFlux, FractionFreqHeating"
        write(u, "(5(L, 1x), 2(es16.8,1x))") addLostPartBool, refluxPartBool, injectionBool, uniformInjectionBool, heatingBool, &
            injectionFlux, FractionFreqHeating
        close(u)
    end subroutine writeParticleInjectionInputs

end module mod_particleInjection
