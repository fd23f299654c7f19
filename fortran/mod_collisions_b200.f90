! =============================================================================
! mod_collisions_b200.f90 -- null-collision MCC on the device for a Fortran host (level-3 integration, INTEGRATION.md 3b).
! The reference keeps reading its tables and building type(nullCollision) objects itself (mod_NullCollision.f90:44-108, 666-939);
! this module hands those objects to the library once and replaces the call
!     call nullCollisionList(j)%generateCollision(particleList, targetParticleList, numberChargedParticles, numberBinaryCollisions, irand, currDel_t)
! (mod_simulation.f90:166-168 and its three twins) by
!     call ipic_generateCollision(nullCollisionList, targetParticleList, j, currDel_t)
! The running totals the diagnostics read (totalAmountCollisions, totalEnergyLoss, totalIncidentEnergy,
! totalNumberCollidableParticles; writeCollisionDiag :611-632) are copied back into the object after every call; resetting them on the
! host (mod_simulation.f90:279-284) is forwarded with ipic_resetCollisionTotals.
! =============================================================================
module mod_collisions_b200
    use iso_fortran_env, only: int32, int64, real64
    use iso_c_binding
    use mod_targetParticle
    use mod_NullCollision
    use ipic_b200_c
    use mod_particleMover, only: ipic_h
    implicit none
    logical, private :: collisionsOnDevice = .false.

contains

    subroutine ipic_attach_collisions(nullCollisionList, targetParticleList)
        ! once, after the first depositJ / solvePotential created the device context: the target particles in list order, then one
        ! collision set per list entry (set j-1 on the device = nullCollisionList(j))
        type(nullCollision), intent(in) :: nullCollisionList(:)
        type(targetParticle), intent(in) :: targetParticleList(:)
        integer :: j
        integer(c_int32_t) :: idx
        if (collisionsOnDevice) return
        do j = 1, size(targetParticleList)
            call ipic_check(ipic_add_target_particle(ipic_h, targetParticleList(j)%mass, targetParticleList(j)%density, &
                targetParticleList(j)%temperature, idx), 'ipic_add_target_particle')
        end do
        do j = 1, size(nullCollisionList)
            call ipic_check(ipic_null_collision_create(ipic_h, nullCollisionList(j)%numberCollisions, nullCollisionList(j)%lengthArrays, &
                nullCollisionList(j)%reducedMass, nullCollisionList(j)%sumMass, nullCollisionList(j)%reducedMassIonization, &
                nullCollisionList(j)%energyArray, nullCollisionList(j)%sigmaArray, nullCollisionList(j)%energyThreshold, &
                nullCollisionList(j)%collisionType, nullCollisionList(j)%reactantsIndx, nullCollisionList(j)%numberProducts, &
                nullCollisionList(j)%productsIndx, int(20261017 + 7919 * j, c_int64_t), idx), 'ipic_null_collision_create')
        end do
        collisionsOnDevice = .true.
    end subroutine ipic_attach_collisions

    subroutine ipic_pull_collision_totals(nc, j)
        type(nullCollision), intent(in out) :: nc
        integer, intent(in) :: j
        type(ipic_collision_stats) :: st
        integer :: k
        call ipic_check(ipic_get_collision_stats(ipic_h, int(j - 1, c_int32_t), st), 'ipic_get_collision_stats')
        do k = 1, nc%numberCollisions
            nc%totalAmountCollisions(k) = st%totalAmountCollisions(k)
            nc%totalEnergyLoss(k) = st%totalEnergyLoss(k)
            nc%totalIncidentEnergy(k) = st%totalIncidentEnergy(k)
        end do
        nc%totalNumberCollidableParticles = st%totalNumberCollidableParticles
    end subroutine ipic_pull_collision_totals

    subroutine ipic_generateCollision(nullCollisionList, targetParticleList, j, del_t)
        ! generateCollision (mod_NullCollision.f90:110-237) of set j over del_t, on the particles where the mover left them
        type(nullCollision), intent(in out) :: nullCollisionList(:)
        type(targetParticle), intent(in) :: targetParticleList(:)
        integer, intent(in) :: j
        real(real64), intent(in) :: del_t
        call ipic_attach_collisions(nullCollisionList, targetParticleList)
        call ipic_check(ipic_generate_collision(ipic_h, int(j - 1, c_int32_t), del_t), 'ipic_generate_collision')
        call ipic_pull_collision_totals(nullCollisionList(j), j)
    end subroutine ipic_generateCollision

    subroutine ipic_resetCollisionTotals(nullCollisionList, j)
        type(nullCollision), intent(in out) :: nullCollisionList(:)
        integer, intent(in) :: j
        if (collisionsOnDevice) call ipic_check(ipic_reset_collision_stats(ipic_h, int(j - 1, c_int32_t)), 'ipic_reset_collision_stats')
        nullCollisionList(j)%totalAmountCollisions = 0
        nullCollisionList(j)%totalEnergyLoss = 0
        nullCollisionList(j)%totalIncidentEnergy = 0
        nullCollisionList(j)%totalNumberCollidableParticles = 0
    end subroutine ipic_resetCollisionTotals

end module mod_collisions_b200
