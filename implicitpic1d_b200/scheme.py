"""Host-side mirror of the density diagnostic of `mod_Scheme` (OpenMP/NGP/src/mod_Scheme.f90:68-158,
OpenMP/CIC/src/mod_Scheme.f90:68-185): `loadParticleDensity` runs the per-species deposit on the device
(`ipic_load_particle_density`: one pass over `l`, 8 B per particle, all-reduced over ranks), `WriteParticleDensity` applies the
reference's O(N_x) epilogue (weight, periodic fold, binomial smoothing, 1/dx normalisation) and writes the unformatted record
`Analysis/dataSet.py` reads.  `Particle.densities` holds the sum of weights of all loads since the last reset (the reference's
SUM(densities, DIM=2)), so time averages accumulate exactly as `loadParticleDensity(..., reset=.false.)` does."""
import os
import numpy as np
from . import _lib
from .domain import CIC


def loadParticleDensity(particleList, world, reset):
    L = _lib.load()
    for p in particleList:
        if p._solver is None:
            raise RuntimeError("particle is not bound to a potentialSolver")
        d = np.zeros(world.NumberXNodes)
        _lib.check(L.ipic_load_particle_density(p._solver._h, p._index, d.ctypes.data_as(_lib._dp)))
        if reset or getattr(p, "densities", None) is None:
            p.densities = d
        else:
            p.densities = p.densities + d


def particleDensityArray(p, world):
    """the `densities` array WriteParticleDensity writes for one species (NGP :96-146, CIC :134-163), in the reference's operation order"""
    Nx, Nh = world.NumberXNodes, world.NumberXHalfNodes
    bc, dx = world.boundaryConditions, world.dx_dl
    D = np.zeros(Nx + 2)                      # 1-based like the reference: D[j] = densities(j, 1)
    D[1:Nx + 1] = p.densities * p.w_p
    dens = np.zeros(Nx + 2)
    if world.scheme != CIC:
        if bc[0] == 3:
            D[1] = D[1] + D[Nx]
            D[Nx] = D[1]
        if world.gridSmoothBool:
            for j in range(1, Nx + 1):
                b = bc[j - 1]
                if b == 0:
                    dens[j] = dens[j] + 0.25 * (D[j - 1] + 2.0 * D[j] + D[j + 1])
                elif b in (1, 4, 2):
                    if j == 1:
                        dens[j] = dens[j] + 0.25 * (2.0 * D[1] + D[2])
                        dens[j + 1] = dens[j + 1] + 0.25 * D[1]
                    else:
                        dens[j] = dens[j] + 0.25 * (2.0 * D[Nx] + D[Nh])
                        dens[j - 1] = dens[j - 1] + 0.25 * D[Nx]
                elif b == 3:
                    dens[j] = dens[j] + 0.25 * (D[Nh] + 2.0 * D[1] + D[2])
        else:
            dens[1:Nx + 1] = D[1:Nx + 1]
        if bc[0] in (1, 2, 4):
            dens[1] = 2.0 * dens[1] / dx[0]
        elif bc[0] == 3:
            dens[1] = 2.0 * dens[1] / (dx[0] + dx[Nh - 1])
        for j in range(2, Nh + 1):
            dens[j] = 2.0 * dens[j] / (dx[j - 2] + dx[j - 1])
        if bc[Nx - 1] in (1, 2, 4):
            dens[Nx] = 2.0 * dens[Nx] / dx[Nh - 1]
        elif bc[Nx - 1] == 3:
            dens[Nx] = dens[1]
    else:
        if world.gridSmoothBool:
            for j in range(1, Nx + 1):
                b = int(bc[j]) - int(bc[j - 1])
                if b == 0:
                    dens[j] = 0.25 * (D[j - 1] + 2.0 * D[j] + D[j + 1])
                elif b in (-2, -1, -4):
                    dens[j] = 0.25 * (D[2] + 3.0 * D[1])
                elif b in (2, 1, 4):
                    dens[j] = 0.25 * (D[Nx - 1] + 3.0 * D[Nx])
                elif b == -3:
                    dens[j] = 0.25 * (D[Nx] + 2.0 * D[1] + D[2])
                elif b == 3:
                    dens[j] = 0.25 * (D[1] + 2.0 * D[Nx] + D[Nx - 1])
            D[1:Nx + 1] = dens[1:Nx + 1]
        dens[1:Nx + 1] = D[1:Nx + 1] / dx
    return dens[1:Nx + 1].copy()


def WriteParticleDensity(particleList, world, CurrentDiagStep, boolAverage, dirName):
    """Density/density_<name>_<step>.dat (or _Average.dat): one Fortran unformatted sequential record of N_x float64.
    Returns the arrays written."""
    os.makedirs(os.path.join(dirName, "Density"), exist_ok=True)
    out = []
    for p in particleList:
        dens = particleDensityArray(p, world)
        tail = "Average" if boolAverage else "%d" % int(CurrentDiagStep)
        with open(os.path.join(dirName, "Density", "density_%s_%s.dat" % (p.name, tail)), "wb") as f:
            marker = np.array([dens.nbytes], dtype="<i4").tobytes()
            f.write(marker + dens.astype("<f8").tobytes() + marker)
        out.append(dens)
    return out
