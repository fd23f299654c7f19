"""implicitpic1d_b200 -- B200-native implicit-PIC inner loop behind the module interface of
nsavard01/ImplicitPic1D (OpenMP/NGP and OpenMP/CIC trees).

Host-side mirror of the reference's Fortran modules; all numerical work happens in the CUDA library
(csrc/, C ABI in include/ipic1d_b200.h).  Names follow the reference:

    mod_domain            -> Domain
    mod_particle          -> Particle
    mod_potentialSolver   -> potentialSolver
    mod_particleMover     -> depositJ, moveParticles, allocateParticleMoverData
    mod_nonLinSolvers     -> initializeSolver, solvePotential, solveDivAmpereAnderson, funcNitsol
    mod_targetParticle    -> targetParticle
    mod_NullCollision     -> nullCollision, readNullCollisionInputs
    mod_Scheme            -> scheme (loadParticleDensity, WriteParticleDensity)
    mod_simulation        -> simulation (solveSimulation, solveSimulationFinalAverage, Histograms, writePhi)
    mod_particleInjection -> particle_injection (addMaxwellianLostParticles, refluxParticles, injectAtBoundary,
                             injectUniformFlux, maxwellianHeating, readInjectionInputs)
"""
from . import constants
from .domain import Domain
from .particle import Particle
from .potential_solver import potentialSolver
from .particle_mover import depositJ, moveParticles, allocateParticleMoverData
from .nonlin_solvers import initializeSolver, solvePotential, solveDivAmpereAnderson, funcNitsol, SolverState, solveJFNK, nitsol, NitsolInfo
from .null_collision import targetParticle, nullCollision, readNullCollisionInputs
from . import particle_injection
from . import scheme
from . import simulation
from ._lib import IpicError, load as load_library, LIB_PATH

__all__ = ["constants", "Domain", "Particle", "potentialSolver", "depositJ", "moveParticles", "allocateParticleMoverData",
           "initializeSolver", "solvePotential", "solveDivAmpereAnderson", "funcNitsol", "SolverState", "solveJFNK", "nitsol", "NitsolInfo", "targetParticle", "nullCollision", "readNullCollisionInputs", "particle_injection", "scheme", "simulation", "IpicError", "load_library", "LIB_PATH"]
