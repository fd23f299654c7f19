"""Host-side mirror of the time loop of `mod_simulation` (OpenMP/SharedModules/src/mod_simulation.f90): `solveSimulation` (:48-308)
and `solveSimulationFinalAverage` (:310-513) with every particle operation on the device -- the implicit step, heating /
re-injection / refluxing / injection, null-collision MCC, and the diagnostics reductions (energies, momentum, density and
temperature dumps, velocity / energy histograms).  The particles never leave HBM; what crosses PCIe per step is the residual norm,
per diagnostic dump a few N_x-sized arrays and scalars.  Files carry the reference's names, headers and record layouts
(`Analysis/dataSet.py` reads them); formatted lines use the reference's edit descriptors (es16.8, I10, ...).

The deck-level state of the reference's modules (T_e, T_i, fractionFreq, numDiagnosticSteps, currentTime, ...) is a `SimulationState`."""
import os
import time
import numpy as np
from . import _lib
from . import particle_injection as pinj
from . import scheme as sch
from .nonlin_solvers import solvePotential, SolverState


class SimulationState:
    """constants.f90 / mod_simulation.f90 module variables the time loop reads"""

    def __init__(self, T_e, T_i, n_ave=0.0, fractionFreq=1.0, numDiagnosticSteps=1, directoryName=".", startSimulationTime=0.0,
                 schemeNum=0, numThread=1, Power=0.0, nu_h=0.0):
        self.T_e, self.T_i, self.n_ave, self.fractionFreq = float(T_e), float(T_i), float(n_ave), float(fractionFreq)
        self.numDiagnosticSteps, self.directoryName = int(numDiagnosticSteps), directoryName
        self.startSimulationTime = self.currentTime = float(startSimulationTime)
        self.schemeNum, self.numThread, self.Power, self.nu_h = int(schemeNum), int(numThread), float(Power), float(nu_h)
        self.inelasticEnergyLoss = self.energyError = self.chargeError = self.gaussError = 0.0
        self.steps = 0
        self.nranks = 1                      # ranks the particles are sharded over (flux-type injections are divided by it)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.nranks = dist.get_world_size()
        except Exception:
            pass


def es(x):
    """Fortran es16.8"""
    return "%16.8E" % x


def _record(path, *arrays):
    """one Fortran unformatted sequential record: 4-byte length, payload, 4-byte length"""
    payload = b"".join(np.ascontiguousarray(a, dtype="<f8").tobytes() for a in arrays)
    marker = np.array([len(payload)], dtype="<i4").tobytes()
    with open(path, "wb") as f:
        f.write(marker + payload + marker)


def generateSaveDirectory(dirName, numberBinaryCollisions=0):
    for sub in ("Density", "ElectronTemperature", "PhaseSpace", "Phi", "Temperature", "BinaryCollisions"):
        os.makedirs(os.path.join(dirName, sub), exist_ok=True)


def writeDomain(world, dirName):
    """Domain%writeDomain (NGP mod_domain.f90:327-343: grid, dx_dl, nodeVol, boundaryConditions; CIC mod_domain.f90:314-328: no nodeVol),
    one unformatted record each"""
    def rec(name, a, dt):
        payload = np.ascontiguousarray(a, dtype=dt).tobytes()
        marker = np.array([len(payload)], dtype="<i4").tobytes()
        with open(os.path.join(dirName, name), "wb") as f:
            f.write(marker + payload + marker)
    rec("domainGrid.dat", world.grid, "<f8")
    rec("domainDxDl.dat", world.dx_dl, "<f8")
    if world.scheme != sch.CIC:
        dx = np.asarray(world.dx_dl, dtype=np.float64)
        nodeVol = np.empty(world.NumberXNodes)                  # makeArraysFromGrid (NGP mod_domain.f90:257-261)
        nodeVol[1:-1] = (dx[:-1] + dx[1:]) / 2.0
        nodeVol[0], nodeVol[-1] = 0.5 * dx[0], 0.5 * dx[-1]
        rec("domainNodeVol.dat", nodeVol, "<f8")
    rec("domainBoundaryConditions.dat", world.boundaryConditions, "<i4")


def writePhi(phi, CurrentDiagStep, boolAverage, dirName):
    """:28-46"""
    os.makedirs(os.path.join(dirName, "Phi"), exist_ok=True)
    name = "phi_Average.dat" if boolAverage else "phi_%d.dat" % int(CurrentDiagStep)
    _record(os.path.join(dirName, "Phi", name), phi)


def _interstep(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, currDel_t):
    """what runs between two implicit steps (:161-168): all of it on the device.  `pinj.numFluxPart` is the whole plasma's figure
    (numThread x the reference's per-thread one); under torch.distributed every rank injects its 1 / world_size share of it."""
    nranks = getattr(st, "nranks", 1)
    flux_all = pinj.numFluxPart
    if nranks > 1:
        pinj.numFluxPart = flux_all / float(nranks)
    try:
        _interstep_body(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, currDel_t)
    finally:
        pinj.numFluxPart = flux_all


def _interstep_body(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, currDel_t):
    if pinj.heatingBool:
        pinj.maxwellianHeating(particleList[0], None, st.fractionFreq, st.T_e, currDel_t, del_t)
    if pinj.addLostPartBool:
        pinj.addMaxwellianLostParticles(particleList, st.T_e, st.T_i, None, world)
    if pinj.refluxPartBool:
        pinj.refluxParticles(particleList, st.T_e, st.T_i, None, world)
    if pinj.injectionBool:
        pinj.injectAtBoundary(particleList, st.T_e, st.T_i, None, world, currDel_t)
    if pinj.uniformInjectionBool:
        pinj.injectUniformFlux(particleList, st.T_e, st.T_i, None, world)
    for nc in nullCollisionList:
        nc.generateCollision(particleList, targetParticleList, len(particleList), len(nullCollisionList), None, currDel_t)


def _bind_collisions(solver, targetParticleList, nullCollisionList):
    for t in targetParticleList:
        t._bind(solver)
    for nc in nullCollisionList:
        nc._bind(solver)


def _collision_loss(nullCollisionList, particleList):
    return sum(0.5 * float(np.sum(nc.totalEnergyLoss)) * particleList[nc.reactantsIndx[0] - 1].w_p for nc in nullCollisionList)


def solveSimulation(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, irand, simulationTime):
    """solveSimulation (:48-308).  Returns the rows written to GlobalDiagnosticData.dat as a list of tuples."""
    d = st.directoryName
    generateSaveDirectory(d, len(nullCollisionList))
    solver._bind(particleList)
    with open(os.path.join(d, "InitialConditions.dat"), "w") as f:
        f.write("Scheme, Number Grid Nodes, T_e, T_i, n_ave, Final Expected Time(s), Delta t(s), FractionFreq, Power(W/m^2), smoothInt, nu_h, "
                "numDiag, numThread, RF_rad_frequency, RF_half_amplitude\n")
        f.write("%6d %6d " % (st.schemeNum, world.NumberXNodes) + " ".join(es(x) for x in (st.T_e, st.T_i, st.n_ave, simulationTime, del_t,
                st.fractionFreq, st.Power)) + " %6d " % int(world.gridSmoothBool) + es(st.nu_h) + " %6d %6d\n" % (st.numDiagnosticSteps, st.numThread))
    pinj.writeParticleInjectionInputs(d)
    with open(os.path.join(d, "ParticleProperties.dat"), "w") as f:
        f.write("Particle Symbol, Particle Mass (kg), Particle Charge (C), Particle Weight (N/m^2), maxIdx\n")
        for p in particleList:
            f.write("%s %s %s %s %10d\n" % (p.name, es(p.mass), es(p.q), es(p.w_p), p.finalIdx))
    with open(os.path.join(d, "TargetProperties.dat"), "w") as f:
        f.write("Target Symbol, Target Mass (kg), Target Density (m^-3), Target Temp (K)\n")
        for t in targetParticleList:
            f.write("%s %s %s %s\n" % (t.name, es(t.mass), es(t.density), es(t.temperature)))
    numCells = min(world.NumberXNodes, world.NumberXHalfNodes)
    accumE = [np.zeros(2) for _ in particleList]                 # accumEnergyLoss / accumWallLoss between two dumps
    accumW = [np.zeros(2) for _ in particleList]
    pfiles = []
    for p in particleList:
        p.writeLocalTemperature(0, d, numCells)
        pf = open(os.path.join(d, "ParticleDiagnostic_%s.dat" % p.name), "a")
        pf.write("Time (s), leftCurrentLoss(A/m^2), rightCurrentLoss(A/m^2), leftPowerLoss(W/m^2), rightPowerLoss(W/m^2), Number Particles, "
                 "Temperature (eV), Ave. Num. SubSteps., Ave. Num. Func. Evals.\n")
        pfiles.append(pf)
    tf = open(os.path.join(d, "SimulationTimeData.dat"), "w")
    tf.write("Elapsed Times(s), Total Solver Time (s), Potential Time (s), Mover Time (s), Collision Time (s), Total Steps\n")
    gf = open(os.path.join(d, "GlobalDiagnosticData.dat"), "a")
    gf.write("Time (s), Collision Loss (W/m^2), ParticleCurrentLoss (A/m^2), ParticlePowerLoss(W/m^2), EnergyTotal (J/m^2), MomentumTotal (kg/s/m), "
             "gaussError (a.u), chargeError (a.u), energyError(a.u), Picard Iteration Number\n")
    st.inelasticEnergyLoss = 0.0
    pinj.energyAddColl = 0.0
    _bind_collisions(solver, targetParticleList, nullCollisionList)
    for nc in nullCollisionList:
        nc.reset_totals()
        nc.writeCollisionProperties(particleList, targetParticleList, d)
    diagTimeDivision = (simulationTime - st.startSimulationTime) / float(st.numDiagnosticSteps)
    diagTime = diagTimeDivision + st.startSimulationTime
    sch.loadParticleDensity(particleList, world, True)
    sch.WriteParticleDensity(particleList, world, 0, False, d)
    writePhi(solver.phi, 0, False, d)
    writeDomain(world, d)                                                  # :147
    st.currentTime = pastDiagTime = st.startSimulationTime
    currDel_t = remainDel_t = del_t
    CurrentDiagStep = 1
    potentialTime = collisionTime = 0.0
    startTotal = time.perf_counter()
    rows = []
    i = 0
    while st.currentTime < simulationTime:
        dump = not (st.currentTime + remainDel_t < diagTime)
        if dump:
            KE_i = sum(p.getTotalKE() for p in particleList)
            solver.depositRho(particleList, world)
            rho_i = solver.rho
        t0 = time.perf_counter()
        remainDel_t, currDel_t = solvePotential(solver, particleList, world, del_t, remainDel_t, currDel_t, st.currentTime)
        potentialTime += time.perf_counter() - t0
        for j, p in enumerate(particleList):
            accumE[j] += p.energyLoss
            accumW[j] += p.wallLoss
        if dump:
            PE_i = solver.getTotalPE(world, False)
            KE_f = sum(p.getTotalKE() + float(np.sum(p.energyLoss)) * p.mass * p.w_p * 0.5 for p in particleList)
            PE_f = solver.getTotalPE(world, True) - solver.getEnergyFromBoundary(world, currDel_t)
            st.energyError = abs((KE_i + PE_i - KE_f - PE_f) / (KE_i + PE_i))
        t0 = time.perf_counter()
        _interstep(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, currDel_t)
        collisionTime += time.perf_counter() - t0
        if dump:
            sch.loadParticleDensity(particleList, world, True)
            sch.WriteParticleDensity(particleList, world, CurrentDiagStep, False, d)
            writePhi(solver.phi_f, CurrentDiagStep, False, d)
            solver.depositRho(particleList, world)
            st.chargeError = solver.getChargeContinuityError(rho_i, world, currDel_t)
            st.gaussError = solver.getError_tridiag_Poisson(world)
            span = st.currentTime + currDel_t - pastDiagTime
            Etotal = solver.getTotalPE(world, True)
            chargeTotal = energyLoss = momentum = 0.0
            for j, p in enumerate(particleList):
                p.writeLocalTemperature(CurrentDiagStep, d, numCells)
                if getattr(st, "writeRestartFiles", True):
                    p.writePhaseSpace(d)                                   # :229: the restart files, rewritten at every dump
                momentum += p.getTotalMomentum()[0] + float(np.sum(p.momentumLoss)) * p.mass * p.w_p
                chargeTotal += float(np.sum(accumW[j])) * p.q * p.w_p
                ke = p.getTotalKE()
                Etotal += ke
                energyLoss += float(np.sum(accumE[j])) * p.w_p * p.mass * 0.5
                n = p.count()
                keAve = ke / p.w_p / n / _lib_e() if n > 0 else 0.0          # getKEAve (mod_particle.f90:193-204), eV
                pfiles[j].write(" ".join(es(x) for x in (st.currentTime + currDel_t, accumW[j][0] * p.q * p.w_p / span, accumW[j][1] * p.q * p.w_p / span,
                                accumE[j][0] * p.mass * p.w_p * 0.5 / span, accumE[j][1] * p.mass * p.w_p * 0.5 / span)) + " %10d " % n
                                + " ".join(es(x) for x in (keAve * 2.0 / 3.0, p.numSubStepsAve, p.numFuncEvalAve)) + "\n")
                accumE[j][:] = 0.0
                accumW[j][:] = 0.0
            st.inelasticEnergyLoss += _collision_loss(nullCollisionList, particleList)
            for nc in nullCollisionList:
                nc.writeCollisionDiag(particleList, targetParticleList, d, span)
            row = (st.currentTime + currDel_t, st.inelasticEnergyLoss / span, chargeTotal / span, energyLoss / span, Etotal, momentum,
                   st.gaussError, st.chargeError, st.energyError, SolverState(solver).iterNumPicard)
            gf.write(" ".join(es(x) for x in row[:9]) + " %4d\n" % row[9])
            rows.append(row)
            ss = SolverState(solver)
            tf.write(" ".join(es(x) for x in (time.perf_counter() - startTotal, potentialTime, ss.potTimer, ss.moverTimer, collisionTime)) + " %8d\n" % (i + 1))
            for nc in nullCollisionList:
                nc.reset_totals()
            CurrentDiagStep += 1
            st.inelasticEnergyLoss = 0.0
            pinj.energyAddColl = 0.0
            pastDiagTime = st.currentTime + currDel_t
            diagTime += diagTimeDivision
        st.currentTime += currDel_t
        i += 1
    for f in pfiles + [tf, gf]:
        f.close()
    st.steps = i
    ss = SolverState(solver)
    with open(os.path.join(d, "SimulationFinalData.dat"), "w") as f:
        f.write("Elapsed Times(s), Solver Time (s), Collision Time (s), Total Steps, Number Adaptive Steps\n")
        f.write(" ".join(es(x) for x in (time.perf_counter() - startTotal, potentialTime, collisionTime)) + " %10d %10d\n" % (i + 1, ss.amountTimeSplits))
    return rows


def _lib_e():
    from . import constants as K
    return K.e


def _allreduce_np(dist, arr, op=None):
    """in-place all-reduce of a float64 numpy array over the ranks (NCCL needs device tensors, gloo takes host ones)"""
    import torch
    t = torch.from_numpy(arr)
    op = dist.ReduceOp.SUM if op is None else op
    if dist.get_backend() == "nccl":
        g = t.cuda()
        dist.all_reduce(g, op=op)
        arr[...] = g.cpu().numpy()
    else:
        dist.all_reduce(t, op=op)
    return arr


def energy_grid(Emin, EMax, binNumber):
    """diffE and E_grid(1:binNumber) of mod_simulation.f90:453-461: logarithmic bins, the grid built by repeated addition of diffE
    (real(binNumber - 1) is a single-precision conversion in the reference, exact for these integers)"""
    import math                                    # libm scalars: numpy's vectorised log / exp may differ from them in the last bit
    diffE = (math.log(EMax) - math.log(Emin)) / float(np.float32(binNumber - 1))
    grid = np.zeros(binNumber)
    partV = math.log(Emin)
    for i in range(binNumber):
        grid[i] = math.exp(partV)
        partV = partV + diffE
    return diffE, grid


# ---- EEDF / velocity histograms of the final averaging window (:433-510) -------------------------------------------------
class Histograms:
    """VMax, EMax, Emin, E_grid, diffE per species and the accumulating VHist(2*binNumber), EHist(binNumber)"""

    def __init__(self, particleList, binNumber, dist=None):
        L = _lib.load()
        self.binNumber = int(binNumber)
        n = len(particleList)
        self.VMax, self.EMax, self.Emin = np.zeros(n), np.zeros(n), np.zeros(n)
        for j, p in enumerate(particleList):
            o = np.zeros(3)
            _lib.check(L.ipic_velocity_extrema(p._solver._h, p._index, o.ctypes.data_as(_lib._dp)))
            self.VMax[j], self.EMax[j], self.Emin[j] = o
        if dist is not None:                                  # ranks combine like the OpenMP reduction(max / min) clauses (:442)
            for arr, op in ((self.VMax, dist.ReduceOp.MAX), (self.EMax, dist.ReduceOp.MAX), (self.Emin, dist.ReduceOp.MIN)):
                _allreduce_np(dist, arr, op)
        self.diffE = np.zeros(n)
        self.E_grid = np.zeros((n, self.binNumber))
        for j in range(n):
            self.diffE[j], self.E_grid[j] = energy_grid(self.Emin[j], self.EMax[j], self.binNumber)
        self.VHist = np.zeros((n, 2 * self.binNumber))
        self.EHist = np.zeros((n, self.binNumber))

    def accumulate(self, particleList):
        """one pass of :480-500 over every species"""
        L = _lib.load()
        for j, p in enumerate(particleList):
            _lib.check(L.ipic_velocity_histograms(p._solver._h, p._index, self.binNumber, self.VMax[j], self.Emin[j], self.EMax[j],
                                                  self.VHist[j].ctypes.data_as(_lib._dp), self.EHist[j].ctypes.data_as(_lib._dp)))

    def write(self, particleList, directoryName, dist=None):
        """Temperature/Temp_<name>_average.dat = (VHist, VMax); Temperature/TempEnergy_<name>_average.dat = (EHist, E_grid m / 2e) (:501-508)"""
        K_e = _lib_e()
        VH, EH = self.VHist, self.EHist
        if dist is not None:
            VH, EH = _allreduce_np(dist, VH.copy()), _allreduce_np(dist, EH.copy())
        os.makedirs(os.path.join(directoryName, "Temperature"), exist_ok=True)
        for j, p in enumerate(particleList):
            _record(os.path.join(directoryName, "Temperature", "Temp_%s_average.dat" % p.name), VH[j], np.array([self.VMax[j]]))
            _record(os.path.join(directoryName, "Temperature", "TempEnergy_%s_average.dat" % p.name), EH[j], self.E_grid[j] * 0.5 * p.mass / K_e)


def solveSimulationFinalAverage(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, irand, averagingTime, binNumber):
    """solveSimulationFinalAverage (:310-513): time-averaged density / phi over `averagingTime`, then the EEDF window (one RF period or
    50 del_t / fractionFreq).  Returns (steps averaged, Histograms)."""
    d = st.directoryName
    generateSaveDirectory(d, len(nullCollisionList))
    solver._bind(particleList)
    _bind_collisions(solver, targetParticleList, nullCollisionList)
    for nc in nullCollisionList:
        nc.reset_totals()
    pinj.energyAddColl = 0.0
    accumE = [np.zeros(2) for _ in particleList]
    accumW = [np.zeros(2) for _ in particleList]
    phi_average = np.zeros(world.NumberXNodes)
    for p in particleList:
        p.densities = np.zeros(world.NumberXNodes)                              # :331-335; every load below accumulates (reset = .false.)
    startTime = st.currentTime
    currDel_t = remainDel_t = del_t
    i = 0
    # wall-loss convergence check (:348-352, 379-386): every 200 del_t / fractionFreq the running mean power to the walls is tested for
    # a relative standard deviation below 1e-6 over the window
    lastCheckTime = startTime
    checkTimeDivision = 200.0 * del_t / st.fractionFreq
    window_loss = []
    RF_ave = 0.0
    bc = world.boundaryConditions
    while st.currentTime - startTime < averagingTime:
        remainDel_t, currDel_t = solvePotential(solver, particleList, world, del_t, remainDel_t, currDel_t, st.currentTime)
        for j, p in enumerate(particleList):
            accumE[j] += p.energyLoss
            accumW[j] += p.wallLoss
        _interstep(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, currDel_t)
        sch.loadParticleDensity(particleList, world, False)
        phi_f = solver.phi_f                                                    # aveRFVoltage(.true., ...) (NGP mod_potentialSolver.f90:469-475)
        phi_average += phi_f
        if bc[0] == 4:
            RF_ave += phi_f[0]
        elif bc[-1] == 4:
            RF_ave += phi_f[-1]
        st.currentTime += currDel_t
        i += 1
        window_loss.append(sum(float(np.sum(p.energyLoss)) for p in particleList) / (st.currentTime - startTime))
        if st.currentTime - lastCheckTime > checkTimeDivision:
            w = np.array(window_loss)
            meanLoss = w.sum() / float(w.size)
            stdLoss = np.sqrt(np.sum((w - meanLoss) ** 2) / float(w.size))
            if meanLoss != 0.0 and stdLoss / meanLoss < 1e-6:
                break
            window_loss = []
            lastCheckTime = st.currentTime
    for p in particleList:
        p.densities = p.densities / float(i)                                    # :389
    phi_average = phi_average / float(i)                                        # aveRFVoltage(.false., ...): phi_f becomes the average (:477-481)
    solver.phi_f = phi_average
    st.RF_ave = RF_ave / float(i) if getattr(solver, "RF_rad_frequency", 0.0) else RF_ave
    writePhi(phi_average, 0, True, d)
    span = st.currentTime - startTime
    chargeLossTotal = sum(float(np.sum(accumW[j])) * p.q_times_wp for j, p in enumerate(particleList))
    ELossTotal = sum(float(np.sum(accumE[j])) * p.mass * p.w_p * 0.5 for j, p in enumerate(particleList))
    st.inelasticEnergyLoss = _collision_loss(nullCollisionList, particleList)
    for nc in nullCollisionList:                                                # BinaryCollisions/<a>_on_<b>/AveCollisionDiag.dat (:401-412)
        nc.writeAveCollisionDiag(particleList, targetParticleList, d, span)
    dens = sch.WriteParticleDensity(particleList, world, 0, True, d)
    # gaussError of the averaged state (:413-417): rho from the averaged densities against the averaged potential
    # (WriteParticleDensity leaves densities(:, 1) = thread-summed weights * w_p, periodic ends folded -- NGP mod_Scheme.f90:102-108, CIC alike)
    rho_ave = np.zeros(world.NumberXNodes)
    for p in particleList:
        D = np.asarray(p.densities, dtype=np.float64) * p.w_p
        if world.scheme != sch.CIC and world.boundaryConditions[0] == 3:
            D = D.copy()
            D[0] = D[0] + D[-1]
            D[-1] = D[0]
        rho_ave = rho_ave + D * p.q
    solver.rho = rho_ave
    st.gaussError = solver.getError_tridiag_Poisson(world)
    with open(os.path.join(d, "GlobalDiagnosticDataAveraged.dat"), "w") as f:
        f.write("Steps Averaged, Averaging Time, Collision Loss (W/m^2), ParticleCurrentLoss (A/m^2), ParticlePowerLoss(W/m^2), gaussError\n")
        f.write("%10d " % i + " ".join(es(x) for x in (span, st.inelasticEnergyLoss / span, chargeLossTotal / span, ELossTotal / span, st.gaussError)) + "\n")
    hist = Histograms(particleList, binNumber)
    # window: one RF period, else 50 del_t / fractionFreq (:463-467)
    window = 50.0 * del_t / st.fractionFreq
    if getattr(solver, "RF_rad_frequency", 0.0):
        window = 2.0 * np.pi / solver.RF_rad_frequency
    startTime = st.currentTime
    while st.currentTime - startTime < window:
        remainDel_t, currDel_t = solvePotential(solver, particleList, world, del_t, remainDel_t, currDel_t, st.currentTime)
        _interstep(st, solver, particleList, targetParticleList, nullCollisionList, world, del_t, currDel_t)
        hist.accumulate(particleList)
        st.currentTime += currDel_t
    hist.write(particleList, d)
    return i, hist, dens
