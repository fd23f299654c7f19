"""Host-side `type(Particle)` (OpenMP/SharedModules/src/mod_particle.f90:15-74).

Holds the species scalars and, optionally, a host copy of phaseSpace in the reference's slab layout
(numThread, finalIdx, 4) == Fortran phaseSpace(4, finalIdx, numThread).  The live particles are device
resident; `upload`/`download` move them through the C ABI."""
import ctypes as C
import numpy as np
from . import _lib


class Particle:
    def __init__(self, mass, q, w_p, N_p, finalIdx, particleName, numThread):
        """particle_constructor (mod_particle.f90:46-74); N_p is particles per thread slab."""
        self.name = particleName
        self.mass, self.q, self.w_p = float(mass), float(q), float(w_p)
        self.q_over_m = self.q / self.mass
        self.q_times_wp = self.q * self.w_p
        self.finalIdx, self.numThread = int(finalIdx), int(numThread)
        self.N_p = np.full(self.numThread, int(N_p), dtype=np.int32)
        self.phaseSpace = None            # host copy, allocated on demand
        self.delIdx = 0
        self.refIdx = 0
        self.wallLoss = np.zeros(2)
        self.energyLoss = np.zeros(2)
        self.momentumLoss = np.zeros(2)
        self.accumEnergyLoss = np.zeros(2)
        self.accumWallLoss = np.zeros(2, dtype=np.int64)
        self.numSubStepsAve = 0.0
        self.numFuncEvalAve = 0.0
        self._solver = None
        self._index = -1

    def initialize_n_ave(self, n_ave, L_domain):
        """mod_particle.f90:76-82"""
        self.w_p = n_ave * L_domain / float(np.sum(self.N_p))
        self.q_times_wp = self.q * self.w_p
        if self._solver is not None:
            _lib.check(_lib.load().ipic_set_weight(self._solver._h, self._index, self.w_p))

    def set_phase_space(self, phaseSpace, N_p):
        ps = np.ascontiguousarray(phaseSpace, dtype=np.float64)
        if ps.shape != (self.numThread, self.finalIdx, 4):
            raise ValueError("phaseSpace must have shape (numThread, finalIdx, 4)")
        self.phaseSpace = ps
        self.N_p = np.ascontiguousarray(N_p, dtype=np.int32).copy()

    # ---- restart files (the reference's particle interchange format)
    def _thread_file(self, dirName, iThread):
        # file name of mod_particle.f90:243 / :375 (write(char_i,'(I3)') + trim(adjustl()) -> plain decimal thread number, 1-based)
        import os
        return os.path.join(dirName, "PhaseSpace", "phaseSpace_%s_thread%d.dat" % (self.name, iThread))

    def writePhaseSpace(self, dirName, download=True):
        """writePhaseSpace (mod_particle.f90:234-249): one unformatted stream file per thread slab holding phaseSpace(:, 1:N_p(t), t),
        i.e. N_p(t) records of (l, v_x, v_y, v_z) as raw little-endian float64.  Pulls the particles from the device first."""
        import os
        if download and self._solver is not None:
            self.download()
        os.makedirs(os.path.join(dirName, "PhaseSpace"), exist_ok=True)
        for t in range(self.numThread):
            np.ascontiguousarray(self.phaseSpace[t, :int(self.N_p[t]), :], dtype="<f8").tofile(self._thread_file(dirName, t + 1))

    def readPhaseSpace(self, restartDirectory, upload=True):
        """The restart branch of readChargedParticleInputs (mod_particle.f90:371-390): every thread reads its own file until end of file;
        N_p(t) = number of complete 4-double records.  A missing file stops the reference ('Does not exist'); here it raises."""
        import os
        if self.phaseSpace is None:
            self.phaseSpace = np.zeros((self.numThread, self.finalIdx, 4))
        for t in range(self.numThread):
            fn = self._thread_file(restartDirectory, t + 1)
            if not os.path.exists(fn):
                raise FileNotFoundError(fn + " Does not exist")
            raw = np.fromfile(fn, dtype="<f8")
            n = raw.size // 4
            if n > self.finalIdx:
                raise ValueError("%s holds %d particles, more than finalIdx = %d" % (fn, n, self.finalIdx))
            self.phaseSpace[t, :n, :] = raw[:4 * n].reshape(n, 4)
            self.N_p[t] = n
        if upload and self._solver is not None:
            self.upload()

    # ---- device residency
    def upload(self):
        if self._solver is None:
            raise RuntimeError("particle is not bound to a potentialSolver")
        L = _lib.load()
        _lib.check(L.ipic_upload_particles(self._solver._h, self._index, self.phaseSpace.ctypes.data_as(_lib._dp),
                                           self.N_p.ctypes.data_as(_lib._ip), self.numThread, self.finalIdx))

    def download(self):
        if self.phaseSpace is None:
            self.phaseSpace = np.zeros((self.numThread, self.finalIdx, 4))
        L = _lib.load()
        _lib.check(L.ipic_download_particles(self._solver._h, self._index, self.phaseSpace.ctypes.data_as(_lib._dp),
                                             self.N_p.ctypes.data_as(_lib._ip), self.numThread, self.finalIdx))
        return self.phaseSpace

    def count(self):
        n = C.c_int64()
        _lib.check(_lib.load().ipic_particle_count(self._solver._h, self._index, C.byref(n)))
        return n.value

    def device_arrays(self):
        """zero-copy torch views (l, v_x, v_y, v_z) of the device-resident arrays; valid until the next call that moves, sorts,
        compacts, uploads or grows the species"""
        import torch
        ptr = [_lib._dp() for _ in range(4)]
        n = C.c_int64()
        _lib.check(_lib.load().ipic_particle_device_pointers(self._solver._h, self._index, *[C.byref(p) for p in ptr], C.byref(n)))

        class _View:
            def __init__(self, addr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (addr, False), "version": 2}
        dev = torch.device("cuda", self._solver.device)
        return tuple(torch.as_tensor(_View(C.cast(p, C.c_void_p).value, n.value), device=dev) for p in ptr)

    # ---- mod_particle.f90:206-230
    def getTotalKE(self):
        out = C.c_double()
        _lib.check(_lib.load().ipic_get_total_KE(self._solver._h, self._index, C.byref(out)))
        return out.value

    def getTotalMomentum(self):
        out = np.zeros(3)
        _lib.check(_lib.load().ipic_get_total_momentum(self._solver._h, self._index, out.ctypes.data_as(_lib._dp)))
        return out

    def localTemperature(self, numCells):
        """the array writeLocalTemperature writes (mod_particle.f90:251-283): per cell j = INT(l) < N_x the temperature
        SUM(v^2) mass / count / 3 / e in eV, 0 where the cell holds no particle"""
        out = np.zeros(int(numCells))
        _lib.check(_lib.load().ipic_local_temperature(self._solver._h, self._index, int(numCells), out.ctypes.data_as(_lib._dp)))
        return out

    def writeLocalTemperature(self, CurrentDiagStep, dirName, numCells):
        """writeLocalTemperature (mod_particle.f90:251-283): Temperature/Temp_<name>_<step>.dat, one Fortran unformatted sequential
        record (4-byte length, numCells float64, 4-byte length) -- what Analysis/dataSet.py reads"""
        import os
        EHist = self.localTemperature(numCells)
        os.makedirs(os.path.join(dirName, "Temperature"), exist_ok=True)
        fn = os.path.join(dirName, "Temperature", "Temp_%s_%d.dat" % (self.name, int(CurrentDiagStep)))
        with open(fn, "wb") as f:
            marker = np.array([EHist.nbytes], dtype="<i4").tobytes()
            f.write(marker + EHist.astype("<f8").tobytes() + marker)
        return EHist

    def refRecord(self):
        """refIdx(T), refRecordIdx(:, T) of the last moveParticles (mod_particle.f90:18-19; mod_particleMover.f90:226-231): per thread slab, the
        1-based post-compaction indices of the particles first reflected at a Neumann wall in the step, in particle order.  The slabs are
        those `download()` writes (N_p is refreshed by it); returns (refIdx per slab, list of int32 arrays)."""
        L = _lib.load()
        n = C.c_int64()
        _lib.check(L.ipic_get_ref_record(self._solver._h, self._index, None, 0, C.byref(n)))
        idx = np.zeros(max(n.value, 1), dtype=np.int32)
        _lib.check(L.ipic_get_ref_record(self._solver._h, self._index, idx.ctypes.data_as(_lib._ip), n.value, C.byref(n)))
        idx = idx[:n.value].astype(np.int64)
        cnt = np.zeros(self.numThread, dtype=np.int32)
        _lib.check(L.ipic_get_ref_record_counts(self._solver._h, self._index, cnt.ctypes.data_as(_lib._ip), self.numThread))
        off = np.concatenate([[0], np.cumsum(self.N_p.astype(np.int64))])
        start = np.concatenate([[0], np.cumsum(cnt.astype(np.int64))])
        rec = [(idx[start[t]:start[t + 1]] - off[t] + 1).astype(np.int32) for t in range(self.numThread)]
        return cnt, rec

    def _refresh_stats(self):
        st = _lib.MoveStats()
        _lib.check(_lib.load().ipic_get_move_stats(self._solver._h, self._index, C.byref(st)))
        self.delIdx, self.refIdx = st.delIdx, st.refIdx
        self.wallLoss = np.array(st.wallLoss[:])
        self.energyLoss = np.array(st.energyLoss[:])
        self.momentumLoss = np.array(st.momentumLoss[:])
        self.accumEnergyLoss = np.array(st.accumEnergyLoss[:])
        self.accumWallLoss = np.array(st.accumWallLoss[:], dtype=np.int64)
        self.numSubStepsAve, self.numFuncEvalAve = st.numSubStepsAve, st.numFuncEvalAve
        self.numSubSteps, self.numFuncEvals = st.numSubSteps, st.numFuncEvals
        self.N_p_total = st.N_p
        return st
