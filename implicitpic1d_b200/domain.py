"""Host-side `type(Domain)`: the grid tables the reference builds in readWorld / constructGrid
(OpenMP/NGP/src/mod_domain.f90:46-265,347-391 ; OpenMP/CIC/src/mod_domain.f90:43-252,332-376).

The reference keeps this on the host (deck reader + a few hundred table entries); only the tables are copied
to the device by potentialSolver.  Arrays are 0-based numpy arrays: element k is Fortran element k+1."""
import math
import numpy as np
from . import constants as K

NGP, CIC, NGP_PERIODIC = 0, 1, 2


class Domain:
    def __init__(self, scheme, NumberXNodes, leftBoundary, rightBoundary):
        if scheme not in (NGP, CIC, NGP_PERIODIC):
            raise ValueError("scheme must be 0 (NGP), 1 (CIC) or 2 (NGP_Periodic, E-field form)")
        self.scheme = scheme
        self._ngp = scheme != CIC          # NGP_Periodic uses the NGP grid (OpenMP/NGP_Periodic/src/mod_domain.f90)
        self.NumberXNodes = int(NumberXNodes)
        self.NumberXHalfNodes = self.NumberXNodes - 1 if scheme != CIC else self.NumberXNodes + 1
        Nx, Nh = self.NumberXNodes, self.NumberXHalfNodes
        nb = Nx if scheme != CIC else Nh
        self.grid = np.arange(1, Nx + 1, dtype=np.float64)
        self.dx_dl = np.ones(Nh if scheme != CIC else Nx)
        self.boundaryConditions = np.zeros(nb, dtype=np.int32)
        self.boundaryConditions[0] = leftBoundary
        self.boundaryConditions[-1] = rightBoundary
        if leftBoundary == 3 or rightBoundary == 3:
            self.boundaryConditions[0] = 3
            self.boundaryConditions[-1] = 3
        self.gridSmoothBool = False
        self.L_domain = float(Nx - 1)
        self.startX, self.endX = 1.0, float(Nx)

    # ---- constructGrid: NGP :104-126 ; CIC :103-126
    def constructGrid(self, del_x, L_domain, gridType):
        Nx, Nh = self.NumberXNodes, self.NumberXHalfNodes
        pi = K.pi
        if self._ngp:
            g = np.zeros(Nx)
            g[Nx - 1] = L_domain
            i = np.arange(2, Nh + 1, dtype=np.float64)     # Fortran index of interior nodes
            if gridType == 0:
                g[1:Nh] = (i - 1) * L_domain / Nh
            elif gridType == 1:
                self._check_debye(del_x, L_domain, Nx)
                g[1:Nh] = L_domain * ((i - 1.0) / (Nx - 1.0) - (1.0 / (Nx - 1.0) - del_x / L_domain)
                                      * np.sin(2 * pi * (i - 1) / (Nx - 1)) / math.sin(2 * pi / (Nx - 1)))
            elif gridType == 2:
                self._check_debye(del_x, L_domain, Nx)
                self._no_periodic()
                g[Nx - 1] = L_domain / 2.0
                g[1:Nh] = L_domain * ((i - 1.0) / (Nx - 1.0) / 2.0 - (1.0 / (Nx - 1.0) / 2.0 - del_x / L_domain)
                                      * np.sin(pi * (i - 1) / (Nx - 1.0)) / math.sin(pi / (Nx - 1.0)))
            elif gridType == 3:
                F = 100.0 * del_x
                if F / L_domain > 0.5:
                    raise ValueError("Debye length too small for exponential increase")
                self._no_periodic()
                k = 2.0 * math.log(L_domain / F - 1.0) / float(Nh)
                A = L_domain / (math.exp(k * float(Nh)) - 1.0)
                g[1:Nh] = A * (np.exp(k * (i - 1)) - 1.0)
            elif gridType == 5:
                self._check_debye(del_x, L_domain, Nx)
                phase = (np.float32(i - 1) / np.float32(Nh)).astype(np.float64)   # real(i-1)/real(N_h) is single precision
                g[1:Nh] = L_domain * (phase + (1.0 - float(Nx) * del_x / L_domain) * np.sin(2.0 * pi * phase) / 2.0 / pi)
            else:
                raise ValueError("Gridtype %d doesn't exist!" % gridType)
            self.grid = g
            self.dx_dl = g[1:] - g[:-1]
            self.startX, self.endX = g[0], g[-1]
            self.L_domain = self.endX - self.startX
        else:
            NN = Nx + 1
            gf = np.zeros(NN)
            gf[NN - 1] = L_domain
            i = np.arange(2, NN, dtype=np.float64)
            if gridType == 0:
                gf[1:NN - 1] = (i - 1) * L_domain / (NN - 1)
            elif gridType == 1:
                self._check_debye(del_x, L_domain, NN)
                gf[1:NN - 1] = L_domain * ((i - 1.0) / (NN - 1.0) - (1.0 / (NN - 1.0) - del_x / L_domain)
                                           * np.sin(2 * pi * (i - 1) / (NN - 1)) / math.sin(2 * pi / (NN - 1)))
            elif gridType == 2:
                self._check_debye(del_x, L_domain, NN)
                gf[NN - 1] = L_domain / 2.0
                gf[1:NN - 1] = L_domain * ((i - 1.0) / (NN - 1.0) / 2.0 - (1.0 / (NN - 1.0) / 2.0 - del_x / L_domain)
                                           * np.sin(pi * (i - 1) / (NN - 1.0)) / math.sin(pi / (NN - 1.0)))
            elif gridType == 5:
                self._check_debye(del_x, L_domain, Nh)
                phase = (np.float32(i - 1) / np.float32(Nx)).astype(np.float64)
                gf[1:NN - 1] = L_domain * (phase + (1.0 - float(Nh) * del_x / L_domain) * np.sin(2.0 * pi * phase) / 2.0 / pi)
            else:
                raise ValueError("Gridtype %d doesn't exist!" % gridType)   # type 3 reads unset memory in the reference
            self.dx_dl = gf[1:] - gf[:-1]
            self.grid = 0.5 * (gf[:-1] + gf[1:])
            self.startX = self.grid[0] - 0.5 * self.dx_dl[0]
            self.endX = self.grid[-1] + 0.5 * self.dx_dl[-1]
            self.L_domain = self.endX - self.startX
        return self

    @staticmethod
    def _check_debye(del_x, L_domain, n):
        if del_x / L_domain >= 1.0 / (float(n) - 1.0):
            raise ValueError("The debyeLength is really large, less nodes needed!")

    def _no_periodic(self):
        if self.boundaryConditions[0] == 3:
            raise ValueError("Mesh is not periodic, cannot have periodic boundary!")

    # ---- getLFromX: NGP :275-296 ; CIC :254-279
    def getLFromX(self, x):
        lo, hi = 1, self.NumberXNodes
        g, dx = self.grid, self.dx_dl
        if self._ngp:
            if x < g[0] or x > g[-1]:
                raise ValueError("x value outside of grid range in getLFromX!")
            while lo != hi - 1:
                mid = (lo + hi) // 2
                if g[mid - 1] <= x:
                    lo = mid
                else:
                    hi = mid
            return lo + (x - g[lo - 1]) / dx[lo - 1]
        if x < g[0] - 0.5 * dx[0] or x > g[-1] + 0.5 * dx[-1]:
            raise ValueError("x value outside of grid range in getLFromX!")
        while lo != hi - 1:
            mid = (lo + hi) // 2
            if g[mid - 1] - 0.5 * dx[mid - 1] <= x:
                lo = mid
            else:
                hi = mid
        if g[lo - 1] + 0.5 * dx[lo - 1] <= x:
            lo = hi
        return lo + 0.5 + (x - g[lo - 1]) / dx[lo - 1]

    # ---- getXFromL: NGP :267-273 ; CIC :281-287
    def getXFromL(self, l):
        il = int(l)
        if self._ngp:
            return self.grid[il - 1] + self.dx_dl[il - 1] * (l - il)
        return self.grid[il - 1] + self.dx_dl[il - 1] * (l - il - 0.5)


def readWorld(scheme, NumberXNodes, L_domain, gridType, del_x_deck, leftBoundary, rightBoundary, smoothInt, T_e, n_ave):
    """readWorld with the deck values passed in (NGP :347-391 ; CIC :332-376)."""
    debyeLength = max(0.1 * K.getDebyeLength(T_e, n_ave), del_x_deck)
    if leftBoundary == 3 or rightBoundary == 3:
        leftBoundary = rightBoundary = 3
    world = Domain(scheme, NumberXNodes, leftBoundary, rightBoundary)
    world.constructGrid(debyeLength, L_domain, gridType)
    world.gridSmoothBool = smoothInt != 0
    return world
