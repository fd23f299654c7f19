"""ctypes binding of the C ABI (include/ipic1d_b200.h).  There is no CPU fallback: if the CUDA library is
missing or a call fails, this raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IPIC1D_B200_LIB: another build of the same library (A/B measurements of kernel variants); default: the in-tree build
LIB_PATH = os.environ.get("IPIC1D_B200_LIB") or os.path.join(_HERE, "lib", "libipic1d_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_STOP, ERR_NOT_CONVERGED, ERR_UNSUPPORTED, ERR_NCCL = range(7)
SCHEME_NGP, SCHEME_CIC, SCHEME_NGP_PERIODIC = 0, 1, 2
F_PHI, F_PHI_F, F_J, F_RHO, F_EFIELD, F_A_TRI, F_B_TRI, F_C_TRI, F_EFIELD_F, F_EFIELD_HALF = range(10)
TRIDIAG_PCR, TRIDIAG_THOMAS, TRIDIAG_SCAN = 0, 1, 2


class MoveStats(C.Structure):
    _fields_ = [("N_p", C.c_int64), ("delIdx", C.c_int64), ("refIdx", C.c_int64), ("numSubSteps", C.c_int64),
                ("numFuncEvals", C.c_int64), ("numSubStepsAve", C.c_double), ("numFuncEvalAve", C.c_double),
                ("wallLoss", C.c_double * 2), ("energyLoss", C.c_double * 2), ("momentumLoss", C.c_double * 2),
                ("accumEnergyLoss", C.c_double * 2), ("accumWallLoss", C.c_int64 * 2)]


class SolverStats(C.Structure):
    _fields_ = [("iterNumPicard", C.c_int32), ("iterNumAdaptiveSteps", C.c_int32), ("amountTimeSplits", C.c_int32),
                ("sweeps", C.c_int64), ("substeps", C.c_int64), ("moverSeconds", C.c_double), ("potSeconds", C.c_double),
                ("lastResidualNorm", C.c_double)]


MAX_COLLISIONS = 16


class CollisionStats(C.Structure):
    _fields_ = [("numberCollisions", C.c_int32), ("lengthArrays", C.c_int32), ("collisionType", C.c_int32 * MAX_COLLISIONS),
                ("energyThreshold", C.c_double * MAX_COLLISIONS), ("sigmaVMax", C.c_double), ("reducedMass", C.c_double),
                ("sumMass", C.c_double), ("reducedMassIonization", C.c_double), ("minEnergy", C.c_double), ("maxEnergy", C.c_double),
                ("P_null", C.c_double), ("numberSelected", C.c_int64), ("ionizations", C.c_int64),
                ("totalAmountCollisions", C.c_int64 * MAX_COLLISIONS), ("totalEnergyLoss", C.c_double * MAX_COLLISIONS),
                ("totalIncidentEnergy", C.c_double * MAX_COLLISIONS), ("totalNumberCollidableParticles", C.c_int64)]


class IpicError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ipic1d_b200 error %d: %s" % (code, msg))
        self.code = code


_d, _i, _i32, _i64, _p = C.c_double, C.c_int, C.c_int32, C.c_int64, C.c_void_p
_dp, _ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)

# every symbol include/ipic1d_b200.h declares, with its signature (tests check the export list against the header)
SIGNATURES = {
    "ipic_version": (C.c_char_p, []),
    "ipic_last_error": (C.c_char_p, []),
    "ipic_device_count": (_i, [C.POINTER(_i)]),
    "ipic_create": (_i, [C.POINTER(_p), _i, _i32, _dp, _dp, _ip, _i, _i]),
    "ipic_destroy": (_i, [_p]),
    "ipic_init_solver": (_i, [_p, _d, _d, _d, _d]),
    "ipic_set_tridiag_mode": (_i, [_p, _i]),
    "ipic_add_species": (_i, [_p, _d, _d, _d, _ip]),
    "ipic_set_weight": (_i, [_p, _i32, _d]),
    "ipic_upload_particles": (_i, [_p, _i32, _dp, _ip, _i32, _i32]),
    "ipic_download_particles": (_i, [_p, _i32, _dp, _ip, _i32, _i32]),
    "ipic_particle_count": (_i, [_p, _i32, C.POINTER(_i64)]),
    "ipic_particle_device_pointers": (_i, [_p, _i32, C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_dp), C.POINTER(_i64)]),
    "ipic_generate_particles": (_i, [_p, _i32, _i64, _d, _d, C.c_uint64]),
    "ipic_generate_particles_dist": (_i, [_p, _i32, _i64, _d, _d, _d, _i, C.c_uint64]),
    "ipic_sort_particles": (_i, [_p, _i32]),
    "ipic_get_field": (_i, [_p, _i, _dp, _i32]),
    "ipic_set_field": (_i, [_p, _i, _dp, _i32]),
    "ipic_deposit_rho": (_i, [_p]),
    "ipic_solve_tridiag_poisson": (_i, [_p, _d]),
    "ipic_solve_initial_potential": (_i, [_p, _d]),
    "ipic_make_const_source_term": (_i, [_p]),
    "ipic_make_half_time_efield": (_i, [_p]),
    "ipic_solve_tridiag_ampere": (_i, [_p, _d]),
    "ipic_set_rf_voltage": (_i, [_p, _d]),
    "ipic_reset_voltage": (_i, [_p]),
    "ipic_set_cic_walls": (_i, [_p, _dp, _d, _dp]),
    "ipic_deposit_J": (_i, [_p, _d]),
    "ipic_move_particles": (_i, [_p, _d]),
    "ipic_get_move_stats": (_i, [_p, _i32, C.POINTER(MoveStats)]),
    "ipic_get_deposit_substeps": (_i, [_p, _i32, C.POINTER(_i64)]),
    "ipic_get_ref_record": (_i, [_p, _i32, _ip, _i64, C.POINTER(_i64)]),
    "ipic_get_ref_record_counts": (_i, [_p, _i32, _ip, _i32]),
    "ipic_initialize_solver": (_i, [_p, _d, _d, _i, _i, _d, _i]),
    "ipic_solve_div_ampere_anderson": (_i, [_p, _d]),
    "ipic_solve_potential": (_i, [_p, _d, _dp, _dp, _d]),
    "ipic_func_nitsol": (_i, [_p, _i32, _dp, _dp, _d]),
    "ipic_solve_jfnk": (_i, [_p, _d]),
    "ipic_get_nitsol_info": (_i, [_p, _ip]),
    "ipic_nitsol_host": (_i, [_i32, _dp, C.CFUNCTYPE(C.c_int, C.c_int32, _dp, _dp, C.c_void_p), C.c_void_p, _d, _d, _d, _i32, _i32, _ip, _ip]),
    "ipic_get_solver_stats": (_i, [_p, C.POINTER(SolverStats)]),
    "ipic_get_norm_history": (_i, [_p, _dp, _i32, _ip]),
    "ipic_get_total_PE": (_i, [_p, _i, _dp]),
    "ipic_get_JdotE": (_i, [_p, _d, _dp]),
    "ipic_get_energy_from_boundary": (_i, [_p, _d, _dp]),
    "ipic_get_error_tridiag_poisson": (_i, [_p, _dp]),
    "ipic_get_charge_continuity_error": (_i, [_p, _dp, _d, _dp]),
    "ipic_get_total_KE": (_i, [_p, _i32, _dp]),
    "ipic_get_total_momentum": (_i, [_p, _i32, _dp]),
    "ipic_local_temperature": (_i, [_p, _i32, _i32, _dp]),
    "ipic_add_target_particle": (_i, [_p, _d, _d, _d, _ip]),
    "ipic_null_collision_create": (_i, [_p, _i32, _i32, _d, _d, _d, _dp, _dp, _dp, _ip, _ip, _ip, _ip, C.c_uint64, _ip]),
    "ipic_generate_collision": (_i, [_p, _i32, _d]),
    "ipic_get_collision_stats": (_i, [_p, _i32, C.POINTER(CollisionStats)]),
    "ipic_reset_collision_stats": (_i, [_p, _i32]),
    "ipic_set_num_to_collide": (_i, [_p, _i32, _i64]),
    "ipic_velocity_extrema": (_i, [_p, _i32, _dp]),
    "ipic_velocity_histograms": (_i, [_p, _i32, _i32, _d, _d, _d, _dp, _dp]),
    "ipic_load_particle_density": (_i, [_p, _i32, _dp]),
    "ipic_set_injection_seed": (_i, [_p, C.c_uint64]),
    "ipic_add_maxwellian_lost_particles": (_i, [_p, _d, _d, _dp, C.POINTER(_i64)]),
    "ipic_reflux_particles": (_i, [_p, _d, _d, _dp]),
    "ipic_inject_at_boundary": (_i, [_p, _d, _d, _d, _d, C.POINTER(_i64)]),
    "ipic_inject_uniform_flux": (_i, [_p, _d, _d, _d, C.POINTER(_i64)]),
    "ipic_maxwellian_heating": (_i, [_p, _i32, _d, _d, C.POINTER(_i64)]),
    "ipic_mcc_permutation": (_i, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]),
    "ipic_nccl_unique_id": (_i, [_p]),
    "ipic_comm_init": (_i, [_p, _i32, _i32, _p]),
    "ipic_set_option": (_i, [_p, C.c_char_p, _d]),
    "ipic_peer_export": (_i, [_p, _p]),
    "ipic_peer_import": (_i, [_p, _i32, _i32, _p]),
    "ipic_set_stream": (_i, [_p, _p]),
    "ipic_get_launch_count": (_i, [_p, C.POINTER(_i64)]),
    "ipic_host_alloc": (_i, [C.POINTER(_p), C.c_uint64]),
    "ipic_host_free": (_i, [_p]),
    "ipic_get_timing": (_i, [_p, C.c_char_p, _dp, C.POINTER(_i64)]),
    "ipic_reset_timing": (_i, [_p]),
    "ipic_synchronize": (_i, [_p]),
}

_lib = None


def load():
    """Loads the CUDA library.  Raises if it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("ipic1d_b200: %s is missing -- build it with __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise IpicError(rc, load().ipic_last_error().decode(errors="replace"))
