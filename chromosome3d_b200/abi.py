"""ctypes mirror of include/chr3d.h (struct layouts and enums only, no library loading).

Kept separate from the loader so that layout checks run without a GPU.
"""
import ctypes as C

LBFGS_MAX_M = 8

OP_INIT, OP_MINIMIZE, OP_MAXWELL, OP_DYNAMICS, OP_SAVE, OP_PICK, OP_CENTER, OP_ENERGY = range(1, 9)
COUPLE_TCOUP, COUPLE_VSCALE = 1, 2
E_OVERALL, E_BON, E_VDW, E_NOE, E_COUNT = 0, 1, 2, 3, 4

ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_CAPACITY, ERR_IO, ERR_SIZE = -1, -2, -3, -4, -5, -6


class SaParams(C.Structure):
    """chr3d_sa_params (include/chr3d.h); defaults come from chr3d_default_params()."""
    _fields_ = [
        ("noe_scale", C.c_double), ("noe_sqconst", C.c_double), ("noe_rswitch", C.c_double),
        ("noe_asymptote", C.c_double), ("noe_low_soft", C.c_int32), ("vdw_min_sep", C.c_int32),
        ("bond_k", C.c_double), ("bond_r0", C.c_double), ("vdw_rmin", C.c_double),
        ("mass", C.c_double), ("fbeta", C.c_double), ("maxwell_temp", C.c_double),
        ("hot_temp", C.c_double), ("hot_step", C.c_int32), ("cool_step", C.c_int32),
        ("hot_dt", C.c_double), ("cool_dt", C.c_double), ("cool_vdw_init", C.c_double),
        ("cool_vdw_finl", C.c_double), ("cool_rad_init", C.c_double), ("cool_rad_finl", C.c_double),
        ("cool_tmpstp", C.c_double), ("pow_step", C.c_int32), ("pow_cycl", C.c_int32),
        ("premin_step", C.c_int32), ("images", C.c_int32), ("lbfgs_m", C.c_int32),
        ("lbfgs_max_ls", C.c_int32), ("lbfgs_c1", C.c_double), ("lbfgs_gtol", C.c_double), ("lbfgs_ftol", C.c_double),
        ("pow_drop", C.c_double), ("seed", C.c_uint64),
        ("angle_d2", C.c_double), ("angle_k", C.c_double),
    ]

    def copy(self):
        out = SaParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(SaParams))
        return out


class Stage(C.Structure):
    """chr3d_stage: one instruction of the annealing stage program."""
    _fields_ = [
        ("op", C.c_int32), ("nstep", C.c_int32), ("flag", C.c_int32), ("pad_", C.c_int32),
        ("dt", C.c_double), ("temp", C.c_double), ("w_noe", C.c_double), ("w_bond", C.c_double),
        ("w_vdw", C.c_double), ("repel", C.c_double), ("drop", C.c_double),
    ]

    def astuple(self):
        return (self.op, self.nstep, self.flag, self.dt, self.temp, self.w_noe, self.w_bond,
                self.w_vdw, self.repel, self.drop)


class Job(C.Structure):
    """chr3d_job: one (matrix, alpha) ensemble."""
    _fields_ = [
        ("IF", C.POINTER(C.c_double)), ("L", C.c_int32), ("min_sep", C.c_int32),
        ("alpha", C.c_double), ("K", C.c_double),
        ("ri", C.POINTER(C.c_int32)), ("rj", C.POINTER(C.c_int32)), ("rd", C.POINTER(C.c_double)),
        ("nr", C.c_int64), ("n_models", C.c_int32), ("pad_", C.c_int32), ("seed", C.c_uint64),
        ("x0", C.POINTER(C.c_double)),
    ]


class Result(C.Structure):
    """chr3d_result: caller-allocated outputs of one job."""
    _fields_ = [
        ("xyz", C.POINTER(C.c_double)), ("energies", C.POINTER(C.c_double)),
        ("force_evals", C.POINTER(C.c_int64)), ("hand", C.POINTER(C.c_int32)),
        ("satisfied", C.POINTER(C.c_int32)), ("sum_dev", C.POINTER(C.c_double)),
        ("n_restraints", C.c_int64), ("device_ms", C.c_double),
    ]


def stage(op, nstep=0, flag=0, dt=0.0, temp=0.0, w_noe=10.0, w_bond=1000.0, w_vdw=1.0, repel=0.85, drop=0.0):
    s = Stage()
    s.op, s.nstep, s.flag, s.dt, s.temp = op, nstep, flag, dt, temp
    s.w_noe, s.w_bond, s.w_vdw, s.repel, s.drop = w_noe, w_bond, w_vdw, repel, drop
    return s


def stage_array(stages):
    arr = (Stage * len(stages))()
    for k, s in enumerate(stages):
        C.memmove(C.byref(arr, k * C.sizeof(Stage)), C.byref(s), C.sizeof(Stage))
    return arr
