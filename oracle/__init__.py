"""Python binding of the fp64 CPU oracle (oracle/chr3d_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from chromosome3d_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libchr3d_oracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(force=False):
    src = os.path.join(_HERE, "chr3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libchr3d_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_dist_to_restraints.restype = C.c_int64
        _lib.orc_energy_forces_bench.restype = C.c_double
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def default_params():
    p = abi.SaParams()
    lib().orc_default_params(C.byref(p))
    return p


def build_program(p):
    n = lib().orc_build_program(C.byref(p), None, 0)
    arr = (abi.Stage * n)()
    lib().orc_build_program(C.byref(p), arr, n)
    return arr


def if_to_dist(IF, alpha, K=11.0):
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    L = IF.shape[0]
    t = np.zeros((L, L), dtype=np.int64)
    rc = lib().orc_if_to_dist(_d(IF), L, C.c_double(alpha), C.c_double(K), t.ctypes.data_as(_lp))
    assert rc == 0
    return t


def dist_to_restraints(tenths, min_sep=5):
    tenths = np.ascontiguousarray(tenths, dtype=np.int64)
    L = tenths.shape[0]
    n = lib().orc_dist_to_restraints(tenths.ctypes.data_as(_lp), L, min_sep, None, None, None, C.c_int64(0))
    ri = np.zeros(n, dtype=np.int32)
    rj = np.zeros(n, dtype=np.int32)
    rd = np.zeros(n, dtype=np.float64)
    lib().orc_dist_to_restraints(tenths.ctypes.data_as(_lp), L, min_sep, _i(ri), _i(rj), _d(rd), C.c_int64(n))
    return ri, rj, rd


def if_to_restraints(IF, alpha, K=11.0, min_sep=5):
    return dist_to_restraints(if_to_dist(IF, alpha, K), min_sep)


def write_dist(path, tenths):
    tenths = np.ascontiguousarray(tenths, dtype=np.int64)
    assert lib().orc_write_dist(path.encode(), tenths.ctypes.data_as(_lp), tenths.shape[0]) == 0


def write_rr(path, ri, rj, rd):
    assert lib().orc_write_rr(path.encode(), _i(ri), _i(rj), _d(rd), C.c_int64(len(ri))) == 0


def write_tbl(path, ri, rj, rd):
    assert lib().orc_write_tbl(path.encode(), _i(ri), _i(rj), _d(rd), C.c_int64(len(ri))) == 0


def energy_forces(xyz, ri, rj, rd, p, st, want_forces=True):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if xyz.ndim == 2:
        xyz = xyz[None]
    m, N, _ = xyz.shape
    e = np.zeros((m, 4))
    f = np.zeros_like(xyz) if want_forces else None
    lib().orc_energy_forces(_d(xyz), m, N, _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), C.byref(p), C.byref(st),
                            _d(e), _d(f) if want_forces else None)
    return e, f


def large_check(IF, alpha, p, st, K=11.0, min_sep=5, gpu_targets=None, x=None, want_forces=True, threads=0):
    """Large structures: cell-by-cell check of the GPU's dense fp32 target matrix and fp64 energies/forces straight from the IF matrix."""
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    N = IF.shape[0]
    bad, rows, worst = C.c_int64(0), C.c_int64(0), C.c_double(0)
    e = np.zeros(4)
    f = np.zeros((N, 3)) if (x is not None and want_forces) else None
    if x is not None:
        x = np.ascontiguousarray(x, dtype=np.float64)
    if gpu_targets is not None:
        gpu_targets = np.ascontiguousarray(gpu_targets, dtype=np.float32)
    if not threads:
        try:
            threads = len(os.sched_getaffinity(0))
        except Exception:
            threads = os.cpu_count() or 1
    rc = lib().orc_large_check(_d(IF), N, C.c_double(alpha), C.c_double(K), int(min_sep), C.byref(p), C.byref(st),
                               gpu_targets.ctypes.data_as(C.POINTER(C.c_float)) if gpu_targets is not None else None,
                               C.byref(bad), C.byref(worst), C.byref(rows), _d(x) if x is not None else None, _d(e),
                               _d(f) if f is not None else None, int(threads))
    assert rc == 0
    return dict(mismatches=bad.value, max_diff=worst.value, rows=rows.value, energies=e, forces=f)


def energy_forces_bench(x, ri, rj, rd, p, st, reps, threads):
    x = np.ascontiguousarray(x, dtype=np.float64)
    e = C.c_double(0)
    sec = lib().orc_energy_forces_bench(_d(x), x.shape[0], _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), C.byref(p),
                                        C.byref(st), reps, threads, C.byref(e))
    return sec, e.value


def random_coil(N, step, seed, model):
    x = np.zeros((N, 3))
    lib().orc_random_coil(N, C.c_double(step), C.c_uint64(seed), model, _d(x))
    return x


def philox(c, k):
    out = (C.c_uint32 * 4)()
    lib().orc_philox(C.c_uint32(c[0]), C.c_uint32(c[1]), C.c_uint32(c[2]), C.c_uint32(c[3]),
                     C.c_uint32(k[0]), C.c_uint32(k[1]), out)
    return list(out)


def run_model(N, ri, rj, rd, p, prog, seed, model, x0=None):
    x = np.zeros((N, 3))
    v = np.zeros((N, 3))
    e = np.zeros(4)
    nev = C.c_int64(0)
    hand = C.c_int(0)
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
    lib().orc_run_model(N, _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), C.byref(p), prog, len(prog),
                        C.c_uint64(seed), model, _d(x0) if x0 is not None else None, _d(x), _d(v), _d(e),
                        C.byref(nev), C.byref(hand))
    return dict(xyz=x, v=v, energies=e, force_evals=nev.value, hand=hand.value)


def run_ensemble(N, ri, rj, rd, p, prog, seed, n_models, first_model=0, x0=None, threads=0):
    xyz = np.zeros((n_models, N, 3))
    e = np.zeros((n_models, 4))
    nev = np.zeros(n_models, dtype=np.int64)
    hand = np.zeros(n_models, dtype=np.int32)
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
    lib().orc_run_ensemble(N, _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), C.byref(p), prog, len(prog),
                           C.c_uint64(seed), n_models, first_model, _d(x0) if x0 is not None else None, threads,
                           _d(xyz), _d(e), nev.ctypes.data_as(_lp), _i(hand))
    return dict(xyz=xyz, energies=e, force_evals=nev, hand=hand)


def max_threads():
    return lib().orc_max_threads()


def assess(xyz, ri, rj, rd):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if xyz.ndim == 2:
        xyz = xyz[None]
    m, N, _ = xyz.shape
    sat = np.zeros(m, dtype=np.int32)
    dev = np.zeros(m)
    lib().orc_assess(_d(xyz), m, N, _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), _i(sat), _d(dev))
    return sat, dev


def rank_models(energies):
    energies = np.ascontiguousarray(energies, dtype=np.float64)
    order = np.zeros(energies.shape[0], dtype=np.int32)
    lib().orc_rank_models(_d(energies), energies.shape[0], _i(order))
    return order


def spearman(IF, xyz, rng=3):
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if xyz.ndim == 2:
        xyz = xyz[None]
    rho = np.zeros(xyz.shape[0])
    lib().orc_spearman(_d(IF), IF.shape[0], _d(xyz), xyz.shape[0], rng, _d(rho))
    return rho


def reduce_model(xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    out = np.zeros(((xyz.shape[0] + 1) // 2, 3))
    lib().orc_reduce_model(_d(xyz), xyz.shape[0], _d(out))
    return out


def write_reduced_pdb(path, xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if lib().orc_write_reduced_pdb(str(path).encode(), _d(xyz), xyz.shape[0]) != 0:
        raise OSError("orc_write_reduced_pdb: cannot write %s" % path)


def model_similarity(xa, xb):
    xa = np.ascontiguousarray(xa, dtype=np.float64)
    xb = np.ascontiguousarray(xb, dtype=np.float64)
    rho, rmsd = C.c_double(0), C.c_double(0)
    lib().orc_model_similarity(_d(xa), xa.shape[0], _d(xb), xb.shape[0], C.byref(rho), C.byref(rmsd))
    return rho.value, rmsd.value


def dg_bounds(N, ri, rj, rd, p):
    ri, rj, rd = np.ascontiguousarray(ri, dtype=np.int32), np.ascontiguousarray(rj, dtype=np.int32), np.ascontiguousarray(rd, dtype=np.float64)
    U, L = np.zeros((N, N)), np.zeros((N, N))
    lib().orc_dg_bounds.restype = C.c_int64
    bad = lib().orc_dg_bounds(N, _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), C.byref(p), _d(U), _d(L))
    return U, L, bad


def dg_embed(N, ri, rj, rd, p, seed, n_models):
    U, L, bad = dg_bounds(N, ri, rj, rd, p)
    xyz, eig = np.zeros((n_models, N, 3)), np.zeros((n_models, 3))
    for m in range(n_models):
        lib().orc_dg_embed_model(N, _d(U), _d(L), C.c_uint64(seed), m, _d(xyz[m]), _d(eig[m]))
    return xyz, eig, bad


def dg_starts(N, ri, rj, rd, p, seed, n_models, dg_scale=100.0, nstep=200, drop=1.0):
    """orc_dg_embed_model + orc_dg_refine: the whole DG stage (C3D:1470-1526) per model."""
    U, L, bad = dg_bounds(N, ri, rj, rd, p)
    xyz, eig, st = np.zeros((n_models, N, 3)), np.zeros((n_models, 3)), np.zeros((n_models, 6))
    for m in range(n_models):
        lib().orc_dg_embed_model(N, _d(U), _d(L), C.c_uint64(seed), m, _d(xyz[m]), _d(eig[m]))
        lib().orc_dg_refine(N, _d(U), _d(L), C.byref(p), C.c_double(dg_scale), nstep, C.c_double(drop), _d(xyz[m]), _d(st[m]))
    keys = ("dgscale", "e_scaled", "e_final", "iterations", "evaluations", "rms_gradient")
    return xyz, eig, [dict(zip(keys, row)) for row in st.tolist()], bad
