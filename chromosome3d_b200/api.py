"""Host-side mirror of the reference's interface for the DGSA hot path, over the C ABI.

Everything numerical happens in libchr3d.so (hand-written CUDA for sm_100a).  This module only
marshals numpy buffers through ctypes, the same way the Perl front end does through its XS/FFI
shim.  There is no CPU fallback: without the built library or without a GPU
every compute call raises Chr3dError.

Reference functions mirrored (names follow chromosome3D.pl):
    calc_len_IF / IF2dist_new / dist2rr / carr2tbl      C3D:110-206, 340-362
    build_models (the CNS dgsa.inp run)                 C3D:254-289, protocol C3D:1573-1829
    assess_dgsa + count_satisfied_tbl_rows + sum_noe_dev C3D:447-600, 769-829
    spearman_IF_pdb.pl                                   SP:32-71
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHR3D_LIBRARY") or os.path.join(_HERE, "libchr3d.so")     # CHR3D_LIBRARY: another build of the same library (kernel variants)
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


class Chr3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libchr3d error %d: %s" % (code, msg))
        self.code = code


def lib():
    """Load libchr3d.so (built in-tree by `make -C chromosome3d_b200/csrc` / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Chr3dError(abi.ERR_NO_DEVICE, "libchr3d.so is not built (run __graft_entry__.build()); there is no CPU path")
        L = C.CDLL(LIB_PATH)
        L.chr3d_version.restype = C.c_char_p
        L.chr3d_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise Chr3dError(rc, lib().chr3d_last_error().decode(errors="replace"))


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def version():
    return lib().chr3d_version().decode()


def device_count():
    return lib().chr3d_device_count()


def init(device=0):
    _check(lib().chr3d_init(device))


def fp32_peak(reps=3):
    v = C.c_double(0)
    _check(lib().chr3d_fp32_peak(reps, C.byref(v)))
    return v.value


def fp32_peak_packed(reps=3):
    v = C.c_double(0)
    _check(lib().chr3d_fp32_peak_packed(reps, C.byref(v)))
    return v.value


def default_params():
    p = abi.SaParams()
    lib().chr3d_default_params(C.byref(p))
    return p


def set_option(name, value=None):
    """chr3d_set_option: launch-shape options of the ensemble kernel ("sa_cluster", "sa_warps", "sa_min_smem", "sa_split") and
    "seqsum" = "serial" (the left-to-right sums as a serial chain instead of in parallel; same bits); None clears."""
    _check(lib().chr3d_set_option(name.encode(), None if value is None else str(value).encode()))


def seq_sum(a):
    """chr3d_seq_sum: the left-to-right fp64 sum of every row of `a` (1-D: one sum), with the bits of `for v in row: s += v`
    (the Perl's `$sum += ...` loops, C3D:132-139, 590-598), evaluated in parallel on the device."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    one = a.ndim == 1
    a2 = a.reshape(1, -1) if one else a.reshape(a.shape[0], -1)
    out = np.zeros(a2.shape[0], dtype=np.float64)
    _check(lib().chr3d_seq_sum(_d(a2), C.c_int64(a2.shape[1]), C.c_int32(a2.shape[0]), _d(out)))
    return float(out[0]) if one else out


def cns_pin_params():
    """chr3d_cns_pin_params: default_params() plus the bead-model stand-ins for what CNS's full-atom poly-GLY chain adds (default
    off): a soft CA(i)-CA(i+2) lower bound (the angle / improper terms keep a real chain from folding back on itself) and a
    larger repel radius (backbone atoms around each CA).  Fitted on the 45 shipped CNS models with the fp64 oracle
    (scripts/fit_cns_pin.py, profiles/r02_cns_pin.md): median IF-Spearman within 0.02 of the shipped model for 42 of 45 (35 without)."""
    p = abi.SaParams()
    lib().chr3d_cns_pin_params(C.byref(p))
    return p


def build_program(p):
    """The dgsa.inp protocol body as a stage program (C3D:1573-1829)."""
    n = lib().chr3d_build_program(C.byref(p), None, 0)
    if n < 0:
        _check(n)
    arr = (abi.Stage * n)()
    lib().chr3d_build_program(C.byref(p), arr, n)
    return arr


# ---------------------------------------------------------------------------------- restraints
def calc_len_IF(path):
    """C3D:164-179: L = number of tokens on the first row."""
    with open(path, "r", newline="") as fh:
        for line in fh:
            toks = line.split()
            if toks:
                return len(toks)
    raise ValueError("empty IF matrix %s" % path)


def IF2dist_new(IF, alpha, K=11.0):
    """C3D:110-162.  Returns (tenths int64 (L,L), n_patched): the `%.1f` text of the .dist file."""
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    L = IF.shape[0]
    t = np.zeros((L, L), dtype=np.int64)
    npatch = C.c_int64(0)
    _check(lib().chr3d_if_to_dist(_d(IF), L, C.c_double(alpha), C.c_double(K), t.ctypes.data_as(_lp), C.byref(npatch)))
    return t, npatch.value


def if_to_restraints(IF, alpha, K=11.0, min_sep=5):
    """dist2rr + carr2tbl (C3D:181-206, 340-362): (ri, rj, rd) in the string-sorted row order."""
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    L = IF.shape[0]
    n = C.c_int64(0)
    _check(lib().chr3d_if_to_restraints(_d(IF), L, C.c_double(alpha), C.c_double(K), min_sep, None, None, None,
                                        C.c_int64(0), C.byref(n)))
    cap = max(n.value, 1)
    ri = np.zeros(cap, dtype=np.int32)
    rj = np.zeros(cap, dtype=np.int32)
    rd = np.zeros(cap, dtype=np.float64)
    if n.value:
        _check(lib().chr3d_if_to_restraints(_d(IF), L, C.c_double(alpha), C.c_double(K), min_sep, _i(ri), _i(rj), _d(rd),
                                            C.c_int64(cap), C.byref(n)))
    return ri[:n.value], rj[:n.value], rd[:n.value]


def write_dist(path, tenths):
    tenths = np.ascontiguousarray(tenths, dtype=np.int64)
    _check(lib().chr3d_write_dist(path.encode(), tenths.ctypes.data_as(_lp), tenths.shape[0]))


def write_rr(path, ri, rj, rd):
    ri, rj, rd = _list(ri, rj, rd)
    _check(lib().chr3d_write_rr(path.encode(), _i(ri), _i(rj), _d(rd), C.c_int64(len(ri))))


def carr2tbl(path, ri, rj, rd):
    ri, rj, rd = _list(ri, rj, rd)
    _check(lib().chr3d_write_tbl(path.encode(), _i(ri), _i(rj), _d(rd), C.c_int64(len(ri))))


def _list(ri, rj, rd):
    return (np.ascontiguousarray(ri, dtype=np.int32), np.ascontiguousarray(rj, dtype=np.int32),
            np.ascontiguousarray(rd, dtype=np.float64))


# ---------------------------------------------------------------------------------- energy
def energy_forces(xyz, ri, rj, rd, p, weights, want_forces=True):
    """CNS `energy` at fixed coordinates: (energies (m,4) [overall,bon,vdw,noe], forces (m,N,3))."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if xyz.ndim == 2:
        xyz = xyz[None]
    m, N, _ = xyz.shape
    ri, rj, rd = _list(ri, rj, rd)
    e = np.zeros((m, 4))
    f = np.zeros_like(xyz) if want_forces else None
    _check(lib().chr3d_energy_forces(_d(xyz), m, N, _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), C.byref(p),
                                     C.byref(weights), _d(e), _d(f) if want_forces else None))
    return e, f


# ---------------------------------------------------------------------------------- SA run
class Job:
    """One (matrix, alpha) ensemble = one chromosome3D.pl invocation (C3D:28-35)."""

    def __init__(self, IF=None, alpha=0.5, K=11.0, min_sep=5, n_models=20, seed=0, restraints=None, N=None, x0=None):
        self.IF = None if IF is None else np.ascontiguousarray(IF, dtype=np.float64)
        self.alpha, self.K, self.min_sep, self.n_models, self.seed = float(alpha), float(K), int(min_sep), int(n_models), int(seed)
        self.restraints = None if restraints is None else _list(*restraints)
        self.N = int(self.IF.shape[0] if self.IF is not None else N)
        self.x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).reshape(self.n_models, self.N, 3)

    def c_struct(self):
        j = abi.Job()
        j.L, j.min_sep, j.alpha, j.K = self.N, self.min_sep, self.alpha, self.K
        j.n_models, j.seed = self.n_models, self.seed
        if self.IF is not None:
            j.IF = _d(self.IF)
        elif self.restraints is not None:
            ri, rj, rd = self.restraints
            j.ri, j.rj, j.rd, j.nr = _i(ri), _i(rj), _d(rd), len(ri)
        if self.x0 is not None:
            j.x0 = _d(self.x0)
        return j


class JobResult:
    def __init__(self, job):
        m, N = job.n_models, job.N
        self.xyz = np.zeros((m, N, 3))
        self.energies = np.zeros((m, 4))
        self.force_evals = np.zeros(m, dtype=np.int64)
        self.hand = np.zeros(m, dtype=np.int32)
        self.satisfied = np.zeros(m, dtype=np.int32)
        self.sum_dev = np.zeros(m)
        self.n_restraints = 0
        self.device_ms = 0.0

    def c_struct(self):
        r = abi.Result()
        r.xyz, r.energies = _d(self.xyz), _d(self.energies)
        r.force_evals = self.force_evals.ctypes.data_as(_lp)
        r.hand, r.satisfied, r.sum_dev = _i(self.hand), _i(self.satisfied), _d(self.sum_dev)
        return r


def _job_arrays(jobs):
    arr = (abi.Job * len(jobs))()
    for k, j in enumerate(jobs):
        arr[k] = j.c_struct()
    return arr


def _collect(results, carr):
    for r, c in zip(results, carr):
        r.n_restraints, r.device_ms = c.n_restraints, c.device_ms
    return results


def build_models(jobs, p=None, program=None):
    """The CNS dgsa run of build_models() (C3D:254-289) for a list of jobs, host buffers in and out."""
    p = p or default_params()
    program = program if program is not None else build_program(p)
    jarr = _job_arrays(jobs)
    results = [JobResult(j) for j in jobs]
    rarr = (abi.Result * len(jobs))(*[r.c_struct() for r in results])
    _check(lib().chr3d_run_ensemble(jarr, len(jobs), C.byref(p), program, len(program), rarr))
    return _collect(results, rarr)


class Ensemble:
    """Device-resident variant: create() uploads and derives targets, run() only launches kernels."""

    def __init__(self, jobs, p=None, program=None):
        self.jobs = jobs
        self.p = p or default_params()
        self.program = program if program is not None else build_program(self.p)
        self._h = C.c_void_p()
        jarr = _job_arrays(jobs)
        _check(lib().chr3d_ensemble_create(jarr, len(jobs), C.byref(self.p), self.program, len(self.program), C.byref(self._h)))

    def run(self, stream=None):
        _check(lib().chr3d_ensemble_run(self._h, C.c_void_p(stream or 0)))

    def fetch(self):
        results = [JobResult(j) for j in self.jobs]
        rarr = (abi.Result * len(self.jobs))(*[r.c_struct() for r in results])
        _check(lib().chr3d_ensemble_fetch(self._h, rarr))
        return _collect(results, rarr)

    def info(self):
        c, t, n = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        rows = C.c_int64(0)
        _check(lib().chr3d_ensemble_info(self._h, C.byref(c), C.byref(t), C.byref(n), C.byref(rows)))
        return dict(cluster=c.value, threads=t.value, launches=n.value, restraint_rows=rows.value)

    def close(self):
        if self._h:
            lib().chr3d_ensemble_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------- assessment
def count_satisfied_and_sum_dev(xyz, ri, rj, rd):
    """count_satisfied_tbl_rows numerator (C3D:464-473) and sum_noe_dev (C3D:592-599) per model."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if xyz.ndim == 2:
        xyz = xyz[None]
    m, N, _ = xyz.shape
    ri, rj, rd = _list(ri, rj, rd)
    sat = np.zeros(m, dtype=np.int32)
    dev = np.zeros(m)
    _check(lib().chr3d_assess(_d(xyz), m, N, _i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), _i(sat), _d(dev)))
    return sat, dev


def rank_models(energies):
    energies = np.ascontiguousarray(energies, dtype=np.float64)
    order = np.zeros(energies.shape[0], dtype=np.int32)
    _check(lib().chr3d_rank_models(_d(energies), energies.shape[0], _i(order)))
    return order


def write_pdb(path, xyz, energies=None, resnames=None, with_conect=False):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    e = None if energies is None else np.ascontiguousarray(energies, dtype=np.float64)
    rn = None if resnames is None else "".join(resnames).encode()
    _check(lib().chr3d_write_pdb(path.encode(), _d(xyz), xyz.shape[0], _d(e) if e is not None else None, rn,
                                 1 if with_conect else 0))


def spearman_IF_pdb(IF, xyz, rng=3):
    """SP:32-71: Spearman rho (unrounded fp64) per model; the script prints '%.3f' of it."""
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    if xyz.ndim == 2:
        xyz = xyz[None]
    rho = np.zeros(xyz.shape[0])
    _check(lib().chr3d_spearman(_d(IF), IF.shape[0], _d(xyz), xyz.shape[0], rng, _d(rho)))
    return rho


# ---------------------------------------------------------------------------------- SURVEY 8(f) "next" rows
def reduce_model(xyz):
    """Pair-averaged model of half the resolution (the shipped *_reduced.pdb files)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    out = np.zeros(((xyz.shape[0] + 1) // 2, 3))
    _check(lib().chr3d_reduce_model(_d(xyz), xyz.shape[0], _d(out)))
    return out


def write_reduced_pdb(path, xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    _check(lib().chr3d_write_reduced_pdb(str(path).encode(), _d(xyz), xyz.shape[0]))


def model_similarity(xyz_a, xyz_b):
    """(Spearman, RMSD) of output_models/similarity.txt between two models' pairwise-distance vectors."""
    a = np.ascontiguousarray(xyz_a, dtype=np.float64)
    b = np.ascontiguousarray(xyz_b, dtype=np.float64)
    rho, rmsd = C.c_double(0), C.c_double(0)
    _check(lib().chr3d_model_similarity(_d(a), a.shape[0], _d(b), b.shape[0], C.byref(rho), C.byref(rmsd)))
    return rho.value, rmsd.value


def read_if_text(path):
    """input/<ID>.txt parsed as IF2dist_new does (C3D:116-129), by the library's threaded parser."""
    L = C.c_int32(0)
    _check(lib().chr3d_read_if_text(str(path).encode(), None, C.c_int64(0), C.byref(L)))
    out = np.zeros((L.value, L.value))
    _check(lib().chr3d_read_if_text(str(path).encode(), _d(out), C.c_int64(out.size), C.byref(L)))
    return out


def write_if_binary(path, IF, elem_bytes=8):
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    _check(lib().chr3d_write_if_binary(str(path).encode(), _d(IF), IF.shape[0], elem_bytes))


def read_if_binary(path):
    L = C.c_int32(0)
    _check(lib().chr3d_read_if_binary(str(path).encode(), None, C.c_int64(0), C.byref(L)))
    out = np.zeros((L.value, L.value))
    _check(lib().chr3d_read_if_binary(str(path).encode(), _d(out), C.c_int64(out.size), C.byref(L)))
    return out


def if_from_coo(bi, bj, v, L, base=0):
    """Sparse (bin_i, bin_j, value) triplets -> dense symmetric IF matrix (scatter on the device)."""
    bi = np.ascontiguousarray(bi, dtype=np.int32)
    bj = np.ascontiguousarray(bj, dtype=np.int32)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros((L, L))
    _check(lib().chr3d_if_from_coo(_i(bi), _i(bj), _d(v), C.c_int64(len(v)), L, base, _d(out)))
    return out


def coo_to_restraints(bi, bj, v, L, alpha, K=11.0, min_sep=5, base=0):
    """chr3d_coo_to_restraints: sparse (bin_i, bin_j, value) triplets -> restraint rows (ri, rj, rd), i-major, without the L x L matrix;
    the same distances as if_to_restraints(if_from_coo(...)).  For Single(restraints=...) / Job(restraints=...)."""
    bi = np.ascontiguousarray(bi, dtype=np.int32)
    bj = np.ascontiguousarray(bj, dtype=np.int32)
    v = np.ascontiguousarray(v, dtype=np.float64)
    cap = max(1, len(v))
    ri, rj, rd = np.zeros(cap, dtype=np.int32), np.zeros(cap, dtype=np.int32), np.zeros(cap)
    n = C.c_int64(0)
    _check(lib().chr3d_coo_to_restraints(_i(bi), _i(bj), _d(v), C.c_int64(len(v)), L, base, C.c_double(alpha), C.c_double(K), min_sep,
                                         _i(ri), _i(rj), _d(rd), C.c_int64(cap), C.byref(n)))
    return ri[:n.value].copy(), rj[:n.value].copy(), rd[:n.value].copy()


def dg_embed(ri, rj, rd, N, n_models=1, seed=0, p=None):
    """DG stage of dgsa.inp (C3D:1470-1526) at bead level: starting coordinates (n_models, N, 3) by metric-matrix
    embedding of the smoothed bounds; also returns the eigenvalues used and the number of triangle-inconsistent pairs.
    Pass the coordinates as Job(x0=...) to anneal from them."""
    ri, rj, rd = _list(ri, rj, rd)
    p = p or default_params()
    xyz = np.zeros((n_models, N, 3))
    eig = np.zeros((n_models, 3))
    bad = C.c_int64(0)
    _check(lib().chr3d_dg_embed(_i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), N, C.byref(p), n_models, C.c_uint64(seed), _d(xyz), _d(eig), C.byref(bad)))
    return xyz, eig, bad.value


def dg_starts(ri, rj, rd, N, n_models=1, seed=0, p=None, dg_scale=100.0, nstep=200, drop=1.0):
    """The whole DG stage (C3D:1470-1526): embedding, `x *= $dgscale`, then md.dg.step = 200 L-BFGS iterations (drop = 1.) on the
    bounds-violation energy alone.  Returns (xyz (n_models, N, 3), eig3, stats rows {dgscale, e_scaled, e_final, iterations,
    evaluations, rms_gradient}, inconsistent pairs); pass xyz as Job(x0=...)."""
    ri, rj, rd = _list(ri, rj, rd)
    p = p or default_params()
    xyz = np.zeros((n_models, N, 3))
    eig = np.zeros((n_models, 3))
    st = np.zeros((n_models, 6))
    bad = C.c_int64(0)
    _check(lib().chr3d_dg_starts(_i(ri), _i(rj), _d(rd), C.c_int64(len(ri)), N, C.byref(p), n_models, C.c_uint64(seed), C.c_double(dg_scale), nstep,
                                 C.c_double(drop), _d(xyz), _d(eig), _d(st), C.byref(bad)))
    keys = ("dgscale", "e_scaled", "e_final", "iterations", "evaluations", "rms_gradient")
    return xyz, eig, [dict(zip(keys, row)) for row in st.tolist()], bad.value


def alpha_sweep(IF, alphas, n_models=20, K=11.0, seed=0, rng=3, p=None):
    """The alpha sweep of perl/chromosome3D_sweep.pl: every (alpha, model) trajectory in ONE call,
    models ranked per alpha like assess_dgsa (int(NOE energy), C3D:617, 802), scored by the IF-distance Spearman (SP:32-71);
    returns per-alpha rows and the selected model (most negative Spearman) with its `rankNN_aXX` name stem."""
    IF = np.ascontiguousarray(IF, dtype=np.float64)
    jobs = [Job(IF=IF, alpha=a, K=K, n_models=n_models, seed=seed) for a in alphas]
    res = build_models(jobs, p)
    rows, best = [], None
    for a, r in zip(alphas, res):
        order = rank_models(r.energies)
        rank_of = {int(m): k + 1 for k, m in enumerate(order)}
        rho = spearman_IF_pdb(IF, r.xyz, rng)
        pick = int(np.argmin(rho))
        rows.append(dict(alpha=float(a), restraints=int(r.n_restraints), best_noe=int(r.energies[order[0], abi.E_NOE]),
                         best_srcc=float(rho[pick]), median_srcc=float(np.median(rho)), picked_model=pick + 1, picked_rank=rank_of[pick]))
        if best is None or rho[pick] < best["srcc"]:
            best = dict(alpha=float(a), model=pick + 1, rank=rank_of[pick], srcc=float(rho[pick]), xyz=r.xyz[pick].copy(),
                        stem="rank%02d_a%02d" % (rank_of[pick], int(a * 10 + 0.5)))
    return rows, best


# ---------------------------------------------------------------------------------- one large structure
def nccl_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the caller broadcasts it, e.g. over torch.distributed)."""
    buf = (C.c_ubyte * 128)()
    _check(lib().chr3d_nccl_unique_id(buf))
    return bytes(buf)


def synthetic_if(N, seed):
    """Synthetic IF matrix of SURVEY 8(d), generated on the device (same definition as datasets.synthetic_if)."""
    out = np.zeros((N, N))
    _check(lib().chr3d_synthetic_if(N, C.c_uint64(seed), _d(out)))
    return out


class Single:
    """One structure of thousands of beads (the large-N engine), optionally one rank of a row-block
    (tile-range) partition over the GPUs of a node.  Same dgsa.inp protocol (C3D:1573-1829)."""

    def __init__(self, IF=None, N=None, alpha=0.5, K=11.0, min_sep=5, synth_seed=0, x0=None, seed=0, p=None, program=None,
                 rank=0, world=1, nccl_id=None, restraints=None):
        self.IF = None if IF is None else np.ascontiguousarray(IF, dtype=np.float64)
        self.N = int(self.IF.shape[0] if self.IF is not None else N)
        self.p = p or default_params()
        self.program = program if program is not None else build_program(self.p)
        self.x0 = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).reshape(self.N, 3)
        self._h = C.c_void_p()
        if restraints is not None:          # sparse restraint table: CSR rows + cell-list repel (chr3d_single_create_sparse)
            ri, rj, rd = restraints
            self._r = (np.ascontiguousarray(ri, dtype=np.int32), np.ascontiguousarray(rj, dtype=np.int32), np.ascontiguousarray(rd, dtype=np.float64))
            _check(lib().chr3d_single_create_sparse(self.N, _i(self._r[0]), _i(self._r[1]), _d(self._r[2]), C.c_int64(len(self._r[0])),
                                                    _d(self.x0) if self.x0 is not None else None, C.c_uint64(seed), C.byref(self.p),
                                                    self.program, len(self.program), C.byref(self._h)))
            return
        idbuf = None
        if nccl_id is not None:
            idbuf = (C.c_ubyte * 128).from_buffer_copy(bytes(nccl_id))
        _check(lib().chr3d_single_create(_d(self.IF) if self.IF is not None else None, self.N, C.c_double(alpha), C.c_double(K),
                                         int(min_sep), C.c_uint64(synth_seed), _d(self.x0) if self.x0 is not None else None,
                                         C.c_uint64(seed), C.byref(self.p), self.program, len(self.program), int(rank), int(world),
                                         idbuf, C.byref(self._h)))

    def comm_handle(self):
        """64-byte CUDA IPC handle of this rank's exchange buffer (world > 1)."""
        buf = (C.c_ubyte * 64)()
        _check(lib().chr3d_single_comm_handle(self._h, buf))
        return bytes(buf)

    def comm_connect(self, handles):
        """handles: the comm_handle() of every rank, in rank order.  Switches the force exchange from NCCL all-reduce
        to the exchange fused into the step kernels (remote stores + flags over NVLink peer memory)."""
        blob = b"".join(bytes(h) for h in handles)
        _check(lib().chr3d_single_comm_connect(self._h, (C.c_ubyte * len(blob)).from_buffer_copy(blob), len(handles)))

    def comm_error(self):
        e = C.c_int32(0)
        _check(lib().chr3d_single_comm_error(self._h, C.byref(e)))
        return e.value

    def run(self, stream=None):
        _check(lib().chr3d_single_run(self._h, C.c_void_p(stream or 0)))

    def eval_bench(self, reps=20, want_energy=True, stream=None):
        """(tile_ms, eval_ms): device time of the pair-tile kernel alone / of a whole force evaluation."""
        a, b = C.c_double(0), C.c_double(0)
        _check(lib().chr3d_single_eval_bench(self._h, C.c_void_p(stream or 0), int(reps), 1 if want_energy else 0, C.byref(a), C.byref(b)))
        return a.value, b.value

    def fetch(self):
        xyz = np.zeros((self.N, 3))
        f = np.zeros((self.N, 3))
        e = np.zeros(4)
        nev, nl, npairs = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        hand = C.c_int32(0)
        ms = C.c_double(0)
        _check(lib().chr3d_single_fetch(self._h, _d(xyz), _d(e), _d(f), C.byref(nev), C.byref(hand), C.byref(ms), C.byref(nl), C.byref(npairs)))
        return dict(xyz=xyz, forces=f, energies=e, force_evals=nev.value, hand=hand.value, device_ms=ms.value, launches=nl.value,
                    restraint_rows=npairs.value)

    def profile(self):
        """Phase clock of the persistent kernel since the last call (ms, CTA 0 of this rank): tile, barrier1, step, barrier2, update."""
        a = (C.c_double * 8)()
        _check(lib().chr3d_single_profile(self._h, a))
        return dict(zip(("tile_ms", "barrier1_ms", "step_ms", "barrier2_ms", "update_ms", "gather_push_ms", "peer_wait_ms", "gather_ms"), [float(v) for v in a[:8]]))

    def sparse_info(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _check(lib().chr3d_single_sparse_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(rebuilds=a.value, csr_entries=b.value, noe_bytes_per_eval=c.value)

    def profile_ctas(self):
        """Per-CTA tile counts (all, near, rows started) and tile-phase milliseconds of the persistent kernel since the last call."""
        cap = 1024
        n = C.c_int32(0)
        t, nn, r = (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_int32 * cap)()
        ms = (C.c_double * cap)()
        _check(lib().chr3d_single_profile_ctas(self._h, cap, C.byref(n), t, nn, r, ms))
        k = min(n.value, cap)
        return dict(tiles=np.array(t[:k]), near=np.array(nn[:k]), rows=np.array(r[:k]), tile_ms=np.array(ms[:k]))

    def targets(self):
        """Test aid: this rank's targets as a dense N*N fp32 matrix (NaN outside the rank's tiles)."""
        out = np.zeros((self.N, self.N), dtype=np.float32)
        _check(lib().chr3d_single_targets(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def info(self):
        nt, mine, tb = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        ctas = C.c_int32(0)
        _check(lib().chr3d_single_info(self._h, C.byref(nt), C.byref(mine), C.byref(ctas), C.byref(tb)))
        return dict(n_tiles=nt.value, tiles_this_rank=mine.value, ctas=ctas.value, target_bytes=tb.value)

    def close(self):
        if self._h:
            lib().chr3d_single_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def group_connect(singles):
    """Virtual ranks: the Single objects (rank g of world n, same process, same device) become one launch group."""
    arr = (C.c_void_p * len(singles))(*[s._h.value for s in singles])
    _check(lib().chr3d_single_group_connect(arr, len(singles)))


def group_run(singles, stream=None):
    arr = (C.c_void_p * len(singles))(*[s._h.value for s in singles])
    _check(lib().chr3d_single_group_run(arr, len(singles), C.c_void_p(stream or 0)))
