"""GPU tests of the sparse force provider (CSR restraint rows + cell-list repel, chr3d_single_create_sparse) against the fp64 oracle
and against the dense tile engine, on a 90 %-sparse synthetic Hi-C matrix (real high-resolution maps are sparse: IF = 0 cells carry
no restraint, C3D:148-152, 196)."""
import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import abi, datasets as ds

pytestmark = pytest.mark.gpu


def sparse_if(N, seed, keep=0.10):
    IF = ds.synthetic_if(N, seed)
    rng = np.random.default_rng(seed)
    m = np.triu(rng.random((N, N)) < keep, 1)
    band = np.triu(np.ones((N, N), bool), 1) & ~np.triu(np.ones((N, N), bool), 8)        # keep the near-diagonal band: it exists in every map
    m = m | band
    m = m | m.T | np.eye(N, dtype=bool)
    return np.where(m, IF, 0.0)


@pytest.fixture(scope="module")
def case():
    N = 1500
    IF = sparse_if(N, 97)
    ri, rj, rd = orc.if_to_restraints(IF, 0.9)
    assert len(ri) < 0.14 * N * N / 2                       # about a tenth of the dense table
    return N, IF, ri, rj, rd


W_CASES = [(2.6, dict(w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)), (1.5, dict(w_noe=1, w_bond=100, w_vdw=20.0, repel=0.5)),
           (3.8, dict(w_noe=4, w_bond=400, w_vdw=0.003, repel=0.9))]


def test_sparse_energy_forces_match_fp64_and_the_dense_engine(gpu_api, case):
    """Every restraint row once from each side, repel through the cell list (compact coils: thousands of repel pairs), bond term:
    energies and forces within 1e-5 of the fp64 evaluation and of the dense tile engine."""
    N, IF, ri, rj, rd = case
    p = gpu_api.default_params()
    for step, wk in W_CASES:
        w = abi.stage(abi.OP_ENERGY, **wk)
        x = orc.random_coil(N, step, 31, 0).astype(np.float32).astype(np.float64)
        prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), w])
        s = gpu_api.Single(N=N, restraints=(ri, rj, rd), x0=x, p=p, program=prog)
        s.run()
        r = s.fetch()
        info = s.sparse_info()
        s.close()
        eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
        assert r["restraint_rows"] == len(ri) and info["csr_entries"] == 2 * len(ri) and info["rebuilds"] >= 1
        assert info["noe_bytes_per_eval"] < 0.3 * (N * N / 2) * 4 * 2          # traffic follows the rows, not N^2
        tot = abs(eo[0, 0])
        assert np.all(np.abs(r["energies"] - eo[0]) <= 1e-5 * np.maximum(np.abs(eo[0]), 1e-3 * tot)), (step, r["energies"], eo)
        assert np.abs(r["forces"] - fo[0]).max() <= 1e-5 * np.abs(fo).max(), step
        if step == 2.6:
            assert eo[0, 2] > 0                                # the repel term is active in this case
            d = gpu_api.Single(IF=IF, alpha=0.9, x0=x, p=p, program=prog)
            d.run()
            rd_ = d.fetch()
            d.close()
            assert np.abs(r["forces"] - rd_["forces"]).max() <= 2e-5 * np.abs(fo).max()
            assert np.allclose(r["energies"], rd_["energies"], rtol=2e-5, atol=2e-5 * tot)


def test_sparse_md_and_minimiser_track_fp64(gpu_api, case):
    """Short hot MD (beads move several skins: the bins are rebuilt on the device along the way) and an L-BFGS step through the
    sparse provider against the oracle."""
    N, IF, ri, rj, rd = case
    p = gpu_api.default_params()
    x0 = orc.random_coil(N, 2.6, 5, 0).astype(np.float32).astype(np.float64)
    w = dict(w_noe=2.0, w_bond=200.0, w_vdw=4.0, repel=0.9)
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                            abi.stage(abi.OP_DYNAMICS, nstep=30, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **w), abi.stage(abi.OP_ENERGY, **w)])
    s = gpu_api.Single(N=N, restraints=(ri, rj, rd), x0=x0, seed=9, p=p, program=prog)
    s.run()
    r = s.fetch()
    info = s.sparse_info()
    s.close()
    o = orc.run_model(N, ri, rj, rd, p, prog, 9, 0, x0)
    assert r["force_evals"] == o["force_evals"] == 31
    assert np.abs(r["xyz"] - o["xyz"]).max() < 2e-4
    assert np.allclose(r["energies"], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
    assert np.abs(o["xyz"] - x0).max() > 0.25 and info["rebuilds"] >= 2          # moved more than half a skin: bins were rebuilt
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=2, drop=2.0e5, **w), abi.stage(abi.OP_CENTER), abi.stage(abi.OP_ENERGY, **w)])
    s = gpu_api.Single(N=N, restraints=(ri, rj, rd), x0=x0, seed=9, p=p, program=prog)
    s.run()
    r = s.fetch()
    s.close()
    o = orc.run_model(N, ri, rj, rd, p, prog, 9, 0, x0)
    assert np.allclose(r["energies"], o["energies"], rtol=1e-4, atol=1e-4 * o["energies"][0])
    assert np.abs(r["xyz"] - o["xyz"]).max() < 1e-2


def test_sparse_rows_are_validated(gpu_api):
    """Rows outside 0 <= i < j < N, non-positive distances and repeated pairs are refused (no out-of-bounds access on the device)."""
    from chromosome3d_b200.api import Chr3dError
    p = gpu_api.default_params()
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_ENERGY)])
    for bad in (([0, 5], [3, 50], [4.0, 5.0]), ([2], [2], [4.0]), ([0, 0], [3, 3], [4.0, 4.1]), ([0], [3], [0.0])):
        with pytest.raises(Chr3dError):
            gpu_api.Single(N=40, restraints=bad, p=p, program=prog)


@pytest.mark.parametrize("L,keep", [(300, 0.08), (2600, 0.01), (9000, 0.002)])
def test_coo_to_restraints_equals_the_dense_derivation(gpu_api, L, keep):
    """Sparse triplets -> rows without the L x L matrix (device arithmetic over the non-zero cells only): the same rows and the same
    distances, bit for bit, as IF2dist_new + dist2rr on the dense matrix (C3D:110-206) — the oracle for the small case, the dense
    CUDA path for the larger ones (9,000 bins: beyond the single-thread sum, the mean is summed row by row in both)."""
    rng = np.random.default_rng(L)
    n = int(keep * L * L / 2)
    bi = rng.integers(0, L, n).astype(np.int32)
    bj = rng.integers(0, L, n).astype(np.int32)
    lo, hi = np.minimum(bi, bj), np.maximum(bi, bj)
    v = np.round(rng.gamma(0.7, 40.0, n) / (1.0 + (hi - lo)) ** 0.8 + 0.01, 4)
    diag = np.arange(L, dtype=np.int32)                       # a full diagonal and first off-diagonal, plus repeats that must add up
    bi = np.concatenate([lo, diag, diag[:-1], lo[:50]]); bj = np.concatenate([hi, diag, diag[:-1] + 1, hi[:50]])
    v = np.concatenate([v, rng.uniform(100, 900, L), rng.uniform(50, 300, L - 1), v[:50]])
    alpha = 0.7
    ri, rj, rd = gpu_api.coo_to_restraints(bi, bj, v, L, alpha)
    assert np.all(rj - ri >= 5) and np.all(rd > 0) and np.all(np.diff(ri.astype(np.int64) * L + rj) > 0)          # i-major, unique
    dense = np.zeros((L, L))
    np.add.at(dense, (bi, bj), v)
    dense = dense + np.triu(dense, 1).T                       # all triplets are upper-triangle here: (i, j) and (j, i) both receive them
    if L <= 300:
        oi, oj, od = orc.if_to_restraints(dense, alpha)
    else:
        oi, oj, od = gpu_api.if_to_restraints(dense, alpha)
    order = np.argsort(oi.astype(np.int64) * L + oj, kind="stable")
    assert len(ri) == len(oi) and np.array_equal(ri, oi[order]) and np.array_equal(rj, oj[order]) and np.array_equal(rd, od[order])
    # one-based indices, and the rows feed the sparse engine
    r1 = gpu_api.coo_to_restraints(bi + 1, bj + 1, v, L, alpha, base=1)
    assert np.array_equal(r1[0], ri) and np.array_equal(r1[2], rd)
    if L == 2600:
        prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)])
        x = orc.random_coil(L, 3.8, 2, 0).astype(np.float32).astype(np.float64)
        s = gpu_api.Single(N=L, restraints=(ri, rj, rd), x0=x, p=gpu_api.default_params(), program=prog)
        s.run()
        r = s.fetch()
        s.close()
        eo, fo = orc.energy_forces(x, ri, rj, rd, gpu_api.default_params(), abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85))
        assert np.allclose(r["energies"], eo[0], rtol=1e-5, atol=1e-5 * abs(eo[0, 0]))
    from chromosome3d_b200.api import Chr3dError
    with pytest.raises(Chr3dError):
        gpu_api.coo_to_restraints([0], [L], [1.0], L, alpha)
    with pytest.raises(Chr3dError):
        gpu_api.coo_to_restraints([0], [9], [1.0], L, 0.0)
    assert len(gpu_api.coo_to_restraints([], [], [], 12, alpha)[0]) == 0
