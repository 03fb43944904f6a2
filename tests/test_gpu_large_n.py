"""GPU parity tests of the large-N (single structure, tile-streaming) engine against the fp64 oracle
and against the small-N engine, through the C ABI (chr3d_single_*).

Tolerances as in test_gpu_parity.py: energies / forces within 1e-5 relative of the fp64 evaluation."""
import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import abi, datasets as ds

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mats():
    return {n: ds.load_if_matrix(ds.input_path(n)) for n in ("chr21_1mb", "chr1_1mb", "chr1_500kb")}


def _eval(gpu_api, IF, alpha, x, w, p=None, rank=0, world=1):
    p = p or gpu_api.default_params()
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), w])
    s = gpu_api.Single(IF=IF, alpha=alpha, x0=x, p=p, program=prog, rank=rank, world=world)
    s.run()
    out = s.fetch()
    out["info"] = s.info()
    s.close()
    return out


W_CASES = [(1.0, dict(w_noe=10, w_bond=1000, w_vdw=1.0, repel=0.85)),
           (0.4, dict(w_noe=1, w_bond=100, w_vdw=20.0, repel=0.5)),
           (0.7, dict(w_noe=4, w_bond=400, w_vdw=0.003, repel=0.9))]


@pytest.mark.parametrize("name,alpha", [("chr21_1mb", 0.5), ("chr1_1mb", 1.1), ("chr1_500kb", 1.1), ("chr1_500kb", 0.5)])
def test_tile_engine_energy_forces_match_fp64(gpu_api, mats, name, alpha):
    """1 tile (N=37), 3 tiles (N=229) and 10 tiles (N=455): every unordered pair once, forces to both beads."""
    IF = mats[name]
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    p = gpu_api.default_params()
    for scale, wk in W_CASES:
        w = abi.stage(abi.OP_ENERGY, **wk)
        x = orc.random_coil(N, 3.8 * scale, 21, 0).astype(np.float32).astype(np.float64)
        r = _eval(gpu_api, IF, alpha, x, w)
        eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
        assert r["restraint_rows"] == len(ri)
        tot = abs(eo[0, 0])
        assert np.all(np.abs(r["energies"] - eo[0]) <= 1e-5 * np.maximum(np.abs(eo[0]), 1e-3 * tot)), (scale, r["energies"], eo)
        assert np.abs(r["forces"] - fo[0]).max() <= 1e-5 * np.abs(fo).max(), scale


def test_tile_engine_matches_small_n_engine_on_synthetic_matrix(gpu_api):
    """Synthetic IF at N=700: ragged last tile (700 = 5*128 + 60), zeros, both engines vs fp64."""
    N = 700
    IF = gpu_api.synthetic_if(N, 20260001)
    ref = ds.synthetic_if(N, 20260001)
    assert np.allclose(IF, ref, rtol=1e-11, atol=0) and np.array_equal(IF == 0, ref == 0) and np.array_equal(IF, IF.T)
    ri, rj, rd = gpu_api.if_to_restraints(IF, 1.0)
    p = gpu_api.default_params()
    w = abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)
    x = orc.random_coil(N, 3.0, 5, 2).astype(np.float32).astype(np.float64)
    r = _eval(gpu_api, IF, 1.0, x, w)
    e_small, f_small = gpu_api.energy_forces(x, ri, rj, rd, p, w)
    eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
    assert r["restraint_rows"] == len(ri)
    assert np.allclose(r["energies"], eo[0], rtol=1e-5, atol=1e-5 * abs(eo[0, 0]))
    assert np.abs(r["forces"] - fo[0]).max() <= 1e-5 * np.abs(fo).max()
    assert np.abs(r["forces"] - f_small[0]).max() <= 2e-5 * np.abs(fo).max()


def test_tile_engine_low_soft_and_edge_shapes(gpu_api):
    """Symmetric-soft low side; N just below / at / above a tile edge; coincident beads."""
    w = abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)
    for N in (127, 128, 129, 257):
        IF = ds.synthetic_if(N, 77 + N)
        ri, rj, rd = orc.if_to_restraints(IF, 0.8)
        for low in (0, 1):
            p = gpu_api.default_params()
            p.noe_low_soft = low
            x = orc.random_coil(N, 2.5, 9, 1).astype(np.float32).astype(np.float64)
            x[N // 2] = x[N // 3]
            r = _eval(gpu_api, IF, 0.8, x, w, p)
            eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
            assert np.allclose(r["energies"], eo[0], rtol=1e-5, atol=1e-5 * abs(eo[0, 0])), (N, low)
            assert np.abs(r["forces"] - fo[0]).max() <= 1e-5 * np.abs(fo).max(), (N, low)


def test_tile_range_shards_sum_to_the_whole(gpu_api, mats):
    """The multi-GPU partition (contiguous tile ranges) evaluated shard by shard on one GPU: partial
    forces and energies of the shards add up to the unsharded evaluation (no pair lost or doubled)."""
    IF = mats["chr1_500kb"]
    N = IF.shape[0]
    w = abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=1.0, repel=0.85)
    x = orc.random_coil(N, 3.0, 4, 0).astype(np.float32).astype(np.float64)
    whole = _eval(gpu_api, IF, 1.1, x, w)
    for world in (2, 3, 8):
        parts = [_eval(gpu_api, IF, 1.1, x, w, rank=r, world=world) for r in range(world)]
        assert sum(q["info"]["tiles_this_rank"] for q in parts) == whole["info"]["n_tiles"]
        assert sum(q["restraint_rows"] for q in parts) == whole["restraint_rows"]
        f = sum(q["forces"] for q in parts)
        e = sum(q["energies"] for q in parts)
        assert np.abs(f - whole["forces"]).max() <= 2e-6 * np.abs(whole["forces"]).max()
        assert np.allclose(e, whole["energies"], rtol=1e-9)


def _run_prog(gpu_api, IF, alpha, prog, x0, seed=9):
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    p = gpu_api.default_params()
    parr = abi.stage_array(prog)
    s = gpu_api.Single(IF=IF, alpha=alpha, x0=x0, seed=seed, p=p, program=parr)
    s.run()
    r = s.fetch()
    s.close()
    o = orc.run_model(N, ri, rj, rd, p, parr, seed, 0, x0)
    return r, o


@pytest.mark.parametrize("name", ["chr1_1mb", "chr1_500kb"])
def test_tile_engine_short_md_tracks_fp64(gpu_api, mats, name):
    IF = mats[name]
    x0 = orc.random_coil(IF.shape[0], 3.8, 5, 0).astype(np.float32).astype(np.float64)
    for flag, dt, temp in ((abi.COUPLE_TCOUP, 0.003, 2000.0), (abi.COUPLE_VSCALE, 0.005, 500.0)):
        w = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
        prog = [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                abi.stage(abi.OP_DYNAMICS, nstep=20, flag=flag, dt=dt, temp=temp, **w), abi.stage(abi.OP_ENERGY, **w)]
        r, o = _run_prog(gpu_api, IF, 0.5, prog, x0)
        assert np.abs(r["xyz"] - o["xyz"]).max() < 1e-4
        assert np.allclose(r["energies"], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
        assert r["force_evals"] == o["force_evals"] == 21


def test_tile_engine_lbfgs_reaches_the_fp64_minimum(gpu_api, mats):
    IF = mats["chr1_1mb"]
    x0 = orc.random_coil(IF.shape[0], 3.8, 5, 0).astype(np.float32).astype(np.float64)
    prog = [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=1, drop=10.0), abi.stage(abi.OP_ENERGY)]
    r, o = _run_prog(gpu_api, IF, 0.5, prog, x0)
    assert np.allclose(r["energies"], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
    assert np.abs(r["xyz"] - o["xyz"]).max() < 1e-4
    prog = [abi.stage(abi.OP_INIT, flag=1)] + [abi.stage(abi.OP_MINIMIZE, nstep=2000, drop=10.0)] * 3 + \
           [abi.stage(abi.OP_CENTER), abi.stage(abi.OP_ENERGY)]
    r, o = _run_prog(gpu_api, IF, 0.5, prog, x0)
    assert abs(r["energies"][0] - o["energies"][0]) <= 2e-3 * o["energies"][0]
    assert np.abs(r["xyz"].mean(0)).max() < 1e-3


def test_tile_engine_full_schedule_statistics(gpu_api, mats):
    """Full dgsa.inp schedule on chr1_1mb with both engines from the same start: same physics, so the
    final NOE energy and IF-Spearman agree to ensemble noise."""
    IF = mats["chr1_1mb"]
    p = gpu_api.default_params()
    p.pow_cycl = 3
    small = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=0.5, n_models=6, seed=82364)], p)[0]
    rho_s = gpu_api.spearman_IF_pdb(IF, small.xyz)
    s = gpu_api.Single(IF=IF, alpha=0.5, seed=82364, p=p)
    s.run()
    r = s.fetch()
    s.close()
    rho_l = gpu_api.spearman_IF_pdb(IF, r["xyz"])[0]
    assert np.isfinite(r["xyz"]).all()
    assert abs(rho_l - np.median(rho_s)) <= 0.03
    assert abs(r["energies"][3] - np.median(small.energies[:, 3])) <= 0.05 * np.median(small.energies[:, 3])
    assert np.abs(r["xyz"].mean(0)).max() < 1e-3


def test_ensemble_api_hands_large_structures_to_the_tile_engine(gpu_api):
    """chr3d_run_ensemble accepts structures beyond the cluster kernel's capacity (here 1,300 beads) and anneals them on
    the tile-streaming engine, model after model: same results as driving that engine directly, scored like any job."""
    N = 1300
    IF = ds.synthetic_if(N, 31337)
    w = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                            abi.stage(abi.OP_DYNAMICS, nstep=15, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **w),
                            abi.stage(abi.OP_MINIMIZE, nstep=5, drop=10.0, **w), abi.stage(abi.OP_CENTER), abi.stage(abi.OP_ENERGY, **w)])
    p = gpu_api.default_params()
    small = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    res = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=0.9, n_models=2, seed=5), gpu_api.Job(IF=small, alpha=0.5, n_models=3, seed=5)], p, prog)
    big = res[0]
    ri, rj, rd = gpu_api.if_to_restraints(IF, 0.9)
    assert big.n_restraints == len(ri) and big.xyz.shape == (2, N, 3)
    for k in range(2):
        eo, _ = orc.energy_forces(big.xyz[k], ri, rj, rd, p, prog[-1])
        assert np.allclose(big.energies[k], eo[0], rtol=2e-5, atol=2e-5 * abs(eo[0, 0])), k     # reported energies belong to the reported coordinates
    assert not np.array_equal(big.xyz[0], big.xyz[1])
    sat, dev = orc.assess(big.xyz, ri, rj, rd)
    assert np.array_equal(sat, big.satisfied) and np.allclose(dev, big.sum_dev, rtol=1e-12)     # > 999 beads: summation order is ours
    assert np.abs(big.xyz.mean(1)).max() < 1e-3 and (big.force_evals > 16).all()
    assert np.isfinite(res[1].energies).all() and res[1].xyz.shape == (3, 37, 3)
    # model 1 through the oracle (fp64, same seed / Philox streams) on the MD part alone (the line search takes discrete decisions)
    prog_md = abi.stage_array([prog[0], prog[1], prog[2], prog[5]])
    md = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=0.9, n_models=2, seed=5)], p, prog_md)[0]
    o = orc.run_model(N, ri, rj, rd, p, prog_md, 5, 1, None)
    assert np.abs(md.xyz[1] - o["xyz"]).max() < 1e-3
    assert np.allclose(md.energies[1], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
