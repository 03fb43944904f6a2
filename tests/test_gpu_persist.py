"""GPU tests of the persistent large-N kernel: the multi-rank protocol (tile-range shards, peer-memory force exchange,
redundant integration) run as VIRTUAL ranks of one cooperative launch on a single GPU, and parity at the benchmarked sizes
(N = 4,985 and N = 24,925) against the fp64 oracle working straight from the IF matrix.

Tolerances (BASELINE.json north_star): energies and forces within 1e-5 relative of the fp64 evaluation on identical
coordinates; the restraint table bit-exact (cell by cell)."""
import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import abi, datasets as ds

pytestmark = pytest.mark.gpu

W = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
PROGS = {
    "md": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
           abi.stage(abi.OP_DYNAMICS, nstep=10, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **W), abi.stage(abi.OP_ENERGY, **W)],
    "cool": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
             abi.stage(abi.OP_DYNAMICS, nstep=6, flag=abi.COUPLE_VSCALE, dt=0.005, temp=500.0, **W),
             abi.stage(abi.OP_DYNAMICS, nstep=6, flag=abi.COUPLE_VSCALE, dt=0.005, temp=475.0, **W), abi.stage(abi.OP_ENERGY, **W)],
    "min": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=4, drop=10.0, **W), abi.stage(abi.OP_ENERGY, **W)],
    "pick": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
             abi.stage(abi.OP_DYNAMICS, nstep=4, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **W), abi.stage(abi.OP_SAVE, **W),
             abi.stage(abi.OP_INIT, flag=-1), abi.stage(abi.OP_MAXWELL, temp=300.0),
             abi.stage(abi.OP_DYNAMICS, nstep=4, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **W), abi.stage(abi.OP_PICK, **W),
             abi.stage(abi.OP_CENTER), abi.stage(abi.OP_ENERGY, **W)],
}


def _one(gpu_api, IF, prog, alpha=1.0, seed=11):
    s = gpu_api.Single(IF=IF, alpha=alpha, seed=seed, p=gpu_api.default_params(), program=prog)
    s.run()
    r = s.fetch()
    r["profile"] = s.profile()
    s.close()
    return r


def _group(gpu_api, IF, prog, world, alpha=1.0, seed=11):
    ss = [gpu_api.Single(IF=IF, alpha=alpha, seed=seed, p=gpu_api.default_params(), program=prog, rank=g, world=world) for g in range(world)]
    try:
        gpu_api.group_connect(ss)
        gpu_api.group_run(ss)
        out = [s.fetch() for s in ss]
        for s, o in zip(ss, out):
            o["info"] = s.info()
            assert s.comm_error() == 0
    finally:
        for s in ss:
            s.close()
    return out


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["md", "cool", "min", "pick"])
def test_virtual_ranks_match_one_rank(gpu_api, world, name):
    """The sharded run (world tile ranges, partial forces exchanged through the peer buffers inside the persistent kernel,
    every rank integrating all beads) against the unsharded run: only the fp32 summation order of the partial forces differs,
    so coordinates agree to 1e-3 A and energies to 1e-5 relative on these short runs; the ranks agree with one another
    bit for bit (they add the same slots in the same order)."""
    N = 700                                     # 6 tile rows, 21 tiles, ragged last tile
    IF = ds.synthetic_if(N, 4242)
    prog = abi.stage_array(PROGS[name])
    one = _one(gpu_api, IF, prog)
    grp = _group(gpu_api, IF, prog, world)
    assert sum(o["info"]["tiles_this_rank"] for o in grp) == 21
    assert sum(o["restraint_rows"] for o in grp) == one["restraint_rows"]
    for o in grp[1:]:
        assert np.array_equal(o["xyz"], grp[0]["xyz"]) and np.array_equal(o["energies"], grp[0]["energies"]) and np.array_equal(o["forces"], grp[0]["forces"])
        assert o["force_evals"] == grp[0]["force_evals"] and o["hand"] == grp[0]["hand"]
    g = grp[0]
    assert g["force_evals"] == one["force_evals"] and g["hand"] == one["hand"]
    assert np.allclose(g["energies"], one["energies"], rtol=1e-5, atol=1e-6 * abs(one["energies"][0])), (g["energies"], one["energies"])
    assert np.abs(g["xyz"] - one["xyz"]).max() < 1e-3


def test_persistent_run_tracks_fp64_and_reports_its_phases(gpu_api):
    """One rank: 20 MD steps against the fp64 oracle (1e-4 A), and the phase clock of the persistent kernel adds up."""
    IF = ds.load_if_matrix(ds.input_path("chr1_500kb"))
    N = IF.shape[0]
    x0 = orc.random_coil(N, 3.8, 5, 0).astype(np.float32).astype(np.float64)
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                            abi.stage(abi.OP_DYNAMICS, nstep=20, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **W), abi.stage(abi.OP_ENERGY, **W)])
    p = gpu_api.default_params()
    s = gpu_api.Single(IF=IF, alpha=0.5, x0=x0, seed=9, p=p, program=prog)
    s.run()
    r = s.fetch()
    prof = s.profile()
    s.close()
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    o = orc.run_model(N, ri, rj, rd, p, prog, 9, 0, x0)
    assert np.abs(r["xyz"] - o["xyz"]).max() < 1e-4
    assert np.allclose(r["energies"], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
    assert r["force_evals"] == o["force_evals"] == 21
    assert r["launches"] <= 12                                # one persistent launch per op (+ small vector kernels)
    assert prof["tile_ms"] > 0 and prof["step_ms"] > 0 and sum(prof.values()) <= r["device_ms"] * 1.05


@pytest.mark.parametrize("N,seed", [(4985, 20260001), (24925, 20260002)])
def test_benchmarked_sizes_against_fp64(gpu_api, N, seed):
    """BASELINE.json configs[3] / configs[4] (synthetic 50 kb / 10 kb chr1): the target table the engine derives on the device
    (beyond 4,096 beads: blocked mean, device pow) equals the oracle's `%.1f` table cell by cell, and energies / forces on a
    random coil agree with the fp64 evaluation to 1e-5."""
    IF = gpu_api.synthetic_if(N, seed)
    p = gpu_api.default_params()
    w = abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)
    x = orc.random_coil(N, 3.8, 3, 0).astype(np.float32).astype(np.float64)
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), w])
    s = gpu_api.Single(IF=IF, alpha=1.1, x0=x, p=p, program=prog)
    s.run()
    r = s.fetch()
    T = s.targets()
    s.close()
    chk = orc.large_check(IF, 1.1, p, w, gpu_targets=T, x=x)
    del T
    assert chk["rows"] == r["restraint_rows"]
    # device pow is not correctly rounded: a cell whose D*10 sits within ~1e-15 relative of a rounding boundary may land on the
    # other side (expected count over 3.1e8 cells: < 1e-3); anything else is a bug
    assert chk["mismatches"] <= 2 and chk["max_diff"] <= 0.1001, chk
    eo, fo = chk["energies"], chk["forces"]
    assert np.allclose(r["energies"], eo, rtol=1e-5, atol=1e-5 * abs(eo[0])), (r["energies"], eo)
    assert np.abs(r["forces"] - fo).max() <= 1e-5 * np.abs(fo).max(), np.abs(r["forces"] - fo).max() / np.abs(fo).max()


def test_virtual_ranks_at_50kb(gpu_api):
    """N = 4,985 (780 tiles) on 8 virtual ranks: the exchange protocol with real block counts (39 bead blocks, bmin skipping)."""
    N = 4985
    p = gpu_api.default_params()
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                            abi.stage(abi.OP_DYNAMICS, nstep=5, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **W),
                            abi.stage(abi.OP_MINIMIZE, nstep=1, drop=1.0e8, **W), abi.stage(abi.OP_ENERGY, **W)])    # (one L-BFGS iteration with a
    # first step large enough for fp32 coordinates to resolve — with CNS's drop = 10 on energies of 4e10 the trial point equals the
    # start and acceptance is decided by rounding noise; a longer minimisation may legitimately take a different number of
    # trials once the fp32 summation order differs)
    one = gpu_api.Single(N=N, alpha=1.1, synth_seed=20260001, seed=5, p=p, program=prog)
    one.run()
    o = one.fetch()
    one.close()
    ss = [gpu_api.Single(N=N, alpha=1.1, synth_seed=20260001, seed=5, p=p, program=prog, rank=g, world=8) for g in range(8)]
    try:
        gpu_api.group_connect(ss)
        gpu_api.group_run(ss)
        rs = [s.fetch() for s in ss]
    finally:
        for s in ss:
            s.close()
    assert sum(r["restraint_rows"] for r in rs) == o["restraint_rows"]
    for r in rs[1:]:
        assert np.array_equal(r["xyz"], rs[0]["xyz"])
    assert rs[0]["force_evals"] == o["force_evals"]
    assert np.allclose(rs[0]["energies"], o["energies"], rtol=2e-5, atol=1e-6 * abs(o["energies"][0]))
    assert np.abs(rs[0]["xyz"] - o["xyz"]).max() < 2e-3
