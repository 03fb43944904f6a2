"""Round-2 parity tests (VERDICT r1 'next round' item 3): the benchmarked ensemble against the fp64 oracle ensemble, and the
mirror-image pick (C3D:1711-1727) on cases with a clear energy gap."""
import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import abi, datasets as ds

pytestmark = pytest.mark.gpu


def test_benchmarked_ensemble_matches_oracle_ensemble(gpu_api):
    """BASELINE.json configs[1] itself — chr1_500kb, alpha = 1.1, 20 models, full dgsa.inp schedule — GPU ensemble against the fp64
    CPU ensemble on the same inputs and seeds: final NOE energy distribution, satisfied-restraint counts and median
    IF-Spearman (north_star tolerance 0.02)."""
    IF = ds.load_if_matrix(ds.input_path("chr1_500kb"))
    N, m, alpha = IF.shape[0], 20, 1.1
    p = gpu_api.default_params()
    res = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=alpha, n_models=m, seed=82364)], p)[0]
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    assert res.n_restraints == len(ri) == 101426
    o = orc.run_ensemble(N, ri, rj, rd, p, orc.build_program(p), 82364, m)
    e, eo = res.energies[:, 3], o["energies"][:, 3]
    assert abs(np.median(e) - np.median(eo)) <= 5e-3 * np.median(eo), (np.median(e), np.median(eo))
    assert abs(e.min() - eo.min()) <= 1e-2 * eo.min() and abs(e.max() - eo.max()) <= 2e-2 * eo.max()
    rho, rho_o = gpu_api.spearman_IF_pdb(IF, res.xyz), orc.spearman(IF, o["xyz"])
    assert abs(np.median(rho) - np.median(rho_o)) <= 0.02, (np.median(rho), np.median(rho_o))
    sat_o, _ = orc.assess(o["xyz"], ri, rj, rd)
    assert abs(np.median(res.satisfied) - np.median(sat_o)) <= max(3, 0.05 * np.median(sat_o)), (np.median(res.satisfied), np.median(sat_o))
    sat_chk, dev_chk = orc.assess(res.xyz, ri, rj, rd)
    assert np.array_equal(sat_chk, res.satisfied) and np.allclose(dev_chk, res.sum_dev, rtol=1e-12)
    assert np.array_equal(res.hand != 0, np.ones(m, bool)) and np.abs(res.xyz.mean(1)).max() < 1e-3


def _pick_prog(long_first):
    w = dict(w_noe=10.0, w_bond=1000.0, w_vdw=0.003, repel=0.9)
    long_min = [abi.stage(abi.OP_MINIMIZE, nstep=3, drop=2.0e5, **w)]        # lowers the energy by several per cent
    none = [abi.stage(abi.OP_MINIMIZE, nstep=0, drop=2.0e5, **w)]            # evaluation only
    a, b = (long_min, none) if long_first else (none, long_min)
    return [abi.stage(abi.OP_INIT, flag=1)] + a + [abi.stage(abi.OP_SAVE, **w), abi.stage(abi.OP_INIT, flag=-1)] + b + \
           [abi.stage(abi.OP_PICK, **w), abi.stage(abi.OP_ENERGY, **w)]


@pytest.mark.parametrize("engine", ["cluster", "tiles"])
@pytest.mark.parametrize("long_first", [True, False])
def test_mirror_image_pick_keeps_the_lower_energy_image(gpu_api, engine, long_first):
    """INIT(+1) ... SAVE, INIT(-1) ... PICK (C3D:1701-1727): one image is minimised, the other is not, so the energies differ by
    several per cent and the choice is unambiguous — the kept image is the minimised one, it equals the oracle's choice, and the
    reported energies / coordinates are those of that image."""
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    p = gpu_api.default_params()
    x0 = orc.random_coil(N, 3.8, 5, 0).astype(np.float32).astype(np.float64)
    prog = abi.stage_array(_pick_prog(long_first))
    o = orc.run_model(N, ri, rj, rd, p, prog, 9, 0, x0)
    if engine == "cluster":
        r = gpu_api.build_models([gpu_api.Job(restraints=(ri, rj, rd), N=N, n_models=1, seed=9, x0=x0[None])], p, prog)[0]
        hand, e, xyz = int(r.hand[0]), r.energies[0], r.xyz[0]
    else:
        s = gpu_api.Single(IF=IF, alpha=0.5, x0=x0, seed=9, p=p, program=prog)
        s.run()
        q = s.fetch()
        s.close()
        hand, e, xyz = q["hand"], q["energies"], q["xyz"]
    want = 1 if long_first else -1
    assert o["hand"] == want and hand == want
    # the unminimised image would have the start energy: the kept one is clearly below it
    w = prog[-1]
    e_start, _ = orc.energy_forces(x0, ri, rj, rd, p, w)
    assert e[0] < 0.99 * e_start[0, 0] and o["energies"][0] < 0.99 * e_start[0, 0]
    assert np.allclose(e, o["energies"], rtol=1e-4, atol=1e-4 * o["energies"][0]), (e, o["energies"])
    assert np.abs(xyz - o["xyz"]).max() < 1e-2
    # image -1 starts from the mirrored coil: its x coordinates keep the mirrored sign pattern
    assert np.sign(np.sum(xyz[:, 0] * x0[:, 0])) == want


@pytest.mark.parametrize("name,alpha", [("chr21_1mb", 0.5), ("chr1_500kb", 0.5)])
def test_angle_stand_in_term_matches_fp64(gpu_api, name, alpha):
    """The CA(i)-CA(i+2) soft lower bound (chr3d_sa_params.angle_d2 / angle_k, the stand-in for CNS's angle term, C3D:1641) in
    both engines against the fp64 oracle: energies (reported with the bonded term) and forces within 1e-5, on a coil compact
    enough that the term is active for many beads."""
    IF = ds.load_if_matrix(ds.input_path(name))
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    p = gpu_api.default_params()
    p.angle_d2, p.angle_k = 6.0, 500.0
    w = abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)
    x = orc.random_coil(N, 2.6, 21, 0).astype(np.float32).astype(np.float64)
    eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
    p0 = gpu_api.default_params()
    e0, _ = orc.energy_forces(x, ri, rj, rd, p0, w)
    assert eo[0, 1] > 1.05 * e0[0, 1]                                       # the term contributes
    e, f = gpu_api.energy_forces(x, ri, rj, rd, p, w)                       # cluster engine
    assert np.allclose(e[0], eo[0], rtol=1e-5, atol=1e-5 * abs(eo[0, 0]))
    assert np.abs(f[0] - fo[0]).max() <= 1e-5 * np.abs(fo).max()
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), w])
    s = gpu_api.Single(IF=IF, alpha=alpha, x0=x, p=p, program=prog)        # tile engine
    s.run()
    r = s.fetch()
    s.close()
    assert np.allclose(r["energies"], eo[0], rtol=1e-5, atol=1e-5 * abs(eo[0, 0]))
    assert np.abs(r["forces"] - fo[0]).max() <= 1e-5 * np.abs(fo).max()
    # and a short trajectory with the term on tracks the oracle
    ww = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                            abi.stage(abi.OP_DYNAMICS, nstep=10, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **ww), abi.stage(abi.OP_ENERGY, **ww)])
    r = gpu_api.build_models([gpu_api.Job(restraints=(ri, rj, rd), N=N, n_models=1, seed=9, x0=x[None])], p, prog)[0]
    o = orc.run_model(N, ri, rj, rd, p, prog, 9, 0, x)
    assert np.abs(r.xyz[0] - o["xyz"]).max() < 1e-4
    assert np.allclose(r.energies[0], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
