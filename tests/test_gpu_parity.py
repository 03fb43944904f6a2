"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical inputs.

Tolerances (BASELINE.json north_star): restraint table and Spearman bit-exact; energies/forces
within 1e-5 relative of the fp64 evaluation; SA ensembles compared on final-energy distribution,
restraint-violation count and median IF-Spearman (0.02)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import abi, datasets as ds

pytestmark = pytest.mark.gpu


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


@pytest.fixture(scope="module")
def mats():
    return {n: ds.load_if_matrix(ds.input_path(n)) for n in ("chr21_1mb", "chr22_1mb", "chr1_1mb", "chr21_500kb", "chr1_500kb")}


# ------------------------------------------------------------------------------- (1) restraints
@pytest.mark.parametrize("name,alpha", [("chr21_1mb", 0.1), ("chr21_1mb", 0.5), ("chr21_1mb", 1.1), ("chr21_1mb", 2.0),
                                        ("chr22_1mb", 0.5), ("chr22_1mb", 1.7), ("chr1_1mb", 1.1), ("chr21_500kb", 1.1),
                                        ("chr1_500kb", 1.1), ("chr1_500kb", 0.5)])
def test_restraint_table_bit_exact(gpu_api, mats, golden_dir, tmp_path, name, alpha):
    """Device table == oracle == the files the real Perl wrote (md5 of .dist/.rr/contact.tbl)."""
    IF = mats[name]
    t, npatched = gpu_api.IF2dist_new(IF, alpha)
    ri, rj, rd = gpu_api.if_to_restraints(IF, alpha)
    to = orc.if_to_dist(IF, alpha)
    ori, orj, ord_ = orc.dist_to_restraints(to)
    assert np.array_equal(t, to)
    assert np.array_equal(ri, ori) and np.array_equal(rj, orj) and np.array_equal(rd, ord_)
    g = json.load(open(os.path.join(golden_dir, "restraints_md5.json")))
    key = "%s@%.1f" % (name, alpha)
    if key in g:
        gpu_api.write_dist(str(tmp_path / "a.dist"), t)
        gpu_api.write_rr(str(tmp_path / "a.rr"), ri, rj, rd)
        gpu_api.carr2tbl(str(tmp_path / "a.tbl"), ri, rj, rd)
        assert _md5(tmp_path / "a.dist") == g[key]["dist"]
        assert _md5(tmp_path / "a.rr") == g[key]["rr"]
        assert _md5(tmp_path / "a.tbl") == g[key]["tbl"]


def test_restraint_edge_cases(gpu_api):
    """zeros (-1.0 cells), exact decimal ties, tiny and ragged matrices, min_sep larger than L."""
    rs = np.random.RandomState(5)
    for L in (2, 5, 6, 33, 64, 97):
        a = rs.lognormal(3.0, 2.0, size=(L, L))
        IF = np.triu(a) + np.triu(a, 1).T
        IF[rs.rand(L, L) < 0.1] = 0.0
        IF = np.triu(IF) + np.triu(IF, 1).T
        for alpha in (0.3, 1.0, 2.0):
            t, _ = gpu_api.IF2dist_new(IF, alpha, 11.0)
            assert np.array_equal(t, orc.if_to_dist(IF, alpha, 11.0)), (L, alpha)
            for ms in (1, 5, L + 3):
                r = gpu_api.if_to_restraints(IF, alpha, 11.0, ms)
                o = orc.dist_to_restraints(orc.if_to_dist(IF, alpha, 11.0), ms)
                assert all(np.array_equal(x, y) for x, y in zip(r, o)), (L, alpha, ms)
    # exact ties of the decimal rounding: D = K/(IF/mean) with IF chosen so that D = x.x5 exactly
    IF = np.full((4, 4), 1.0)
    IF[0, 1] = IF[1, 0] = 8.0      # alpha=1: mean = (14+16)/16, D = 11*mean/8 ...
    t, npatched = gpu_api.IF2dist_new(IF, 1.0, 11.0)
    assert np.array_equal(t, orc.if_to_dist(IF, 1.0, 11.0))
    IF = np.full((2, 2), 4.0)       # D = 0.25*... : K=0.25 -> D = 0.25 exactly -> "%.1f" -> 0.2 (half-even)
    t, npatched = gpu_api.IF2dist_new(IF, 1.0, 0.25)
    assert np.array_equal(t, orc.if_to_dist(IF, 1.0, 0.25)) and t[0, 0] == 2 and npatched >= 1


# ------------------------------------------------------------------------------- (2) energy / forces
@pytest.mark.parametrize("name,alpha", [("chr21_1mb", 0.5), ("chr21_1mb", 1.1), ("chr1_1mb", 0.5), ("chr1_500kb", 1.1)])
def test_energy_forces_match_fp64(gpu_api, mats, name, alpha):
    """NOE soft-square + bond + repel on identical coordinates: <= 1e-5 relative (per term and total,
    max force component), for compact and extended chains and for the hot / cool / final weights."""
    IF = mats[name]
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    p = gpu_api.default_params()
    cases = [(1.0, abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=1.0, repel=0.85)),
             (0.4, abi.stage(abi.OP_ENERGY, w_noe=1, w_bond=100, w_vdw=20.0, repel=0.5)),
             (0.7, abi.stage(abi.OP_ENERGY, w_noe=4, w_bond=400, w_vdw=0.003, repel=0.9))]
    for scale, w in cases:
        x = np.stack([orc.random_coil(N, 3.8 * scale, 11, m) for m in range(3)]).astype(np.float32).astype(np.float64)
        e, f = gpu_api.energy_forces(x, ri, rj, rd, p, w)
        eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
        tot = np.abs(eo[:, 0:1])
        assert np.all(np.abs(e - eo) <= 1e-5 * np.maximum(np.abs(eo), 1e-3 * tot)), (scale, e, eo)
        assert np.abs(f - fo).max() <= 1e-5 * np.abs(fo).max(), scale


def test_energy_forces_general_soft_square(gpu_api, mats):
    """Non-default soft-square constants (b != 0) and the symmetric-soft low side use the exact piecewise form."""
    IF = mats["chr21_1mb"]
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    w = abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)
    x = orc.random_coil(N, 3.0, 3, 1).astype(np.float32).astype(np.float64)
    for asym, rsw, low in ((1.0, 0.5, 0), (3.0, 2.0, 1), (2.0, 1.0, 1)):
        p = gpu_api.default_params()
        p.noe_asymptote, p.noe_rswitch, p.noe_low_soft = asym, rsw, low
        e, f = gpu_api.energy_forces(x, ri, rj, rd, p, w)
        eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
        assert np.allclose(e, eo, rtol=1e-5, atol=1e-5 * abs(eo[0, 0]))
        assert np.abs(f - fo).max() <= 1e-5 * np.abs(fo).max()


def test_energy_restraint_table_edge_cases(gpu_api):
    """Empty table, a single restraint, coincident beads."""
    p = gpu_api.default_params()
    w = abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)
    x = orc.random_coil(9, 3.8, 1, 0).astype(np.float32).astype(np.float64)
    empty = (np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
    e, f = gpu_api.energy_forces(x, *empty, p, w)
    eo, fo = orc.energy_forces(x, *empty, p, w)
    assert np.allclose(e, eo, rtol=1e-5, atol=1e-9) and np.allclose(f, fo, rtol=1e-4, atol=1e-5 * max(1.0, np.abs(fo).max()))
    one = (np.array([0], np.int32), np.array([8], np.int32), np.array([7.5]))
    x[3] = x[6]      # coincident non-bonded beads: repel active at r = 0
    e, f = gpu_api.energy_forces(x, *one, p, w)
    eo, fo = orc.energy_forces(x, *one, p, w)
    assert np.allclose(e, eo, rtol=1e-5, atol=1e-9) and np.abs(f - fo).max() <= 1e-5 * np.abs(fo).max()


# ------------------------------------------------------------------------------- (3) integrator / minimiser
def _run_both(gpu_api, IF, alpha, prog, seed=9, N=None):
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    p = gpu_api.default_params()
    x0 = orc.random_coil(N, 3.8, 5, 0).astype(np.float32).astype(np.float64)
    parr = abi.stage_array(prog)
    r = gpu_api.build_models([gpu_api.Job(restraints=(ri, rj, rd), N=N, n_models=1, seed=seed, x0=x0[None])], p, parr)[0]
    o = orc.run_model(N, ri, rj, rd, p, parr, seed, 0, x0)
    return r, o


@pytest.mark.parametrize("name", ["chr21_1mb", "chr1_1mb"])
def test_short_md_tracks_fp64(gpu_api, mats, name):
    """20 Verlet steps (Maxwell start, both coupling modes): coordinates within 1e-4 A, energies 1e-5."""
    for flag, dt, temp in ((abi.COUPLE_TCOUP, 0.003, 2000.0), (abi.COUPLE_VSCALE, 0.005, 500.0)):
        w = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
        prog = [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                abi.stage(abi.OP_DYNAMICS, nstep=20, flag=flag, dt=dt, temp=temp, **w), abi.stage(abi.OP_ENERGY, **w)]
        r, o = _run_both(gpu_api, mats[name], 0.5, prog)
        assert np.abs(r.xyz[0] - o["xyz"]).max() < 1e-4
        assert np.allclose(r.energies[0], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
        assert r.force_evals[0] == o["force_evals"] == 21


@pytest.mark.parametrize("name", ["chr21_1mb", "chr1_1mb"])
def test_lbfgs_reaches_the_fp64_minimum(gpu_api, mats, name):
    """L-BFGS (vector-free form on the device, textbook two-loop in the oracle): same first steps,
    same converged energy."""
    # first iteration = steepest descent with the drop-sized step: must agree tightly.  (From the second
    # iteration on the fp32 curvature pair y = g' - g of that tiny step carries ~1e-1 relative noise, so
    # paths differ until both converge to the same minimum.)
    prog = [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=1, drop=10.0), abi.stage(abi.OP_ENERGY)]
    r, o = _run_both(gpu_api, mats[name], 0.5, prog)
    assert np.allclose(r.energies[0], o["energies"], rtol=1e-5, atol=1e-5 * o["energies"][0])
    assert np.abs(r.xyz[0] - o["xyz"]).max() < 1e-4
    prog = [abi.stage(abi.OP_INIT, flag=1)] + [abi.stage(abi.OP_MINIMIZE, nstep=2000, drop=10.0)] * 3 + \
           [abi.stage(abi.OP_CENTER), abi.stage(abi.OP_ENERGY)]
    r, o = _run_both(gpu_api, mats[name], 0.5, prog)
    assert abs(r.energies[0, 0] - o["energies"][0]) <= 2e-3 * o["energies"][0]
    assert np.abs(r.xyz[0].mean(0)).max() < 1e-3


@pytest.mark.parametrize("name", ["chr21_1mb", "chr1_1mb"])
def test_lbfgs_iterations_track_fp64(gpu_api, mats, name):
    """The minimiser itself follows the oracle iteration by iteration (same trial counts, energies to 2e-5, coordinates to 2e-3 A
    over 20 iterations) as soon as the first step is one fp32 coordinates can resolve.  With CNS's drop = 10 on energies of
    1e6-1e7 the first step is drop / |g|^2 ~ 1e-5 A = two ulps of a coordinate, the curvature pair of that step is rounding
    noise in any fp32 engine (profiles/r02_lbfgs_tracking.log), and only the converged state is comparable (test above)."""
    for drop, e_tol, x_tol in ((1.0e4, 2e-5, 2e-3), (1.0e6, 2e-6, 2e-4)):
        for k in (5, 20):
            prog = [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=k, drop=drop), abi.stage(abi.OP_ENERGY)]
            r, o = _run_both(gpu_api, mats[name], 0.5, prog)
            assert r.force_evals[0] == o["force_evals"]
            assert abs(r.energies[0, 0] - o["energies"][0]) <= e_tol * o["energies"][0], (drop, k)
            assert np.abs(r.xyz[0] - o["xyz"]).max() < x_tol, (drop, k)


def test_mirror_image_pick(gpu_api, mats):
    """INIT(+1)...SAVE, INIT(-1)...PICK keeps the lower-energy image (C3D:1711-1727)."""
    w = dict(w_noe=10.0, w_bond=1000.0, w_vdw=0.003, repel=0.9)
    body = [abi.stage(abi.OP_MINIMIZE, nstep=20, drop=0.0, **w)]
    prog = [abi.stage(abi.OP_INIT, flag=1)] + body + [abi.stage(abi.OP_SAVE, **w), abi.stage(abi.OP_INIT, flag=-1)] + body + \
           [abi.stage(abi.OP_PICK, **w), abi.stage(abi.OP_ENERGY, **w)]
    r, o = _run_both(gpu_api, mats["chr21_1mb"], 0.5, prog)
    assert r.hand[0] in (1, -1)
    # energies of the two images are nearly degenerate; whichever was kept must be what was reported
    assert np.isfinite(r.energies[0]).all()


# ------------------------------------------------------------------------------- (3) full ensembles
@pytest.mark.parametrize("name,alpha,m", [("chr21_1mb", 0.5, 10), ("chr22_1mb", 0.5, 8), ("chr1_1mb", 0.5, 8), ("chr21_1mb", 1.1, 10)])
def test_sa_ensemble_matches_oracle_ensemble(gpu_api, mats, name, alpha, m):
    """Full dgsa.inp schedule, GPU ensemble vs fp64 CPU ensemble on the same inputs: final NOE energy
    distribution (median within 0.5 %), satisfied-restraint counts, median IF-Spearman within 0.02."""
    IF = mats[name]
    N = IF.shape[0]
    p = gpu_api.default_params()
    res = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=alpha, n_models=m, seed=82364)], p)[0]
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    o = orc.run_ensemble(N, ri, rj, rd, p, orc.build_program(p), 82364, m)
    assert res.n_restraints == len(ri)
    assert abs(np.median(res.energies[:, 3]) - np.median(o["energies"][:, 3])) <= 5e-3 * np.median(o["energies"][:, 3])
    assert abs(res.energies[:, 3].min() - o["energies"][:, 3].min()) <= 1e-2 * o["energies"][:, 3].min()
    rho, rho_o = gpu_api.spearman_IF_pdb(IF, res.xyz), orc.spearman(IF, o["xyz"])
    assert abs(np.median(rho) - np.median(rho_o)) <= 0.02
    sat_o, _ = orc.assess(o["xyz"], ri, rj, rd)
    assert abs(np.median(res.satisfied) - np.median(sat_o)) <= max(3, 0.05 * np.median(sat_o))
    # the run reports what the reference's own scorer would report for these coordinates
    sat_chk, dev_chk = orc.assess(res.xyz, ri, rj, rd)
    assert np.array_equal(sat_chk, res.satisfied) and np.array_equal(dev_chk, res.sum_dev)
    assert np.abs(res.xyz.mean(1)).max() < 1e-3                      # centred (C3D:1805-1816)
    assert res.force_evals.max() < 50000                             # minimiser stops (ftol/gtol)


def test_sa_against_shipped_cns_models(gpu_api, mats):
    """Weak pin against CNS itself: at alpha=0.5 the ensemble reproduces the statistics
    of the shipped CNS models (IF-Spearman within 0.02, bond lengths, radius of gyration)."""
    anchors = {"chr21_1mb": ("chr21_1mb_rank07_a11.pdb", -0.845), "chr1_1mb": ("chr1_1mb_rank09_a11.pdb", -0.862)}
    for name, (pdb, rho_ship) in anchors.items():
        IF = mats[name]
        res = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=0.5, n_models=10, seed=82364)])[0]
        rho = gpu_api.spearman_IF_pdb(IF, res.xyz)
        assert abs(np.median(rho) - rho_ship) <= 0.02, (name, np.median(rho))
        ship = ds.read_ca_pdb(os.path.join(ds.shipped_models_dir(), pdb))
        rg = lambda x: np.sqrt(((x - x.mean(0)) ** 2).sum(1).mean())
        assert abs(np.mean([rg(x) for x in res.xyz]) - rg(ship)) < 0.1 * rg(ship)


def test_runs_are_reproducible_and_seed_dependent(gpu_api, mats):
    IF = mats["chr21_1mb"]
    p = gpu_api.default_params()
    p.pow_cycl = 2
    a = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=0.5, n_models=3, seed=7)], p)[0]
    b = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=0.5, n_models=3, seed=7)], p)[0]
    c = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=0.5, n_models=3, seed=8)], p)[0]
    assert np.array_equal(a.xyz, b.xyz) and np.array_equal(a.energies, b.energies)
    assert not np.array_equal(a.xyz, c.xyz)


def test_many_jobs_one_call(gpu_api, mats):
    """Several (matrix, alpha) jobs of different size in one call (the config-3 pattern)."""
    p = gpu_api.default_params()
    p.pow_cycl = 1
    jobs = [gpu_api.Job(IF=mats["chr21_1mb"], alpha=0.5, n_models=3, seed=1), gpu_api.Job(IF=mats["chr22_1mb"], alpha=0.7, n_models=2, seed=2),
            gpu_api.Job(IF=mats["chr21_500kb"], alpha=0.5, n_models=2, seed=3)]
    res = gpu_api.build_models(jobs, p)
    single = gpu_api.build_models([jobs[1]], p)[0]
    assert [r.xyz.shape for r in res] == [(3, 37, 3), (2, 35, 3), (2, 73, 3)]
    assert all(np.isfinite(r.energies).all() for r in res)
    assert np.array_equal(res[1].xyz, single.xyz)


# ------------------------------------------------------------------------------- (4) assessment
def test_assess_bit_exact(gpu_api, mats, golden_dir):
    """count/sum_dev on the device == real assess_dgsa output (golden) == oracle, incl. the 1000 A cap."""
    g = json.load(open(os.path.join(golden_dir, "assess.json")))
    for name, rec in g.items():
        IF = ds.load_if_matrix(ds.input_path(name))
        ri, rj, rd = orc.if_to_restraints(IF, float(rec["alpha"]))
        xyz = np.array(rec["xyz"])
        sat, dev = gpu_api.count_satisfied_and_sum_dev(xyz, ri, rj, rd)
        for k in range(rec["m"]):
            row = rec["table"]["%s_matrix_%d" % (name, k + 1)]
            assert sat[k] == row["count"] and "%.2f" % dev[k] == row["sum_dev"]
        so, do = orc.assess(xyz, ri, rj, rd)
        assert np.array_equal(sat, so) and np.array_equal(dev, do)
    # shipped CNS model known answer: 68/528, 2955.67
    IF = mats["chr21_1mb"]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    xyz = ds.read_ca_pdb(os.path.join(ds.shipped_models_dir(), "chr21_1mb_rank07_a11.pdb"))
    sat, dev = gpu_api.count_satisfied_and_sum_dev(xyz, ri, rj, rd)
    assert sat[0] == 68 and "%.2f" % dev[0] == "2955.67"


def test_assess_random_coordinates_large_table(gpu_api, mats):
    IF = mats["chr1_500kb"]
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    rs = np.random.RandomState(2)
    xyz = np.stack([orc.random_coil(N, 3.8, 1, m) * 0.3 + rs.normal(0, 1e-4, (N, 3)) for m in range(3)])
    sat, dev = gpu_api.count_satisfied_and_sum_dev(xyz, ri, rj, rd)
    so, do = orc.assess(xyz, ri, rj, rd)
    assert np.array_equal(sat, so) and np.array_equal(dev, do)


# ------------------------------------------------------------------------------- (5) Spearman
def test_spearman_bit_exact(gpu_api, mats, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "spearman.json")))
    for key, val in g.items():
        name, pdb, rng = key.split("|")
        IF = ds.load_if_matrix(ds.input_path(name))
        xyz = ds.read_ca_pdb(os.path.join(ds.shipped_models_dir(), pdb))
        rho = gpu_api.spearman_IF_pdb(IF, xyz, int(rng))
        assert "%.3f" % rho[0] == val, key
        assert rho[0] == orc.spearman(IF, xyz, int(rng))[0], key        # fp64 bit-identical


def test_spearman_ties_and_edges(gpu_api):
    rs = np.random.RandomState(3)
    L = 40
    IF = rs.randint(0, 6, size=(L, L)).astype(np.float64)     # heavy ties, zeros
    IF = np.triu(IF) + np.triu(IF, 1).T
    xyz = np.round(rs.normal(0, 3, (4, L, 3)), 1)             # coarse grid -> tied distances
    xyz[1] = xyz[0]
    for rng in (0, 1, 3, 39):
        assert np.array_equal(gpu_api.spearman_IF_pdb(IF, xyz, rng), orc.spearman(IF, xyz, rng)), rng
    assert np.isnan(gpu_api.spearman_IF_pdb(IF, xyz, 40)).all()   # SP:38-41
    const = np.ones((L, L))
    assert np.array_equal(gpu_api.spearman_IF_pdb(const, xyz, 3), np.zeros(4))   # zero variance -> 0


def test_seeded_random_coil_start_equals_oracle(gpu_api):
    """Trajectories without caller-supplied coordinates start from the Philox-seeded random coil; the device builds
    it with the same additions in the same order as the oracle (bit-identical after the fp32 conversion)."""
    for N, m, seed in ((37, 3, 82364), (455, 2, 7), (900, 1, 99)):
        prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1)])
        ri, rj, rd = np.array([0], np.int32), np.array([N - 1], np.int32), np.array([10.0])
        r = gpu_api.build_models([gpu_api.Job(restraints=(ri, rj, rd), N=N, n_models=m, seed=seed)], gpu_api.default_params(), prog)[0]
        for k in range(m):
            ref = orc.random_coil(N, 3.8, seed, k).astype(np.float32).astype(np.float64)
            assert np.array_equal(r.xyz[k], ref), (N, k)


def test_images_side_by_side_equal_the_sequential_run(gpu_api, mats):
    """The host may run the two hot-MD images of a model as trajectories of their own and resume at PICK (api.cu,
    ensemble_create).  With the launch shape pinned, that run must be bit-identical to the one-launch run: same Philox
    streams, same reduction orders, same stage program."""
    IF = mats["chr21_500kb"]
    job = lambda: [gpu_api.Job(IF=IF, alpha=0.5, n_models=3, seed=11)]
    try:
        gpu_api.set_option("sa_cluster", 2); gpu_api.set_option("sa_warps", 8)
        e = gpu_api.Ensemble(job()); split_launches = e.info()["launches"]; e.close()
        a = gpu_api.build_models(job())[0]
        gpu_api.set_option("sa_split", 0)
        e = gpu_api.Ensemble(job()); one_launch = e.info()["launches"]; e.close()
        b = gpu_api.build_models(job())[0]
    finally:
        for k in ("sa_cluster", "sa_warps", "sa_split"):
            gpu_api.set_option(k, None)
    assert split_launches == 2 and one_launch == 1
    assert np.array_equal(a.xyz, b.xyz) and np.array_equal(a.energies, b.energies)
    assert np.array_equal(a.force_evals, b.force_evals) and np.array_equal(a.hand, b.hand)
    assert set(a.hand.tolist()) <= {1, -1}


def test_co_resident_clusters_complete(gpu_api):
    """Regression for a phase race in the cluster exchange ("Findings"): with several CTAs per SM the skew between
    the CTAs of a cluster is large enough to expose any pair of exchange phases that lacks a separating dependency; the shape
    4 CTAs x 4 warps with 120 trajectories used to hang 3 times out of 5.  Runs in child processes: a hang ends in a trap
    that kills the CUDA context."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from chromosome3d_b200 import api, datasets as ds\n"
            "api.init(0)\n"
            "for k, v in (('sa_min_smem', 0), ('sa_split', 0), ('sa_cluster', 4), ('sa_warps', 4)): api.set_option(k, v)\n"
            "IF = ds.load_if_matrix(ds.input_path('chr9_1mb'))\n"
            "p = api.default_params(); p.hot_step = p.cool_step = 60; p.pow_step = 20; p.premin_step = 5\n"
            "r = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=120, seed=82364)], p)[0]\n"
            "print('ok', int(r.force_evals.min()))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    for _ in range(4):
        r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr[-400:]


@pytest.mark.parametrize("pin", [False, True])
def test_sa_against_all_shipped_cns_models(gpu_api, pin):
    """Weak pin against CNS itself, all 45 shipped models (each one CNS model selected by the authors) against 10-model
    ensembles at the script's defaults alpha 0.5, K 11 — one batched call.  Plain bead model (fp64 oracle, scripts/fit_cns_pin.py;
    the B200 ensembles agree, profiles/r01c_shipped_models_vs_ensembles.txt): median IF-Spearman within 0.02 of the shipped model
    for 35 chromosomes, within 0.05 for 44.  With the default-off CNS stand-ins (cns_pin_params(): soft CA(i)-CA(i+2) lower bound,
    repel radius 5.5 A): 42 of 45 within 0.02.  The two that stay at +0.05 (chr22_1mb, chr21_500kb; 35 and 73 beads) sit where
    the Spearman changes by 0.12 per 0.1 of alpha (profiles/r02_cns_pin.md): a 10 % difference in overall scale explains them."""
    import glob
    rg = lambda x: np.sqrt(((x - x.mean(0)) ** 2).sum(1).mean())
    jobs, ships = [], []
    p = gpu_api.cns_pin_params() if pin else gpu_api.default_params()
    for name in ds.list_inputs():
        hits = glob.glob(os.path.join(ds.shipped_models_dir(), "%s_rank*_a11.pdb" % name))
        if not hits:
            continue
        IF = ds.load_if_matrix(ds.input_path(name))
        ship = ds.read_ca_pdb(hits[0])
        if ship.shape[0] == IF.shape[0]:
            ships.append((IF, ship)); jobs.append(gpu_api.Job(IF=IF, alpha=0.5, n_models=10, seed=82364))
    assert len(jobs) == 45
    res = gpu_api.build_models(jobs, p=p)
    d_rho = np.array([np.median(gpu_api.spearman_IF_pdb(IF, r.xyz, 3)) - gpu_api.spearman_IF_pdb(IF, ship, 3)[0] for (IF, ship), r in zip(ships, res)])
    d_rg = np.array([np.mean([rg(x) for x in r.xyz]) / rg(ship) - 1 for (IF, ship), r in zip(ships, res)])
    print("pin=%s within 0.02: %d, within 0.05: %d, mean %+.4f, max %.3f" % (pin, (abs(d_rho) <= 0.02).sum(), (abs(d_rho) <= 0.05).sum(), d_rho.mean(), abs(d_rho).max()))
    assert abs(d_rho.mean()) < 0.03 and (abs(d_rho) <= 0.02).sum() >= (39 if pin else 30) and (abs(d_rho) <= 0.055).sum() >= 42 and abs(d_rho).max() < 0.08
    assert abs(d_rg.mean()) < 0.08 and abs(d_rg).max() < 0.25


def test_restraint_rows_are_validated(gpu_api):
    """Caller-supplied rows index device arrays: out-of-range beads, i == j, non-positive distances are refused everywhere; the
    dense-table paths (annealing, energy / forces) also refuse a pair listed twice."""
    from chromosome3d_b200.api import Chr3dError
    x = orc.random_coil(30, 3.8, 1, 0)
    ok = (np.array([0, 3], dtype=np.int32), np.array([5, 9], dtype=np.int32), np.array([4.0, 6.0]))
    gpu_api.count_satisfied_and_sum_dev(x[None], *ok)
    for bad in ((np.array([0, 3]), np.array([5, 30]), np.array([4.0, 6.0])), (np.array([-1]), np.array([5]), np.array([4.0])),
                (np.array([4]), np.array([4]), np.array([4.0])), (np.array([1]), np.array([5]), np.array([0.0]))):
        with pytest.raises(Chr3dError):
            gpu_api.count_satisfied_and_sum_dev(x[None], *bad)
        with pytest.raises(Chr3dError):
            gpu_api.build_models([gpu_api.Job(restraints=bad, N=30, n_models=1)])
    with pytest.raises(Chr3dError):                            # bonded neighbours: the dense table cannot keep repel off such a row
        gpu_api.build_models([gpu_api.Job(restraints=(np.array([3]), np.array([4]), np.array([3.8])), N=30, n_models=1)])
    dup = (np.array([0, 5, 2]), np.array([5, 0, 9]), np.array([4.0, 4.5, 6.0]))
    gpu_api.count_satisfied_and_sum_dev(x[None], *dup)                              # row-wise scoring counts every row, like assess_dgsa (C3D:769-880)
    with pytest.raises(Chr3dError):
        gpu_api.build_models([gpu_api.Job(restraints=dup, N=30, n_models=1)])
