"""CPU: the oracle against the golden vectors produced by the REAL reference Perl
(tests/golden/make_golden.py).  These tests pin the oracle; the GPU tests then compare the CUDA
path with the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import datasets as ds


def _md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def _golden(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


def _cases(golden_dir, small_only=True):
    g = _golden(golden_dir, "restraints_md5.json")
    keys = sorted(g)
    if small_only:
        keys = [k for k in keys if k.split("@")[0].endswith("_1mb")] + [k for k in keys if k.startswith("chr21_500kb")]
    return g, keys


def test_restraint_tables_match_real_perl(golden_dir, tmp_path):
    """.dist / .rr / contact.tbl byte-identical to chromosome3D.pl (C3D:110-206, 340-362)."""
    g, keys = _cases(golden_dir)
    assert len(keys) >= 40
    cache = {}
    for key in keys:
        name, alpha = key.split("@")
        if name not in cache:
            cache[name] = ds.load_if_matrix(ds.input_path(name))
        IF = cache[name]
        t = orc.if_to_dist(IF, float(alpha))
        ri, rj, rd = orc.dist_to_restraints(t)
        assert len(ri) == g[key]["rows"], key
        orc.write_dist(str(tmp_path / "a.dist"), t)
        orc.write_rr(str(tmp_path / "a.rr"), ri, rj, rd)
        orc.write_tbl(str(tmp_path / "a.tbl"), ri, rj, rd)
        assert _md5(tmp_path / "a.dist") == g[key]["dist"], key
        assert _md5(tmp_path / "a.rr") == g[key]["rr"], key
        assert _md5(tmp_path / "a.tbl") == g[key]["tbl"], key


def test_restraint_table_large_matrix(golden_dir, tmp_path):
    """chr1_500kb (L=455, 101,426 rows at alpha=1.1): the BASELINE config-2 table."""
    g = _golden(golden_dir, "restraints_md5.json")
    IF = ds.load_if_matrix(ds.input_path("chr1_500kb"))
    t = orc.if_to_dist(IF, 1.1)
    ri, rj, rd = orc.dist_to_restraints(t)
    assert len(ri) == 101426 == g["chr1_500kb@1.1"]["rows"]
    orc.write_tbl(str(tmp_path / "a.tbl"), ri, rj, rd)
    assert _md5(tmp_path / "a.tbl") == g["chr1_500kb@1.1"]["tbl"]


def test_contact_tbl_text_fixture(golden_dir, tmp_path):
    """The committed contact.tbl written by the real Perl (chr21_1mb, alpha 1.1), line by line."""
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    ri, rj, rd = orc.if_to_restraints(IF, 1.1)
    orc.write_tbl(str(tmp_path / "a.tbl"), ri, rj, rd)
    want = open(os.path.join(golden_dir, "chr21_1mb_a1.1_contact.tbl")).read()
    assert open(tmp_path / "a.tbl").read() == want
    assert want.splitlines()[0] == "assign45 (resid   1 and name ca) (resid  10 and name ca) 177.90 0.00 0.00"


def test_assess_matches_real_perl(golden_dir):
    """count_satisfied_tbl_rows / sum_noe_dev / ranking against the real assess_dgsa (C3D:769-829)."""
    g = _golden(golden_dir, "assess.json")
    for name, rec in g.items():
        IF = ds.load_if_matrix(ds.input_path(name))
        ri, rj, rd = orc.if_to_restraints(IF, float(rec["alpha"]))
        xyz = np.array(rec["xyz"])
        sat, dev = orc.assess(xyz, ri, rj, rd)
        ID = name + "_matrix"
        for k in range(rec["m"]):
            row = rec["table"]["%s_%d" % (ID, k + 1)]
            assert row["total"] == len(ri)
            assert sat[k] == row["count"], (name, k)
            assert "%.2f" % dev[k] == row["sum_dev"], (name, k)
        order = orc.rank_models(np.array(rec["energies"]))
        want = [int(nm.rsplit("_", 1)[1]) - 1 for _, nm in rec["ranks"]]
        assert list(order[:len(want)]) == want


def test_spearman_matches_real_script(golden_dir):
    """spearman_IF_pdb.pl (SP:32-71) run for real on shipped models, '%.3f' output."""
    g = _golden(golden_dir, "spearman.json")
    assert len(g) >= 10
    for key, val in g.items():
        name, pdb, rng = key.split("|")
        IF = ds.load_if_matrix(ds.input_path(name))
        xyz = ds.read_ca_pdb(os.path.join(ds.shipped_models_dir(), pdb))
        assert xyz.shape[0] == IF.shape[0]
        rho = orc.spearman(IF, xyz, int(rng))[0]
        assert "%.3f" % rho == val, key


def test_shipped_model_known_answer():
    """the shipped chr21_1mb model scores 68/528 and 2955.67 against the alpha=0.5 table."""
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    xyz = ds.read_ca_pdb(os.path.join(ds.shipped_models_dir(), "chr21_1mb_rank07_a11.pdb"))
    sat, dev = orc.assess(xyz, ri, rj, rd)
    assert sat[0] == 68 and "%.2f" % dev[0] == "2955.67"


def test_forces_are_gradient_of_energy():
    """Oracle self-check: analytic forces vs central finite differences (<=1e-6 relative)."""
    from chromosome3d_b200 import abi
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    p = orc.default_params()
    rs = np.random.RandomState(0)
    for w, low_soft in ((abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85), 0),
                        (abi.stage(abi.OP_ENERGY, w_noe=2, w_bond=100, w_vdw=20.0, repel=1.5), 1)):
        p.noe_low_soft = low_soft
        x = orc.random_coil(N, 3.8, 3, 0) * 0.6
        e, f = orc.energy_forces(x, ri, rj, rd, p, w)
        h = 1e-5
        for _ in range(12):
            i, c = rs.randint(N), rs.randint(3)
            xp, xm = x.copy(), x.copy()
            xp[i, c] += h
            xm[i, c] -= h
            ep = orc.energy_forces(xp, ri, rj, rd, p, w, want_forces=False)[0][0, 0]
            em = orc.energy_forces(xm, ri, rj, rd, p, w, want_forces=False)[0][0, 0]
            fd = -(ep - em) / (2 * h)
            assert abs(fd - f[0, i, c]) <= 1e-6 * max(1.0, np.abs(f).max()), (i, c, fd, f[0, i, c])


def test_program_follows_dgsa_inp():
    """Stage program = the protocol of dgsa.inp (C3D:1573-1829): counts, temperatures, weights."""
    from chromosome3d_b200 import abi
    p = orc.default_params()
    prog = orc.build_program(p)
    ops = [s.op for s in prog]
    dyn = [s for s in prog if s.op == abi.OP_DYNAMICS]
    hot = [s for s in dyn if s.flag == abi.COUPLE_TCOUP]
    cool = [s for s in dyn if s.flag == abi.COUPLE_VSCALE]
    assert [s.nstep for s in hot] == [125, 125, 125, 500, 125] * 2          # C3D:1644-1700, two images
    assert len(cool) == 81 and all(s.nstep == 12 for s in cool)              # C3D:1740-1782
    assert [s.temp for s in cool][:3] == [2000.0, 1975.0, 1950.0] and cool[-1].temp == 0.0
    assert cool[0].repel == 0.80 and abs(cool[1].repel - (1.0 - 0.15 / 80)) < 1e-12 and abs(cool[-1].repel - 0.85) < 1e-9
    assert cool[0].w_vdw == 0.003 and abs(cool[-1].w_vdw - 4.0) < 1e-9
    assert ops.count(abi.OP_MINIMIZE) == 4 + 10 and prog[-3].nstep == 15000 and prog[-3].drop == 10.0
    assert [s.w_noe for s in hot[:5]] == [1.0, 2.0, 2.0, 4.0, 10.0]           # weights * 0.1 .. 1.0 times noe scale 10
    assert ops[-2:] == [abi.OP_CENTER, abi.OP_ENERGY] and abi.OP_PICK in ops and abi.OP_SAVE in ops


def test_oracle_sa_reproduces_shipped_statistics():
    """Ensemble-level anchor: at alpha=0.5 the bead-level SA lands on the
    statistics of the shipped CNS model of chr21_1mb (rho -0.845, Rg 18.1, bonds 3.8-4.8 A)."""
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    p = orc.default_params()
    res = orc.run_ensemble(N, ri, rj, rd, p, orc.build_program(p), 82364, 4)
    rho = orc.spearman(IF, res["xyz"])
    assert abs(np.median(rho) - (-0.845)) < 0.02
    x = res["xyz"]
    bl = np.linalg.norm(x[:, 1:] - x[:, :-1], axis=2)
    assert 3.6 < bl.min() and bl.max() < 5.2
    rg = np.sqrt(((x - x.mean(1, keepdims=True)) ** 2).sum(2).mean(1))
    assert abs(rg.mean() - 18.1) < 1.5


def test_spearman_against_an_independent_implementation():
    """The rank-correlation arithmetic of the un-vendored Statistics::RankCorrelation 0.1204 (SP:10-13, 65-69: average ranks for
    ties, Pearson on the ranks) is restated in the oracle; scipy.stats.spearmanr is an independent implementation of the same
    published definition.  Same pair selection as the script (all ordered pairs with |i-j| >= range, SP:42-60), distances as the
    '%.3f' text the script compares: the two must agree to rounding noise, ties (zero IF cells, equal distances) included."""
    stats = pytest.importorskip("scipy.stats")
    for name, pdb in (("chr21_1mb", "chr21_1mb_rank07_a11.pdb"), ("chr22_500kb", None), ("chr1_1mb", None)):
        IF = ds.load_if_matrix(ds.input_path(name))
        L = IF.shape[0]
        if pdb:
            xyz = ds.read_ca_pdb(os.path.join(ds.shipped_models_dir(), pdb))
        else:
            xyz = np.round(orc.random_coil(L, 3.8, 11, 0), 3)          # 3-decimal coordinates, as a PDB file holds them
        for rng in (3, 1):
            a, b = [], []
            for i in range(L):
                for j in range(L):
                    if abs(i - j) < rng:
                        continue
                    a.append(IF[i, j])
                    b.append(float("%.3f" % np.sqrt(((xyz[i] - xyz[j]) ** 2).sum())))
            want = stats.spearmanr(a, b).statistic
            got = orc.spearman(IF, xyz, rng)[0]
            assert abs(got - want) < 1e-12, (name, rng, got, want)


def test_energy_terms_against_a_vectorised_restatement():
    """A second restatement of the three terms, written from the formulas of SURVEY App. A.2-A.4 as whole-array numpy expressions
    (the C oracle loops pair by pair): soft-square NOE with the general (a, b, c) tail, harmonic bead bond, quartic repel over
    all pairs at least vdw_min_sep apart.  Energies must agree to 1e-12 relative, forces (analytic, vectorised) to 1e-10."""
    from chromosome3d_b200 import abi
    for name, alpha in (("chr21_1mb", 0.5), ("chr22_500kb", 1.1)):
        IF = ds.load_if_matrix(ds.input_path(name))
        N = IF.shape[0]
        ri, rj, rd = orc.if_to_restraints(IF, alpha)
        for (w_noe, w_bond, w_vdw, repel, asym, low_soft, scale) in ((10.0, 1000.0, 4.0, 0.85, 2.0, 0, 1.0), (2.0, 100.0, 20.0, 1.5, 0.5, 1, 0.5),
                                                                     (1.0, 100.0, 0.003, 0.9, 3.0, 0, 2.0)):
            p = orc.default_params()
            p.noe_asymptote, p.noe_low_soft = asym, low_soft
            st = abi.stage(abi.OP_ENERGY, w_noe=w_noe, w_bond=w_bond, w_vdw=w_vdw, repel=repel)
            x = orc.random_coil(N, 3.8, 5, 0) * scale
            e, f = orc.energy_forces(x, ri, rj, rd, p, st)
            C, rsw, c = p.noe_sqconst, p.noe_rswitch, p.noe_asymptote
            b = (c - 2.0 * C * rsw) * rsw * rsw                       # App. A.2 with e = 2, s = 1
            a = C * rsw * rsw - b / rsw - c * rsw
            F = np.zeros_like(x)
            # NOE
            d = x[ri] - x[rj]
            R = np.sqrt((d * d).sum(1))
            delta = R - rd
            u = np.abs(delta)
            tail = (u > rsw) & ((delta > 0) | bool(low_soft))
            uu = np.where(tail, u, 1.0)
            e_noe = w_noe * np.where(tail, a + b / uu + c * u, C * u * u).sum()
            dEdu = np.where(tail, c - b / (uu * uu), 2.0 * C * u)
            g = (w_noe * dEdu * np.sign(delta) / R)[:, None] * d
            np.add.at(F, ri, -g)
            np.add.at(F, rj, g)
            # bond
            d = x[:-1] - x[1:]
            R = np.sqrt((d * d).sum(1))
            e_bond = w_bond * ((R - p.bond_r0) ** 2).sum()
            g = (2.0 * w_bond * (R - p.bond_r0) / R)[:, None] * d
            F[:-1] -= g
            F[1:] += g
            # repel
            i, j = np.triu_indices(N, k=p.vdw_min_sep)
            d = x[i] - x[j]
            r2 = (d * d).sum(1)
            ov = np.maximum(0.0, (repel * p.vdw_rmin) ** 2 - r2)
            e_vdw = w_vdw * (ov * ov).sum()
            g = (-4.0 * w_vdw * ov)[:, None] * d
            np.add.at(F, i, -g)
            np.add.at(F, j, g)
            want = np.array([e_noe + e_bond + e_vdw, e_bond, e_vdw, e_noe])
            assert np.allclose(e[0], want, rtol=1e-12, atol=1e-9), (name, e[0], want)
            assert np.abs(f[0] - F).max() <= 1e-10 * max(1.0, np.abs(F).max()), np.abs(f[0] - F).max()
