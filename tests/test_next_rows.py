"""SURVEY 8(f) "next" rows: cross-resolution tools (reduced models, similarity.txt) and matrix loaders.

The generating programs are not in the reference tree, but their OUTPUTS are (output_models/*_reduced.pdb,
similarity.txt; shipped here as data/shipped_models.tar.xz), so both the oracle and the product are pinned on them:
23 reduced files byte for byte, 19 (Spearman, RMSD) blocks to 1e-12."""
import glob
import os
import re

import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import api, datasets as ds


def _pairs():
    d = ds.shipped_models_dir()
    return [(r.replace("_reduced.pdb", ".pdb"), r) for r in sorted(glob.glob(os.path.join(d, "*_reduced.pdb")))]


def _similarity_blocks():
    d = ds.shipped_models_dir()
    txt = open(os.path.join(d, "similarity.txt")).read().replace("\r", "")
    out = []
    for name, sp, rm in re.findall(r"(\S+)\nSpearman correlation: (\S+)\nRMSD: (\S+)", txt):
        one_mb = glob.glob(os.path.join(d, "%s_1mb_rank*_a11.pdb" % name.split("_")[0]))[0]
        out.append((one_mb, os.path.join(d, name + ".pdb"), float(sp), float(rm)))
    return out


def test_oracle_reduced_models_byte_exact(tmp_path):
    pairs = _pairs()
    assert len(pairs) == 23
    for src, red in pairs:
        x = orc.reduce_model(ds.read_ca_pdb(src))
        orc.write_reduced_pdb(tmp_path / "o.pdb", x)
        assert open(tmp_path / "o.pdb", "rb").read() == open(red, "rb").read(), red


def test_oracle_similarity_matches_shipped_numbers():
    blocks = _similarity_blocks()
    assert len(blocks) == 19
    for one_mb, half_mb, sp, rm in blocks:
        rho, rmsd = orc.model_similarity(ds.read_ca_pdb(one_mb), orc.reduce_model(ds.read_ca_pdb(half_mb)))
        assert abs(rho - sp) < 1e-12 and abs(rmsd - rm) < 1e-11, half_mb


def test_reduced_pdb_writer_byte_exact(tmp_path):
    """Host-side writer of the product (own implementation of the HALF_UP shortest-digits formatting)."""
    for src, red in _pairs():
        api.write_reduced_pdb(tmp_path / "p.pdb", orc.reduce_model(ds.read_ca_pdb(src)))
        assert open(tmp_path / "p.pdb", "rb").read() == open(red, "rb").read(), red
    # odd bead count: the trailing bead is copied; values that are not ties; negative zero keeps its sign like the tool's
    x = np.array([[1.0005, -2.0005, 0.0], [3.0, 4.0, -0.0004], [5.5555, 6.25, 7.125]])
    api.write_reduced_pdb(tmp_path / "q.pdb", x)
    orc.write_reduced_pdb(tmp_path / "r.pdb", x)
    assert open(tmp_path / "q.pdb", "rb").read() == open(tmp_path / "r.pdb", "rb").read()


@pytest.mark.parametrize("name", ["chr21_1mb", "chr22_500kb", "chr1_500kb"])
def test_text_loader_matches_reference_parse(name):
    """chr3d_read_if_text == the split-on-whitespace parse of IF2dist_new (C3D:116-129), bit for bit."""
    a = api.read_if_text(ds.input_path(name))
    b = ds.load_if_matrix(ds.input_path(name))
    assert a.shape == b.shape and np.array_equal(a, b)


def test_text_loader_ragged_and_errors(tmp_path):
    p = tmp_path / "m.txt"
    p.write_bytes(b"\r\n1 2.5 3e-1 \r\n4 5\r\n7 8 9 10\r\n11 12 13\r\n")       # blank first line, short row, long row, extra row
    a = api.read_if_text(p)
    assert a.shape == (3, 3) and np.array_equal(a, np.array([[1, 2.5, 0.3], [4, 5, 0], [7, 8, 9]]))
    (tmp_path / "e.txt").write_bytes(b"\n\n")
    with pytest.raises(api.Chr3dError):
        api.read_if_text(tmp_path / "e.txt")
    (tmp_path / "s.txt").write_bytes(b"1 2 3\n4 5 6\n")
    with pytest.raises(api.Chr3dError):
        api.read_if_text(tmp_path / "s.txt")
    with pytest.raises(api.Chr3dError):
        api.read_if_text(tmp_path / "missing.txt")


def test_binary_container_round_trip(tmp_path):
    IF = ds.load_if_matrix(ds.input_path("chr21_500kb"))
    api.write_if_binary(tmp_path / "a.bin", IF)
    assert np.array_equal(api.read_if_binary(tmp_path / "a.bin"), IF)
    api.write_if_binary(tmp_path / "b.bin", IF, 4)
    assert np.array_equal(api.read_if_binary(tmp_path / "b.bin"), IF.astype(np.float32).astype(np.float64))
    (tmp_path / "c.bin").write_bytes(b"not a matrix at all")
    with pytest.raises(api.Chr3dError):
        api.read_if_binary(tmp_path / "c.bin")


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_reduce_model_bit_exact(gpu_api):
    for src, _ in _pairs()[:6]:
        x = ds.read_ca_pdb(src)
        assert np.array_equal(gpu_api.reduce_model(x), orc.reduce_model(x))
    x = np.random.RandomState(5).randn(7, 3) * 30
    assert np.array_equal(gpu_api.reduce_model(x), orc.reduce_model(x))


@pytest.mark.gpu
def test_gpu_similarity_matches_shipped_numbers_and_oracle(gpu_api):
    for one_mb, half_mb, sp, rm in _similarity_blocks():
        a, b = ds.read_ca_pdb(one_mb), gpu_api.reduce_model(ds.read_ca_pdb(half_mb))
        rho, rmsd = gpu_api.model_similarity(a, b)
        orho, ormsd = orc.model_similarity(a, b)
        assert abs(rho - sp) < 1e-12 and abs(rmsd - rm) < 1e-11, half_mb
        assert abs(rho - orho) < 1e-13 and abs(rmsd - ormsd) < 1e-11
    # ties: integer coordinates on a lattice produce many equal distances (average ranks)
    rs = np.random.RandomState(2)
    a, b = rs.randint(0, 4, size=(40, 3)).astype(float), rs.randint(0, 4, size=(40, 3)).astype(float)
    rho, rmsd = gpu_api.model_similarity(a, b)
    orho, ormsd = orc.model_similarity(a, b)
    assert abs(rho - orho) < 1e-13 and abs(rmsd - ormsd) < 1e-12


@pytest.mark.gpu
def test_gpu_coo_scatter(gpu_api):
    rs = np.random.RandomState(3)
    L = 300
    dense = np.triu(rs.rand(L, L) * (rs.rand(L, L) < 0.2))
    bi, bj = np.nonzero(dense)
    v = dense[bi, bj]
    want = dense + np.triu(dense, 1).T
    assert np.array_equal(gpu_api.if_from_coo(bi, bj, v, L), want)
    assert np.array_equal(gpu_api.if_from_coo(bi + 1, bj + 1, v, L, base=1), want)
    with pytest.raises(gpu_api.Chr3dError):
        gpu_api.if_from_coo([0], [L], [1.0], L)
    assert np.array_equal(gpu_api.if_from_coo([], [], [], 4), np.zeros((4, 4)))


# ------------------------------------------------------------------------------------------ DG embedding (8(f) rank 1)
def _exact_case():
    X = ds.read_ca_pdb(os.path.join(ds.shipped_models_dir(), "chr21_1mb_rank07_a11.pdb"))
    iu = np.triu_indices(len(X), 1)
    return X, iu[0].astype(np.int32), iu[1].astype(np.int32), np.linalg.norm(X[iu[0]] - X[iu[1]], axis=1)


def test_oracle_dg_recovers_a_consistent_structure():
    """Metric data (all pair distances of a real model): no inconsistent pair, rank-3 metric matrix, every distance back to 1e-10;
    the eigenvalues equal those of the double-centred squared-distance matrix."""
    X, ri, rj, rd = _exact_case()
    xyz, eig, bad = orc.dg_embed(len(X), ri, rj, rd, orc.default_params(), 7, 2)
    assert bad == 0
    D0 = np.linalg.norm(X[:, None] - X[None], axis=2)
    for m in range(2):
        D = np.linalg.norm(xyz[m][:, None] - xyz[m][None], axis=2)
        assert np.abs(D - D0).max() < 1e-10
    n = len(X)
    J = np.eye(n) - 1.0 / n
    w = np.linalg.eigvalsh(-0.5 * J @ (D0 ** 2) @ J)[::-1][:3]
    assert np.allclose(eig[0], w, rtol=1e-10)


def test_oracle_dg_on_hic_bounds():
    """Hi-C targets are not metric: smoothing shortens upper bounds, the lower pass meets them; the embedded start already
    correlates with the IF matrix far better than a random coil does."""
    IF = ds.load_if_matrix(ds.input_path("chr21_500kb"))
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 1.1)
    p = orc.default_params()
    U, L, bad = orc.dg_bounds(N, ri, rj, rd, p)
    assert bad > 0 and np.all(L <= U + 1e-12) and np.allclose(U, U.T) and np.all(np.diag(U) == 0)
    assert np.all(U[ri, rj] <= rd + 1e-12)                       # shortest paths never exceed the direct target
    k = np.arange(N - 1)
    assert np.all(U[k, k + 1] <= p.bond_r0 + 1e-12)
    xyz, eig, _ = orc.dg_embed(N, ri, rj, rd, p, 82364, 1)
    assert eig[0, 0] >= eig[0, 1] >= eig[0, 2] > 0
    assert orc.spearman(IF, xyz[0], 3)[0] < orc.spearman(IF, orc.random_coil(N, 3.8, 82364, 0), 3)[0] - 0.1


@pytest.mark.gpu
def test_gpu_dg_embed_bit_identical_to_oracle(gpu_api):
    X, ri, rj, rd = _exact_case()
    xyz, eig, bad = gpu_api.dg_embed(ri, rj, rd, len(X), n_models=2, seed=7)
    oxyz, oeig, obad = orc.dg_embed(len(X), ri, rj, rd, orc.default_params(), 7, 2)
    assert bad == obad == 0 and np.array_equal(eig, oeig) and np.array_equal(xyz, oxyz)
    for name, alpha in (("chr21_500kb", 1.1), ("chr22_1mb", 0.5), ("chr1_1mb", 1.1)):
        IF = ds.load_if_matrix(ds.input_path(name))
        ri, rj, rd = orc.if_to_restraints(IF, alpha)
        xyz, eig, bad = gpu_api.dg_embed(ri, rj, rd, IF.shape[0], n_models=3, seed=82364)
        oxyz, oeig, obad = orc.dg_embed(IF.shape[0], ri, rj, rd, orc.default_params(), 82364, 3)
        assert bad == obad and np.array_equal(eig, oeig) and np.array_equal(xyz, oxyz), name


def test_oracle_dg_refine_lowers_the_bounds_energy():
    """C3D:1516-1521 restated: x *= dgscale, then the 200-step L-BFGS on the bounds-violation energy alone.  A consistent set of
    bounds is minimised to (numerically) zero; on Hi-C bounds the energy drops and the restrained distances move closer."""
    X, ri, rj, rd = _exact_case()
    p = orc.default_params()
    xyz, eig, st, bad = orc.dg_starts(len(X), ri, rj, rd, p, 7, 1)
    assert bad == 0 and abs(st[0]["dgscale"] - 1.0) < 1e-6 and st[0]["e_final"] <= st[0]["e_scaled"] and st[0]["e_final"] < 1e-12
    IF = ds.load_if_matrix(ds.input_path("chr22_1mb"))
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    x0, _, _ = orc.dg_embed(N, ri, rj, rd, p, 82364, 1)
    xyz, eig, st, bad = orc.dg_starts(N, ri, rj, rd, p, 82364, 1)
    assert 1.0 < st[0]["dgscale"] < 1.5                          # the projection onto three dimensions shrinks the structure
    assert st[0]["e_final"] < 0.8 * st[0]["e_scaled"] and 1 <= st[0]["iterations"] <= 200 and st[0]["evaluations"] > st[0]["iterations"]
    dev = lambda x: np.sqrt(np.mean((np.linalg.norm(x[ri] - x[rj], axis=1) ** 2 - rd ** 2) ** 2))
    assert dev(xyz[0]) < dev(x0[0])
    z, _, st0, _ = orc.dg_starts(N, ri, rj, rd, p, 82364, 1, nstep=0)
    assert np.array_equal(z[0], x0[0] * st0[0]["dgscale"]) and st0[0]["iterations"] == 0 and st0[0]["e_scaled"] == st[0]["e_scaled"]


@pytest.mark.gpu
def test_gpu_dg_starts_bit_identical_to_oracle(gpu_api):
    """chr3d_dg_starts (embedding + dgscale + 200 L-BFGS iterations on the bounds energy, one CTA per model, fp64, fixed
    summation order) against orc_dg_refine: coordinates and statistics bit for bit."""
    p = orc.default_params()
    X, ri, rj, rd = _exact_case()
    xyz, eig, st, bad = gpu_api.dg_starts(ri, rj, rd, len(X), n_models=2, seed=7)
    oxyz, oeig, ost, obad = orc.dg_starts(len(X), ri, rj, rd, p, 7, 2)
    assert bad == obad == 0 and np.array_equal(xyz, oxyz) and st == ost
    for name, alpha, nstep in (("chr21_500kb", 1.1, 200), ("chr22_1mb", 0.5, 200), ("chr1_1mb", 1.1, 25)):
        IF = ds.load_if_matrix(ds.input_path(name))
        ri, rj, rd = orc.if_to_restraints(IF, alpha)
        xyz, eig, st, bad = gpu_api.dg_starts(ri, rj, rd, IF.shape[0], n_models=2, seed=82364, nstep=nstep)
        oxyz, oeig, ost, obad = orc.dg_starts(IF.shape[0], ri, rj, rd, p, 82364, 2, nstep=nstep)
        assert bad == obad and np.array_equal(eig, oeig) and st == ost and np.array_equal(xyz, oxyz), (name, st, ost)
        assert st[0]["e_final"] < st[0]["e_scaled"]
    # a sparse table leaves slack between the bounds: the models differ, each one is refined on its own
    N = 60
    Xs = orc.random_coil(N, 3.0, 5, 0)
    k = np.array([(i, j) for i in range(N) for j in range(i + 2, N) if (i * 7 + j * 3) % 5 == 0], dtype=np.int32)
    dd = np.linalg.norm(Xs[k[:, 0]] - Xs[k[:, 1]], axis=1)
    xyz, eig, st, bad = gpu_api.dg_starts(k[:, 0], k[:, 1], dd, N, n_models=3, seed=11)
    oxyz, oeig, ost, obad = orc.dg_starts(N, k[:, 0], k[:, 1], dd, p, 11, 3)
    assert np.array_equal(xyz, oxyz) and st == ost and not np.array_equal(xyz[0], xyz[1])


@pytest.mark.gpu
def test_gpu_annealing_from_dg_start(gpu_api):
    """The embedded coordinates feed the annealer through Job.x0; the ensemble ends at least as well (IF-Spearman) as from coils."""
    IF = ds.load_if_matrix(ds.input_path("chr21_500kb"))
    N = IF.shape[0]
    ri, rj, rd = gpu_api.if_to_restraints(IF, 1.1)
    x0, _, _ = gpu_api.dg_embed(ri, rj, rd, N, n_models=6, seed=82364)
    a = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=1.1, n_models=6, seed=82364, x0=x0)])[0]
    b = gpu_api.build_models([gpu_api.Job(IF=IF, alpha=1.1, n_models=6, seed=82364)])[0]
    ra, rb = np.median(gpu_api.spearman_IF_pdb(IF, a.xyz, 3)), np.median(gpu_api.spearman_IF_pdb(IF, b.xyz, 3))
    assert np.all(np.isfinite(a.energies)) and ra < rb + 0.02


@pytest.mark.gpu
def test_gpu_alpha_sweep_python_mirror(gpu_api):
    IF = ds.load_if_matrix(ds.input_path("chr22_1mb"))
    rows, best = gpu_api.alpha_sweep(IF, [0.4, 0.5, 0.6], n_models=3, seed=5)
    assert [r["alpha"] for r in rows] == [0.4, 0.5, 0.6]
    for r in rows:
        assert r["restraints"] == len(orc.if_to_restraints(IF, r["alpha"])[0]) and 1 <= r["picked_rank"] <= 3
    assert best["srcc"] == min(r["best_srcc"] for r in rows) and best["srcc"] < -0.7
    assert re.fullmatch(r"rank0[123]_a0[456]", best["stem"])
    assert "%.3f" % orc.spearman(IF, best["xyz"], 3)[0] == "%.3f" % best["srcc"]


# ------------------------------------------------------------------------------------------ small / degenerate inputs (CPU)
def test_oracle_dg_tiny_chain_without_restraints():
    """Three beads, bonds only: the bounds are (r0, r0, [repel floor, 2 r0]); three points always embed exactly."""
    p = orc.default_params()
    empty_i, empty_d = np.zeros(0, dtype=np.int32), np.zeros(0)
    U, L, bad = orc.dg_bounds(3, empty_i, empty_i, empty_d, p)
    assert bad == 0 and U[0, 1] == U[1, 2] == p.bond_r0 and U[0, 2] == 2 * p.bond_r0
    assert L[0, 2] == pytest.approx(p.vdw_rmin * p.cool_rad_finl) and L[0, 1] == p.bond_r0
    xyz, eig, _ = orc.dg_embed(3, empty_i, empty_i, empty_d, p, 3, 4)
    for m in range(4):
        d01, d12, d02 = (np.linalg.norm(xyz[m][a] - xyz[m][b]) for a, b in ((0, 1), (1, 2), (0, 2)))
        assert d01 == pytest.approx(p.bond_r0, abs=1e-9) and d12 == pytest.approx(p.bond_r0, abs=1e-9)
        assert L[0, 2] - 1e-9 <= d02 <= U[0, 2] + 1e-9
    assert len({round(float(np.linalg.norm(x[0] - x[2])), 6) for x in xyz}) > 1        # the free distance differs between models


def test_reduced_pdb_single_bead_and_pairs(tmp_path):
    one = np.array([[1.2345, -0.0005, 10.0]])
    api.write_reduced_pdb(tmp_path / "a.pdb", one)
    assert open(tmp_path / "a.pdb", "rb").read() == b"\r\nATOM      1   CA MET B1          1.235  -0.001  10.000  0.20 10.00\r\nEND\r\n"
    two = orc.reduce_model(np.array([[0.001, 0.0, 0.0], [0.002, 0.0, 0.0], [5.0, 5.0, 5.0]]))
    assert two.shape == (2, 3) and two[0, 0] == 0.0015 and np.array_equal(two[1], [5.0, 5.0, 5.0])
    orc.write_reduced_pdb(tmp_path / "b.pdb", two)
    api.write_reduced_pdb(tmp_path / "c.pdb", two)
    txt = open(tmp_path / "b.pdb", "rb").read()
    assert txt == open(tmp_path / "c.pdb", "rb").read() and b"   0.002   0.000   0.000" in txt and txt.count(b"CONECT") == 1


def test_text_loader_whitespace_variants(tmp_path):
    p = tmp_path / "t.txt"
    p.write_bytes(b"1\t2  3 \r\n  4 5\t6\r\n7 8 9")            # tabs, leading blanks, no newline at the end
    assert np.array_equal(api.read_if_text(p), np.array([[1.0, 2, 3], [4, 5, 6], [7, 8, 9]]))
    p.write_bytes(b"1 x\n2 3\n")
    with pytest.raises(api.Chr3dError):
        api.read_if_text(p)
