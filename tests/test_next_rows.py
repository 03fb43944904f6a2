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
