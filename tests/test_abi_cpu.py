"""CPU: the C-ABI library loads and exports every symbol include/chr3d.h declares; host-side logic
that needs no device (defaults, stage program, ranking, writers) matches the oracle; compute entry
points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "chr3d.h")).read()
    names = set(re.findall(r"\b(chr3d_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 20
    lib = api.lib()
    for n in sorted(names):
        assert hasattr(lib, n), "libchr3d.so does not export %s" % n


def test_struct_layouts_match_header():
    """sizeof of the ctypes mirrors == what the C compiler laid out (checked through the oracle build,
    which includes the same header)."""
    p, po = api.default_params(), orc.default_params()
    assert bytes(p) == bytes(po)
    assert C.sizeof(abi.Stage) == 72 and C.sizeof(abi.SaParams) == 240
    sizes = [C.c_int32(0) for _ in range(4)]
    api.lib().chr3d_abi_sizes(*[C.byref(v) for v in sizes])        # what the library itself was compiled with
    assert [v.value for v in sizes] == [C.sizeof(abi.SaParams), C.sizeof(abi.Stage), C.sizeof(abi.Job), C.sizeof(abi.Result)]


def test_stage_program_equals_oracle_program():
    for images, pow_cycl in ((2, 10), (1, 3)):
        p = api.default_params()
        p.images, p.pow_cycl = images, pow_cycl
        a, b = api.build_program(p), orc.build_program(p)
        assert len(a) == len(b)
        assert all(bytes(x) == bytes(y) for x, y in zip(a, b))
    assert len(api.build_program(api.default_params())) == 113


def test_rank_models_int_truncation_and_ties():
    e = np.zeros((5, 4))
    e[:, 3] = [10.9, 10.2, 3.7, -2.5, 3.1]       # int(): 10 10 3 -2 3  (C3D:617)
    assert list(api.rank_models(e)) == [3, 2, 4, 0, 1] == list(orc.rank_models(e))


def test_text_writers_match_oracle(tmp_path):
    rs = np.random.RandomState(1)
    L = 23
    t = rs.randint(-1, 5000, size=(L, L)).astype(np.int64)
    t[t < 0] = -10
    ri, rj, rd = orc.dist_to_restraints(t)
    for nm, fa, fo, args in (("d", api.write_dist, orc.write_dist, (t,)), ("r", api.write_rr, orc.write_rr, (ri, rj, rd)),
                             ("t", api.carr2tbl, orc.write_tbl, (ri, rj, rd))):
        fa(str(tmp_path / (nm + "1")), *args)
        fo(str(tmp_path / (nm + "2")), *args)
        assert open(tmp_path / (nm + "1"), "rb").read() == open(tmp_path / (nm + "2"), "rb").read()


def test_pdb_writer_layout(tmp_path, golden_dir):
    """Final model layout = what reindex_chain/add_connect_rows leave (C3D:208-215, 831-862), checked
    byte for byte against the _model1.pdb the real Perl wrote (tests/golden/assess.json)."""
    import json
    g = json.load(open(os.path.join(golden_dir, "assess.json")))["chr21_1mb"]
    best = int(g["ranks"][0][1].rsplit("_", 1)[1]) - 1
    xyz = np.array(g["xyz"])[best]
    api.write_pdb(str(tmp_path / "m.pdb"), xyz, None, None, True)
    assert open(tmp_path / "m.pdb").read() == g["model1_pdb"]
    api.write_pdb(str(tmp_path / "t.pdb"), xyz, np.array([1.5, 2.5, 3.5, 42.7]))
    txt = open(tmp_path / "t.pdb").read().splitlines()
    assert any(re.match(r"^REMARK noe\s*=\s*42\.7", l) for l in txt) and txt[-1] == "END"


@pytest.mark.skipif(api.device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_device():
    IF = np.ones((8, 8))
    with pytest.raises(api.Chr3dError) as ei:
        api.IF2dist_new(IF, 0.5)
    assert ei.value.code == abi.ERR_NO_DEVICE
    with pytest.raises(api.Chr3dError):
        api.spearman_IF_pdb(IF, np.zeros((8, 3)))
    with pytest.raises(api.Chr3dError):
        api.build_models([api.Job(IF=IF, n_models=1)])


def test_set_option_accepts_only_the_launch_shape_options():
    """chr3d_set_option is host-only bookkeeping: the four launch-shape options of the ensemble kernel, integers, clearable."""
    for k in ("sa_cluster", "sa_warps", "sa_min_smem", "sa_split"):
        api.set_option(k, 2)
        api.set_option(k, None)
    with pytest.raises(api.Chr3dError):
        api.set_option("sa_cluster", "two")
    with pytest.raises(api.Chr3dError):
        api.set_option("CHR3D_LN_DEBUG", 1)          # development switches are not options (compiled in only with -DCHR3D_DEV)
    api.set_option("seqsum", "serial")               # the left-to-right sums as a serial chain (cross-checks); nothing else is a value
    api.set_option("seqsum", None)
    with pytest.raises(api.Chr3dError):
        api.set_option("seqsum", "tree")


def test_seq_sum_needs_a_device():
    """chr3d_seq_sum has no CPU path either: bad arguments are refused, and without a GPU the call fails loudly."""
    import ctypes as C
    out = (C.c_double * 1)()
    assert api.lib().chr3d_seq_sum(None, C.c_int64(5), C.c_int32(1), out) == abi.ERR_ARG
    if api.device_count() <= 0:
        with pytest.raises(api.Chr3dError) as ei:
            api.seq_sum([1.0, 2.0])
        assert ei.value.code == abi.ERR_NO_DEVICE
