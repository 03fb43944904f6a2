"""The Perl front end (perl/chromosome3D.pl, perl/spearman_IF_pdb.pl) over the C ABI.

CPU part: the binding loads (XS shim when FFI::Platypus is absent), host-only entry points work, the
CLI keeps the reference's options, and without a GPU the run fails loudly (no CPU fallback).
GPU part: a whole run on chr21_1mb reproduces the reference's output layout; the restraint files are
byte-identical to what the real Perl wrote (md5 fixtures); the printed assessment equals the oracle's."""
import hashlib
import json
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import oracle as orc
from chromosome3d_b200 import api, datasets as ds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PERL = os.path.join(ROOT, "perl")
pytestmark = pytest.mark.skipif(shutil.which("perl") is None, reason="perl not installed")


def _perl(args, **kw):
    return subprocess.run(["perl", "-I", PERL] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, **kw)


def test_binding_loads_and_host_entry_points_work():
    code = ('use Chr3D; my $p = Chr3D::default_params(); my ($g, $n) = Chr3D::build_program($p);'
            'my @e = (0,0,0,10.9, 0,0,0,10.2, 0,0,0,3.7, 0,0,0,-2.5, 0,0,0,3.1);'
            'print join(" ", Chr3D::backend(), $n, length($p), @{Chr3D::rank_models(pack("d*", @e), 5)}), "\\n";')
    r = _perl(["-e", code])
    assert r.returncode == 0, r.stderr
    backend, n, plen, *order = r.stdout.split()
    assert backend in ("XS", "FFI::Platypus") and int(n) == 113 and int(plen) == 240
    assert [int(o) for o in order] == [3, 2, 4, 0, 1]            # int() truncation + index tie-break (C3D:617, 823)


def test_cli_keeps_reference_options(tmp_path):
    r = _perl([os.path.join(PERL, "chromosome3D.pl"), "-h"])
    assert r.returncode == 1 and "ERROR!!" in r.stdout and "-i" in r.stdout and "-a" in r.stdout      # C3D:37, 2532, 2556
    r = _perl([os.path.join(PERL, "chromosome3D.pl"), "-o", str(tmp_path / "o")])
    assert r.returncode == 1 and "required" in r.stdout
    r = _perl([os.path.join(PERL, "chromosome3D.pl"), "-x", "1"])
    assert r.returncode == 1


@pytest.mark.skipif(api.device_count() > 0, reason="only meaningful without a GPU")
def test_front_end_fails_loudly_without_gpu(tmp_path):
    r = _perl([os.path.join(PERL, "chromosome3D.pl"), "-if", ds.input_path("chr21_1mb"), "-o", str(tmp_path / "o"), "-m", "2"])
    assert r.returncode != 0 and "no usable CUDA device" in (r.stderr + r.stdout)


def _md5(p):
    return hashlib.md5(open(p, "rb").read()).hexdigest()


@pytest.mark.gpu
def test_front_end_whole_run_layout_and_parity(tmp_path, golden_dir):
    out = tmp_path / "chr21"
    m = 7
    r = _perl([os.path.join(PERL, "chromosome3D.pl"), "-i", ds.input_path("chr21_1mb"), "-o", str(out), "-a", "0.5", "-m", str(m)])
    assert r.returncode == 0, r.stderr + r.stdout
    ID = "chr21_1mb_matrix"
    g = json.load(open(os.path.join(golden_dir, "restraints_md5.json")))["chr21_1mb@0.5"]
    assert _md5(out / (ID + ".dist")) == g["dist"] and _md5(out / (ID + ".rr")) == g["rr"] and _md5(out / "contact.tbl") == g["tbl"]
    assert open(out / (ID + ".txt"), "rb").read() == open(ds.input_path("chr21_1mb"), "rb").read()
    files = sorted(os.listdir(out))
    assert [f for f in files if "_model" in f] == ["%s_model%d.pdb" % (ID, k) for k in range(1, 6)]
    assert len([f for f in files if re.match(ID + r"_\d+\.pdb$", f)]) == m - 5
    assert "contact_violation.txt" in files and "model_info.log" in files
    # stdout: banners + table
    assert "L          : 37" in r.stdout and "Restraints : 528 lines in tbl file" in r.stdout
    rows = re.findall(r"^(\d+)/528\s+(\d+\.\d\d)\s+(\S+)\s*$", r.stdout, re.M)
    assert len(rows) == m
    picks = re.findall(r"^model(\d)\.pdb <= \./(\S+)\.pdb$", r.stdout, re.M)
    assert [int(k) for k, _ in picks] == [1, 2, 3, 4, 5]
    # the printed numbers are what the reference's scorer gives for the written coordinates
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    final = {name: "%s_model%s.pdb" % (ID, k) for k, name in picks}
    noe = {}
    for line in open(out / "model_info.log"):
        mm = re.match(r"^(?:\./(\S+)\.pdb)?REMARK (\w+)\s*=\s*(\S+)", line)
        if mm and mm.group(1):
            cur = mm.group(1)
        if mm and mm.group(2) == "noe":
            noe[cur] = float(mm.group(3))
    assert len(noe) == m
    for sat, dev, name in rows:
        xyz = ds.read_ca_pdb(str(out / final.get(name, name + ".pdb")))
        s, d = orc.assess(xyz, ri, rj, rd)
        assert int(sat) == s[0] and dev == "%.2f" % d[0], name
    order = [name for _, _, name in rows]                       # worst first by int(noe)
    assert [int(noe[n]) for n in order] == sorted((int(v) for v in noe.values()), reverse=True)
    assert [name for _, name in picks] == order[::-1][:5]
    # final model layout
    txt = open(out / (ID + "_model1.pdb")).read().split("\n")
    assert re.match(r"^ATOM      1  CA  [A-Z]{3}     1    [ \-\d\.]{24}  1\.00  0\.00           C  $", txt[0])
    assert txt[37] == "" and txt[38] == "CONECT    1    2" and txt[-2] == "END" and len(txt[0]) == 80
    viol = open(out / "contact_violation.txt").read().split("\n")
    assert len(viol) - 1 == m * (528 + 2) and viol[0].startswith("#NOE violation check")
    # Spearman script on the produced models == oracle
    r2 = _perl([os.path.join(PERL, "spearman_IF_pdb.pl"), ds.input_path("chr21_1mb"), str(out)])
    assert r2.returncode == 0, r2.stderr
    lines = r2.stdout.strip().split("\n")
    assert lines[0] == "SRCC\tPDB" and len(lines) == 1 + m
    for ln in lines[1:]:
        v, f = ln.split("\t")
        assert v == "%.3f" % orc.spearman(IF, ds.read_ca_pdb(f))[0]
    assert [float(l.split("\t")[0]) for l in lines[1:]] == sorted((float(l.split("\t")[0]) for l in lines[1:]), reverse=True)


# ------------------------------------------------------------------------------- SURVEY 8(f) rows through the Perl binding
def test_perl_reduced_writer_and_text_loader_host_side(tmp_path):
    """Host-only entry points through the Perl binding: the threaded text parser gives the same bytes as the Perl-side
    parser, and the reduced-PDB writer reproduces a shipped file from (oracle-averaged) coordinates."""
    src = os.path.join(ds.shipped_models_dir(), "chr21_500kb_rank04_a11.pdb")
    red = orc.reduce_model(ds.read_ca_pdb(src))
    red.tofile(tmp_path / "red.f64")
    code = ('use Chr3D; my ($a, $L) = Chr3D::read_if_text($ARGV[0]); my ($b, $L2) = Chr3D::read_if_matrix($ARGV[0]);'
            'open(my $f, "<:raw", $ARGV[1]) or die; local $/; my $x = <$f>; close $f;'
            'Chr3D::write_reduced_pdb($ARGV[2], $x, length($x) / 24); print "$L $L2 ", ($a eq $b ? "same" : "diff"), "\\n";')
    r = _perl(["-e", code, ds.input_path("chr22_1mb"), str(tmp_path / "red.f64"), str(tmp_path / "out.pdb")])
    assert r.returncode == 0, r.stderr
    L, L2, same = r.stdout.split()
    assert L == L2 and same == "same"
    assert open(tmp_path / "out.pdb", "rb").read() == open(src.replace(".pdb", "_reduced.pdb"), "rb").read()


@pytest.mark.gpu
def test_perl_similarity_tool_reproduces_shipped_block(tmp_path):
    d = ds.shipped_models_dir()
    r = _perl([os.path.join(PERL, "similarity_pdb.pl"), os.path.join(d, "chr21_1mb_rank07_a11.pdb"), os.path.join(d, "chr21_500kb_rank04_a11.pdb")])
    assert r.returncode == 0, r.stderr
    txt = open(os.path.join(d, "similarity.txt")).read().replace("\r", "")
    m = re.search(r"chr21_500kb_rank04_a11\nSpearman correlation: (\S+)\nRMSD: (\S+)", txt)
    lines = r.stdout.splitlines()
    assert lines[0] == "chr21_500kb_rank04_a11"
    assert abs(float(lines[1].split(": ")[1]) - float(m.group(1))) < 1e-12 and abs(float(lines[2].split(": ")[1]) - float(m.group(2))) < 1e-11
    r = _perl([os.path.join(PERL, "similarity_pdb.pl"), "--reduce", os.path.join(d, "chr21_500kb_rank04_a11.pdb"), str(tmp_path / "r.pdb")])
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "r.pdb", "rb").read() == open(os.path.join(d, "chr21_500kb_rank04_a11_reduced.pdb"), "rb").read()


@pytest.mark.gpu
def test_alpha_sweep_driver(tmp_path):
    """Three alphas x four models in one batched call; the selected file follows the shipped naming and layout and its
    Spearman equals the oracle's on the written coordinates."""
    out = tmp_path / "sweep"
    r = _perl([os.path.join(PERL, "chromosome3D_sweep.pl"), "-i", ds.input_path("chr21_1mb"), "-o", str(out), "-m", "4", "--alphas", "0.5,1.1,1.5", "--keep-all"])
    assert r.returncode == 0, r.stderr + r.stdout
    rows = [l.split("\t") for l in open(out / "alpha_sweep.txt").read().splitlines() if not l.startswith("#")]
    assert [row[0] for row in rows] == ["0.5", "1.1", "1.5"]
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    for row in rows:
        assert int(row[1]) == len(orc.if_to_restraints(IF, float(row[0]))[0])
    sel = re.search(r"Selected\s+: (\S+)\.pdb", r.stdout).group(1)
    assert re.fullmatch(r"chr21_1mb_rank\d\d_a(05|11|15)", sel)
    kept = sorted(f for f in os.listdir(out) if re.fullmatch(r"chr21_1mb_rank\d\d_a\d\d\.pdb", f))
    assert len(kept) == 3 and sel + ".pdb" in kept
    xyz = ds.read_ca_pdb(str(out / (sel + ".pdb")))
    best = min(float(row[3]) for row in rows)
    assert "%.3f" % orc.spearman(IF, xyz, 3)[0] == "%.3f" % best
    shipped = open(os.path.join(ds.shipped_models_dir(), "chr21_1mb_rank07_a11.pdb")).read().splitlines()
    mine = open(out / (sel + ".pdb")).read().splitlines()
    assert len(mine) == len(shipped) and [l[:30] for l in mine] == [l[:30] for l in shipped]        # same records, same columns
    red = open(out / (sel + "_reduced.pdb"), "rb").read()
    assert red.startswith(b"\r\nATOM      1   CA MET B1") and red.endswith(b"END\r\n") and red.count(b"ATOM") == 19


@pytest.mark.gpu
def test_front_end_dg_start(tmp_path):
    """--start dg: the DG stage (C3D:1470-1526) supplies the starting coordinates; the run keeps the reference's layout."""
    out = tmp_path / "dg"
    r = _perl([os.path.join(PERL, "chromosome3D.pl"), "-i", ds.input_path("chr21_1mb"), "-o", str(out), "-m", "3", "--start", "dg"])
    assert r.returncode == 0, r.stderr + r.stdout
    assert re.search(r"DG start\s+: embedded 3 models \(\d+ pairs", r.stdout)
    m = re.search(r"DG start\s+: dgscale ([0-9.]+), (\d+) iterations .* E_dg ([0-9.e+]+) -> ([0-9.e+]+)", r.stdout)
    assert m and 1.0 < float(m.group(1)) < 1.5 and 1 <= int(m.group(2)) <= 200 and float(m.group(4)) < float(m.group(3))
    assert all(os.path.exists(out / ("chr21_1mb_matrix_model%d.pdb" % k)) for k in (1, 2, 3))
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    xyz = ds.read_ca_pdb(str(out / "chr21_1mb_matrix_model1.pdb"))
    assert orc.spearman(IF, xyz, 3)[0] < -0.8


@pytest.mark.gpu
def test_front_end_cns_pin(tmp_path):
    """--pin: the run anneals with chr3d_cns_pin_params (soft CA(i)-CA(i+2) lower bound, repel radius 5.5 A) and equals the Python
    mirror's run with the same parameters; no CA(i)-CA(i+2) distance of the best model falls far below the bound."""
    from chromosome3d_b200 import api
    out = tmp_path / "pin"
    r = _perl([os.path.join(PERL, "chromosome3D.pl"), "-i", ds.input_path("chr21_1mb"), "-o", str(out), "-m", "3", "--pin"])
    assert r.returncode == 0, r.stderr + r.stdout
    IF = ds.load_if_matrix(ds.input_path("chr21_1mb"))
    xyz = ds.read_ca_pdb(str(out / "chr21_1mb_matrix_model1.pdb"))
    api.init(0)
    res = api.build_models([api.Job(IF=IF, alpha=0.5, n_models=3, seed=82364)], p=api.cns_pin_params())[0]
    best = int(api.rank_models(res.energies)[0])
    assert np.abs(np.round(res.xyz[best], 3) - xyz).max() <= 1.5e-3          # the PDB holds three decimals
    d2 = np.linalg.norm(xyz[2:] - xyz[:-2], axis=1)
    assert d2.min() > 5.0
