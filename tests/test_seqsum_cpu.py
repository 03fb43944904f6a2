"""The arithmetic of the parallel left-to-right fp64 sum (chromosome3d_b200/csrc/seqsum_core.h, used by seq_sum_exact_kernel for
the Perl's `$sum += ...` loops, C3D:132-139 and C3D:590-598), compiled for the host and checked against the scalar loop."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_windowed_integer_scan_equals_the_scalar_loop(tmp_path):
    exe = str(tmp_path / "seqsum_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "seqsum_host_check.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "all ok" in out.stdout
    # real data folds in windows: the Hi-C-like case must not degrade to a pass per term
    hic = [ln for ln in out.stdout.splitlines() if ln.startswith("ok   hic_pow") and "window=8192" in ln]
    assert hic and all(int(ln.split("iterations=")[1]) < 200 for ln in hic), hic
