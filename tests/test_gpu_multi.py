"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): the sharded single-structure run over NCCL
reproduces the one-GPU run, and ensemble shards need no communication."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from chromosome3d_b200 import abi, api, datasets as ds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); api.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 700
IF = ds.synthetic_if(N, 4242)
p = api.default_params()
w = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
progs = {"md": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                abi.stage(abi.OP_DYNAMICS, nstep=10, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **w), abi.stage(abi.OP_ENERGY, **w)],
         "min": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=4, drop=10.0, **w), abi.stage(abi.OP_ENERGY, **w)]}
out = {}
for name, st in progs.items():
    prog = abi.stage_array(st)
    box = [api.nccl_unique_id() if rank == 0 else None]      # a unique id serves one communicator
    dist.broadcast_object_list(box, src=0)
    s = api.Single(IF=IF, alpha=1.0, seed=11, p=p, program=prog, rank=rank, world=world, nccl_id=box[0])
    if os.environ.get("CHR3D_TEST_P2P") == "1":        # fused peer-memory exchange instead of NCCL all-reduce
        hs = [None] * world
        dist.all_gather_object(hs, s.comm_handle())
        s.comm_connect(hs)
    s.run(); r = s.fetch(); info = s.info(); cerr = s.comm_error(); s.close()
    assert cerr == 0, "peer exchange timed out"
    rows = torch.tensor([float(r["restraint_rows"])], device="cuda"); dist.all_reduce(rows)
    if rank == 0:
        one = api.Single(IF=IF, alpha=1.0, seed=11, p=p, program=prog)
        one.run(); o = one.fetch(); one.close()
        out[name] = dict(e=r["energies"].tolist(), e1=o["energies"].tolist(), dx=float(np.abs(r["xyz"] - o["xyz"]).max()),
                         rows=rows.item(), rows1=o["restraint_rows"], evals=r["force_evals"], evals1=o["force_evals"], tiles=info["tiles_this_rank"])
    dist.barrier()
if rank == 0:
    print("RESULT " + json.dumps(out))
dist.barrier(); dist.destroy_process_group()
'''


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("p2p", ["0", "1"])
def test_sharded_single_structure_matches_one_gpu(tmp_path, p2p):
    """10 MD steps, and 4 L-BFGS iterations, on 2 ranks (tile-range shards; force exchange by NCCL all-reduce or by
    the exchange fused into the step kernels) vs 1 GPU: only the fp32 summation order of the partial forces
    differs, so coordinates agree to 1e-3 A and energies to 1e-5 relative (short runs: hot MD is chaotic and
    the line search takes discrete decisions, so long runs diverge legitimately)."""
    import json
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "2951%s" % p2p, str(script)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, CHR3D_TEST_P2P=p2p))
    assert r.returncode == 0, r.stdout[-3000:]
    line = [l for l in r.stdout.split("\n") if l.startswith("RESULT ")][0]
    both = json.loads(line[7:])
    for name, o in both.items():
        assert o["rows"] == o["rows1"] and o["evals"] == o["evals1"], (name, o)
        assert np.allclose(o["e"], o["e1"], rtol=1e-5, atol=1e-6 * abs(o["e1"][0])), (name, o)
        assert o["dx"] < 1e-3, (name, o)


def test_sweep_driver_shards_alphas_over_two_gpus(tmp_path):
    """perl/chromosome3D_sweep.pl --gpus 2: alphas dealt to one child process per GPU, no communication; the merged table and
    the selected model equal the one-GPU run (same seeds, same launch shapes per job)."""
    import re
    import shutil
    from chromosome3d_b200 import api, datasets as ds
    if api.device_count() < 2 or shutil.which("perl") is None:
        pytest.skip("needs 2 GPUs and perl")
    perl = os.path.join(ROOT, "perl")
    outs = {}
    for tag, extra in (("one", []), ("two", ["--gpus", "2"])):
        out = tmp_path / tag
        r = subprocess.run(["perl", "-I", perl, os.path.join(perl, "chromosome3D_sweep.pl"), "-i", ds.input_path("chr22_1mb"), "-o", str(out), "-m", "3",
                            "--alphas", "0.4,0.5,0.6,0.7"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert r.returncode == 0, r.stderr + r.stdout
        outs[tag] = (open(out / "alpha_sweep.txt").read(), re.findall(r"Selected\s+: (\S+)\.pdb", r.stdout)[-1], sorted(os.listdir(out)))
    # the children anneal 6 trajectories per call instead of 12, so the launch shapes (hence the reduction orders) may differ:
    # same alphas and restraint counts, statistically the same scores
    rows = {t: [l.split("\t") for l in outs[t][0].splitlines() if not l.startswith("#")] for t in outs}
    assert [r[0] for r in rows["one"]] == [r[0] for r in rows["two"]] == ["0.4", "0.5", "0.6", "0.7"]
    assert [r[1] for r in rows["one"]] == [r[1] for r in rows["two"]]
    for a, b in zip(rows["one"], rows["two"]):
        assert abs(float(a[3]) - float(b[3])) < 0.05 and abs(float(a[4]) - float(b[4])) < 0.05
    for t in outs:
        sel, files = outs[t][1], outs[t][2]
        assert re.fullmatch(r"chr22_1mb_rank0[123]_a0[4567]", sel)
        assert files == ["alpha_sweep.txt", sel + ".pdb", sel + "_reduced.pdb"], files
