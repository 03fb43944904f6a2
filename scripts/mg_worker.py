#!/usr/bin/env python3
"""Trajectory equality of the sharded single structure (run under torch.distributed.run with N ranks, one per GPU): short MD,
cooling and L-BFGS programs on a 700-bead and a 4,985-bead structure, sharded run (peer-memory exchange inside the persistent
kernel when CHR3D_TEST_P2P=1, NCCL all-reduce baseline otherwise) against the one-GPU run.  Prints one RESULT line."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
from chromosome3d_b200 import abi, api, datasets as ds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
api.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p2p = os.environ.get("CHR3D_TEST_P2P") == "1"
p = api.default_params()
w = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
progs = {"md": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                abi.stage(abi.OP_DYNAMICS, nstep=10, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **w), abi.stage(abi.OP_ENERGY, **w)],
         "cool": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                  abi.stage(abi.OP_DYNAMICS, nstep=6, flag=abi.COUPLE_VSCALE, dt=0.005, temp=500.0, **w),
                  abi.stage(abi.OP_DYNAMICS, nstep=6, flag=abi.COUPLE_VSCALE, dt=0.005, temp=475.0, **w), abi.stage(abi.OP_ENERGY, **w)],
         "min": [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=4, drop=10.0, **w), abi.stage(abi.OP_ENERGY, **w)]}
out = {}
for N, seed in ((700, 4242), (4985, 20260001)):
    IF = ds.synthetic_if(N, seed) if N <= 1000 else None
    for name, st in progs.items():
        prog = abi.stage_array(st)
        nccl_id = None
        if not p2p:
            box = [api.nccl_unique_id() if rank == 0 else None]      # a unique id serves one communicator
            dist.broadcast_object_list(box, src=0)
            nccl_id = box[0]
        kw = dict(IF=IF, alpha=1.0) if IF is not None else dict(N=N, alpha=1.1, synth_seed=seed)
        s = api.Single(seed=11, p=p, program=prog, rank=rank, world=world, nccl_id=nccl_id, **kw)
        if p2p:                                           # exchange fused into the persistent kernel
            hs = [None] * world
            dist.all_gather_object(hs, s.comm_handle())
            s.comm_connect(hs)
        s.run()
        r = s.fetch()
        info = s.info()
        cerr = s.comm_error()
        s.close()
        assert cerr == 0, "peer exchange timed out"
        rows = torch.tensor([float(r["restraint_rows"])], device="cuda")
        dist.all_reduce(rows)
        # every rank holds the same coordinates (bit for bit): compare checksums
        chk = torch.tensor([float(np.abs(r["xyz"]).sum()), float(r["energies"][0])], dtype=torch.float64, device="cuda")
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if rank == 0:
            one = api.Single(seed=11, p=p, program=prog, **kw)
            one.run()
            o = one.fetch()
            one.close()
            out["%d_%s" % (N, name)] = dict(e=r["energies"].tolist(), e1=o["energies"].tolist(), dx=float(np.abs(r["xyz"] - o["xyz"]).max()),
                                            rel_de=float(abs(r["energies"][0] - o["energies"][0]) / abs(o["energies"][0])),
                                            rows=rows.item(), rows1=o["restraint_rows"], evals=r["force_evals"], evals1=o["force_evals"],
                                            tiles_rank0=info["tiles_this_rank"], ranks_identical=bool(torch.equal(lo, hi)))
        dist.barrier()
if rank == 0:
    print("RESULT " + json.dumps(dict(world=world, exchange="peer memory (persistent kernel)" if p2p else "nccl all-reduce (legacy path)", cases=out)))
dist.barrier()
dist.destroy_process_group()
