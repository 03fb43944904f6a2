#!/usr/bin/env python3
"""Development: event trace of the persistent kernel for a short MD op of the 24,925-bead structure on N ranks (torchrun), printed
for rank 0 and the last rank as microseconds relative to the first event.  Needs the -DCHR3D_DEV build (CHR3D_LIBRARY)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
from chromosome3d_b200 import abi, api
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); api.init(local)
dist.init_process_group("nccl" if world > 1 else "gloo", **({"device_id": torch.device("cuda", local)} if world > 1 else {}))
W = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                        abi.stage(abi.OP_DYNAMICS, nstep=40, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **W)])
NB = int(os.environ.get("TRACE_N", "24925"))
s = api.Single(N=NB, alpha=1.1, synth_seed=20260002, seed=5, p=api.default_params(), program=prog, rank=rank, world=world)
if world > 1:
    hs = [None] * world
    dist.all_gather_object(hs, s.comm_handle()); s.comm_connect(hs)
for _ in range(3):
    s.run(); s.fetch()
buf = (C.c_ulonglong * 2048)()
assert api.lib().chr3d_single_trace(s._h, buf) == 0
t = np.array(buf[:], dtype=np.uint64).reshape(2, 512, 2)
dist.barrier()
for r in (0, world - 1):
    if rank == r:
        for w in (0, 1):
            ev = t[w]; n = int((ev[:, 1] > 0).sum())
            t0 = int(ev[0, 1])
            # evaluations 10..13 of the op
            starts = [k for k in range(n) if ev[k, 0] == 1]
            if len(starts) > 14:
                a, b = starts[10], starts[14]
                print("rank %d cta %s: " % (r, "0" if w == 0 else "mid") + " ".join("%d@%.1f" % (int(ev[k, 0]), (int(ev[k, 1]) - int(ev[a, 1])) / 1e3) for k in range(a, b)), flush=True)
    dist.barrier()
s.close()
dist.destroy_process_group()
