
import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
from chromosome3d_b200 import abi, api, datasets as ds
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); api.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
box = [api.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
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
