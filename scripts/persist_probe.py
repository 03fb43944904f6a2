#!/usr/bin/env python3
"""One short DYNAMICS op of the 24,925-bead structure (a single ln_persist_kernel launch of NSTEP + 1 force evaluations), run
twice: the command behind the ncu capture of the persistent kernel (profiles/*_ln_persist_kernel_ncu.md)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from chromosome3d_b200 import abi, api

api.init(0)
N, nstep = int(os.environ.get("N", "24925")), int(os.environ.get("NSTEP", "16"))
w = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                        abi.stage(abi.OP_DYNAMICS, nstep=nstep, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **w)])
s = api.Single(N=N, alpha=1.1, synth_seed=20260002, seed=3, p=api.default_params(), program=prog)
for _ in range(2):
    s.run()
    r = s.fetch()
    print(json.dumps(dict(N=N, force_evals=int(r["force_evals"]), device_ms=r["device_ms"], ms_per_evaluation=r["device_ms"] / (nstep + 1), profile=s.profile())), flush=True)
s.close()
