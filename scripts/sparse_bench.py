#!/usr/bin/env python3
"""Per-evaluation cost of the sparse force provider (CSR restraint rows + cell-list repel) against the dense tile engine at
N = 24,925 beads: restraint tables of 10 %, 1 % and 0.1 % of the pairs (real 10 kb Hi-C maps are in the 0.1-1 % range), hot MD
from a consistent random-walk structure.  Prints one JSON line per density (and the dense engine's number for the full table)."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from chromosome3d_b200 import abi, api

api.init(0)
N = int(os.environ.get("N", "24925"))
nstep = int(os.environ.get("NSTEP", "200"))
p = api.default_params()
w = dict(w_noe=2.0, w_bond=200.0, w_vdw=4.0, repel=0.9)
prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                        abi.stage(abi.OP_DYNAMICS, nstep=nstep, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **w), abi.stage(abi.OP_ENERGY, **w)])
rng = np.random.default_rng(7)
# the hidden structure: a 3.8 A random walk, restraint distances are its distances, so the run starts at a consistent state
step = rng.normal(size=(N, 3)); step *= 3.8 / np.linalg.norm(step, axis=1, keepdims=True)
X = np.cumsum(step, axis=0)
for dens in [float(v) for v in os.environ.get("DENSITY", "0.1,0.01,0.001").split(",")]:
    t0 = time.time()
    ri, rj = [], []
    for r0 in range(0, N, 512):
        r1 = min(N, r0 + 512)
        m = rng.random((r1 - r0, N)) < dens
        ii, jj = np.nonzero(m)
        ii += r0
        k = jj >= ii + 5
        ri.append(ii[k].astype(np.int32)); rj.append(jj[k].astype(np.int32))
    ri, rj = np.concatenate(ri), np.concatenate(rj)
    rd = np.maximum(np.linalg.norm(X[ri] - X[rj], axis=1), 1.0)
    t1 = time.time()
    s = api.Single(N=N, restraints=(ri, rj, rd), x0=X, seed=3, p=p, program=prog)
    t2 = time.time()
    s.run(); s.run()
    r = s.fetch()
    info = s.sparse_info()
    s.close()
    ev = nstep + 2                                  # the opening evaluation of DYNAMICS, one per step, ENERGY
    print(json.dumps(dict(N=N, density=dens, restraint_rows=int(len(ri)), build_list_s=round(t1 - t0, 2), create_s=round(t2 - t1, 2), force_evals=int(ev),
                          ms_per_evaluation=r["device_ms"] / ev, pair_evals_per_s=len(ri) * ev / (r["device_ms"] * 1e-3), rebuilds=info["rebuilds"],
                          noe_bytes_per_eval=info["noe_bytes_per_eval"], noe_GBps_if_alone=info["noe_bytes_per_eval"] / (r["device_ms"] / ev * 1e-3) / 1e9)), flush=True)
    del ri, rj, rd
if os.environ.get("DENSE", "1") == "1":
    s = api.Single(N=N, alpha=1.1, synth_seed=20260002, x0=X, seed=3, p=p, program=prog)
    s.run(); s.run()
    r = s.fetch()
    s.close()
    print(json.dumps(dict(N=N, density="dense tile engine (synthetic IF, every pair restrained)", restraint_rows=int(r["restraint_rows"]), force_evals=int(r["force_evals"]),
                          ms_per_evaluation=r["device_ms"] / (nstep + 2))), flush=True)
