#!/usr/bin/env python3
"""How long does the fp32 L-BFGS of the ensemble kernel track the fp64 oracle?  chr1_1mb / chr21_1mb, alpha 0.5, from the same
coil: energy and coordinates after k = 1..8 iterations for CNS's drop = 10 and for a first step that fp32 gradients resolve."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import oracle as orc
from chromosome3d_b200 import abi, api, datasets as ds

api.init(0)
p = api.default_params()
for name in ("chr21_1mb", "chr1_1mb"):
    IF = ds.load_if_matrix(ds.input_path(name))
    N = IF.shape[0]
    ri, rj, rd = orc.if_to_restraints(IF, 0.5)
    x0 = orc.random_coil(N, 3.8, 5, 0).astype(np.float32).astype(np.float64)
    for drop in (10.0, 1.0e4, 1.0e6):
        for k in (1, 2, 3, 4, 5, 6, 8, 12, 20, -1):
            mins = [abi.stage(abi.OP_MINIMIZE, nstep=k, drop=drop)] if k > 0 else [abi.stage(abi.OP_MINIMIZE, nstep=2000, drop=drop)] * 3     # -1: to convergence
            prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1)] + mins + [abi.stage(abi.OP_ENERGY)])
            r = api.build_models([api.Job(restraints=(ri, rj, rd), N=N, n_models=1, seed=9, x0=x0[None])], p, prog)[0]
            o = orc.run_model(N, ri, rj, rd, p, prog, 9, 0, x0)
            print("%s drop %.0e k %2d  E gpu %.6e oracle %.6e rel %.2e  max|dx| %.2e  evals %d / %d" % (
                name, drop, k, r.energies[0, 0], o["energies"][0], abs(r.energies[0, 0] - o["energies"][0]) / o["energies"][0],
                np.abs(r.xyz[0] - o["xyz"]).max(), r.force_evals[0], o["force_evals"]), flush=True)
