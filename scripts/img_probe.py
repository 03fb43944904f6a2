#!/usr/bin/env python3
"""Development probe: would running the two hot-phase images as separate (smaller-cluster) trajectories pay?"""
import os, sys, time
sys.path.insert(0, ".")
from chromosome3d_b200 import api, datasets as ds
api.init(0)
IF = ds.load_if_matrix(ds.input_path("chr1_500kb"))
nm = int(sys.argv[1])
p = api.default_params()
p.images = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for rep in range(3):
    res = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=nm, seed=82364)], p)[0]
print("models", nm, "images", p.images, "cluster", os.environ.get("CHR3D_SA_CLUSTER"), "warps", os.environ.get("CHR3D_SA_WARPS"), "device_ms", res.device_ms, "evals/model", float(res.force_evals.mean()))
