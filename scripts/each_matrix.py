#!/usr/bin/env python3
"""Development: one 20-model ensemble per input matrix, each in its own process (a launch failure kills the context)."""
import subprocess, sys
sys.path.insert(0, ".")
from chromosome3d_b200 import datasets as ds
code = """
import sys; sys.path.insert(0, '.')
from chromosome3d_b200 import api, datasets as ds
api.init(0)
IF = ds.load_if_matrix(ds.input_path(sys.argv[1]))
e = api.Ensemble([api.Job(IF=IF, alpha=1.1, n_models=int(sys.argv[2]), seed=82364)])
print(sys.argv[1], IF.shape[0], e.info())
r = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=int(sys.argv[2]), seed=82364)])[0]
print('   ok', r.device_ms)
"""
names = sys.argv[2:] or ds.list_inputs()
for n in names:
    r = subprocess.run([sys.executable, "-c", code, n, sys.argv[1]], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    print(r.stdout.strip() if r.returncode == 0 else "%s FAILED: %s | %s" % (n, r.stdout.strip(), r.stderr.strip().splitlines()[-1][:200]), flush=True)
