#!/bin/bash
# Development: rebuild with -DSA_TIMING into a scratch library and print the per-phase cycle breakdown.
set -e
cd chromosome3d_b200/csrc
mkdir -p /tmp/satim
for f in api restraints sa_ensemble assess microbench large_n; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr -DCHR3D_WITH_NCCL -DSA_TIMING -c -o /tmp/satim/$f.o $f.cu &
done; wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /tmp/satim/libchr3d.so /tmp/satim/*.o -Xcompiler -fPIC -ldl
cd ../..
cp chromosome3d_b200/libchr3d.so /tmp/satim/libchr3d.orig.so
cp /tmp/satim/libchr3d.so chromosome3d_b200/libchr3d.so
python - <<'PY'
import ctypes as C, numpy as np, os, sys
sys.path.insert(0, ".")
from chromosome3d_b200 import api, datasets as ds
api.init(0)
IF = ds.load_if_matrix(ds.input_path("chr1_500kb"))
p = api.default_params()
res = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=20, seed=82364)], p)[0]
out = (C.c_ulonglong * 32)()
api.lib().chr3d_debug_sa_timing(out)
names = ["control (MD)", "pair phase", "wait forces", "gather+bond", "kick+publish", "wait coords", "lambda+update", "L-BFGS digest"]
for w, nm in ((0, "owner warp 0"), (1, "last warp")):
    a = list(out[16 * w:16 * w + 11]); ne = a[10]
    print(nm, "evals", ne, "device_ms", res.device_ms)
    tot = sum(a[:8])
    for k in range(8):
        print("   %-20s %8.0f cycles/eval  %5.1f%%" % (names[k], a[k] / max(ne, 1), 100.0 * a[k] / max(tot, 1)))
    print("   total %.0f cycles/eval = %.2f us at 1.965 GHz" % (tot / max(ne, 1), tot / max(ne, 1) / 1965.0))
PY
cp /tmp/satim/libchr3d.orig.so chromosome3d_b200/libchr3d.so
