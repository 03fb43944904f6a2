import sys, time, os
sys.path.insert(0, ".")
import numpy as np
from chromosome3d_b200 import api, datasets as ds
api.init(0)
IF = ds.load_if_matrix(ds.input_path("chr1_500kb"))
p = api.default_params()
for k in range(4):
    t0 = time.perf_counter()
    r = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=20, seed=82364)], p)[0]
    print("e2e call %d: %.1f ms (device %.1f ms)" % (k, 1e3 * (time.perf_counter() - t0), r.device_ms), file=sys.stderr)
