import sys, time
sys.path.insert(0, ".")
from chromosome3d_b200 import api, datasets as ds
api.init(0)
for name in ("chr1_500kb", "chr21_1mb"):
    IF = ds.load_if_matrix(ds.input_path(name))
    for a in (1.1, 0.5):
        api.IF2dist_new(IF, a)
        t0 = time.perf_counter(); t, npatch = api.IF2dist_new(IF, a); dt = time.perf_counter() - t0
        print(name, a, "n_patched", npatch, "ms %.2f" % (1e3 * dt))
