"""One chr3d_run_ensemble call of the benchmarked config (chr1_500kb, alpha 1.1, 20 models, host buffers in, host buffers out),
repeated, with the left-to-right sums evaluated in parallel (default) and as a serial chain (chr3d_set_option seqsum=serial):
the same-box A/B behind the e2e line, and — under `ncu --metrics gpu__time_duration.sum` — the launch list of one call.
No torch: the process starts in a second."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from chromosome3d_b200 import api, datasets  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
api.init(0)
IF = datasets.load_if_matrix(datasets.input_path("chr1_500kb"))
p = api.default_params()


def timed(n):
    out = []
    for _ in range(n):
        t0 = time.perf_counter()
        res = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=20, seed=82364)], p)[0]
        out.append(1e3 * (time.perf_counter() - t0))
    return out, res


timed(2)
rows = {}
ref = None
for mode in ("parallel", "serial", "parallel", "serial"):
    api.set_option("seqsum", "serial" if mode == "serial" else None)
    ms, res = timed(reps)
    rows.setdefault(mode, []).extend(ms)
    key = (res.xyz.tobytes(), res.sum_dev.tobytes(), res.satisfied.tobytes(), res.energies.tobytes())
    if ref is None:
        ref = key
    assert key == ref, "results differ between the parallel and the serial evaluation of the sums"
api.set_option("seqsum", None)
print(json.dumps({"workload": "chr1_500kb_a1.1_m20 through chr3d_run_ensemble, host buffers", "calls_per_mode": 2 * reps,
                  "ms_median": {k: float(np.median(v)) for k, v in rows.items()}, "ms_min": {k: float(np.min(v)) for k, v in rows.items()},
                  "device_ms_last": float(res.device_ms), "identical_results": True}))
