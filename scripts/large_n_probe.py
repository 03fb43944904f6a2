#!/usr/bin/env python3
"""Development probe (GPU box): force-evaluation throughput of the large-N tile engine."""
import json
import sys
import time

sys.path.insert(0, ".")
from chromosome3d_b200 import api  # noqa: E402

api.init(0)
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
print(json.dumps({"fp32_peak_tflops": api.fp32_peak(3), "fp32_peak_packed_tflops": api.fp32_peak_packed(3)}), flush=True)
for N in [int(a) for a in sys.argv[1:]] or [4985, 24925]:
    t0 = time.time()
    s = api.Single(N=N, alpha=1.1, synth_seed=20260001 + (N > 10000))
    t1 = time.time()
    info = s.info()
    for want_e in (True, False):
        tile_ms, eval_ms = s.eval_bench(reps=20, want_energy=want_e)
        pairs = N * (N - 1) // 2
        print(json.dumps({"N": N, "want_energy": want_e, "create_s": round(t1 - t0, 2), "tile_ms": tile_ms, "eval_ms": eval_ms,
                          "tiles": info["n_tiles"], "ctas": info["ctas"], "target_MB": info["target_bytes"] / 1e6,
                          "GBs_tile": info["target_bytes"] / tile_ms / 1e6, "frac_hbm": info["target_bytes"] / tile_ms / 1e6 / peaks["hbm_gbs"],
                          "pair_evals_per_s_eval": pairs / eval_ms * 1e3}), flush=True)
    s.close()
