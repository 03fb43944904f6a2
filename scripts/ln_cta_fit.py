#!/usr/bin/env python3
"""Cost model of the large-N tile walk: per-CTA tile-phase times of the persistent kernel (chr3d_single_profile_ctas) fitted to
time = a*far_tiles + b*near_tiles + c*rows_started.  Prints b/a and c/a (LN_COST_NEAR, LN_COST_ROW of large_n.cu) and the
spread of the per-CTA times (max/mean): what the slowest CTA costs every evaluation at the first grid barrier."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from chromosome3d_b200 import abi, api
api.init(0)
W = dict(w_noe=2.0, w_bond=200.0, w_vdw=0.01, repel=0.9)
for N, seed in ((24925, 20260002), (4985, 20260001)):
    prog = abi.stage_array([abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0),
                            abi.stage(abi.OP_DYNAMICS, nstep=60, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, **W)])
    s = api.Single(N=N, alpha=1.1, synth_seed=seed, seed=5, p=api.default_params(), program=prog)
    s.run(); s.fetch(); s.profile_ctas(); s.profile()
    s.run(); r = s.fetch(); pc = s.profile_ctas(); ph = s.profile()
    s.close()
    far = pc["tiles"] - pc["near"]
    A = np.stack([far, pc["near"], pc["rows"]], 1).astype(float)
    coef, res, rk, sv = np.linalg.lstsq(A, pc["tile_ms"], rcond=None)
    evals = 61
    print(json.dumps(dict(N=N, ctas=int(len(far)), tiles_min=int(pc["tiles"].min()), tiles_max=int(pc["tiles"].max()), rows_max=int(pc["rows"].max()),
                          us_per_far_tile=1e3 * coef[0] / evals, near_over_far=coef[1] / coef[0], row_over_far=coef[2] / coef[0],
                          cta_ms_mean=float(pc["tile_ms"].mean()), cta_ms_max=float(pc["tile_ms"].max()), spread=float(pc["tile_ms"].max() / pc["tile_ms"].mean()),
                          resid_rms=float(np.sqrt(np.mean((A @ coef - pc["tile_ms"]) ** 2))), phases=ph, device_ms=r["device_ms"])), flush=True)
    order = np.argsort(-pc["tile_ms"])[:5]
    print("  slowest CTAs:", [(int(c), int(pc["tiles"][c]), int(pc["near"][c]), int(pc["rows"][c]), round(float(pc["tile_ms"][c]), 3)) for c in order], flush=True)
    order = np.argsort(pc["tile_ms"])[:5]
    print("  fastest CTAs:", [(int(c), int(pc["tiles"][c]), int(pc["near"][c]), int(pc["rows"][c]), round(float(pc["tile_ms"][c]), 3)) for c in order], flush=True)
