mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_next_rows.py -m gpu -x -q > gpurun_out/pytest_next.log 2>&1; echo "pytest_exit=$?"
tail -25 gpurun_out/pytest_next.log
python - <<'PY'
import sys, time; sys.path.insert(0, '.')
import numpy as np
from chromosome3d_b200 import api, datasets as ds
api.init(0)
IF = ds.load_if_matrix(ds.input_path("chr1_500kb")); N = IF.shape[0]
ri, rj, rd = api.if_to_restraints(IF, 1.1)
for rep in range(2):
    t = time.time(); x0, eig, bad = api.dg_embed(ri, rj, rd, N, n_models=20, seed=82364); print("dg_embed 20 models N=455: %.1f ms, bad %d, eig %s" % ((time.time() - t) * 1e3, bad, eig[0]))
a = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=20, seed=82364, x0=x0)])[0]
b = api.build_models([api.Job(IF=IF, alpha=1.1, n_models=20, seed=82364)])[0]
for nm, r in (("dg start", a), ("coil start", b)):
    rho = api.spearman_IF_pdb(IF, r.xyz, 3)
    print(nm, "device_ms %.1f evals %.0f noe median %.0f min %.0f srcc median %.3f best %.3f" % (r.device_ms, r.force_evals.mean(), np.median(r.energies[:, 3]), r.energies[:, 3].min(), np.median(rho), rho.min()))
PY
