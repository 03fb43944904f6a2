#!/usr/bin/env python3
"""Weak CNS pin: ensembles of the fp64 oracle on all 45 matrices with a shipped CNS model,
for a grid of the bead-model stand-in parameters (CA(i)-CA(i+2) lower bound d2 / k2, repel radius R_min), scored against the
shipped models: median IF-Spearman difference, radius of gyration.  CPU only (the oracle is the checker, not the product)."""
import glob, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import oracle as orc
from chromosome3d_b200 import datasets as ds

rg = lambda x: np.sqrt(((x - x.mean(0)) ** 2).sum(1).mean())
alpha = float(os.environ.get("ALPHA", "0.5"))
n_models = int(os.environ.get("MODELS", "10"))
cases = []
for name in ds.list_inputs():
    hits = glob.glob(os.path.join(ds.shipped_models_dir(), "%s_rank*_a11.pdb" % name))
    if not hits:
        continue
    IF = ds.load_if_matrix(ds.input_path(name))
    ship = ds.read_ca_pdb(hits[0])
    if ship.shape[0] != IF.shape[0]:
        continue
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    cases.append((name, IF, ship, ri, rj, rd, orc.spearman(IF, ship)[0], rg(ship)))
grid = [tuple(float(v) for v in g.split(",")) for g in sys.argv[1:]] or [(0.0, 0.0, 4.73)]
for d2, k2, rmin in grid:
    p = orc.default_params()
    p.angle_d2, p.angle_k, p.vdw_rmin = d2, k2, rmin
    prog = orc.build_program(p)
    t0 = time.time()
    rows = []
    for name, IF, ship, ri, rj, rd, rho_ship, rg_ship in cases:
        res = orc.run_ensemble(IF.shape[0], ri, rj, rd, p, prog, 82364, n_models, 0, None, 0)
        rho = orc.spearman(IF, res["xyz"])
        rows.append((name, IF.shape[0], rho_ship, float(np.median(rho)), rg_ship, float(np.mean([rg(x) for x in res["xyz"]]))))
    d = np.array([r[3] - r[2] for r in rows]); g = np.array([r[5] / r[4] - 1 for r in rows])
    out = dict(alpha=alpha, d2=d2, k2=k2, rmin=rmin, n=len(rows), mean_d=float(d.mean()), within_002=int((abs(d) <= 0.02).sum()), within_005=int((abs(d) <= 0.05).sum()),
               max_d=float(abs(d).max()), rg_mean=float(g.mean()), rg_max=float(abs(g).max()), seconds=time.time() - t0,
               worst=[(rows[k][0], round(float(d[k]), 3)) for k in np.argsort(-abs(d))[:6]])
    print(json.dumps(out), flush=True)
    if os.environ.get("ROWS"):
        for r in rows:
            print("%-12s N=%3d srcc shipped %.3f ours median %.3f | Rg shipped %.1f ours %.1f" % r, flush=True)
