#!/usr/bin/env python3
"""Do DG starts change the annealed ensembles?    For every matrix with a shipped CNS model: the fp64
oracle's ensemble (dgsa.inp schedule, 10 models, alpha = 0.5) annealed from random coils and from the DG stage's output
(embedding, dgscale, 200-step bounds minimisation), scored by the IF-distance Spearman (SP:32-71), the best NOE energy and the
Spearman against the shipped model.  CPU only (the oracle is the checker, not the product).  Writes a markdown table."""
import glob, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import oracle as orc
from chromosome3d_b200 import datasets as ds

alpha, n_models = float(os.environ.get("ALPHA", "0.5")), int(os.environ.get("MODELS", "10"))
p = orc.default_params()
prog = orc.build_program(p)
print("| matrix | N | shipped srcc | coil: median srcc | DG: median srcc | coil: best NOE | DG: best NOE | DG dgscale | DG E_dg scaled -> final | DG stage iterations |")
print("|---|---|---|---|---|---|---|---|---|---|")
dc, dd, t0 = [], [], time.time()
for name in ds.list_inputs():
    hits = glob.glob(os.path.join(ds.shipped_models_dir(), "%s_rank*_a11.pdb" % name))
    if not hits:
        continue
    IF = ds.load_if_matrix(ds.input_path(name))
    ship = ds.read_ca_pdb(hits[0])
    N = IF.shape[0]
    if ship.shape[0] != N:
        continue
    ri, rj, rd = orc.if_to_restraints(IF, alpha)
    rho_ship = orc.spearman(IF, ship)[0]
    coil = orc.run_ensemble(N, ri, rj, rd, p, prog, 82364, n_models, 0, None, 0)
    x0, _, st, _ = orc.dg_starts(N, ri, rj, rd, p, 82364, n_models)
    dg = orc.run_ensemble(N, ri, rj, rd, p, prog, 82364, n_models, 0, x0, 0)
    rc, rg_ = np.median(orc.spearman(IF, coil["xyz"])), np.median(orc.spearman(IF, dg["xyz"]))
    dc.append(rc - rho_ship); dd.append(rg_ - rho_ship)
    print("| %s | %d | %.3f | %.3f | %.3f | %.1f | %.1f | %.3f | %.3g -> %.3g | %d |" % (name, N, rho_ship, rc, rg_, coil["energies"][:, 1].min(), dg["energies"][:, 1].min(),
          st[0]["dgscale"], st[0]["e_scaled"], st[0]["e_final"], st[0]["iterations"]), flush=True)
dc, dd = np.array(dc), np.array(dd)
print("\n%d matrices, alpha %.1f, %d models each, %.0f s of oracle time.  Median-srcc difference to the shipped model: coil starts mean %+.4f (|d| <= 0.02 for %d), "
      "DG starts mean %+.4f (|d| <= 0.02 for %d); largest |coil - DG| difference %.4f." % (len(dc), alpha, n_models, time.time() - t0, dc.mean(), int((abs(dc) <= 0.02).sum()),
                                                                                       dd.mean(), int((abs(dd) <= 0.02).sum()), float(abs(dc - dd).max())))
