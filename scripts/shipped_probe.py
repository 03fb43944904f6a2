#!/usr/bin/env python3
"""All shipped CNS models against 10-model ensembles of this engine at the script's defaults (alpha 0.5, K 11):
IF-Spearman, radius of gyration, mean bond length."""
import glob, os, sys
sys.path.insert(0, ".")
import numpy as np
from chromosome3d_b200 import api, datasets as ds
api.init(0)
rg = lambda x: np.sqrt(((x - x.mean(0)) ** 2).sum(1).mean())
names, jobs, ships = [], [], []
for name in ds.list_inputs():
    hits = glob.glob(os.path.join(ds.shipped_models_dir(), "%s_rank*_a11.pdb" % name))
    if not hits:
        continue
    IF = ds.load_if_matrix(ds.input_path(name))
    ship = ds.read_ca_pdb(hits[0])
    if ship.shape[0] != IF.shape[0]:
        continue
    names.append(name); ships.append((IF, ship)); jobs.append(api.Job(IF=IF, alpha=0.5, n_models=10, seed=82364))
res = api.build_models(jobs)
d_rho, d_rg = [], []
for name, (IF, ship), r in zip(names, ships, res):
    rho_ship = api.spearman_IF_pdb(IF, ship, 3)[0]
    rho = api.spearman_IF_pdb(IF, r.xyz, 3)
    mine_rg = np.mean([rg(x) for x in r.xyz])
    d_rho.append(np.median(rho) - rho_ship); d_rg.append(mine_rg / rg(ship) - 1)
    print("%-12s N=%3d srcc shipped %.3f ours median %.3f best %.3f | Rg shipped %.1f ours %.1f" % (name, IF.shape[0], rho_ship, np.median(rho), rho.min(), rg(ship), mine_rg))
d_rho, d_rg = np.array(d_rho), np.array(d_rg)
print("models %d | median-srcc difference: mean %.3f, |d|<=0.02: %d, |d|<=0.05: %d, max |d| %.3f | Rg ratio-1: mean %.3f max |.| %.3f" %
      (len(names), d_rho.mean(), (abs(d_rho) <= 0.02).sum(), (abs(d_rho) <= 0.05).sum(), abs(d_rho).max(), d_rg.mean(), abs(d_rg).max()))
