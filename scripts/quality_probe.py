import sys; sys.path.insert(0, '.')
import numpy as np
from chromosome3d_b200 import api, datasets as ds
api.init(0)
rg = lambda x: np.sqrt(((x - x.mean(0)) ** 2).sum(1).mean())
names = sys.argv[1:] or ["chr21_1mb", "chr1_1mb", "chr21_500kb", "chr1_500kb"]
for name in names:
    IF = ds.load_if_matrix(ds.input_path(name))
    for alpha in (0.5, 0.8, 1.1):
        r = api.build_models([api.Job(IF=IF, alpha=alpha, n_models=10, seed=82364)])[0]
        rho = api.spearman_IF_pdb(IF, r.xyz, 3)
        b = np.linalg.norm(r.xyz[:, 1:] - r.xyz[:, :-1], axis=2)
        ri, rj, rd = api.if_to_restraints(IF, alpha)
        print("%s a=%.1f srcc median %.3f best %.3f | Rg %.1f bond mean %.2f max %.1f | target d median %.1f max %.0f | noe %.3g" %
              (name, alpha, np.median(rho), rho.min(), np.mean([rg(x) for x in r.xyz]), b.mean(), b.max(), np.median(rd), rd.max(), np.median(r.energies[:, 3])), flush=True)
