"""Developer check run on the GPU box: exercises every subsystem against the oracle and prints diagnostics."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as orc
from chromosome3d_b200 import abi, api, datasets as ds

def section(name):
    print("\n=== %s ===" % name, flush=True)

def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
    sys.stdout.flush()

api.init(0)
print(api.version())
mats = {n: ds.load_if_matrix(ds.input_path(n)) for n in ("chr21_1mb", "chr22_1mb", "chr1_1mb", "chr1_500kb")}

def t_restraints():
    section("restraints")
    for name, alpha in [("chr21_1mb", 0.5), ("chr21_1mb", 1.1), ("chr21_1mb", 2.0), ("chr1_500kb", 1.1), ("chr1_500kb", 0.3)]:
        IF = mats[name]
        t0 = time.time(); t, npatch = api.IF2dist_new(IF, alpha); ri, rj, rd = api.if_to_restraints(IF, alpha); dt = time.time() - t0
        to = orc.if_to_dist(IF, alpha); ori, orj, ord_ = orc.dist_to_restraints(to)
        print(name, alpha, "tenths equal", np.array_equal(t, to), "ndiff", int((t != to).sum()), "patched", npatch,
              "rows", len(ri), len(ori), "list equal", np.array_equal(ri, ori) and np.array_equal(rj, orj) and np.array_equal(rd, ord_), "t=%.3fs" % dt)

def t_energy():
    section("energy/forces")
    p = api.default_params()
    for name, alpha in [("chr21_1mb", 0.5), ("chr1_1mb", 1.1), ("chr1_500kb", 1.1)]:
        IF = mats[name]; N = IF.shape[0]
        ri, rj, rd = orc.if_to_restraints(IF, alpha)
        for scale, w in [(1.0, abi.stage(abi.OP_ENERGY, w_noe=10, w_bond=1000, w_vdw=4.0, repel=0.85)),
                         (0.3, abi.stage(abi.OP_ENERGY, w_noe=1, w_bond=100, w_vdw=20.0, repel=0.5))]:
            x = np.stack([orc.random_coil(N, 3.8 * scale, 11, m) for m in range(3)]).astype(np.float32).astype(np.float64)
            e, f = api.energy_forces(x, ri, rj, rd, p, w)
            eo, fo = orc.energy_forces(x, ri, rj, rd, p, w)
            print(name, alpha, scale, "E rel", np.abs(e - eo).max(0) / np.maximum(np.abs(eo).max(0), 1e-30), "F rel", np.abs(f - fo).max() / np.abs(fo).max())

def t_short():
    section("short-horizon MD / minimise parity")
    p = api.default_params()
    for name, alpha in [("chr21_1mb", 0.5), ("chr1_1mb", 0.5)]:
        IF = mats[name]; N = IF.shape[0]
        ri, rj, rd = orc.if_to_restraints(IF, alpha)
        x0 = orc.random_coil(N, 3.8, 5, 0).astype(np.float32).astype(np.float64)
        for label, prog in [
            ("md_tcoup", [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0), abi.stage(abi.OP_DYNAMICS, nstep=20, flag=abi.COUPLE_TCOUP, dt=0.003, temp=2000.0, w_noe=1.0, w_bond=100.0, w_vdw=0.01, repel=0.9), abi.stage(abi.OP_ENERGY, w_noe=1.0, w_bond=100.0, w_vdw=0.01, repel=0.9)]),
            ("md_vscale", [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MAXWELL, temp=300.0), abi.stage(abi.OP_DYNAMICS, nstep=20, flag=abi.COUPLE_VSCALE, dt=0.005, temp=500.0), abi.stage(abi.OP_ENERGY)]),
            ("min30", [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=30, drop=10.0), abi.stage(abi.OP_ENERGY)]),
            ("min300", [abi.stage(abi.OP_INIT, flag=1), abi.stage(abi.OP_MINIMIZE, nstep=300, drop=10.0), abi.stage(abi.OP_CENTER), abi.stage(abi.OP_ENERGY)]),
        ]:
            parr = abi.stage_array(prog)
            r = api.build_models([api.Job(restraints=(ri, rj, rd), N=N, n_models=1, seed=9, x0=x0[None])], p, parr)[0]
            o = orc.run_model(N, ri, rj, rd, p, parr, 9, 0, x0)
            dx = np.abs(r.xyz[0] - o["xyz"]).max()
            print(name, label, "max|dx|=%.3e" % dx, "E gpu", r.energies[0], "E cpu", o["energies"], "evals", r.force_evals[0], o["force_evals"])

def t_full():
    section("full SA ensembles")
    for name, alpha, m in [("chr21_1mb", 0.5, 8), ("chr22_1mb", 0.5, 8), ("chr1_1mb", 0.5, 8), ("chr21_1mb", 1.1, 8), ("chr1_500kb", 1.1, 20), ("chr1_500kb", 0.5, 20)]:
        IF = mats[name]; N = IF.shape[0]
        p = api.default_params()
        ens = api.Ensemble([api.Job(IF=IF, alpha=alpha, n_models=m, seed=82364)], p)
        t0 = time.time(); ens.run(); res = ens.fetch()[0]; dt = time.time() - t0
        info = ens.info(); ens.close()
        ri, rj, rd = orc.if_to_restraints(IF, alpha)
        rho = api.spearman_IF_pdb(IF, res.xyz); rho_o = orc.spearman(IF, res.xyz)
        sat_o, dev_o = orc.assess(res.xyz, ri, rj, rd)
        x = res.xyz; bl = np.linalg.norm(x[:, 1:] - x[:, :-1], axis=2)
        rg = np.sqrt(((x - x.mean(1, keepdims=True)) ** 2).sum(2).mean(1))
        pe = float(res.n_restraints) * res.force_evals.sum()
        print(name, alpha, "N", N, "cfg", info, "wall %.3fs dev %.1fms" % (dt, res.device_ms), "evals", res.force_evals.mean(), "pair-evals/s %.3e" % (pe / (res.device_ms * 1e-3)),
              "\n   Enoe med %.1f" % np.median(res.energies[:, 3]), "E", res.energies[0], "hand", res.hand.tolist(),
              "\n   bond mean/max %.2f %.2f" % (bl.mean(), bl.max()), "Rg %.1f" % rg.mean(), "rho med %.3f [%.3f %.3f]" % (np.median(rho), rho.min(), rho.max()),
              "rho==oracle", np.array_equal(rho, rho_o), "sat", res.satisfied.tolist()[:4], "sat==oracle", np.array_equal(sat_o, res.satisfied),
              "dev==oracle", np.array_equal(dev_o, res.sum_dev), np.abs(dev_o - res.sum_dev).max())
        if name == "chr21_1mb":
            t0 = time.time(); o = orc.run_ensemble(N, ri, rj, rd, p, orc.build_program(p), 82364, m); dto = time.time() - t0
            print("   oracle: Enoe med %.1f" % np.median(o["energies"][:, 3]), "evals", o["force_evals"].mean(), "rho med %.3f" % np.median(orc.spearman(IF, o["xyz"])), "t %.2fs" % dto)

for fn in (t_restraints, t_energy, t_short, t_full):
    run(fn)
print("done")
