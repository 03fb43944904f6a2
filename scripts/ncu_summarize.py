#!/usr/bin/env python3
"""Turns the ncu artefacts of scripts/r2_final.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
usage: ncu_summarize.py <tag>"""
import collections, csv, io, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02e"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    c = collections.Counter()
    for r in data:
        for k in cols:
            if r[k].isdigit():
                c[hdr[k]] += int(r[k])
    tot = sum(c.values())
    return [(k, v, 100.0 * v / max(tot, 1)) for k, v in c.most_common(8)], len(data)


def kernel_summary(rep, name, note):
    m = raw(rep)
    st, nsass = stalls(rep)
    js = {"report": os.path.basename(rep), "kernel": m.get("Kernel Name", ("?", ""))[0], "metrics": {k: m[k][0] for k in WANT if k in m},
          "units": {k: m[k][1] for k in WANT if k in m}, "stall_samples_top": st, "sass_instructions": nsass}
    rd = float(m.get("dram__bytes_read.sum", ("0", ""))[0] or 0); wr = float(m.get("dram__bytes_write.sum", ("0", ""))[0] or 0)
    unit = m.get("dram__bytes_read.sum", ("", "byte"))[1].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
    unit_w = m.get("dram__bytes_write.sum", ("", "byte"))[1].lower()
    scale_w = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit_w, 1.0)
    js["dram_bytes_per_launch"] = rd * scale + wr * scale_w
    with open(os.path.join(P, "%s_%s_ncu.json" % (tag, name)), "w") as f:
        json.dump(js, f, indent=1)
    lines = ["# %s — ncu `--set full --clock-control none --import-source on` capture of `%s`" % (tag, name), "", note, "",
             "| metric | value | unit |", "|---|---|---|"]
    for k in WANT:
        if k in m:
            lines.append("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
    lines += ["", "DRAM traffic per launch (read + write): %.4g bytes" % js["dram_bytes_per_launch"], "",
              "Warp-stall samples (all warps, whole launch):", "", "| reason | samples | share |", "|---|---|---|"]
    lines += ["| %s | %d | %.1f %% |" % s for s in st]
    open(os.path.join(P, "%s_%s_ncu.md" % (tag, name)), "w").write("\n".join(lines) + "\n")
    return js


def launches():
    src = os.path.join(G, "%s_launches_bench.csv" % tag)
    rows = [r for r in csv.reader(open(src)) if len(r) > 5 and r[0].isdigit()]
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric Name, Unit, Value
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0]
        unit, val = r[-2], float(r[-1].replace(",", ""))
        ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    out = ["# %s — kernel launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` under ncu" % tag,
           "", "`ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv` (times are cold-cache and serialised; the SHARE is what matters).",
           "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| %s | %d | %.3f | %.1f %% |" % (k, n, ms, 100.0 * ms / max(tot, 1e-12)))
    open(os.path.join(P, "%s_launches_bench.md" % tag), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copy(src, os.path.join(P, "%s_launches_bench.csv" % tag))


launches()
sa = kernel_summary(os.path.join(G, "%s_sa_kernel.ncu-rep" % tag), "sa_kernel",
                    "Command: `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-n` (full dgsa.inp schedule, chr1_500kb alpha=1.1, 20 models); 4th sa_kernel launch captured.")
ln = kernel_summary(os.path.join(G, "%s_ln_tile_kernel.ncu-rep" % tag), "ln_tile_kernel",
                    "Command: `python scripts/large_n_probe.py 24925` (synthetic 10 kb matrix, N=24,925, 19,110 tiles = 1.25 GB of targets); 4th ln_tile_kernel launch captured.")
summary = {"dram_bytes_per_launch": sa["dram_bytes_per_launch"], "ln_tile_dram_bytes_per_launch": ln["dram_bytes_per_launch"], "tag": tag}
if os.path.exists(os.path.join(G, "%s_ln_persist_kernel.ncu-rep" % tag)):
    pk = kernel_summary(os.path.join(G, "%s_ln_persist_kernel.ncu-rep" % tag), "ln_persist_kernel",
                        "Command: `python scripts/persist_probe.py` (N=24,925, one DYNAMICS op of 16 steps = ONE cooperative launch of 17 force evaluations: tile phase, grid barrier / ready counters, gather + Verlet step, per evaluation); 2nd launch captured.")
    summary["ln_persist_dram_bytes_per_launch"] = pk["dram_bytes_per_launch"]
    summary["ln_persist_evaluations_per_launch"] = 17
    print(open(os.path.join(P, "%s_ln_persist_kernel_ncu.md" % tag)).read())
json.dump(summary, open(os.path.join(P, "sa_kernel_ncu_summary.json"), "w"), indent=1)
print(open(os.path.join(P, "%s_launches_bench.md" % tag)).read())
print(open(os.path.join(P, "%s_sa_kernel_ncu.md" % tag)).read())
print(open(os.path.join(P, "%s_ln_tile_kernel_ncu.md" % tag)).read())
