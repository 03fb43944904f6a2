"""Source-level breakdown of an ncu capture of sa_kernel: share of the warp-stall samples and of the executed instructions per
region of sa_ensemble.cu, and the mbarrier wait sites one by one.

    ncu -i gpurun_out/r02e_sa_kernel.ncu-rep --page source --csv --print-source cuda,sass > /tmp/sa_cs.csv
    python scripts/ncu_source_regions.py /tmp/sa_cs.csv > profiles/r02e_sa_kernel_regions.md
"""
import csv
import sys
from collections import Counter


def load(path):
    seq, cur, curfile, i_s, i_i = [], None, None, None, None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            curfile = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            i_s, i_i = r.index("# Samples"), r.index("Instructions Executed")
            continue
        if r[0] != "":
            try:
                cur = (curfile, int(r[0]))
            except ValueError:
                pass
            continue
        if len(r) < 4 or r[2] in ("...", "", "-"):
            continue
        try:
            seq.append((cur, r[3].strip(), int(r[i_s]), int(r[i_i]), int(r[2], 16)))
        except ValueError:
            pass
    seq.sort(key=lambda s: s[4])
    return seq


# line ranges of sa_ensemble.cu at the commit the capture was taken from (r02e)
REGIONS = [("pair arithmetic (pair_sym2 and the packed intrinsics it calls)", lambda f, l: not f.startswith("sa_") or 474 <= l <= 494),
           ("pair loop around it (tile bookkeeping, target and coordinate loads)", lambda f, l: 500 <= l <= 565 or 282 <= l <= 286),
           ("row / column reductions in the warp, st.async of the partial forces", lambda f, l: 602 <= l <= 645 or 297 <= l <= 309),
           ("mbarrier waits (all sites)", lambda f, l: 317 <= l <= 346),
           ("gather of the partial forces + bond term (owner threads)", lambda f, l: 646 <= l <= 706),
           ("Verlet step, kinetic energy, thermostat, publishing coordinates", lambda f, l: 906 <= l <= 950 or 358 <= l <= 380 or 271 <= l <= 276),
           ("all-reduce of owner columns", lambda f, l: 385 <= l <= 425),
           ("L-BFGS bookkeeping", lambda f, l: 709 <= l <= 893)]


def main():
    seq = load(sys.argv[1])
    tot_s, tot_i = sum(s[2] for s in seq), sum(s[3] for s in seq)
    cs, ci = Counter(), Counter()
    for line, _, smp, ins, _ in seq:
        if not line:
            continue
        name = next((n for n, f in REGIONS if f(line[0], line[1])), "everything else (stage program, set-up, SAVE / PICK, output)")
        cs[name] += smp
        ci[name] += ins
    print("| region of `sa_kernel` | warp-stall samples | instructions executed |\n|---|---|---|")
    for name, v in cs.most_common():
        print("| %s | %.1f %% | %.1f %% |" % (name, 100.0 * v / tot_s, 100.0 * ci[name] / tot_i))
    print("\nWait sites (`mbarrier.try_wait` loops; the line after each loop tells which wait it is):\n")
    print("| site | samples | instructions | spin iterations |\n|---|---|---|---|")
    skip = {342, 323, 318, 306, 297, 151, 320, 321, 345}
    for k, s in enumerate(seq):
        if "TRYWAIT" not in s[1]:
            continue
        ss, ii = sum(x[2] for x in seq[k - 1:k + 9]), sum(x[3] for x in seq[k - 1:k + 9])
        if ss < 0.002 * tot_s:
            continue
        nxt = [x[0][1] for x in seq[k + 9:k + 80] if x[0] and x[0][0].startswith("sa_") and x[0][1] not in skip]
        what = "partial forces (MB_F)" if nxt and 660 <= nxt[0] <= 670 else ("coordinates + kinetic energy (MB_X, dynamics)" if nxt and nxt[0] == 930 else
               ("all-reduce scalars (MB_R)" if nxt and 415 <= nxt[0] <= 420 else "coordinates (MB_X, minimiser / install)"))
        print("| %s | %.2f %% | %.2f %% | %d |" % (what, 100.0 * ss / tot_s, 100.0 * ii / tot_i, seq[k][3]))


if __name__ == "__main__":
    main()
