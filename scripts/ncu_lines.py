#!/usr/bin/env python3
"""Aggregates the stall samples of an ncu report per CUDA source line (SASS offsets mapped through nvdisasm -g).
usage: ncu_lines.py report.ncu-rep object.o mangled_kernel_substring [top]"""
import csv, io, re, subprocess, sys, collections
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [k for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
dis = subprocess.run(["nvdisasm", "-g", "-c", obj], stdout=subprocess.PIPE, text=True).stdout
# find the kernel's section, map instruction index -> line
lines = []; cur = None; inside = False
for l in dis.split("\n"):
    if l.startswith(".text."):
        inside = kern in l
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass rows", len(data), "disasm instr", len(lines))
agg = collections.Counter(); inst = collections.Counter(); st = collections.defaultdict(collections.Counter)
tot = 0
for k, r in enumerate(data):
    s = int(r[si]); tot += s
    ln = lines[k] if k < len(lines) else None
    agg[ln] += s; inst[ln] += int(r[ii])
    for c in stall_cols:
        v = int(r[c]) if r[c].isdigit() else 0
        if v: st[ln][hdr[c]] += v
print("total samples", tot)
src = {}
for ln, s in agg.most_common(top):
    if ln and ln[0] not in src:
        try: src[ln[0]] = open("chromosome3d_b200/csrc/" + ln[0]).read().split("\n")
        except Exception: src[ln[0]] = []
    text = src[ln[0]][ln[1] - 1].strip()[:90] if ln and src.get(ln[0]) and ln[1] - 1 < len(src[ln[0]]) else ""
    tops = ",".join("%s:%d" % (a.replace("stall_", ""), b) for a, b in st[ln].most_common(3))
    print("%5.1f%% inst %9d %s:%s  %s   [%s]" % (100.0 * s / tot, inst[ln], ln[0] if ln else "?", ln[1] if ln else "?", text, tops))
