#!/usr/bin/env python3
"""Regenerates tests/golden/*.json by running the REAL reference Perl from /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py [--full]

What runs for real:
  * chromosome3D.pl  IF2dist_new / dist2rr / carr2tbl (C3D:110-206, 340-362) -> md5 of .dist/.rr/contact.tbl
  * chromosome3D.pl  assess_dgsa (C3D:769-829) on CNS-style trial PDBs written by a stub `cns_solve`
  * spearman_IF_pdb.pl with two stand-in modules for its un-vendored dependencies
    (`confold` rebuilt from C3D:620-729 semantics, `Statistics::RankCorrelation` = average ranks +
    Pearson on ranks): pins pair selection, %.3f distance text and output format, not the module.
The scratch copy of chromosome3D.pl differs from the reference in two lines only: the CNS path
(C3D:13) and the bash-only `&>` redirection (C3D:296), both irrelevant to the functions pinned.
Nothing is written inside /root/reference or the repo except the JSON/PDB fixtures next to this file.
"""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
SCRATCH = "/tmp/c3d_golden"


def md5(path):
    return hashlib.md5(open(path, "rb").read()).hexdigest()


def setup():
    shutil.rmtree(SCRATCH, ignore_errors=True)
    cns = os.path.join(SCRATCH, "cns")
    os.makedirs(os.path.join(cns, "libraries", "toppar"))
    os.makedirs(os.path.join(cns, "intel-x86_64bit-linux", "bin"))
    open(os.path.join(cns, "cns_solve_env.sh"), "w").write("# stub\n")
    open(os.path.join(cns, "libraries", "toppar", "protein.param"), "w").write("! stub\n")
    stub = os.path.join(cns, "intel-x86_64bit-linux", "bin", "cns_solve")
    open(stub, "w").write(
        "#!/bin/bash\n# stub engine: drains the CNS script, provides the files the Perl checks for\n"
        "cat > /dev/null\ntouch extended.pdb extended.mtf\n"
        "if [ -n \"$C3D_STUB_MODELS\" ] && [ -d \"$C3D_STUB_MODELS\" ]; then cp $C3D_STUB_MODELS/*.pdb . ; touch dg_sub_embed_1.pdb; fi\n")
    os.chmod(stub, 0o755)
    src = open(os.path.join(REF, "chromosome3D.pl")).read()
    src = src.replace('"/home/bap54/chr3D_projects/project01/programs/cns_solve_1.3"', '"%s"' % cns)
    src = src.replace('system("$command &> $log");', 'system("$command > $log 2>&1");')
    open(os.path.join(SCRATCH, "c3d.pl"), "w").write(src)
    # stand-in modules for spearman_IF_pdb.pl
    mod = os.path.join(SCRATCH, "mods")
    os.makedirs(os.path.join(mod, "Statistics"))
    open(os.path.join(mod, "confold.pm"), "w").write(r'''package confold;
use strict; use warnings; use Exporter 'import';
our @EXPORT = qw(load_pdb ca_xyz_pdb calc_dist);
sub load_pdb { my $d = shift; my @l = <$d/*.pdb>; return @l; }
sub ca_xyz_pdb { my $f = shift; my %x = (); open F, $f or die $!;
  while(<F>){ next if $_ !~ m/^ATOM/; my $an = substr($_,12,4); $an =~ s/\s+//g; next unless $an eq "CA";
    my $rn = substr($_,22,5); $rn =~ s/\s+//g; my $x = substr($_,30,8); my $y = substr($_,38,8); my $z = substr($_,46,8);
    $x =~ s/\s+//g; $y =~ s/\s+//g; $z =~ s/\s+//g; $x{$rn} = "$x $y $z"; } close F; return %x; }
sub calc_dist { my @a = split(/\s+/, shift); my @b = split(/\s+/, shift);
  return sprintf "%.3f", sqrt(($a[0]-$b[0])**2+($a[1]-$b[1])**2+($a[2]-$b[2])**2); }
1;
''')
    open(os.path.join(mod, "Statistics", "RankCorrelation.pm"), "w").write(r'''package Statistics::RankCorrelation;
use strict; use warnings;
sub new { my ($c, $x, $y, %o) = @_; my $s = { x => $x, y => $y }; return bless $s, $c; }
sub _rank { my $v = shift; my @idx = sort { $v->[$a] <=> $v->[$b] } 0..$#$v; my @r; my $i = 0;
  while ($i <= $#idx) { my $j = $i; $j++ while ($j < $#idx && $v->[$idx[$j+1]] == $v->[$idx[$i]]);
    my $avg = (($i+1)+($j+1))/2; $r[$idx[$_]] = $avg for ($i..$j); $i = $j+1; } return \@r; }
sub spearman { my $s = shift; my $rx = _rank($s->{x}); my $ry = _rank($s->{y}); my $n = scalar @$rx;
  my ($xs,$ys) = (0,0); $xs += $_ for @$rx; $ys += $_ for @$ry; my ($xm,$ym) = ($xs/$n, $ys/$n);
  my ($sx,$sy,$sxy) = (0,0,0);
  for my $i (0..$n-1) { $sx += ($rx->[$i]-$xm)**2; $sy += ($ry->[$i]-$ym)**2; $sxy += ($rx->[$i]-$xm)*($ry->[$i]-$ym); }
  return $sx*$sy ? $sxy/sqrt($sx*$sy) : 0; }
1;
''')
    sp = open(os.path.join(REF, "spearman_IF_pdb.pl")).read()
    sp = sp.replace("use lib '/home/bap54/chr3D_projects/project01/scripts';", "")
    sp = sp.replace("use lib '/home/bap54/chr3D_projects/project01/scripts/Statistics-RankCorrelation-0.1204/lib';", "")
    open(os.path.join(SCRATCH, "spearman.pl"), "w").write(sp)


def run_c3d(matrix, alpha, out, m=1, models_dir=None):
    env = dict(os.environ)
    if models_dir:
        env["C3D_STUB_MODELS"] = models_dir
    r = subprocess.run(["perl", os.path.join(SCRATCH, "c3d.pl"), "-i", matrix, "-o", out, "-a", str(alpha), "-m", str(m)],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return r


def golden_restraints(full):
    names = sorted(f[:-len("_matrix.txt")] for f in os.listdir(os.path.join(REF, "input")))
    alphas_all = ["%.1f" % (0.1 * k) for k in range(1, 21)]
    plan = []
    for n in names:
        if n in ("chr21_1mb", "chr22_1mb"):
            al = alphas_all
        elif n.endswith("_1mb") or full:
            al = ["0.5", "1.1"] if not full else alphas_all
        elif n in ("chr1_500kb", "chr21_500kb", "chr13_500kb"):
            al = ["0.5", "1.1"]
        else:
            al = ["1.1"]
        plan += [(n, a) for a in al]
    out = {}
    for k, (n, a) in enumerate(plan):
        d = os.path.join(SCRATCH, "run_%s_%s" % (n, a))
        run_c3d(os.path.join(REF, "input", n + "_matrix.txt"), a, d)
        ID = n + "_matrix"
        rows = sum(1 for _ in open(os.path.join(d, "contact.tbl")))
        out["%s@%s" % (n, a)] = dict(rows=rows, dist=md5(os.path.join(d, ID + ".dist")), rr=md5(os.path.join(d, ID + ".rr")),
                                     tbl=md5(os.path.join(d, "contact.tbl")))
        if n == "chr21_1mb" and a == "1.1":
            shutil.copy(os.path.join(d, "contact.tbl"), os.path.join(HERE, "chr21_1mb_a1.1_contact.tbl"))
        shutil.rmtree(d)
        print("restraints %d/%d %s a=%s rows=%d" % (k + 1, len(plan), n, a, rows), flush=True)
    json.dump(out, open(os.path.join(HERE, "restraints_md5.json"), "w"), indent=1, sort_keys=True)


def write_cns_style_pdb(path, xyz, e4, resnames):
    """What a CNS trial structure looks like to the Perl: REMARK energies, several atoms per residue, END."""
    with open(path, "w") as f:
        f.write("REMARK FILENAME=\"%s\"\n" % os.path.basename(path))
        for name, v in zip(("overall", "bon", "vdw", "noe"), e4):
            f.write("REMARK %s = %.4f\n" % (name, v))
        serial = 0
        for i, (x, y, z) in enumerate(xyz):
            for an, off in (("N", -0.7), ("CA", 0.0), ("C", 0.7)):
                serial += 1
                f.write("ATOM  %5d  %-3s %3s  %4d    %8.3f%8.3f%8.3f  1.00  0.00           %s  \n"
                        % (serial, an, resnames[i], i + 1, x + off, y + off if an != "CA" else y, z, an[0]))
        f.write("END\n")


def golden_assess():
    import numpy as np
    import oracle as orc
    from chromosome3d_b200 import datasets as ds
    refseq = re.search(r'my \$REFSEQUENCE = "([A-Z]+)"', open(os.path.join(REF, "chromosome3D.pl")).read()).group(1)
    aa = dict(A="ALA", N="ASN", C="CYS", Q="GLN", H="HIS", L="LEU", M="MET", P="PRO", T="THR", Y="TYR", R="ARG", D="ASP",
              E="GLU", G="GLY", I="ILE", K="LYS", F="PHE", S="SER", W="TRP", V="VAL")
    out = {}
    for name, alpha, m in (("chr21_1mb", "0.5", 6), ("chr22_1mb", "1.1", 4)):
        IF = ds.load_if_matrix(os.path.join(REF, "input", name + "_matrix.txt"))
        N = IF.shape[0]
        ri, rj, rd = orc.if_to_restraints(IF, float(alpha))
        p = orc.default_params()
        prog = orc.build_program(p)
        res = orc.run_ensemble(N, ri, rj, rd, p, prog, 4242, m)
        mdir = os.path.join(SCRATCH, "models_%s" % name)
        os.makedirs(mdir, exist_ok=True)
        ID = name + "_matrix"
        xyz = np.round(res["xyz"], 3)
        # make model 2 a noisy copy so that energies (and ranks) differ clearly
        xyz[1] += np.random.RandomState(1).normal(0, 0.3, xyz[1].shape).round(3)
        e = res["energies"].copy()
        e[:, 3] += np.arange(m) * 7.3
        for k in range(m):
            write_cns_style_pdb(os.path.join(mdir, "%s_%d.pdb" % (ID, k + 1)), xyz[k], e[k], [aa[c] for c in refseq[:N]])
        d = os.path.join(SCRATCH, "assess_%s" % name)
        r = run_c3d(os.path.join(REF, "input", name + "_matrix.txt"), alpha, d, m=m, models_dir=mdir)
        table = {}
        for line in r.stdout.splitlines():
            mm = re.match(r"^(-?\d+)/(\d+)\s+(-?[\d.]+)\s+(\S+)\s*$", line)
            if mm:
                table[mm.group(4)] = dict(count=int(mm.group(1)), total=int(mm.group(2)), sum_dev=mm.group(3))
        ranks = re.findall(r"^model(\d+)\.pdb <= \./(\S+)\.pdb", r.stdout, re.M)
        final = open(os.path.join(d, ID + "_model1.pdb")).read()
        out[name] = dict(alpha=alpha, m=m, xyz=xyz.tolist(), energies=e.tolist(), table=table,
                         ranks=[[int(a), b] for a, b in ranks], model1_pdb=final,
                         stdout_tail=r.stdout[-1500:], returncode=r.returncode)
        print("assess", name, table, ranks, "rc", r.returncode)
    json.dump(out, open(os.path.join(HERE, "assess.json"), "w"), indent=1, sort_keys=True)


def golden_spearman():
    out = {}
    pairs = [("chr21_1mb", "chr21_1mb_rank07_a11.pdb"), ("chr22_1mb", "chr22_1mb_rank02_a11.pdb"),
             ("chr1_1mb", "chr1_1mb_rank09_a11.pdb"), ("chr1_500kb", "chr1_500kb_rank03_a11.pdb"),
             ("chr13_500kb", "chr13_500kb_rank10_a11.pdb")]
    avail = set(os.listdir(os.path.join(REF, "output_models")))
    for name, pdb in pairs:
        if pdb not in avail:
            cand = [f for f in avail if f.startswith(name + "_rank") and f.endswith("_a11.pdb")]
            if not cand:
                continue
            pdb = sorted(cand)[0]
        for rng in ("3", "1", "10"):
            r = subprocess.run(["perl", "-I", os.path.join(SCRATCH, "mods"), os.path.join(SCRATCH, "spearman.pl"),
                                os.path.join(REF, "input", name + "_matrix.txt"), os.path.join(REF, "output_models", pdb), rng],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            lines = r.stdout.strip().splitlines()
            val = lines[1].split("\t")[0] if len(lines) > 1 else None
            out["%s|%s|%s" % (name, pdb, rng)] = val
            print("spearman", name, pdb, rng, val, flush=True)
            if name == "chr1_500kb":
                break
    json.dump(out, open(os.path.join(HERE, "spearman.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    setup()
    what = [a for a in sys.argv[1:] if not a.startswith("--")] or ["restraints", "assess", "spearman"]
    if "assess" in what:
        golden_assess()
    if "spearman" in what:
        golden_spearman()
    if "restraints" in what:
        golden_restraints("--full" in sys.argv)
