#!/usr/bin/perl -w
# chromosome3D.pl — front end of the B200-native Chromosome3D hot path.
#
# Same command line, input format and output directory layout as the reference script
# (/root/reference/chromosome3D.pl, cited as C3D:N), but the distance-geometry simulated-annealing
# run is done in-process by libchr3d.so (CUDA, sm_100a) through Chr3D.pm instead of writing CNS
# scripts and shelling out to `cns_solve` (C3D:234-289).  What is kept:
#   options      -i <IF matrix> -o <dir> [-k 11] [-a 0.5] [-m 20] [-h]          (C3D:28-35; `-if` accepted, RM:57)
#   side files   <ID>.txt <ID>.dist <ID>.rr contact.tbl                          (C3D:59, 87-89)
#   models       <ID>_<k>.pdb trial structures with REMARK energies, then CA-only renumbered files with
#                CONECT rows, best five renamed <ID>_model{1..5}.pdb             (C3D:813-828)
#   reports      stdout banners and the assessment table, contact_violation.txt, model_info.log (C3D:478-483, 804-810, 869)
# What is gone: the fake-protein scaffolding (fasta, input.seq, gseq.inp, extn.inp, dgsa.inp, job.sh) and the CNS
# install checks (C3D:81-84) — there is no CNS.
use strict;
use warnings;
use Getopt::Long;
use File::Basename;
use File::Copy qw(copy);
use Cwd qw(abs_path);
use FindBin;
use lib $FindBin::Bin;
use Chr3D;

my ($help, $dir_out, $KSCALING, $ALPHA, $mcount, $file_if, $seed, $device, $range, $start, $pin);
($KSCALING, $ALPHA, $mcount) = (11, 0.5, 20);          # C3D:18-21
my $min_sep  = 5;                                       # $SEPARATION, C3D:20
my $DISTRELAX = 0.5;                                    # C3D:74
Getopt::Long::Configure("no_auto_abbrev");
GetOptions("h" => \$help, "o=s" => \$dir_out, "k=i" => \$KSCALING, "a=s" => \$ALPHA, "m=i" => \$mcount,
           "i|if=s" => \$file_if, "seed=i" => \$seed, "device=i" => \$device, "start=s" => \$start, "pin" => \$pin) or usage("bad option");
$start //= "coil";
usage("--start must be coil or dg") if $start ne "coil" && $start ne "dg";
usage("") if $help;
usage("-i <IF matrix> is required") if !defined $file_if;
usage("-o <output directory> is required") if !defined $dir_out;
usage("input matrix $file_if does not exist") if !-f $file_if;
usage("-m must be at least 1") if $mcount < 1;

my $program = basename($0);
system("mkdir", "-p", $dir_out) == 0 or die "cannot create $dir_out\n";
$dir_out = abs_path($dir_out);
$file_if = abs_path($file_if);
print "Start Time : " . localtime() . " [$program]\n";
print "Input      : $file_if\n";
print "Output Dir : $dir_out\n";
print "Scaling(K) : $KSCALING\n";
print "Alpha      : $ALPHA\n";
print "Conversion : D = $KSCALING * mean(IF^$ALPHA) / IF^$ALPHA\n";      # what C3D:139-152 computes (the banner at C3D:53 misstates it)
unlink grep { -f $_ } glob("$dir_out/*");               # the reference starts from an empty directory (C3D:56)
my $ID = basename($file_if, ".txt");
copy($file_if, "$dir_out/$ID.txt") or die "cannot copy $file_if: $!\n";
chdir $dir_out or die "cannot chdir $dir_out: $!\n";

Chr3D::init($device // 0);
my ($IF, $L) = Chr3D::read_if_matrix("$ID.txt");
print "L          : $L\n";

# ---- IF -> restraints (C3D:87-89) --------------------------------------------------------------
my $tenths = Chr3D::if_to_dist($IF, $L, $ALPHA + 0, $KSCALING + 0);
Chr3D::write_dist("$ID.dist", $tenths, $L);
my ($ri, $rj, $rd, $nr) = Chr3D::if_to_restraints($IF, $L, $ALPHA + 0, $KSCALING + 0, $min_sep);
Chr3D::write_rr("$ID.rr", $ri, $rj, $rd, $nr);
Chr3D::write_tbl("contact.tbl", $ri, $rj, $rd, $nr);
print "Restraints : $nr lines in tbl file\n";
die "no restraints: nothing to fold\n" if $nr < 1;

# ---- the DGSA run (replaces build_extended/build_models, C3D:234-289) --------------------------
print "\n(B) Build models using libchr3d dgsa (" . Chr3D::version() . ", binding: " . Chr3D::backend() . ")..\n";
my $x0;      # --start dg: the DG stage of dgsa.inp (C3D:1470-1526): embedding of the smoothed bounds, dgscale, 200-step bounds-only minimisation
if ($start eq "dg") {
    ($x0, my $bad, my $st) = Chr3D::dg_starts($ri, $rj, $rd, $nr, $L, $mcount, $seed // 0);
    print "DG start   : embedded $mcount models ($bad pairs with triangle-inconsistent bounds)\n";
    printf "DG start   : dgscale %.3f, %d iterations of the bounds-only minimisation (md.dg.step 200), E_dg %.4g -> %.4g\n", $st->[0], $st->[3], $st->[1], $st->[2];
}
my $res = Chr3D::run_ensemble(IF => $IF, L => $L, alpha => $ALPHA + 0, K => $KSCALING + 0, min_sep => $min_sep,
                              n_models => $mcount, seed => $seed // 0, x0 => $x0, params => ($pin ? Chr3D::cns_pin_params() : undef));
my @energies = unpack("d" . (4 * $mcount), $res->{energies});
my $xyz_bytes = 8 * 3 * $L;
my @pdb;
for my $k (1 .. $mcount) {      # trial structures <ID>_k.pdb with the REMARK block get_cns_energy() reads (C3D:602-618, 1818-1828)
    my $xyz = substr($res->{xyz}, ($k - 1) * $xyz_bytes, $xyz_bytes);
    my $e4  = substr($res->{energies}, ($k - 1) * 32, 32);
    Chr3D::write_pdb("./${ID}_$k.pdb", $xyz, $L, $e4, 0);
    push @pdb, "./${ID}_$k.pdb";
}
die "model ${ID}_$mcount.pdb was not written\n" if !-f "${ID}_$mcount.pdb";       # success test of C3D:266

# ---- assessment (assess_dgsa, C3D:769-829) ------------------------------------------------------
print "(C) Assess models..\n";
my @noe = map { int($energies[4 * $_ + 3]) } 0 .. $mcount - 1;                    # int() truncation, C3D:617
my @best_first  = @{ Chr3D::rank_models($res->{energies}, $mcount) };             # ascending, ties by model index
my @worst_first = reverse @best_first;
my @sat = unpack("l$mcount", $res->{satisfied});
my @dev = unpack("d$mcount", $res->{sum_dev});
print "\nNOE_SATISFIED(\x{c2}\x{b1}${DISTRELAX}A)  SUM_OF_DEVIATIONS>= 0.2  PDB\n";
open(my $viol, ">", "contact_violation.txt") or die $!;
my @tbl_rows = tbl_rows_for_log($ri, $rj, $rd, $nr);
for my $m (@worst_first) {
    printf "%-9s             %-9s                %-25s\n", "$sat[$m]/$nr", sprintf("%.2f", $dev[$m]), basename($pdb[$m], ".pdb");
    # contact_violation.txt (C3D:475-483): flag, deviation, measured distance, restraint row; flagged rows first
    my ($flag, $d, $dist) = Chr3D::assess_rows(substr($res->{xyz}, $m * $xyz_bytes, $xyz_bytes), $L, $ri, $rj, $rd, $nr);
    my @f = unpack("l$nr", $flag); my @dd = unpack("d$nr", $d); my @ds = unpack("d$nr", $dist);
    print $viol "#NOE violation check; $pdb[$m] against contact.tbl\n";
    print $viol "#violation-flag, deviation, actual-measurement, Input-NOE-restraint\n";
    for my $want (1, 0) {
        for my $k (0 .. $nr - 1) {
            next if $f[$k] != $want;
            print $viol sprintf("%3s\t%.2f\t%.2f # %s\n", $f[$k], $dd[$k], $ds[$k], $tbl_rows[$k]);
        }
    }
}
close $viol;

print "\nremoving non-CA ATOM rows and backing up REMARK rows..\n";
open(my $info, ">>", "model_info.log") or die $!;
for my $m (@worst_first) {      # REMARK block -> model_info.log (C3D:869-879); CA-only, renumbered, CONECT rows (C3D:208-215, 831-862)
    open(my $in, "<", $pdb[$m]) or die $!;
    print $info $pdb[$m];
    while (my $l = <$in>) { print $info $l if $l =~ /^REMARK/; }
    close $in;
    Chr3D::write_pdb($pdb[$m], substr($res->{xyz}, $m * $xyz_bytes, $xyz_bytes), $L, undef, 1);
}
close $info;
print "\n";
my $rank = 1;
for my $m (@best_first) {       # C3D:822-828
    print "model$rank.pdb <= $pdb[$m]\n";
    rename($pdb[$m], "${ID}_model$rank.pdb") or die "rename: $!";
    last if ++$rank > 5;
}
print "\nFinished [$program]: " . localtime() . "\n";
exit 0;

# restraint rows as the violation log prints them: the tbl row with its parentheses blanked (C3D:475, 497-500)
sub tbl_rows_for_log {
    my ($ri, $rj, $rd, $nr) = @_;
    my @i = unpack("l$nr", $ri); my @j = unpack("l$nr", $rj); my @d = unpack("d$nr", $rd);
    return map { sprintf("assign45  resid %3d and name ca   resid %3d and name ca  %.2f 0.00 0.00", $i[$_] + 1, $j[$_] + 1, $d[$_]) } 0 .. $nr - 1;
}

sub usage {
    my ($msg) = @_;
    print "ERROR!! $msg\n" if defined $msg;                # the reference prints this line even for -h (C3D:2532)
    print <<"EOT";
Chromosome3D (libchr3d / B200 engine) - chromosome structure from a Hi-C interaction-frequency matrix
by distance-geometry simulated annealing.

Usage:  $0 -i <IF matrix> -o <output directory> [options]
  -i   input interaction-frequency matrix, L rows of L space separated numbers   (-if is accepted too)
  -o   output directory (created; existing files in it are removed)
  -k   scaling constant K of D = K * mean(IF^a) / IF^a     [11]
  -a   alpha, exponent of the IF to distance conversion    [0.5]
  -m   number of models to build                           [20]
  -h   this message
  --seed <n>    annealing seed [82364]     --device <n>   CUDA device [0]
  --start coil|dg   starting coordinates: seeded random coil [default] or distance-geometry embedding (the DG stage of dgsa.inp)
  --pin         bead-level stand-ins for the angle terms and backbone atoms of CNS's poly-GLY chain (closer to the shipped CNS models)
EOT
    exit 1;
}
