#!/usr/bin/perl -w
# chromosome3D_sweep.pl — alpha sweep, model selection and the output_models/ naming (SURVEY 8(f) rank 2).
#
# The reference ships only the RESULT of this step: output_models/<chr>_<res>_rankNN_aXX.pdb (46 files, all "a11") and, for
# the 500 kb models, <...>_reduced.pdb; the script that swept alpha (chromosome3D.pl takes one alpha per run, C3D:32),
# picked a model and named it was never published.  This driver does that step on top of libchr3d:
#   * one chr3d_run_ensemble call holds every (alpha, model) trajectory of the sweep — the GPU batches them all;
#   * per alpha the models are ranked by int(NOE energy) exactly as assess_dgsa does (C3D:617, 802, 823);
#   * every model is scored by spearman_IF_pdb.pl's IF-vs-distance Spearman (SP:32-71, range 3);
#   * the selected model is the one with the most negative Spearman over the whole sweep; it is written as
#     <ID>_rankNN_aXX.pdb (NN = its energy rank inside its alpha's ensemble, XX = alpha*10, two digits; layout of the
#     shipped files: CA-only, CONECT rows) plus <ID>_rankNN_aXX_reduced.pdb (pair-averaged, SURVEY App. B.6);
#   * alpha_sweep.txt lists, per alpha, restraint count, best energy, best/median Spearman and the model picked.
use strict;
use warnings;
use Getopt::Long;
use File::Basename;
use Cwd qw(abs_path);
use FindBin;
use lib $FindBin::Bin;
use Chr3D;

my ($help, $dir_out, $K, $m, $file_if, $seed, $device, $alphas, $range, $keep_all, $gpus, $part) = (0, undef, 11, 20, undef, 0, 0, "0.1:2.0:0.1", 3, 0, 1, undef);
Getopt::Long::Configure("no_auto_abbrev");
GetOptions("h" => \$help, "o=s" => \$dir_out, "k=i" => \$K, "m=i" => \$m, "i|if=s" => \$file_if, "seed=i" => \$seed, "device=i" => \$device,
           "alphas=s" => \$alphas, "range=i" => \$range, "keep-all" => \$keep_all, "gpus=i" => \$gpus, "part=i" => \$part) or usage();
usage() if $help || !defined $file_if || !defined $dir_out || !-f $file_if || $m < 1;
my @alpha;
if ($alphas =~ /^([\d.]+):([\d.]+):([\d.]+)$/) { my $n = int(($2 - $1) / $3 + 0.5); @alpha = map { sprintf("%.1f", $1 + $3 * $_) + 0 } 0 .. $n; }
else { @alpha = map { $_ + 0 } split(/,/, $alphas); }
usage() if !@alpha;

system("mkdir", "-p", $dir_out) == 0 or die "cannot create $dir_out\n";
$dir_out = abs_path($dir_out); $file_if = abs_path($file_if);
my $ID = basename($file_if, ".txt"); $ID =~ s/_matrix$//;

# --gpus N: the alphas are dealt round-robin to N child processes, one per GPU (ensemble sharding, no collectives: the
# reference's own parallelism is test.sh backgrounding independent runs, TS:4-12); the parent merges their tables and picks
# the overall best model.
if ($gpus > 1 && !defined $part) {
    my @pids;
    for my $g (0 .. $gpus - 1) {
        my @mine = map { $alpha[$_] } grep { $_ % $gpus == $g } 0 .. $#alpha;
        next if !@mine;
        my $pid = fork(); die "fork: $!" if !defined $pid;
        if ($pid == 0) {
            my @cmd = ($^X, "-I", $FindBin::Bin, "$FindBin::Bin/" . basename($0), "-i", $file_if, "-o", $dir_out, "-k", $K, "-m", $m, "--seed", $seed,
                       "--device", $g, "--alphas", join(",", @mine), "--range", $range, "--part", $g, ($keep_all ? ("--keep-all") : ()));
            exec(@cmd) or die "exec: $!";
        }
        push @pids, $pid;
    }
    my $failed = 0;
    for my $pid (@pids) { waitpid($pid, 0); $failed ||= $? != 0; }
    die "a sweep worker failed\n" if $failed;
    my (@rows, $best);
    for my $g (0 .. $gpus - 1) {
        next if !-f "$dir_out/alpha_sweep.part$g.txt";
        open(my $fh, "<", "$dir_out/alpha_sweep.part$g.txt") or die $!;
        while (my $l = <$fh>) { push @rows, $l if $l !~ /^#/; }
        close $fh; unlink "$dir_out/alpha_sweep.part$g.txt";
        open($fh, "<", "$dir_out/best.part$g.txt") or die $!;
        my ($name, $rho) = split(/\t/, scalar <$fh>); chomp $rho; close $fh; unlink "$dir_out/best.part$g.txt";
        $best = { name => $name, rho => $rho } if !defined $best || $rho < $best->{rho};
    }
    @rows = sort { (split(/\t/, $a))[0] <=> (split(/\t/, $b))[0] } @rows;
    open(my $tab, ">", "$dir_out/alpha_sweep.txt") or die $!;
    print $tab "#alpha\trestraints\tbest_noe\tbest_srcc\tmedian_srcc\tpicked_model\tpicked_rank\n", @rows;
    close $tab;
    if (!$keep_all) {       # only the overall winner stays
        for my $f (glob("$dir_out/${ID}_rank*_a*.pdb")) { unlink $f if index($f, $best->{name}) < 0; }
    }
    printf "Selected   : %s.pdb  (SRCC %.3f, %d GPUs)\n", $best->{name}, $best->{rho}, $gpus;
    exit 0;
}
Chr3D::init($device);
my ($IF, $L) = Chr3D::read_if_text($file_if);
print "Input      : $file_if (L = $L)\nAlphas     : @alpha\nModels     : $m per alpha, " . ($m * @alpha) . " trajectories in one call\n";

my @jobs = map { { IF => $IF, L => $L, alpha => $_, K => $K + 0, min_sep => 5, n_models => $m, seed => $seed } } @alpha;
my $t0 = time;
my @res = Chr3D::run_ensembles(\@jobs);
printf "Annealed   : %d trajectories (%s, binding %s), device time %.1f ms\n", $m * @alpha, Chr3D::version(), Chr3D::backend(), $res[0]{device_ms};

my $MET = "MET" x $L;      # the shipped rankNN_aXX.pdb files name every residue MET
open(my $tab, ">", defined $part ? "$dir_out/alpha_sweep.part$part.txt" : "$dir_out/alpha_sweep.txt") or die $!;
print $tab "#alpha\trestraints\tbest_noe\tbest_srcc\tmedian_srcc\tpicked_model\tpicked_rank\n";
my ($best, $xyz_bytes) = (undef, 8 * 3 * $L);
for my $ai (0 .. $#alpha) {
    my $r = $res[$ai];
    my @order = @{ Chr3D::rank_models($r->{energies}, $m) };        # best first
    my %rank_of; $rank_of{ $order[$_] } = $_ + 1 for 0 .. $#order;
    my @rho = @{ Chr3D::spearman($IF, $L, $r->{xyz}, $m, $range) };
    my @e = unpack("d" . (4 * $m), $r->{energies});
    my $pick = 0; for my $k (1 .. $m - 1) { $pick = $k if $rho[$k] < $rho[$pick]; }      # most negative SRCC, first wins ties
    my @sorted = sort { $a <=> $b } @rho;
    my $median = $m % 2 ? $sorted[($m - 1) / 2] : 0.5 * ($sorted[$m / 2 - 1] + $sorted[$m / 2]);
    printf $tab "%.1f\t%d\t%d\t%.3f\t%.3f\t%d\t%d\n", $alpha[$ai], $r->{n_restraints}, int($e[4 * $order[0] + 3]), $rho[$pick], $median, $pick + 1, $rank_of{$pick};
    if ($keep_all) {
        my $name = sprintf("%s_rank%02d_a%02d", $ID, $rank_of{$pick}, int($alpha[$ai] * 10 + 0.5));
        Chr3D::write_pdb("$dir_out/$name.pdb", substr($r->{xyz}, $pick * $xyz_bytes, $xyz_bytes), $L, undef, 1, $MET);
    }
    $best = { a => $ai, k => $pick, rho => $rho[$pick], rank => $rank_of{$pick} } if !defined $best || $rho[$pick] < $best->{rho};
}
close $tab;
my $name = sprintf("%s_rank%02d_a%02d", $ID, $best->{rank}, int($alpha[$best->{a}] * 10 + 0.5));
my $xyz = substr($res[$best->{a}]{xyz}, $best->{k} * $xyz_bytes, $xyz_bytes);
Chr3D::write_pdb("$dir_out/$name.pdb", $xyz, $L, undef, 1, $MET);
my ($red, $nred) = Chr3D::reduce_model($xyz, $L);
Chr3D::write_reduced_pdb("$dir_out/${name}_reduced.pdb", $red, $nred);
if (defined $part) { open(my $bf, ">", "$dir_out/best.part$part.txt") or die $!; print $bf "$name\t$best->{rho}\n"; close $bf; }
printf "%s : %s.pdb  (alpha %.1f, model %d, energy rank %d, SRCC %.3f)  [%d s]\n", (defined $part ? "Part $part best" : "Selected  "), $name, $alpha[$best->{a}], $best->{k} + 1, $best->{rank}, $best->{rho}, time - $t0;
exit 0;

sub usage {
    print <<"EOT";
Usage: $0 -i <IF matrix> -o <output directory> [-k 11] [-m 20] [--alphas 0.1:2.0:0.1 | --alphas 0.5,1.1] [--range 3] [--keep-all] [--gpus N]
  sweeps alpha, anneals -m models per alpha in one GPU batch, selects the model with the best IF-distance Spearman and
  writes <ID>_rankNN_aXX.pdb, <ID>_rankNN_aXX_reduced.pdb and alpha_sweep.txt into the output directory
EOT
    exit 1;
}
