#!/usr/bin/perl -w
# spearman_IF_pdb.pl — IF-vs-model-distance Spearman scoring through libchr3d.so.
#
# Same command line and output as the reference script (/root/reference/spearman_IF_pdb.pl, SP:N):
#   spearman_IF_pdb.pl <IF matrix> <pdb file or directory of *.pdb> [range=3]
# prints "SRCC\tPDB" and one "%.3f\t<pdb>" row per model, best (largest) first (SP:73-76).  The un-vendored
# modules the reference needs (confold, Statistics::RankCorrelation, SP:10-13) are not used: pair selection,
# "%.3f" distances, average ranks and the correlation are computed on the GPU in exact integer arithmetic.
use strict;
use warnings;
use FindBin;
use lib $FindBin::Bin;
use Chr3D;

my ($mat, $pdb, $range) = @ARGV;
if (!defined $mat || !defined $pdb) { print "Usage: $0 <IF matrix> <pdb|dir> [range=3]\n"; exit 1; }
$range = 3 if !defined $range;                                  # SP:20-22
die "matrix $mat does not exist\n" if !-f $mat;
my @list = -d $pdb ? sort glob("$pdb/*.pdb") : ($pdb);          # SP:24-31
die "no pdb files under $pdb\n" if !@list;

Chr3D::init(0);
my ($IF, $L) = Chr3D::read_if_matrix($mat);
my %srcc;
for my $f (@list) {
    my ($xyz, $n) = Chr3D::read_ca_pdb($f);
    if ($range >= $n) { print "Spearman Correlation coefficient = -\n"; exit 0; }      # SP:38-41
    die "$f has $n CA atoms, the matrix has $L rows\n" if $n != $L;                    # size check, SP:64
    my $rho = Chr3D::spearman($IF, $L, $xyz, 1, $range)->[0];
    $srcc{$f} = sprintf("%.3f", $rho);                                                  # SP:67
}
print "SRCC\tPDB\n";
for my $f (sort { $srcc{$b} <=> $srcc{$a} or $a cmp $b } keys %srcc) { print "$srcc{$f}\t$f\n"; }
