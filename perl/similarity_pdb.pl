#!/usr/bin/perl -w
# similarity_pdb.pl — the scorer behind output_models/similarity.txt (SURVEY App. B.6; its program is not in the reference tree).
#   similarity_pdb.pl <1mb model.pdb> <500kb model.pdb>     the 500 kb model is pair-averaged first (unrounded means)
#   similarity_pdb.pl --reduce <model.pdb> <out_reduced.pdb> only writes the pair-averaged model in the shipped dialect
# Prints one block in the shipped format: name, "Spearman correlation: v", "RMSD: v", empty line.
use strict;
use warnings;
use File::Basename;
use FindBin;
use lib $FindBin::Bin;
use Chr3D;

Chr3D::init(0);
if (@ARGV == 3 && $ARGV[0] eq "--reduce") {
    my ($x, $n) = Chr3D::read_ca_pdb($ARGV[1]);
    my ($r, $m) = Chr3D::reduce_model($x, $n);
    Chr3D::write_reduced_pdb($ARGV[2], $r, $m);
    exit 0;
}
die "usage: $0 <1mb model.pdb> <500kb model.pdb> | --reduce <model.pdb> <out_reduced.pdb>\n" if @ARGV != 2;
my ($xa, $na) = Chr3D::read_ca_pdb($ARGV[0]);
my ($xh, $nh) = Chr3D::read_ca_pdb($ARGV[1]);
my ($xb, $nb) = Chr3D::reduce_model($xh, $nh);
my ($rho, $rmsd) = Chr3D::model_similarity($xa, $na, $xb, $nb);
print basename($ARGV[1], ".pdb"), "\n", "Spearman correlation: ", Chr3D::shortest($rho), "\n", "RMSD: ", Chr3D::shortest($rmsd), "\n\n";
