package Chr3D;
# Chr3D.pm — Perl binding of libchr3d.so (include/chr3d.h), the in-process replacement for the
# `cns_solve < dgsa.inp` run that chromosome3D.pl's build_models() shells out to (C3D:254-289).
#
# Backend: FFI::Platypus when it is installed (what BASELINE.json:north_star asks for), otherwise the
# XS shim perl/xs/Chr3D.xs (built by perl/xs/build.sh).  Both bind the same C symbols; pointers are
# addresses of packed Perl string buffers, the caller allocates every buffer (as the ABI requires).
use strict;
use warnings;
use File::Basename qw(dirname);
use Cwd qw(abs_path);
use Carp qw(confess);

our $HERE    = dirname(abs_path(__FILE__));
our $LIBPATH = $ENV{CHR3D_LIB} || abs_path("$HERE/../chromosome3d_b200") . "/libchr3d.so";
our $BACKEND;
my $ffi;

my %SIG = (    # name => [ret, args] in FFI::Platypus type names (pointers = opaque)
    chr3d_version          => ['string', []],
    chr3d_last_error       => ['string', []],
    chr3d_device_count     => ['int',    []],
    chr3d_init             => ['int',    ['int']],
    chr3d_shutdown         => ['int',    []],
    chr3d_default_params   => ['void',   ['opaque']],
    chr3d_cns_pin_params   => ['void',   ['opaque']],
    chr3d_build_program    => ['int',    ['opaque', 'opaque', 'int']],
    chr3d_if_to_dist       => ['int',    ['opaque', 'sint32', 'double', 'double', 'opaque', 'opaque']],
    chr3d_if_to_restraints => ['int',    ['opaque', 'sint32', 'double', 'double', 'sint32', 'opaque', 'opaque', 'opaque', 'sint64', 'opaque']],
    chr3d_write_dist       => ['int',    ['string', 'opaque', 'sint32']],
    chr3d_write_rr         => ['int',    ['string', 'opaque', 'opaque', 'opaque', 'sint64']],
    chr3d_write_tbl        => ['int',    ['string', 'opaque', 'opaque', 'opaque', 'sint64']],
    chr3d_run_ensemble     => ['int',    ['opaque', 'sint32', 'opaque', 'opaque', 'sint32', 'opaque']],
    chr3d_assess           => ['int',    ['opaque', 'sint32', 'sint32', 'opaque', 'opaque', 'opaque', 'sint64', 'opaque', 'opaque']],
    chr3d_assess_rows      => ['int',    ['opaque', 'sint32', 'opaque', 'opaque', 'opaque', 'sint64', 'opaque', 'opaque', 'opaque']],
    chr3d_rank_models      => ['int',    ['opaque', 'sint32', 'opaque']],
    chr3d_write_pdb        => ['int',    ['string', 'opaque', 'sint32', 'opaque', 'opaque', 'sint32']],
    chr3d_spearman         => ['int',    ['opaque', 'sint32', 'opaque', 'sint32', 'sint32', 'opaque']],
    chr3d_dg_embed         => ['int',    ['opaque', 'opaque', 'opaque', 'sint64', 'sint32', 'opaque', 'sint32', 'uint64', 'opaque', 'opaque', 'opaque']],
    chr3d_dg_starts        => ['int',    ['opaque', 'opaque', 'opaque', 'sint64', 'sint32', 'opaque', 'sint32', 'uint64', 'double', 'sint32', 'double', 'opaque', 'opaque', 'opaque', 'opaque']],
    chr3d_reduce_model     => ['int',    ['opaque', 'sint32', 'opaque']],
    chr3d_write_reduced_pdb => ['int',   ['string', 'opaque', 'sint32']],
    chr3d_model_similarity => ['int',    ['opaque', 'sint32', 'opaque', 'sint32', 'opaque', 'opaque']],
    chr3d_read_if_text     => ['int',    ['string', 'opaque', 'sint64', 'opaque']],
    chr3d_write_if_binary  => ['int',    ['string', 'opaque', 'sint32', 'sint32']],
    chr3d_read_if_binary   => ['int',    ['string', 'opaque', 'sint64', 'opaque']],
);

sub _load {
    return if $BACKEND;
    confess "libchr3d.so not found at $LIBPATH (build it with `make -C chromosome3d_b200/csrc`); there is no CPU path" if not -f $LIBPATH;
    if (!$ENV{CHR3D_FORCE_XS} && eval { require FFI::Platypus; require FFI::Platypus::Buffer; 1 }) {
        $ffi = FFI::Platypus->new(api => 1);
        $ffi->lib($LIBPATH);
        no strict 'refs';
        for my $name (keys %SIG) {
            my $f = $ffi->function($name => $SIG{$name}[1] => $SIG{$name}[0]);
            *{"Chr3D::_c_$name"} = sub { $f->call(@_) };
        }
        *Chr3D::ptr = sub { my ($p, $len) = FFI::Platypus::Buffer::scalar_to_buffer($_[0]); $p };
        $BACKEND = "FFI::Platypus";
        return;
    }
    my $so = "$HERE/xs/auto/Chr3D/XS/XS.so";
    if (not -f $so) {
        system("bash", "$HERE/xs/build.sh") == 0 or confess "cannot build the XS shim (perl/xs/build.sh) and FFI::Platypus is not installed";
    }
    require DynaLoader;
    my $lib = DynaLoader::dl_load_file($so, 0) or confess "dl_load_file($so): " . DynaLoader::dl_error();
    my $boot = DynaLoader::dl_find_symbol($lib, "boot_Chr3D__XS") or confess "boot_Chr3D__XS not found in $so";
    my $xs = DynaLoader::dl_install_xsub("Chr3D::XS::bootstrap", $boot, $so);
    &$xs("Chr3D::XS");
    no strict 'refs';
    for my $name (keys %SIG) { *{"Chr3D::_c_$name"} = \&{"Chr3D::XS::$name"}; }
    *Chr3D::ptr = \&Chr3D::XS::ptr;
    $BACKEND = "XS";
}

sub _check {
    my ($rc, $what) = @_;
    confess "libchr3d: $what failed ($rc): " . _c_chr3d_last_error() if $rc != 0;
}
sub _buf { my $n = shift; my $b = "\0" x $n; return $b; }
sub _p { return defined $_[0] ? ptr($_[0]) : 0; }

sub backend      { _load(); return $BACKEND; }
sub version      { _load(); return _c_chr3d_version(); }
sub device_count { _load(); return _c_chr3d_device_count(); }
sub init         { _load(); _check(_c_chr3d_init($_[0] // 0), "chr3d_init"); }

# chr3d_sa_params is kept as an opaque 240-byte buffer; the few fields the front end sets are patched by offset.
use constant PARAMS_BYTES => 240;
use constant STAGE_BYTES  => 72;
sub default_params { _load(); my $p = _buf(PARAMS_BYTES); _c_chr3d_default_params(ptr($p)); return $p; }
# the defaults plus the default-off stand-ins for CNS's poly-GLY chain terms (chr3d_cns_pin_params, DESIGN.md section 2)
sub cns_pin_params { _load(); my $p = _buf(PARAMS_BYTES); _c_chr3d_cns_pin_params(ptr($p)); return $p; }
sub build_program {
    my ($p) = @_;
    my $n = _c_chr3d_build_program(ptr($p), 0, 0);
    confess "chr3d_build_program: " . _c_chr3d_last_error() if $n < 0;
    my $prog = _buf($n * STAGE_BYTES);
    _c_chr3d_build_program(ptr($p), ptr($prog), $n);
    return ($prog, $n);
}

# IF2dist_new (C3D:110-162): returns the packed int64 tenths of the `%.1f` text
sub if_to_dist {
    my ($IF, $L, $alpha, $K) = @_;
    _load();
    my $t = _buf(8 * $L * $L); my $np = _buf(8);
    _check(_c_chr3d_if_to_dist(ptr($IF), $L, $alpha, $K, ptr($t), ptr($np)), "chr3d_if_to_dist");
    return $t;
}
# dist2rr + carr2tbl (C3D:181-206, 340-362): packed (int32 i, int32 j, double d) columns in C3D:203 order
sub if_to_restraints {
    my ($IF, $L, $alpha, $K, $min_sep) = @_;
    _load();
    my $n = _buf(8);
    _check(_c_chr3d_if_to_restraints(ptr($IF), $L, $alpha, $K, $min_sep, 0, 0, 0, 0, ptr($n)), "chr3d_if_to_restraints");
    my $nr = unpack("q", $n); my $cap = $nr > 0 ? $nr : 1;
    my ($ri, $rj, $rd) = (_buf(4 * $cap), _buf(4 * $cap), _buf(8 * $cap));
    _check(_c_chr3d_if_to_restraints(ptr($IF), $L, $alpha, $K, $min_sep, ptr($ri), ptr($rj), ptr($rd), $cap, ptr($n)), "chr3d_if_to_restraints") if $nr > 0;
    return ($ri, $rj, $rd, $nr);
}
sub write_dist { _check(_c_chr3d_write_dist($_[0], ptr($_[1]), $_[2]), "chr3d_write_dist"); }
sub write_rr   { _check(_c_chr3d_write_rr($_[0], ptr($_[1]), ptr($_[2]), ptr($_[3]), $_[4]), "chr3d_write_rr"); }
sub write_tbl  { _check(_c_chr3d_write_tbl($_[0], ptr($_[1]), ptr($_[2]), ptr($_[3]), $_[4]), "chr3d_write_tbl"); }

# build_models (C3D:254-289): anneal m models of one matrix.  Returns a hash of packed buffers.
# One call for any number of (matrix, alpha) jobs: the library batches every trajectory of every job into one launch.
# jobs: list of hashes { IF, L, alpha, K, min_sep, n_models, seed }; returns a list of result hashes.
sub run_ensembles {
    my ($jobs, $params) = @_;
    _load();
    my $p = $params // default_params();
    my ($prog, $nst) = build_program($p);
    my (@r, $jobbuf, $resbuf);
    ($jobbuf, $resbuf) = ("", "");
    for my $a (@$jobs) {
        my ($L, $m) = ($a->{L}, $a->{n_models});
        my %r = (xyz => _buf(8 * $m * $L * 3), energies => _buf(8 * 4 * $m), force_evals => _buf(8 * $m), hand => _buf(4 * $m),
                 satisfied => _buf(4 * $m), sum_dev => _buf(8 * $m));
        push @r, \%r;
        # struct chr3d_job / chr3d_result (include/chr3d.h), x86-64 layout
        # x0 (optional): packed doubles n_models*L*3, e.g. from dg_embed(); the scalar must stay alive until the call returns
        $jobbuf .= pack("Q l l d d Q Q Q q l l Q Q", ptr($a->{IF}), $L, $a->{min_sep} // 5, $a->{alpha}, $a->{K} // 11, 0, 0, 0, 0, $m, 0, $a->{seed} // 0,
                        defined $a->{x0} ? ptr($a->{x0}) : 0);
        $resbuf .= pack("Q Q Q Q Q Q q d", ptr($r{xyz}), ptr($r{energies}), ptr($r{force_evals}), ptr($r{hand}), ptr($r{satisfied}), ptr($r{sum_dev}), 0, 0.0);
    }
    _check(_c_chr3d_run_ensemble(ptr($jobbuf), scalar(@$jobs), ptr($p), ptr($prog), $nst, ptr($resbuf)), "chr3d_run_ensemble");
    my $rsz = length($resbuf) / @$jobs;
    for my $k (0 .. $#r) {
        my @tail = unpack("Q6 q d", substr($resbuf, $k * $rsz, $rsz));
        $r[$k]{n_restraints} = $tail[6]; $r[$k]{device_ms} = $tail[7];
    }
    return @r;
}
sub run_ensemble {
    my (%a) = @_;    # IF, L, alpha, K, min_sep, n_models, seed, params (opaque, optional)
    my ($r) = run_ensembles([\%a], $a{params});
    return $r;
}
sub assess_rows {
    my ($xyz_model, $N, $ri, $rj, $rd, $nr) = @_;
    _load();
    my $cap = $nr > 0 ? $nr : 1;
    my ($flag, $dev, $dist) = (_buf(4 * $cap), _buf(8 * $cap), _buf(8 * $cap));
    _check(_c_chr3d_assess_rows(ptr($xyz_model), $N, ptr($ri), ptr($rj), ptr($rd), $nr, ptr($flag), ptr($dev), ptr($dist)), "chr3d_assess_rows");
    return ($flag, $dev, $dist);
}
sub assess {
    my ($xyz, $m, $N, $ri, $rj, $rd, $nr) = @_;
    _load();
    my ($sat, $dev) = (_buf(4 * $m), _buf(8 * $m));
    _check(_c_chr3d_assess(ptr($xyz), $m, $N, ptr($ri), ptr($rj), ptr($rd), $nr, ptr($sat), ptr($dev)), "chr3d_assess");
    return ([unpack("l$m", $sat)], [unpack("d$m", $dev)]);
}
sub rank_models {
    my ($energies, $m) = @_;
    _load();
    my $o = _buf(4 * $m);
    _check(_c_chr3d_rank_models(ptr($energies), $m, ptr($o)), "chr3d_rank_models");
    return [unpack("l$m", $o)];
}
sub write_pdb {
    my ($path, $xyz_model, $N, $e4, $with_conect, $resnames) = @_;     # resnames: 3*N characters, default = the reference sequence (C3D:93-98)
    _load();
    _check(_c_chr3d_write_pdb($path, ptr($xyz_model), $N, defined $e4 ? ptr($e4) : 0, defined $resnames ? ptr($resnames) : 0, $with_conect ? 1 : 0), "chr3d_write_pdb");
}
sub spearman {
    my ($IF, $L, $xyz, $m, $range) = @_;
    _load();
    my $rho = _buf(8 * $m);
    _check(_c_chr3d_spearman(ptr($IF), $L, ptr($xyz), $m, $range, ptr($rho)), "chr3d_spearman");
    return [unpack("d$m", $rho)];
}

# ---- SURVEY 8(f) "next" rows: DG start, cross-resolution tools and matrix containers ------------
# DG stage of dgsa.inp (C3D:1470-1526) at bead level: (packed xyz of n_models*N*3 doubles, inconsistent-pair count)
sub dg_embed {
    my ($ri, $rj, $rd, $nr, $N, $m, $seed, $params) = @_;
    _load();
    my $p = $params // default_params();
    my ($xyz, $eig, $bad) = (_buf(8 * 3 * $N * $m), _buf(8 * 3 * $m), _buf(8));
    _check(_c_chr3d_dg_embed(ptr($ri), ptr($rj), ptr($rd), $nr, $N, ptr($p), $m, $seed // 0, ptr($xyz), ptr($eig), ptr($bad)), "chr3d_dg_embed");
    return ($xyz, unpack("q", $bad));
}
# The whole DG stage (C3D:1470-1526): embedding, x *= $dgscale, md.dg.step = 200 L-BFGS iterations (drop = 1.) on the bounds energy with
# md.dg.scale = 100.  Returns (packed xyz, inconsistent pairs, [dgscale, E scaled, E final, iterations, evaluations, rms gradient] of model 1)
sub dg_starts {
    my ($ri, $rj, $rd, $nr, $N, $m, $seed, $params) = @_;
    _load();
    my $p = $params // default_params();
    my ($xyz, $eig, $st, $bad) = (_buf(8 * 3 * $N * $m), _buf(8 * 3 * $m), _buf(8 * 6 * $m), _buf(8));
    _check(_c_chr3d_dg_starts(ptr($ri), ptr($rj), ptr($rd), $nr, $N, ptr($p), $m, $seed // 0, 100.0, 200, 1.0, ptr($xyz), ptr($eig), ptr($st), ptr($bad)), "chr3d_dg_starts");
    return ($xyz, unpack("q", $bad), [unpack("d6", $st)]);
}
sub reduce_model {      # (packed xyz, N) -> (packed xyz of ceil(N/2) beads, n)
    my ($xyz, $N) = @_;
    _load();
    my $m = int(($N + 1) / 2);
    my $out = _buf(8 * 3 * $m);
    _check(_c_chr3d_reduce_model(ptr($xyz), $N, ptr($out)), "chr3d_reduce_model");
    return ($out, $m);
}
sub write_reduced_pdb { _load(); _check(_c_chr3d_write_reduced_pdb($_[0], ptr($_[1]), $_[2]), "chr3d_write_reduced_pdb"); }
sub model_similarity {  # (xyz_a, n_a, xyz_b, n_b) -> (spearman, rmsd)
    my ($xa, $na, $xb, $nb) = @_;
    _load();
    my ($rho, $rmsd) = (_buf(8), _buf(8));
    _check(_c_chr3d_model_similarity(ptr($xa), $na, ptr($xb), $nb, ptr($rho), ptr($rmsd)), "chr3d_model_similarity");
    return (unpack("d", $rho), unpack("d", $rmsd));
}
sub read_if_text {      # the library's threaded parser of input/<ID>.txt: (packed doubles, L)
    my ($path) = @_;
    _load();
    my $Lb = _buf(4);
    _check(_c_chr3d_read_if_text($path, 0, 0, ptr($Lb)), "chr3d_read_if_text");
    my $L = unpack("l", $Lb);
    my $buf = _buf(8 * $L * $L);
    _check(_c_chr3d_read_if_text($path, ptr($buf), $L * $L, ptr($Lb)), "chr3d_read_if_text");
    return ($buf, $L);
}
sub write_if_binary { _load(); _check(_c_chr3d_write_if_binary($_[0], ptr($_[1]), $_[2], $_[3] // 8), "chr3d_write_if_binary"); }
sub read_if_binary {
    my ($path) = @_;
    _load();
    my $Lb = _buf(4);
    _check(_c_chr3d_read_if_binary($path, 0, 0, ptr($Lb)), "chr3d_read_if_binary");
    my $L = unpack("l", $Lb);
    my $buf = _buf(8 * $L * $L);
    _check(_c_chr3d_read_if_binary($path, ptr($buf), $L * $L, ptr($Lb)), "chr3d_read_if_binary");
    return ($buf, $L);
}
# shortest decimal text that reads back as the same double (what the similarity.txt writer printed)
sub shortest { my $v = shift; for my $p (15, 16, 17) { my $s = sprintf("%.${p}g", $v); return $s if $s + 0 == $v; } return sprintf("%.17g", $v); }

# IF matrix text -> (packed doubles, L).  L = token count of the first row (calc_len_IF, C3D:164-179); rows are split on
# whitespace, which also eats the CR of the CRLF line ends (C3D:119-122).
sub read_if_matrix {
    my ($path) = @_;
    open(my $fh, "<", $path) or confess "cannot open $path: $!";
    my (@rows, $L);
    while (my $line = <$fh>) {
        $line =~ s/^\s+//;
        my @t = split(/\s+/, $line);
        next if !@t;
        $L //= scalar(@t);
        push @rows, \@t;
    }
    close $fh;
    confess "empty IF matrix $path" if !$L;
    confess "IF matrix $path has " . scalar(@rows) . " rows for $L columns" if @rows < $L;
    my $buf = "";
    for my $i (0 .. $L - 1) {
        my @r = @{ $rows[$i] };
        push @r, (0) x ($L - @r) if @r < $L;
        $buf .= pack("d$L", @r[0 .. $L - 1]);
    }
    return ($buf, $L);
}

# CA coordinates of a PDB as its 3-decimal text fields give them (columns of C3D:674-691), packed doubles
sub read_ca_pdb {
    my ($path) = @_;
    open(my $fh, "<", $path) or confess "cannot open $path: $!";
    my ($buf, $n) = ("", 0);
    while (my $l = <$fh>) {
        next if $l !~ /^ATOM/;
        my $name = substr($l, 12, 4); $name =~ s/\s+//g;
        next if $name ne "CA";
        $buf .= pack("d3", substr($l, 30, 8) + 0, substr($l, 38, 8) + 0, substr($l, 46, 8) + 0);
        $n++;
    }
    close $fh;
    return ($buf, $n);
}

1;
