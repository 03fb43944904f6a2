/*
 * Chr3D.xs — XS shim binding the C ABI of libchr3d.so (include/chr3d.h) into Perl.
 *
 * Used by perl/Chr3D.pm when FFI::Platypus is not installed.  It adds no logic: every XSUB is the
 * C function of the same name with pointer arguments passed as integers (addresses of Perl string
 * buffers obtained with Chr3D::XS::ptr), exactly what FFI::Platypus' `opaque` type passes.
 * Replaces the process boundary `cns_solve < dgsa.inp` of chromosome3D.pl (C3D:254-289).
 */
#define PERL_NO_GET_CONTEXT
#include "EXTERN.h"
#include "perl.h"
#include "XSUB.h"
#include <stdint.h>
#include "chr3d.h"

#define P(t, v) ((t)(uintptr_t)(v))

MODULE = Chr3D::XS    PACKAGE = Chr3D::XS

PROTOTYPES: DISABLE

UV
ptr(sv)
    SV* sv
  CODE:
    /* address of the string buffer of a Perl scalar (the scalar must stay alive and unmodified) */
    if (!SvPOK(sv)) croak("Chr3D::XS::ptr: not a string buffer");
    RETVAL = (UV)(uintptr_t)SvPVX(sv);
  OUTPUT:
    RETVAL

const char*
chr3d_version()

const char*
chr3d_last_error()

int
chr3d_device_count()

int
chr3d_init(device)
    int device

int
chr3d_shutdown()

void
chr3d_default_params(p)
    UV p
  CODE:
    chr3d_default_params(P(chr3d_sa_params*, p));

void
chr3d_cns_pin_params(p)
    UV p
  CODE:
    chr3d_cns_pin_params(P(chr3d_sa_params*, p));

int
chr3d_build_program(p, out, cap)
    UV p
    UV out
    int cap
  CODE:
    RETVAL = chr3d_build_program(P(const chr3d_sa_params*, p), P(chr3d_stage*, out), cap);
  OUTPUT:
    RETVAL

int
chr3d_if_to_dist(IF, L, alpha, K, tenths, n_patched)
    UV IF
    int L
    double alpha
    double K
    UV tenths
    UV n_patched
  CODE:
    RETVAL = chr3d_if_to_dist(P(const double*, IF), L, alpha, K, P(int64_t*, tenths), P(int64_t*, n_patched));
  OUTPUT:
    RETVAL

int
chr3d_if_to_restraints(IF, L, alpha, K, min_sep, out_i, out_j, out_d, cap, out_n)
    UV IF
    int L
    double alpha
    double K
    int min_sep
    UV out_i
    UV out_j
    UV out_d
    IV cap
    UV out_n
  CODE:
    RETVAL = chr3d_if_to_restraints(P(const double*, IF), L, alpha, K, min_sep, P(int32_t*, out_i), P(int32_t*, out_j),
                                    P(double*, out_d), (int64_t)cap, P(int64_t*, out_n));
  OUTPUT:
    RETVAL

int
chr3d_write_dist(path, tenths, L)
    const char* path
    UV tenths
    int L
  CODE:
    RETVAL = chr3d_write_dist(path, P(const int64_t*, tenths), L);
  OUTPUT:
    RETVAL

int
chr3d_write_rr(path, ri, rj, rd, n)
    const char* path
    UV ri
    UV rj
    UV rd
    IV n
  CODE:
    RETVAL = chr3d_write_rr(path, P(const int32_t*, ri), P(const int32_t*, rj), P(const double*, rd), (int64_t)n);
  OUTPUT:
    RETVAL

int
chr3d_write_tbl(path, ri, rj, rd, n)
    const char* path
    UV ri
    UV rj
    UV rd
    IV n
  CODE:
    RETVAL = chr3d_write_tbl(path, P(const int32_t*, ri), P(const int32_t*, rj), P(const double*, rd), (int64_t)n);
  OUTPUT:
    RETVAL

int
chr3d_run_ensemble(jobs, njobs, p, program, nstages, results)
    UV jobs
    int njobs
    UV p
    UV program
    int nstages
    UV results
  CODE:
    RETVAL = chr3d_run_ensemble(P(const chr3d_job*, jobs), njobs, P(const chr3d_sa_params*, p), P(const chr3d_stage*, program),
                                nstages, P(chr3d_result*, results));
  OUTPUT:
    RETVAL

int
chr3d_assess(xyz, n_models, N, ri, rj, rd, nr, satisfied, sum_dev)
    UV xyz
    int n_models
    int N
    UV ri
    UV rj
    UV rd
    IV nr
    UV satisfied
    UV sum_dev
  CODE:
    RETVAL = chr3d_assess(P(const double*, xyz), n_models, N, P(const int32_t*, ri), P(const int32_t*, rj), P(const double*, rd),
                          (int64_t)nr, P(int32_t*, satisfied), P(double*, sum_dev));
  OUTPUT:
    RETVAL

int
chr3d_assess_rows(xyz, N, ri, rj, rd, nr, flag, dev, dist)
    UV xyz
    int N
    UV ri
    UV rj
    UV rd
    IV nr
    UV flag
    UV dev
    UV dist
  CODE:
    RETVAL = chr3d_assess_rows(P(const double*, xyz), N, P(const int32_t*, ri), P(const int32_t*, rj), P(const double*, rd),
                               (int64_t)nr, P(int32_t*, flag), P(double*, dev), P(double*, dist));
  OUTPUT:
    RETVAL

int
chr3d_rank_models(energies, n_models, order)
    UV energies
    int n_models
    UV order
  CODE:
    RETVAL = chr3d_rank_models(P(const double*, energies), n_models, P(int32_t*, order));
  OUTPUT:
    RETVAL

int
chr3d_write_pdb(path, xyz, N, energies4, resnames, with_conect)
    const char* path
    UV xyz
    int N
    UV energies4
    UV resnames
    int with_conect
  CODE:
    RETVAL = chr3d_write_pdb(path, P(const double*, xyz), N, P(const double*, energies4), P(const char*, resnames), with_conect);
  OUTPUT:
    RETVAL

int
chr3d_spearman(IF, L, xyz, n_models, range, rho)
    UV IF
    int L
    UV xyz
    int n_models
    int range
    UV rho
  CODE:
    RETVAL = chr3d_spearman(P(const double*, IF), L, P(const double*, xyz), n_models, range, P(double*, rho));
  OUTPUT:
    RETVAL

int
chr3d_reduce_model(xyz, N, out_xyz)
    UV xyz
    int N
    UV out_xyz
  CODE:
    RETVAL = chr3d_reduce_model(P(const double*, xyz), N, P(double*, out_xyz));
  OUTPUT:
    RETVAL

int
chr3d_write_reduced_pdb(path, xyz, n)
    const char* path
    UV xyz
    int n
  CODE:
    RETVAL = chr3d_write_reduced_pdb(path, P(const double*, xyz), n);
  OUTPUT:
    RETVAL

int
chr3d_model_similarity(xyz_a, n_a, xyz_b, n_b, spearman, rmsd)
    UV xyz_a
    int n_a
    UV xyz_b
    int n_b
    UV spearman
    UV rmsd
  CODE:
    RETVAL = chr3d_model_similarity(P(const double*, xyz_a), n_a, P(const double*, xyz_b), n_b, P(double*, spearman), P(double*, rmsd));
  OUTPUT:
    RETVAL

int
chr3d_read_if_text(path, out, cap, L)
    const char* path
    UV out
    IV cap
    UV L
  CODE:
    RETVAL = chr3d_read_if_text(path, P(double*, out), (int64_t)cap, P(int32_t*, L));
  OUTPUT:
    RETVAL

int
chr3d_write_if_binary(path, IF, L, elem_bytes)
    const char* path
    UV IF
    int L
    int elem_bytes
  CODE:
    RETVAL = chr3d_write_if_binary(path, P(const double*, IF), L, elem_bytes);
  OUTPUT:
    RETVAL

int
chr3d_read_if_binary(path, out, cap, L)
    const char* path
    UV out
    IV cap
    UV L
  CODE:
    RETVAL = chr3d_read_if_binary(path, P(double*, out), (int64_t)cap, P(int32_t*, L));
  OUTPUT:
    RETVAL

int
chr3d_dg_embed(ri, rj, rd, nr, N, p, n_models, seed, xyz_out, eig3_out, n_inconsistent)
    UV ri
    UV rj
    UV rd
    IV nr
    int N
    UV p
    int n_models
    UV seed
    UV xyz_out
    UV eig3_out
    UV n_inconsistent
  CODE:
    RETVAL = chr3d_dg_embed(P(const int32_t*, ri), P(const int32_t*, rj), P(const double*, rd), (int64_t)nr, N, P(const chr3d_sa_params*, p),
                            n_models, (uint64_t)seed, P(double*, xyz_out), P(double*, eig3_out), P(int64_t*, n_inconsistent));
  OUTPUT:
    RETVAL

int
chr3d_dg_starts(ri, rj, rd, nr, N, p, n_models, seed, dg_scale, nstep, drop, xyz_out, eig3_out, stats_out, n_inconsistent)
    UV ri
    UV rj
    UV rd
    IV nr
    int N
    UV p
    int n_models
    UV seed
    double dg_scale
    int nstep
    double drop
    UV xyz_out
    UV eig3_out
    UV stats_out
    UV n_inconsistent
  CODE:
    RETVAL = chr3d_dg_starts(P(const int32_t*, ri), P(const int32_t*, rj), P(const double*, rd), (int64_t)nr, N, P(const chr3d_sa_params*, p),
                             n_models, (uint64_t)seed, dg_scale, nstep, drop, P(double*, xyz_out), P(double*, eig3_out), P(double*, stats_out),
                             P(int64_t*, n_inconsistent));
  OUTPUT:
    RETVAL
