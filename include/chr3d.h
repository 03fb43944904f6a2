/*
 * chr3d.h — C ABI of libchr3d.so, the B200-native replacement for the CNS `dgsa.inp`
 * distance-geometry simulated-annealing run that chromosome3D.pl drives.
 *
 * Citation shorthand:  C3D:N = /root/reference/chromosome3D.pl line N
 *                      SP:N  = /root/reference/spearman_IF_pdb.pl line N
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; caller allocates every output buffer.
 *   - every call returns 0 on success, <0 on error; chr3d_last_error() gives the text
 *     (thread local).  The library never exits/aborts and never falls back to a CPU path:
 *     without a usable CUDA device every compute entry point returns CHR3D_ERR_NO_DEVICE.
 *   - bead / residue indices are 0-based in the ABI (the text writers print them 1-based,
 *     as C3D:198 does).
 *   - distances in Angstrom, energies in kcal/mol, time in ps, temperature in K.
 */
#ifndef CHR3D_H
#define CHR3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHR3D_OK              0
#define CHR3D_ERR_ARG        -1
#define CHR3D_ERR_NO_DEVICE  -2
#define CHR3D_ERR_CUDA       -3
#define CHR3D_ERR_CAPACITY   -4
#define CHR3D_ERR_IO         -5
#define CHR3D_ERR_SIZE       -6

/* ------------------------------------------------------------------------------------------
 * Physics / protocol parameters.  Defaults = the literals chromosome3D.pl interpolates into
 * dgsa.inp (C3D:65-74, 1078-1126, 1395-1426) plus the bead-level restatement of the CNS
 * NOE soft-square / bond / repel terms.
 * ---------------------------------------------------------------------------------------- */
typedef struct chr3d_sa_params {
    /* NOE soft-square term (class defaults of CNS `readdata`, C3D:1407) */
    double  noe_scale;      /* md.cool.noe = md.pow.noe = $con_wt = 10      C3D:66,1111,1120 */
    double  noe_sqconst;    /* sqconstant 1                                                  */
    double  noe_rswitch;    /* rswitch 1.0 A                                                 */
    double  noe_asymptote;  /* asymptote 2.0 (slope of the linear tail)                      */
    int32_t noe_low_soft;   /* 0: square-well below the lower bound, 1: soft on both sides   */
    int32_t vdw_min_sep;    /* repel acts on |i-j| >= this (2: bonded neighbours excluded)    */
    /* bead bond term (stand-in for the peptide backbone between consecutive CA atoms)      */
    double  bond_k;         /* 1000 kcal/mol/A^2                                             */
    double  bond_r0;        /* 3.8 A                                                         */
    /* repel term  E = w_vdw * max(0,(s*rmin)^2 - r^2)^2      rexp=2 irexp=2 rcon=1  C3D:1398 */
    double  vdw_rmin;       /* 4.73 A                                                        */
    /* dynamics */
    double  mass;           /* do (mass=100) (all)                                  C3D:1416 */
    double  fbeta;          /* do (fbeta=10) (all), 1/ps                            C3D:1415 */
    double  maxwell_temp;   /* maxwell(0.5)                                   C3D:1646-1648 */
    double  hot_temp;       /* md.hot.temp 2000                                     C3D:1093 */
    int32_t hot_step;       /* md.hot.step 1000                                     C3D:1095 */
    int32_t cool_step;      /* md.cool.step 1000                                    C3D:1097 */
    double  hot_dt;         /* md.hot.ss 0.003 ps                                   C3D:1099 */
    double  cool_dt;        /* md.cool.ss 0.005 ps                                  C3D:1101 */
    double  cool_vdw_init;  /* 0.003                                                C3D:1103 */
    double  cool_vdw_finl;  /* 4.0                                                  C3D:1105 */
    double  cool_rad_init;  /* $rep1 1.0                                         C3D:70,1107 */
    double  cool_rad_finl;  /* $rep2 0.85                                        C3D:68,1109 */
    double  cool_tmpstp;    /* 25 K                                                 C3D:1117 */
    int32_t pow_step;       /* $mini 15000                                       C3D:71,1124 */
    int32_t pow_cycl;       /* 10                                                   C3D:1126 */
    int32_t premin_step;    /* minimize lbfgs nstep=100 (run twice)           C3D:1640,1642 */
    int32_t images;         /* 2: both mirror images are annealed, lower energy kept C3D:1605 */
    /* L-BFGS (CNS's own constants are not published in the reference tree: ours, documented) */
    int32_t lbfgs_m;        /* history length, <= CHR3D_LBFGS_MAX_M                          */
    int32_t lbfgs_max_ls;   /* max Armijo halvings per iteration                             */
    double  lbfgs_c1;       /* Armijo constant                                               */
    double  lbfgs_gtol;     /* stop when rms gradient component < gtol                       */
    double  lbfgs_ftol;     /* stop when a quasi-Newton step lowers E by <= ftol*max(|E|,1)  */
    double  pow_drop;       /* drop=10.0 expected first-step energy decrease       C3D:1802 */
    uint64_t seed;          /* md.seed 82364                                        C3D:980 */
    /* Stand-in for the terms the bead model drops (CNS adds `angl`, then `impr dihe`, C3D:1641, 1643):
     * a soft lower bound on the CA(i)-CA(i+2) distance, E = w* . angle_k . max(0, angle_d2 - r)^2, scaled by the
     * stage's `weights *` factor like the bond term and reported with it (REMARK bon).  Default off (angle_d2 = 0).  */
    double  angle_d2;       /* A; 0 = off                                                    */
    double  angle_k;        /* kcal/mol/A^2                                                  */
} chr3d_sa_params;

#define CHR3D_LBFGS_MAX_M 8

/* One instruction of the annealing "stage program" (the bead-level restatement of the
 * protocol body of dgsa.inp, C3D:1573-1829).  Built by chr3d_build_program().            */
enum {
    CHR3D_OP_INIT     = 1,  /* x = x0 with x-coordinate * flag (image +1/-1), v = 0  C3D:1606-1609 */
    CHR3D_OP_MINIMIZE = 2,  /* L-BFGS, nstep iterations max                  C3D:1640,1642,1802 */
    CHR3D_OP_MAXWELL  = 3,  /* v ~ Maxwell(temp)                                   C3D:1646-1648 */
    CHR3D_OP_DYNAMICS = 4,  /* nstep Verlet steps, dt, bath temp; flag = coupling  C3D:1650-1700,1766 */
    CHR3D_OP_SAVE     = 5,  /* store x,v,E (image 1)                               C3D:1701-1708 */
    CHR3D_OP_PICK     = 6,  /* keep the lower-energy of current / saved            C3D:1711-1727 */
    CHR3D_OP_CENTER   = 7,  /* translate geometric centre to origin                C3D:1805-1816 */
    CHR3D_OP_ENERGY   = 8   /* evaluate energy with this stage's weights (report)  C3D:1818     */
};
enum { CHR3D_COUPLE_TCOUP = 1, CHR3D_COUPLE_VSCALE = 2 };

typedef struct chr3d_stage {
    int32_t op;
    int32_t nstep;
    int32_t flag;
    int32_t pad_;
    double  dt;       /* ps */
    double  temp;     /* K */
    double  w_noe;    /* `weights *` factor x NOE scale                                    */
    double  w_bond;   /* `weights *` factor x bond_k                                       */
    double  w_vdw;    /* vdw weight x rcon                                                 */
    double  repel;    /* nbonds repel= (radius scale s)                                    */
    double  drop;     /* minimize drop=                                                    */
} chr3d_stage;

/* energies reported per model, the four terms get_cns_energy() knows (C3D:604-606) */
enum { CHR3D_E_OVERALL = 0, CHR3D_E_BON = 1, CHR3D_E_VDW = 2, CHR3D_E_NOE = 3, CHR3D_E_COUNT = 4 };

/* One (matrix, alpha) job = one chromosome3D.pl invocation's worth of models (C3D:254-289). */
typedef struct chr3d_job {
    /* either an IF matrix (targets derived on the device, C3D:110-206) ...               */
    const double*  IF;        /* L*L row-major, or NULL                                    */
    int32_t        L;         /* beads (calc_len_IF, C3D:164)                              */
    int32_t        min_sep;   /* $SEPARATION 5 (C3D:20,195)                                */
    double         alpha;     /* -a (C3D:32)                                               */
    double         K;         /* -k (C3D:31)                                               */
    /* ... or an explicit restraint table (rows of contact.tbl, C3D:360), used when IF==NULL */
    const int32_t* ri;
    const int32_t* rj;
    const double*  rd;
    int64_t        nr;
    /* ensemble */
    int32_t        n_models;  /* -m (C3D:33) */
    int32_t        pad_;
    uint64_t       seed;      /* 0 = use params->seed */
    const double*  x0;        /* optional start coordinates n_models*L*3; NULL = seeded random coil */
} chr3d_job;

typedef struct chr3d_result {
    double*  xyz;          /* n_models*L*3 final coordinates (centred)                     */
    double*  energies;     /* n_models*CHR3D_E_COUNT                                       */
    int64_t* force_evals;  /* n_models: energy+force evaluations performed                 */
    int32_t* hand;         /* n_models: +1/-1 image kept (C3D:1717-1721)                   */
    int32_t* satisfied;    /* n_models: count_satisfied_tbl_rows numerator (C3D:464-473)   */
    double*  sum_dev;      /* n_models: sum_noe_dev (C3D:592-599), unrounded               */
    int64_t  n_restraints; /* rows of the restraint table (out)                            */
    double   device_ms;    /* kernel time of the run (out)                                 */
} chr3d_result;

/* ---------------------------------- lifecycle ------------------------------------------- */
const char* chr3d_version(void);
const char* chr3d_last_error(void);
int  chr3d_device_count(void);
int  chr3d_init(int device);          /* bind this host thread/process to a CUDA device   */
int  chr3d_shutdown(void);

void chr3d_default_params(chr3d_sa_params* p);
/* the same plus the default-off stand-ins for CNS's poly-GLY chain terms (angle_d2 = 6.0, angle_k = 100, vdw_rmin = 5.5), fitted on the
 * 45 shipped CNS models: 42 of 45 within 0.02 of the shipped IF-Spearman instead of 35                       */
void chr3d_cns_pin_params(chr3d_sa_params* p);
/* sizeof of the ABI structs as this build of the library sees them: a binding checks its own layout against these */
void chr3d_abi_sizes(int32_t* sa_params, int32_t* stage, int32_t* job, int32_t* result);
/* Writes the stage program for `p`; returns the number of stages (or <0).  cap==0: size query. */
int  chr3d_build_program(const chr3d_sa_params* p, chr3d_stage* out, int cap);

/* Measurement aid: FP32 FMA throughput of the device in TFLOP/s (roofline denominator of the
 * FP32-bound small-N kernel).                                                              */
int chr3d_fp32_peak(int reps, double* tflops);
/* Same chain issued as packed FFMA2 (fma.rn.f32x2), the instruction form of the large-N pair term. */
int chr3d_fp32_peak_packed(int reps, double* tflops);

/* ------------------------- (1) IF -> restraints  C3D:110-206, 340-362 -------------------- */
/* IF2dist_new: tenths[i*L+j] = the `%.1f` text of D as an integer number of tenths
 * (-10 for the "-1.0" cells).  n_patched (optional) = cells whose decimal rounding could not
 * be certified from the device result and were re-derived on the host with libm pow.       */
int chr3d_if_to_dist(const double* IF, int32_t L, double alpha, double K,
                     int64_t* tenths, int64_t* n_patched);
/* dist2rr + carr2tbl: rows in the string-sorted key order of C3D:203.  cap==0: size query. */
int chr3d_if_to_restraints(const double* IF, int32_t L, double alpha, double K, int32_t min_sep,
                           int32_t* out_i, int32_t* out_j, double* out_d,
                           int64_t cap, int64_t* out_n);
int chr3d_write_dist(const char* path, const int64_t* tenths, int32_t L);          /* <ID>.dist  C3D:156-161 */
int chr3d_write_rr  (const char* path, const int32_t* ri, const int32_t* rj,
                     const double* rd, int64_t n);                                 /* <ID>.rr    C3D:203-205 */
int chr3d_write_tbl (const char* path, const int32_t* ri, const int32_t* rj,
                     const double* rd, int64_t n);                                 /* contact.tbl C3D:360    */

/* ----------------- (2) energy + forces on fixed coordinates (CNS `energy`) --------------- */
/* xyz: n_models*N*3.  energies: n_models*4.  forces: n_models*N*3 (= -dE/dx), may be NULL. */
int chr3d_energy_forces(const double* xyz, int32_t n_models, int32_t N,
                        const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr,
                        const chr3d_sa_params* p, const chr3d_stage* weights,
                        double* energies, double* forces);

/* ------------------- (3) the SA run: replaces build_models() C3D:254-289 ----------------- */
/* Host-buffer call (the e2e path): upload, anneal every model of every job, download.     */
int chr3d_run_ensemble(const chr3d_job* jobs, int32_t njobs, const chr3d_sa_params* p,
                       const chr3d_stage* program, int32_t nstages, chr3d_result* results);

/* Device-resident variant: create uploads and derives targets, run only launches kernels on
 * `stream` (a cudaStream_t, may be NULL), fetch downloads.                                 */
typedef struct chr3d_ensemble chr3d_ensemble;
int chr3d_ensemble_create(const chr3d_job* jobs, int32_t njobs, const chr3d_sa_params* p,
                          const chr3d_stage* program, int32_t nstages, chr3d_ensemble** out);
int chr3d_ensemble_run(chr3d_ensemble* e, void* stream);
int chr3d_ensemble_fetch(chr3d_ensemble* e, chr3d_result* results);
int chr3d_ensemble_info(const chr3d_ensemble* e, int32_t* cluster_size, int32_t* threads,
                        int32_t* n_launches, int64_t* restraint_rows_total);
int chr3d_ensemble_destroy(chr3d_ensemble* e);

/* ------------- (3b) one large structure, optionally sharded over the GPUs of a node -------- */
/* The same dgsa.inp run (C3D:254-289, protocol C3D:1573-1829) for one structure of thousands of
 * beads, far beyond the reference's 663-bead cap (C3D:93-94).  The upper triangle of the target
 * matrix lives in HBM as 128x128 fp32 tiles; rank `rank` of `world` owns a contiguous range of
 * the tile list (row blocks balanced by pair count) and only derives/stores that part.  One process
 * per GPU.  Every DYNAMICS / MINIMIZE / ENERGY op of the stage program is ONE persistent cooperative
 * kernel per rank that runs all force evaluations of the op (tile pipeline, grid barriers, integrator
 * or L-BFGS decisions on the device); with world > 1 the partial forces are exchanged inside that
 * kernel over NVLink peer memory (chr3d_single_comm_handle / _connect).  Without a peer connection
 * the ranks fall back to the multi-launch path with one NCCL all-reduce per evaluation
 * (chr3d_nccl_unique_id() on rank 0, broadcast by the caller): the baseline, not the product path.
 *   IF        : host N*N row-major, or NULL with synth_seed != 0 (synthetic IF generated on the
 *               device; parsing 25k^2 decimals is not part of the hot path)
 *   x0        : optional N*3 start coordinates, NULL = seeded random coil                     */
typedef struct chr3d_single chr3d_single;
int chr3d_single_create(const double* IF, int32_t N, double alpha, double K, int32_t min_sep,
                        uint64_t synth_seed, const double* x0, uint64_t seed,
                        const chr3d_sa_params* p, const chr3d_stage* program, int32_t nstages,
                        int32_t rank, int32_t world, const void* nccl_id /* 128 bytes or NULL */,
                        chr3d_single** out);
int chr3d_single_run(chr3d_single* h, void* stream);
/* Measurement aid: `reps` force evaluations at the start coordinates, device-timed.
 * tile_ms = the stand-alone pair-tile kernel, eval_ms = a whole evaluation as the integrator runs it
 * (inside the persistent kernel: tile phase, barriers, gather, peer exchange, integration).      */
int chr3d_single_eval_bench(chr3d_single* h, void* stream, int32_t reps, int32_t want_energy,
                            double* tile_ms, double* eval_ms);
int chr3d_single_fetch(chr3d_single* h, double* xyz, double* energies4, double* forces,
                       int64_t* force_evals, int32_t* hand, double* device_ms,
                       int64_t* n_launches, int64_t* restraint_pairs);
int chr3d_single_info(const chr3d_single* h, int64_t* n_tiles, int64_t* tiles_this_rank,
                      int32_t* ctas, int64_t* target_bytes);
/* Phase clock of the persistent kernel since the last call, milliseconds as seen by CTA 0 of this rank, 8 entries:
 * [0] tile phases, [1] first grid barrier (+ energy publication), [2] step phases (one rank: gather + integration;
 * several ranks: summation + integration), [3] second barrier (+ decisions), [4] L-BFGS update + third barrier,
 * [5] push to the peers, [6] wait for the peers' partial forces, [7] gather of this rank's partials (several
 * ranks).  Reading resets it. */
int chr3d_single_profile(chr3d_single* h, double* ms8);
/* Per-CTA view of the tile walk (arrays of `cap` entries, *n = CTAs of this rank): tiles, tiles on the slow "near" path, tile
 * rows started, and the milliseconds every CTA spent in tile phases of the persistent kernel since the last call (reset by
 * reading).  This is the data the cost model of the tile partition is fitted to.                                        */
int chr3d_single_profile_ctas(chr3d_single* h, int32_t cap, int32_t* n, int32_t* tiles, int32_t* near_tiles, int32_t* rows,
                              double* tile_ms);
/* Development builds (-DCHR3D_DEV) only: raw (event, ns) trace of two CTAs of the last persistent launch, 2 x 512 pairs. */
int chr3d_single_trace(chr3d_single* h, unsigned long long* out2048);
/* Test aid: this rank's targets as a dense row-major N*N fp32 matrix (cells of other ranks' tiles and the
 * diagonal: NaN; sentinels as stored: -1 no restraint, -2 no restraint and no repel).                  */
int chr3d_single_targets(chr3d_single* h, float* out);
int chr3d_single_destroy(chr3d_single* h);
int chr3d_nccl_unique_id(void* out128);
/* Fused peer-memory force exchange (world > 1): export this rank's 64-byte CUDA IPC handle, gather the handles of
 * all ranks in rank order, connect.  Afterwards the per-evaluation exchange runs inside the persistent kernel over
 * NVLink (remote stores + epoch flags).  A wait that cannot complete (a peer died) makes the kernel give up after
 * ~20 s; chr3d_single_fetch / _eval_bench then return CHR3D_ERR_CUDA and chr3d_single_comm_error reports the code. */
int chr3d_single_comm_handle(chr3d_single* h, void* out64);
int chr3d_single_comm_connect(chr3d_single* h, const void* handles /* world*64 bytes */, int32_t world);
int chr3d_single_comm_error(chr3d_single* h, int32_t* err);
/* Virtual ranks: `n` structures created in this process on this device as rank g of world n become ONE launch
 * group — the persistent kernel hosts all of them in one cooperative grid (SMs/n CTAs each), their "peer" buffers
 * are plain device pointers.  The multi-rank protocol on a single GPU (used by the parity tests).               */
int chr3d_single_group_connect(chr3d_single** hs, int32_t n);
int chr3d_single_group_run(chr3d_single** hs, int32_t n, void* stream);
/* Sparse restraint tables (real high-resolution Hi-C: most bin pairs have IF = 0 and no restraint, C3D:148-152, 196): one
 * structure from explicit rows (0 <= ri < rj < N, rd > 0, no pair twice — the rows of contact.tbl / chr3d_if_to_restraints).  The
 * NOE term walks a CSR of the rows (8 B per entry and evaluation instead of 4 B for every one of the N^2/2 pairs), the repel
 * term uses a cell list that is rebuilt on the device when a bead has moved more than half the skin (CNS: nbonds cutnb /
 * tolerance, C3D:1395-1404, 1729-1738); run / fetch / destroy are the chr3d_single_* calls.  One GPU.                      */
int chr3d_single_create_sparse(int32_t N, const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr,
                               const double* x0, uint64_t seed, const chr3d_sa_params* p, const chr3d_stage* program,
                               int32_t nstages, chr3d_single** out);
int chr3d_single_sparse_info(chr3d_single* h, int64_t* n_rebuilds, int64_t* csr_entries, int64_t* noe_bytes_per_eval);
/* Synthetic IF matrix of SURVEY 8(d) (N*N doubles, generated on the device, copied back).     */
int chr3d_synthetic_if(int32_t N, uint64_t seed, double* out);

/* --------------- (4) ranking, violation statistics, PDB  C3D:447-600, 769-880 ------------ */
/* xyz are rounded to the 3-decimal PDB text first (C3D:727 works on the text fields).     */
int chr3d_assess(const double* xyz, int32_t n_models, int32_t N,
                 const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr,
                 int32_t* satisfied, double* sum_dev);
/* Rows of contact_violation.txt for ONE model (C3D:458-475): violation flag, deviation, "%.3f" distance. */
int chr3d_assess_rows(const double* xyz, int32_t N,
                      const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr,
                      int32_t* flag, double* dev, double* dist);
/* order[k] = model index of rank k (ascending int(noe), ties by model index; C3D:617,802,823) */
int chr3d_rank_models(const double* energies, int32_t n_models, int32_t* order);
/* CNS-style trial structure <ID>_k.pdb: REMARK energies + one CA ATOM per bead + END      */
int chr3d_write_pdb(const char* path, const double* xyz, int32_t N, const double* energies4,
                    const char* resnames /* N*3 chars or NULL -> REFSEQUENCE */, int32_t with_conect);

/* --------------------------- (5) Spearman  SP:32-71 -------------------------------------- */
int chr3d_spearman(const double* IF, int32_t L, const double* xyz, int32_t n_models,
                   int32_t range, double* rho);

/* ------------------------- 8(f) "next" rows: callers and formats either side of the path -------------------------
 * The programs behind output_models "_reduced.pdb" files and similarity.txt are not in the reference tree; these entry points
 * reproduce their shipped outputs (formats: SURVEY App. B.6).                                                           */
/* bead k of the reduced model = mean of beads 2k, 2k+1 (a trailing unpaired bead is copied); out_xyz: ceil(N/2)*3        */
int chr3d_reduce_model(const double* xyz, int32_t N, double* out_xyz);
/* the "_reduced.pdb" dialect: empty first line, CRLF, chain B, "%.3f" with HALF_UP on the shortest decimal digits        */
int chr3d_write_reduced_pdb(const char* path, const double* xyz, int32_t n);
/* similarity.txt: Spearman of the upper-triangle distance vectors of two models (first min(n_a, n_b) beads) and the RMSE
 * between them after scaling b's distances by mean(da)/mean(db)                                                          */
int chr3d_model_similarity(const double* xyz_a, int32_t n_a, const double* xyz_b, int32_t n_b, double* spearman, double* rmsd);
/* input/<ID>.txt as IF2dist_new parses it (C3D:116-129, 164-179); cap == 0: only *L is returned                           */
int chr3d_read_if_text(const char* path, double* out, int64_t cap, int32_t* L);
/* raw container for matrices beyond decimal text: "CHR3DIF1", int32 L, int32 elem_bytes (4|8), L*L little-endian values    */
int chr3d_write_if_binary(const char* path, const double* IF, int32_t L, int32_t elem_bytes);
int chr3d_read_if_binary(const char* path, double* out, int64_t cap, int32_t* L);
/* sparse (bin_i, bin_j, value) triplets, base 0 or 1, -> dense symmetric L*L matrix (device scatter; repeats add up)        */
int chr3d_if_from_coo(const int32_t* bi, const int32_t* bj, const double* v, int64_t nnz, int32_t L, int32_t base, double* out);

/* sparse triplets -> restraint rows (IF2dist_new + dist2rr, C3D:110-206) without forming the L x L matrix: the arithmetic runs on
 * the device over the non-zero cells only and gives the rows of chr3d_if_to_restraints on the dense matrix chr3d_if_from_coo would
 * build (same distances, bit for bit), in i-major row order instead of the reference's string order.  alpha > 0.  ri/rj/rd: capacity
 * `cap` rows (nnz is always enough); cap = 0 only counts.  Feed the rows to chr3d_single_create_sparse or chr3d_job.ri/rj/rd.        */
int chr3d_coo_to_restraints(const int32_t* bi, const int32_t* bj, const double* v, int64_t nnz, int32_t L, int32_t base,
                            double alpha, double K, int32_t min_sep, int32_t* ri, int32_t* rj, double* rd, int64_t cap, int64_t* n_out);

/* DG stage of dgsa.inp (C3D:1470-1526) at bead level: bounds from the restraint table (0-based ri < rj), triangle smoothing,
 * per-model trial distances, metric-matrix embedding (three largest positive eigenpairs).  xyz_out: n_models*N*3 starting
 * coordinates — pass them as chr3d_job.x0; eig3_out (n_models*3) and n_inconsistent are optional diagnostics.  N <= 1600.   */
int chr3d_dg_embed(const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr, int32_t N, const chr3d_sa_params* p,
                   int32_t n_models, uint64_t seed, double* xyz_out, double* eig3_out, int64_t* n_inconsistent);

/* The whole DG stage of dgsa.inp (C3D:1470-1526): the embedding above, then `do (x = x * $dgscale)` (C3D:1516-1518) and
 * `minimize lbfgs nstep=&md.dg.step drop=1.` (C3D:1519-1521; md.dg.step = 200, C3D:1090) on the bounds-violation energy alone
 * (`flags exclude * include dg`, C3D:1495) with weight dg_scale (md.dg.scale = 100, exponent 2, C3D:1078-1080).  The template
 * fit that follows in the script (C3D:1604-1621) is the identity for one-atom residues; the enantiomer loop is
 * the INIT(+1/-1) ... SAVE / PICK pair of the stage program.  stats_out (optional): 6 doubles per model {dgscale, E after
 * scaling, E final, iterations, energy evaluations, rms gradient}. */
int chr3d_dg_starts(const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr, int32_t N, const chr3d_sa_params* p,
                    int32_t n_models, uint64_t seed, double dg_scale, int32_t nstep, double drop,
                    double* xyz_out, double* eig3_out, double* stats_out, int64_t* n_inconsistent);

/* Host-only (no device): the cost-balanced tile-range split chr3d_single_create uses for N beads on `world` ranks (SURVEY 8(e):
 * upper-triangle layout, contiguous ranges; a diagonal tile and the start of a tile row cost more than a far tile).
 * rank_cut[world + 1]; cta_cut[world * (ctas + 1)] and cost[world] are optional. */
int chr3d_single_tile_split(int32_t N, int32_t world, int32_t ctas, int64_t* rank_cut, int64_t* cta_cut, double* cost);

/* Launch-shape options of the ensemble kernel, process-wide, read when an ensemble is created (unset = the cost model of the
 * library decides): "sa_cluster" = CTAs per trajectory (1..8), "sa_warps" = warps per CTA (4..12), "sa_min_smem" = least shared
 * memory a CTA asks for in bytes (0 lets several CTAs share an SM), "sa_split" = "0" never runs the two hot-MD images of a model
 * (C3D:1605-1708) as trajectories of their own.  value NULL or "" clears the option.  (The reference has no counterpart: its
 * parallelism is one cns_solve process per run, test.sh.)  "seqsum" = "serial" evaluates the left-to-right fp64 sums (below) as
 * a serial chain instead of in parallel — same bits, for cross-checks.  Everything else that used to be an environment switch
 * is compiled in only with -DCHR3D_DEV. */
int chr3d_set_option(const char* name, const char* value);

/* The left-to-right fp64 sum `for (k) s += a[k]` of each of nseg segments of n_per terms (host arrays), evaluated on the device
 * with the bits of the scalar loop: what the restraint derivation uses for the mean of IF**alpha (`$sum += ...`, C3D:132-139)
 * and the assessment for the sum of deviations (C3D:590-598).  Parallel inside a binade of the running sum (integer prefix
 * scans), one real addition per tie / carry / irregular term; csrc/seqsum_core.h states the argument. */
int chr3d_seq_sum(const double* a, int64_t n_per, int32_t nseg, double* out);

#ifdef __cplusplus
}
#endif
#endif /* CHR3D_H */
