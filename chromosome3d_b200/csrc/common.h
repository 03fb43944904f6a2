// common.h — shared host-side helpers of libchr3d.so (error text, CUDA checks, device structs)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "../../include/chr3d.h"

#define CHR3D_API extern "C" __attribute__((visibility("default")))

// AKMA unit system of CNS/X-PLOR
#define CHR3D_TIMFAC 0.04888821
#define CHR3D_KBOLTZ 1.987191e-3

namespace chr3d {

// Tuning options set through chr3d_set_option (launch shape of the ensemble kernel; see include/chr3d.h) are looked up first.
// Development switches (probes, traces, debug variants of the large-N kernels) are read from the environment only in a
// -DCHR3D_DEV build (`make DEV=1`); the release library ignores the environment, so a stray variable cannot change what a
// benchmark measures.
const char* option_lookup(const char* name);       // api.cu; nullptr when the option is not set
static inline const char* dev_getenv(const char* name) {
    if (const char* v = option_lookup(name)) return v;
#ifdef CHR3D_DEV
    return getenv(name);
#else
    return nullptr;
#endif
}

void set_error(const char* fmt, ...);
int  require_device();                 // 0 or CHR3D_ERR_NO_DEVICE, sets the error text

#define CHR3D_CUDA(call)                                                                         \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            chr3d::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return CHR3D_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

// RAII device buffer.  Stream-ordered allocation from the device's default memory pool (kept warm by
// require_device(): release threshold = never), because cudaMalloc/cudaFree churn dominated the host-side
// cost of one chr3d_run_ensemble call.  Allocation and release are ordered on the legacy default stream,
// which synchronises with every blocking stream the kernels may be launched on.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMallocAsync((void**)&p, count * sizeof(T), (cudaStream_t)0);
    }
    void release() { if (p) cudaFreeAsync(p, (cudaStream_t)0); p = nullptr; n = 0; }
};

// ---- device-side views of the physics parameters / stage program (fp32 compute) -------------
struct DevParams {
    float C, rsw, casym, a, b;     // NOE soft-square constants
    float lo_clamp;                // -rsw if the low side is soft, -inf otherwise
    float r0;                      // bond length
    float mass, inv_mass;
    float fbeta;                   // 1/ps
    float kboltz;
    int   m, max_ls;
    float c1, gtol, ftol;
    int   general_noe;             // b != 0: use the exact piecewise formula
    float ang_d2, ang_ratio;       // CA(i)-CA(i+2) lower bound (0 = off), angle_k / bond_k
};
struct DevStage {
    int   op, nstep, flag, pad_;
    float dt_akma, dt_ps, temp, w_noe, w_bond, w_vdw, R2, drop;
};

void make_dev_params(const chr3d_sa_params& p, DevParams& d);
void make_dev_stage(const chr3d_sa_params& p, const chr3d_stage& s, DevStage& d);

// pitch (floats) of the dense target matrix
static inline int target_pitch(int N) { return (N + 31) & ~31; }

// ---- restraints.cu ---------------------------------------------------------------------------
// Derives, on the device, the `%.1f` tenths of D = K/(IF^alpha/mean) for every cell (d_tenths, L*L
// int64) and a certification flag count.  d_IF: L*L doubles on the device.
// flagged_host: indices of cells whose decimal rounding is not certified (empty when more than
// CHR3D_MAX_FLAGGED cells are flagged: the caller then re-derives every cell).
#define CHR3D_MAX_FLAGGED 4096
int dev_if_to_tenths(const double* d_IF, int L, double alpha, double K, long long* d_tenths,
                     long long* n_uncertain_host, std::vector<long long>* flagged_host, cudaStream_t s);
int dev_cells_to_tenths(const double* d_vals, long long n, const long long* d_row_ptr, int L, double alpha, double K, long long* d_tenths,
                        long long* n_uncertain_host, std::vector<long long>* flagged_host, cudaStream_t s);
// Dense fp32 target matrix [N][pitch] from tenths: d>0 restraint, -1 none, -2 none + no repel.
int dev_tenths_to_targets(const long long* d_tenths, int L, int min_sep, int vdw_min_sep, float* d_T,
                          cudaStream_t s);
int dev_list_to_targets(const int32_t* d_ri, const int32_t* d_rj, const double* d_rd, long long nr, int N,
                        int vdw_min_sep, float* d_T, cudaStream_t s);
// Lexicographically ordered restraint list (C3D:203) from tenths; d_count receives the row count.
int dev_tenths_to_list(const long long* d_tenths, int L, int min_sep, int32_t* d_ri, int32_t* d_rj,
                       double* d_rd, long long cap, long long* n_host, cudaStream_t s);

// ---- api.cu ----------------------------------------------------------------------------------
// IF (host) -> `%.1f` tenths on the device with certified decimal rounding (bit-exact with the Perl)
int if_to_tenths_certified(const double* IF, int L, double alpha, double K, DevBuf<double>& dIF,
                           DevBuf<long long>& dT, long long* n_patched, cudaStream_t s);
// largest L for which the single-structure engine derives its targets through the certified table
#define CHR3D_EXACT_TABLE_MAX_L 4096

// ---- sa_ensemble.cu --------------------------------------------------------------------------
struct TrajDesc {
    const float* T;        // tile-packed lane-major targets (sa_pack_tiles)
    const float* x0;       // N*3 start coordinates
    float*  x_out;         // N*3
    float*  f_out;         // N*3 or nullptr (forces of the last evaluation)
    float*  scratch;       // 6*N floats (SAVE slot)
    double* e_out;         // 4
    long long* nevals_out; // 1
    int*    hand_out;      // 1
    int     N, pitch, model, pad_;
    unsigned long long seed;
    // stage range [k0, k1) of the program this trajectory executes (k1 == 0: the whole program).  The two hot-MD images of
    // a model (C3D:1605-1708) are independent until PICK, so the host may run them as two trajectories side by side
    // and let a third one resume at PICK:
    int     k0, k1;
    int     resume_image, pad2_;      // image sign of the state in state_in (the trajectory that resumes at PICK)
    const float*  state_in;           // != nullptr: start from these 6N floats (x then v) instead of an INIT op
    float*        state_out;          // != nullptr: x and v at the end of the range (6N floats)
    const double* esaved_in;          // energy of the SAVE slot, for the resumed PICK
    double*       esaved_out;         // != nullptr: the energy SAVE recorded
};
struct SaLaunchCfg { int cluster; int threads; int stage_targets; size_t smem; double cost; };   // stage_targets: resident tile rounds; cost: modelled cycles per force evaluation of the whole batch
// dense [N][pitch] targets -> 32x32 tile-packed lane-major layout of the SA kernel
int sa_pack_tiles(const float* d_T, int N, int pitch, float* d_Tt, cudaStream_t s);
size_t sa_tiles_floats(int N);
int sa_choose_config(int N, int n_traj, int m, SaLaunchCfg* cfg);
int sa_launch(const TrajDesc* d_traj, int n_traj, int Nmax, const DevParams& P, const DevStage* d_prog, int nstages,
              const SaLaunchCfg& cfg, cudaStream_t s);
int dev_random_coil(float* d_x0, int N, int n_models, int first_model, double step, unsigned long long seed,
                    cudaStream_t s);

// ---- assess.cu -------------------------------------------------------------------------------
int dev_assess(const double* d_xyz, int n_models, int N, const int32_t* d_ri, const int32_t* d_rj,
               const double* d_rd, long long nr, int32_t* d_sat, double* d_dev, cudaStream_t s);
int dev_assess_rows(const double* d_xyz, int N, const int32_t* d_ri, const int32_t* d_rj, const double* d_rd, long long nr,
                    int32_t* d_flag, double* d_dev, double* d_dist, cudaStream_t s);
int dev_spearman(const double* d_IF, int L, const double* d_xyz, int n_models, int range, double* rho_host,
                 cudaStream_t s);
int dev_model_similarity(const double* d_xa, const double* d_xb, int n, double* rho_host, double* rmsd_host, cudaStream_t s);

}  // namespace chr3d

// ---- large_n.cu (internal entry points used by the ensemble API for structures beyond the cluster kernel's capacity)
struct chr3d_single;
int chr3d_single_create_dense(const float* d_T, int pitch, int32_t N, uint64_t seed, const chr3d_sa_params* p, const chr3d_stage* program,
                              int32_t nstages, long long n_rows, chr3d_single** out);
int chr3d_single_run_model(chr3d_single* h, int model, const float* d_x0, void* stream, float* d_xout, double* d_e4, long long* d_nevals, int* d_hand);
