// api.cu — the C ABI of libchr3d.so (include/chr3d.h): host orchestration around the kernels.
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <time.h>
#include <algorithm>
#include <map>
#include <memory>
#include <vector>
#include <mutex>
#include "common.h"

namespace chr3d {

// ---- options (chr3d_set_option): a small process-wide table, consulted by dev_getenv before the environment ----------------
static std::mutex g_opt_mu;
static std::map<std::string, std::string>& opt_table() { static std::map<std::string, std::string> t; return t; }
const char* option_lookup(const char* name) {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    auto& t = opt_table();
    auto it = t.find(name);
    return it == t.end() ? nullptr : it->second.c_str();
}

static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
}
static int g_device = -1;
// CHR3D_TRACE=1: host-side phase timings on stderr (development aid)
struct Trace {
    bool on; double t0; const char* what;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
    explicit Trace(const char* w) : what(w) { static int en = dev_getenv("CHR3D_TRACE") ? 1 : 0; on = en; if (on) t0 = now(); }
    void lap(const char* tag) { if (on) { cudaDeviceSynchronize(); double t = now(); fprintf(stderr, "[chr3d] %s/%s %.3f ms\n", what, tag, 1e3 * (t - t0)); t0 = t; } }
};
int require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no usable CUDA device (%s); libchr3d has no CPU path", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return CHR3D_ERR_NO_DEVICE;
    }
    if (g_device < 0) { g_device = 0; cudaGetDevice(&g_device); }
    static int pool_ready = -1;
    if (pool_ready != g_device) {           // keep freed blocks in the pool instead of returning them to the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, g_device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        pool_ready = g_device;
    }
    return 0;
}

void make_dev_params(const chr3d_sa_params& p, DevParams& d) {
    memset(&d, 0, sizeof d);
    d.C = (float)p.noe_sqconst; d.rsw = (float)p.noe_rswitch; d.casym = (float)p.noe_asymptote;
    double b = (p.noe_asymptote - 2.0 * p.noe_sqconst * p.noe_rswitch) * p.noe_rswitch * p.noe_rswitch;
    double a = p.noe_sqconst * p.noe_rswitch * p.noe_rswitch - b / p.noe_rswitch - p.noe_asymptote * p.noe_rswitch;
    d.a = (float)a; d.b = (float)b;
    d.general_noe = fabs(b) > 1e-12 ? 1 : 0;
    d.lo_clamp = p.noe_low_soft ? -(float)p.noe_rswitch : -INFINITY;
    d.r0 = (float)p.bond_r0; d.mass = (float)p.mass; d.inv_mass = (float)(1.0 / p.mass);
    d.fbeta = (float)p.fbeta; d.kboltz = (float)CHR3D_KBOLTZ;
    d.m = std::max(1, std::min((int)p.lbfgs_m, CHR3D_LBFGS_MAX_M)); d.max_ls = p.lbfgs_max_ls;
    d.c1 = (float)p.lbfgs_c1; d.gtol = (float)p.lbfgs_gtol; d.ftol = (float)p.lbfgs_ftol;
    d.ang_d2 = (p.angle_d2 > 0 && p.bond_k > 0 && p.angle_k > 0) ? (float)p.angle_d2 : 0.f; d.ang_ratio = d.ang_d2 > 0.f ? (float)(p.angle_k / p.bond_k) : 0.f;
}
void make_dev_stage(const chr3d_sa_params& p, const chr3d_stage& s, DevStage& d) {
    memset(&d, 0, sizeof d);
    d.op = s.op; d.nstep = s.nstep; d.flag = s.flag;
    d.dt_ps = (float)s.dt; d.dt_akma = (float)(s.dt / CHR3D_TIMFAC); d.temp = (float)s.temp;
    d.w_noe = (float)s.w_noe; d.w_bond = (float)s.w_bond; d.w_vdw = (float)s.w_vdw;
    double r = s.repel * p.vdw_rmin; d.R2 = (float)(r * r); d.drop = (float)s.drop;
}

// ---- exact host derivation of single cells (decimal-rounding certification, see restraints.cu)
static long long parse_tenths(const char* s) {
    int neg = 0; if (*s == '-') { neg = 1; s++; }
    long long v = 0;
    for (; *s && *s != '.'; s++) { if (*s < '0' || *s > '9') return 0; v = v * 10 + (*s - '0'); }
    if (*s == '.') s++;
    v = v * 10 + ((*s >= '0' && *s <= '9') ? (*s - '0') : 0);
    return neg ? -v : v;
}
static int patch_cells(const double* IF, int L, double alpha, double K, long long* d_tenths,
                       const std::vector<long long>& cells, bool all, cudaStream_t s) {
    const size_t n = (size_t)L * L;
    double sum = 0.0;
    for (size_t k = 0; k < n; k++) sum += pow(IF[k], alpha);          // C3D:132-137
    const double mean = sum / ((double)L * (double)L);
    auto cell = [&](size_t k) -> long long {
        double q = pow(IF[k], alpha) / mean;
        if (q == 0) return -10;
        char buf[512]; snprintf(buf, sizeof buf, "%.1f", K / q);
        return parse_tenths(buf);
    };
    if (all) {
        std::vector<long long> t(n);
        for (size_t k = 0; k < n; k++) t[k] = cell(k);
        CHR3D_CUDA(cudaMemcpyAsync(d_tenths, t.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, s));
        CHR3D_CUDA(cudaStreamSynchronize(s));
    } else {
        for (long long k : cells) {
            long long v = cell((size_t)k);
            CHR3D_CUDA(cudaMemcpyAsync(d_tenths + k, &v, sizeof(long long), cudaMemcpyHostToDevice, s));
            CHR3D_CUDA(cudaStreamSynchronize(s));
        }
    }
    return 0;
}
// IF (host) -> tenths on the device, certified
int if_to_tenths_certified(const double* IF, int L, double alpha, double K, DevBuf<double>& dIF,
                           DevBuf<long long>& dT, long long* n_patched, cudaStream_t s) {
    const size_t n = (size_t)L * L;
    CHR3D_CUDA(dIF.alloc(n)); CHR3D_CUDA(dT.alloc(n));
    CHR3D_CUDA(cudaMemcpyAsync(dIF.p, IF, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    long long nu = 0; std::vector<long long> flagged;
    int rc = dev_if_to_tenths(dIF.p, L, alpha, K, dT.p, &nu, &flagged, s);
    if (rc) return rc;
    if (nu > 0) { rc = patch_cells(IF, L, alpha, K, dT.p, flagged, nu > CHR3D_MAX_FLAGGED, s); if (rc) return rc; }
    if (n_patched) *n_patched = nu;
    return 0;
}

// Sparse triplets -> restraint rows without ever forming the L x L matrix (matrix input at scale).  The
// triplets are symmetrised and ordered on the host (data-format work, like the text parser): every (i, j, v) counts for (i, j)
// and (j, i), repeated cells add up in input order.  The arithmetic — IF**alpha, the mean in the reference's order, K / (IF**alpha /
// mean), the certified `%.1f` rounding — runs on the device over the non-zero cells only (dev_cells_to_tenths); flagged cells are
// re-derived with libm exactly as for a dense matrix.  Rows come out in i-major order (not the string order of `sort keys`,
// C3D:203: the engines do not care, chr3d_write_rr / _tbl of these rows would differ from the reference's files in row order only).
CHR3D_API int chr3d_coo_to_restraints(const int32_t* bi, const int32_t* bj, const double* v, int64_t nnz, int32_t L, int32_t base,
                                      double alpha, double K, int32_t min_sep, int32_t* ri, int32_t* rj, double* rd, int64_t cap, int64_t* n_out) {
    if (L < 2 || nnz < 0 || (nnz > 0 && (!bi || !bj || !v)) || (base != 0 && base != 1) || !n_out || !(alpha > 0) || (cap > 0 && (!ri || !rj || !rd))) {
        set_error("chr3d_coo_to_restraints: bad arguments (alpha must be positive: 0**alpha of the absent cells has to be 0)"); return CHR3D_ERR_ARG;
    }
    if (int rc = require_device()) return rc;
    struct Ent { unsigned long long key; double v; };
    std::vector<Ent> ent; ent.reserve((size_t)nnz * 2);
    for (int64_t k = 0; k < nnz; k++) {
        const long long i = (long long)bi[k] - base, j = (long long)bj[k] - base;
        if (i < 0 || j < 0 || i >= L || j >= L) { set_error("chr3d_coo_to_restraints: bin index outside [%d, %d] in triplet %lld", base, L - 1 + base, (long long)k); return CHR3D_ERR_ARG; }
        if (!(v[k] >= 0.0)) { set_error("chr3d_coo_to_restraints: triplet %lld has a negative or NaN value", (long long)k); return CHR3D_ERR_ARG; }
        ent.push_back({(unsigned long long)i * (unsigned long long)L + (unsigned long long)j, v[k]});
        if (i != j) ent.push_back({(unsigned long long)j * (unsigned long long)L + (unsigned long long)i, v[k]});
    }
    std::stable_sort(ent.begin(), ent.end(), [](const Ent& a, const Ent& b) { return a.key < b.key; });
    std::vector<unsigned long long> key; std::vector<double> val;
    for (size_t k = 0; k < ent.size(); k++) {
        if (!key.empty() && key.back() == ent[k].key) val.back() += ent[k].v;
        else { key.push_back(ent[k].key); val.push_back(ent[k].v); }
    }
    {   // cells that are (or sum to) zero are absent cells
        size_t w = 0;
        for (size_t k = 0; k < key.size(); k++) if (val[k] != 0.0) { key[w] = key[k]; val[w] = val[k]; w++; }
        key.resize(w); val.resize(w);
    }
    const long long n = (long long)key.size();
    std::vector<long long> row_ptr((size_t)L + 1, 0);
    for (long long k = 0; k < n; k++) row_ptr[(size_t)(key[(size_t)k] / (unsigned long long)L) + 1]++;
    for (int i = 0; i < L; i++) row_ptr[(size_t)i + 1] += row_ptr[(size_t)i];
    DevBuf<double> dv; DevBuf<long long> dptr, dT;
    CHR3D_CUDA(dv.alloc((size_t)std::max<long long>(n, 1))); CHR3D_CUDA(dT.alloc((size_t)std::max<long long>(n, 1))); CHR3D_CUDA(dptr.alloc((size_t)L + 1));
    if (n > 0) CHR3D_CUDA(cudaMemcpy(dv.p, val.data(), sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    CHR3D_CUDA(cudaMemcpy(dptr.p, row_ptr.data(), sizeof(long long) * ((size_t)L + 1), cudaMemcpyHostToDevice));
    long long nu = 0; std::vector<long long> flagged;
    if (int rc = dev_cells_to_tenths(dv.p, n, dptr.p, L, alpha, K, dT.p, &nu, &flagged, 0)) return rc;
    std::vector<long long> t((size_t)std::max<long long>(n, 1));
    if (n > 0) CHR3D_CUDA(cudaMemcpy(t.data(), dT.p, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost));
    if (nu > 0) {   // decimal-rounding certification failed for these cells: libm, the reference's own arithmetic
        double sum = 0.0;
        for (long long k = 0; k < n; k++) sum += pow(val[(size_t)k], alpha);
        const double mean = sum / ((double)L * (double)L);
        auto cell = [&](size_t k) -> long long {
            const double q = pow(val[k], alpha) / mean;
            if (q == 0) return -10;
            char buf[512]; snprintf(buf, sizeof buf, "%.1f", K / q);
            return parse_tenths(buf);
        };
        if (nu > CHR3D_MAX_FLAGGED) { for (long long k = 0; k < n; k++) t[(size_t)k] = cell((size_t)k); }
        else for (long long k : flagged) t[(size_t)k] = cell((size_t)k);
    }
    int64_t rows = 0;
    for (long long k = 0; k < n; k++) {
        const long long i = (long long)(key[(size_t)k] / (unsigned long long)L), j = (long long)(key[(size_t)k] % (unsigned long long)L);
        if (j <= i || j - i < min_sep || t[(size_t)k] <= 0) continue;             // C3D:195-196
        if (rows < cap) { ri[rows] = (int32_t)i; rj[rows] = (int32_t)j; rd[rows] = (double)t[(size_t)k] / 10.0; }
        rows++;
    }
    *n_out = rows;
    if (cap > 0 && rows > cap) { set_error("chr3d_coo_to_restraints: %lld rows, capacity %lld", (long long)rows, (long long)cap); return CHR3D_ERR_SIZE; }
    return 0;
}

static const char* REFSEQ =   // the fake protein sequence CNS is given, one residue per bead (C3D:93)
    "RSEDWQCPRTPYAASRDFDVKYVVPSFSAGGLVQAMVTYEGDRNESAVFVAIRNRLHVLGPDLKSVQSLATGPAGDPGCQTCAACGPGPHGPPGDTDTKVLVLDPALPALVSCGSSLQGRCFLHDLEPQGTAVHLAAPACLFSAHHNRPDDCPDCVASPLGTRVTVVEQGQASYFYVASSLDAAVAGSFSPRSVSIRRLKADASGFAPGFVALSVLPKHLVSYSIEYVHSFHTGAFVYFLTVQPASVTDDPSALHTRLARLSATEPELGDYRELVLDCRFAPKLVPRGSPEGGQPYPVLQVAHSAPVGAQLATELSIAEGQEVLFGVFVTGKDGGPGVGPNSVVCAFPIDLLDTLIDEGVERCCESPVHPGLRRGLDFFQSPSFCPNPPGLEALSPNTSCRHFPLLVSSSFSRVDLFNGLLGPVQVTALYVTRLDNVTVAHMGTMDGRILQVELVRSLNYLLYVSNFSLGDSGQPVQRDVSRLGDHLLFASGDQVFQVPIQGPGCRHFLTCGRCLRAWHFMGCGWCGNMCGQQKECPGSWQQDHCPPKLTEFHPHSGPLRGSTRLTLCGSNFYLHPSGLVPEGTHQVTVGQSPCRPLPKDSSKLRPVPRKDFVEEFECELEPLGTQAVGPTNVSLTVTNMPPGKHFRVDGTSVLRGFSFMETG";
static const char* aa3(char c) {
    switch (c) {   // %AA3TO1 reversed (C3D:77-78)
    case 'A': return "ALA"; case 'N': return "ASN"; case 'C': return "CYS"; case 'Q': return "GLN"; case 'H': return "HIS";
    case 'L': return "LEU"; case 'M': return "MET"; case 'P': return "PRO"; case 'T': return "THR"; case 'Y': return "TYR";
    case 'R': return "ARG"; case 'D': return "ASP"; case 'E': return "GLU"; case 'G': return "GLY"; case 'I': return "ILE";
    case 'K': return "LYS"; case 'F': return "PHE"; case 'S': return "SER"; case 'W': return "TRP"; case 'V': return "VAL";
    }
    return "GLY";
}

// ------------------------------------------------------------------------------------------------
// ensemble object
// ------------------------------------------------------------------------------------------------
struct JobDev {
    int L = 0, n_models = 0; long long nr = 0; unsigned long long seed = 0;
    DevBuf<float> T, Tt, x0, xout, scratch;
    DevBuf<double> e, xyz_d, dev;
    DevBuf<long long> nev;
    DevBuf<int> hand, sat;
    DevBuf<int32_t> ri, rj; DevBuf<double> rd;
    SaLaunchCfg cfg;
    // split run (see ensemble_create): the two hot-MD images of every model anneal side by side as trajectories of their own
    // (launch shape cfg_img), a third trajectory resumes at PICK with the shape cfg
    bool split = false; SaLaunchCfg cfg_img;
    DevBuf<float> st_img; DevBuf<double> esaved; DevBuf<long long> nev_a, nev_b; DevBuf<int> hand_tmp;
    int traj0 = 0;
    chr3d_single* big = nullptr;      // structures beyond the cluster kernel's capacity run on the tile-streaming engine
    bool has_x0 = false;
    ~JobDev();
};
struct LaunchGroup { SaLaunchCfg cfg; int traj0, n_traj, Nmax, phase; };

JobDev::~JobDev() { if (big) chr3d_single_destroy(big); }

__global__ void f2d_kernel(const float* __restrict__ a, double* __restrict__ b, long long n) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) b[k] = (double)a[k];
}
__global__ void d2f_kernel(const double* __restrict__ a, float* __restrict__ b, long long n) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) b[k] = (float)a[k];
}

}  // namespace chr3d

using namespace chr3d;

struct chr3d_ensemble {
    std::vector<std::unique_ptr<JobDev>> jobs;
    std::vector<LaunchGroup> groups;
    DevBuf<TrajDesc> traj;
    DevBuf<DevStage> prog;
    DevParams P;
    int nstages = 0, n_traj = 0, n_launches = 0;
    long long rows_total = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evmid = nullptr;
    std::vector<cudaStream_t> gstreams; std::vector<cudaEvent_t> gevents;
    bool ran = false;
    ~chr3d_ensemble() {
        if (ev0) cudaEventDestroy(ev0); if (ev1) cudaEventDestroy(ev1); if (evmid) cudaEventDestroy(evmid);
        for (cudaEvent_t ev : gevents) if (ev) cudaEventDestroy(ev);
        for (cudaStream_t st : gstreams) if (st) cudaStreamDestroy(st);
    }
};

// ---------------------------------- lifecycle ---------------------------------------------------
CHR3D_API const char* chr3d_version(void) { return "chr3d 0.1 (sm_100a)"; }
CHR3D_API const char* chr3d_last_error(void) { return g_err.c_str(); }
CHR3D_API int chr3d_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
CHR3D_API int chr3d_init(int device) {
    int n = chr3d_device_count();
    if (n <= 0) { set_error("no usable CUDA device; libchr3d has no CPU path"); return CHR3D_ERR_NO_DEVICE; }
    if (device < 0 || device >= n) { set_error("device %d out of range (0..%d)", device, n - 1); return CHR3D_ERR_ARG; }
    CHR3D_CUDA(cudaSetDevice(device));
    CHR3D_CUDA(cudaFree(0));
    g_device = device;
    return 0;
}
CHR3D_API int chr3d_shutdown(void) { g_device = -1; return 0; }

CHR3D_API void chr3d_default_params(chr3d_sa_params* p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->noe_scale = 10.0; p->noe_sqconst = 1.0; p->noe_rswitch = 1.0; p->noe_asymptote = 2.0; p->noe_low_soft = 0;
    p->vdw_min_sep = 2; p->bond_k = 1000.0; p->bond_r0 = 3.8; p->vdw_rmin = 4.73;
    p->mass = 100.0; p->fbeta = 10.0; p->maxwell_temp = 0.5;
    p->hot_temp = 2000.0; p->hot_step = 1000; p->cool_step = 1000; p->hot_dt = 0.003; p->cool_dt = 0.005;
    p->cool_vdw_init = 0.003; p->cool_vdw_finl = 4.0; p->cool_rad_init = 1.0; p->cool_rad_finl = 0.85; p->cool_tmpstp = 25.0;
    p->pow_step = 15000; p->pow_cycl = 10; p->premin_step = 100; p->images = 2;
    p->lbfgs_m = 5; p->lbfgs_max_ls = 20; p->lbfgs_c1 = 1e-4; p->lbfgs_gtol = 1e-3; p->lbfgs_ftol = 1e-7; p->pow_drop = 10.0;
    p->seed = 82364;
    p->angle_d2 = 0.0; p->angle_k = 0.0;       // the CA(i)-CA(i+2) lower bound is off by default
}
CHR3D_API void chr3d_abi_sizes(int32_t* sa_params, int32_t* stage, int32_t* job, int32_t* result) {
    if (sa_params) *sa_params = (int32_t)sizeof(chr3d_sa_params);
    if (stage) *stage = (int32_t)sizeof(chr3d_stage);
    if (job) *job = (int32_t)sizeof(chr3d_job);
    if (result) *result = (int32_t)sizeof(chr3d_result);
}

// The dgsa.inp protocol body as a stage program (C3D:1573-1829).
CHR3D_API int chr3d_build_program(const chr3d_sa_params* p, chr3d_stage* out, int cap) {
    if (!p) { set_error("chr3d_build_program: null params"); return CHR3D_ERR_ARG; }
    std::vector<chr3d_stage> v;
    auto emit = [&](int op, int nstep, int flag, double dt, double temp, double wstar, double wvdw, double repel, double drop) {
        chr3d_stage s; memset(&s, 0, sizeof s);
        s.op = op; s.nstep = nstep; s.flag = flag; s.dt = dt; s.temp = temp;
        s.w_noe = wstar * p->noe_scale; s.w_bond = wstar * p->bond_k; s.w_vdw = wvdw; s.repel = repel; s.drop = drop;
        v.push_back(s);
    };
    const int n_img = p->images >= 2 ? 2 : 1;
    const int hot8 = p->hot_step / 8, hot2 = p->hot_step / 2;                   // C3D:1644-1645
    struct Hot { int n; double w, vdw, rep; };
    const Hot hot[5] = { {hot8, 0.1, 20.0, 0.5}, {hot8, 0.2, 20.0, 0.5},        // C3D:1649, 1659
                         {hot8, 0.2, 0.01, 0.9}, {hot2, 0.4, 0.003, 0.9},       // C3D:1669-1670, 1680-1681
                         {hot8, 1.0, 0.003, 0.9} };                             // C3D:1691
    for (int im = 0; im < n_img; im++) {
        const int sign = im ? -1 : 1;                                           // C3D:1605
        emit(CHR3D_OP_INIT, 0, sign, 0, 0, 1.0, 20.0, 0.5, 0);                   // C3D:1606-1639
        for (int r = 0; r < 2; r++) emit(CHR3D_OP_MINIMIZE, p->premin_step, 0, 0, 0, 1.0, 20.0, 0.5, 0);   // C3D:1640,1642
        emit(CHR3D_OP_MAXWELL, 0, 0, 0, p->maxwell_temp, 0.1, 20.0, 0.5, 0);     // C3D:1646-1648
        for (const Hot& h : hot) emit(CHR3D_OP_DYNAMICS, h.n, CHR3D_COUPLE_TCOUP, p->hot_dt, p->hot_temp, h.w, h.vdw, h.rep, 0);
        if (sign == 1 && n_img == 2) emit(CHR3D_OP_SAVE, 0, 0, 0, 0, 1.0, 0.003, 0.9, 0);       // C3D:1701-1708
    }
    if (n_img == 2) emit(CHR3D_OP_PICK, 0, 0, 0, 0, 1.0, 0.003, 0.9, 0);         // C3D:1711-1727
    const int ncycle = (int)((p->hot_temp - 0.0) / p->cool_tmpstp);             // C3D:1741
    const int nstep = p->cool_step / ncycle;                                    // C3D:1742
    const double vdw_step = pow(p->cool_vdw_finl / p->cool_vdw_init, 1.0 / ncycle);   // C3D:1743
    const double rad_step = (p->cool_rad_init - p->cool_rad_finl) / ncycle;     // C3D:1744
    double radius = p->cool_rad_init, rep = 0.80, bath = p->hot_temp, kv = p->cool_vdw_init;   // C3D:1745,1731,1757-1758
    for (int ic = 0; ic <= ncycle; ic++) {                                      // C3D:1759-1760
        emit(CHR3D_OP_DYNAMICS, nstep, CHR3D_COUPLE_VSCALE, p->cool_dt, bath, 1.0, kv, rep, 0);   // C3D:1762-1774
        radius = std::max(p->cool_rad_finl, radius - rad_step); rep = radius;   // C3D:1775-1776
        kv = std::min(p->cool_vdw_finl, kv * vdw_step);                         // C3D:1777
        bath -= p->cool_tmpstp;                                                 // C3D:1778
    }
    for (int cyc = 0; cyc < p->pow_cycl; cyc++)                                 // C3D:1798-1803
        emit(CHR3D_OP_MINIMIZE, p->pow_step, 0, 0, 0, 1.0, 1.0, rep, p->pow_drop);
    emit(CHR3D_OP_CENTER, 0, 0, 0, 0, 1.0, 1.0, rep, 0);                         // C3D:1805-1816
    emit(CHR3D_OP_ENERGY, 0, 0, 0, 0, 1.0, 1.0, rep, 0);                         // C3D:1818
    if (out) for (int k = 0; k < (int)v.size() && k < cap; k++) out[k] = v[k];
    return (int)v.size();
}

// ------------------------------ (1) IF -> restraints --------------------------------------------
CHR3D_API int chr3d_if_to_dist(const double* IF, int32_t L, double alpha, double K, int64_t* tenths, int64_t* n_patched) {
    if (!IF || !tenths || L < 1) { set_error("chr3d_if_to_dist: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    DevBuf<double> dIF; DevBuf<long long> dT; long long np = 0;
    if (int rc = if_to_tenths_certified(IF, L, alpha, K, dIF, dT, &np, 0)) return rc;
    CHR3D_CUDA(cudaMemcpy(tenths, dT.p, sizeof(long long) * (size_t)L * L, cudaMemcpyDeviceToHost));
    if (n_patched) *n_patched = np;
    return 0;
}
CHR3D_API int chr3d_if_to_restraints(const double* IF, int32_t L, double alpha, double K, int32_t min_sep,
                                     int32_t* out_i, int32_t* out_j, double* out_d, int64_t cap, int64_t* out_n) {
    if (!IF || L < 1 || !out_n) { set_error("chr3d_if_to_restraints: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    DevBuf<double> dIF; DevBuf<long long> dT;
    if (int rc = if_to_tenths_certified(IF, L, alpha, K, dIF, dT, nullptr, 0)) return rc;
    long long n = 0;
    if (cap <= 0 || !out_i || !out_j || !out_d) {
        if (int rc = dev_tenths_to_list(dT.p, L, min_sep, nullptr, nullptr, nullptr, 0, &n, 0)) return rc;
        *out_n = n; return 0;
    }
    DevBuf<int32_t> ri, rj; DevBuf<double> rd;
    CHR3D_CUDA(ri.alloc((size_t)cap)); CHR3D_CUDA(rj.alloc((size_t)cap)); CHR3D_CUDA(rd.alloc((size_t)cap));
    if (int rc = dev_tenths_to_list(dT.p, L, min_sep, ri.p, rj.p, rd.p, cap, &n, 0)) return rc;
    *out_n = n;
    if (n > cap) { set_error("chr3d_if_to_restraints: %lld rows exceed cap %lld", n, (long long)cap); return CHR3D_ERR_CAPACITY; }
    CHR3D_CUDA(cudaMemcpy(out_i, ri.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    CHR3D_CUDA(cudaMemcpy(out_j, rj.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    CHR3D_CUDA(cudaMemcpy(out_d, rd.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return 0;
}

static void put_tenths(std::string& s, long long t) {
    char b[48];
    if (t < 0) snprintf(b, sizeof b, "-%lld.%lld ", (-t) / 10, (-t) % 10); else snprintf(b, sizeof b, "%lld.%lld ", t / 10, t % 10);
    s += b;
}
static int write_all(const char* path, const std::string& s) {
    FILE* f = fopen(path, "w");
    if (!f) { set_error("cannot open %s for writing", path); return CHR3D_ERR_IO; }
    size_t w = fwrite(s.data(), 1, s.size(), f);
    fclose(f);
    if (w != s.size()) { set_error("short write to %s", path); return CHR3D_ERR_IO; }
    return 0;
}
CHR3D_API int chr3d_write_dist(const char* path, const int64_t* tenths, int32_t L) {
    if (!path || !tenths) { set_error("chr3d_write_dist: bad arguments"); return CHR3D_ERR_ARG; }
    std::string s; s.reserve((size_t)L * L * 8);
    for (int i = 0; i < L; i++) { for (int j = 0; j < L; j++) put_tenths(s, tenths[(size_t)i * L + j]); s += '\n'; }   // C3D:156-161
    return write_all(path, s);
}
CHR3D_API int chr3d_write_rr(const char* path, const int32_t* ri, const int32_t* rj, const double* rd, int64_t n) {
    if (!path || (n > 0 && (!ri || !rj || !rd))) { set_error("chr3d_write_rr: bad arguments"); return CHR3D_ERR_ARG; }
    std::string s; char b[128];
    for (int64_t k = 0; k < n; k++) { snprintf(b, sizeof b, "%d %d %.2f %.2f 1.0\n", ri[k] + 1, rj[k] + 1, rd[k], rd[k]); s += b; }   // C3D:204
    return write_all(path, s);
}
CHR3D_API int chr3d_write_tbl(const char* path, const int32_t* ri, const int32_t* rj, const double* rd, int64_t n) {
    if (!path || (n > 0 && (!ri || !rj || !rd))) { set_error("chr3d_write_tbl: bad arguments"); return CHR3D_ERR_ARG; }
    std::string s; char b[160];
    for (int64_t k = 0; k < n; k++) {   // C3D:352-360 with C2 == C3
        snprintf(b, sizeof b, "assign45 (resid %3d and name %2s) (resid %3d and name %2s) %.2f %.2f %.2f\n",
                 ri[k] + 1, "ca", rj[k] + 1, "ca", (rd[k] + rd[k]) / 2, (rd[k] - rd[k]) / 2, (rd[k] - rd[k]) / 2);
        s += b;
    }
    return write_all(path, s);
}

// ------------------------------ ensemble --------------------------------------------------------
// Rows index beads on the device (coordinates, the dense target table): every row is checked on the host.  `unique`: the
// dense-table paths also refuse a pair that is listed twice (the scatter would keep whichever row landed last).
static int upload_list(JobDev& jd, const int32_t* ri, const int32_t* rj, const double* rd, long long nr, cudaStream_t s, int N, bool unique, const char* who, int vdw_min_sep = 0) {
    for (long long k = 0; k < nr; k++)
        if (ri[k] < 0 || rj[k] < 0 || ri[k] >= N || rj[k] >= N || ri[k] == rj[k] || !(rd[k] > 0.0)) {
            set_error("%s: restraint row %lld (%d, %d, %g) is not a pair of distinct beads 0 .. %d with a positive distance", who, k, ri[k], rj[k], rd[k], N - 1);
            return CHR3D_ERR_ARG;
        }
    // the dense table encodes "restraint" as "restraint + repel": a row between beads closer than vdw_min_sep (bonded neighbours, which
    // the reference's 1-2 exclusion and the oracle keep free of repel) cannot be represented there
    for (long long k = 0; k < nr && vdw_min_sep > 0; k++)
        if (abs(ri[k] - rj[k]) < vdw_min_sep) {
            set_error("%s: restraint row %lld joins beads %d and %d, closer than vdw_min_sep = %d (the dense table would add the repel term to it; the sparse engine, chr3d_single_create_sparse, takes such rows)", who, k, ri[k], rj[k], vdw_min_sep);
            return CHR3D_ERR_ARG;
        }
    if (unique && nr > 1) {
        std::vector<unsigned long long> key((size_t)nr);
        for (long long k = 0; k < nr; k++) { const unsigned a = (unsigned)std::min(ri[k], rj[k]), b = (unsigned)std::max(ri[k], rj[k]); key[(size_t)k] = ((unsigned long long)a << 32) | b; }
        std::sort(key.begin(), key.end());
        for (size_t k = 1; k < key.size(); k++)
            if (key[k] == key[k - 1]) { set_error("%s: the pair (%u, %u) is listed more than once", who, (unsigned)(key[k] >> 32), (unsigned)key[k]); return CHR3D_ERR_ARG; }
    }
    jd.nr = nr;
    CHR3D_CUDA(jd.ri.alloc((size_t)std::max<long long>(nr, 1))); CHR3D_CUDA(jd.rj.alloc((size_t)std::max<long long>(nr, 1)));
    CHR3D_CUDA(jd.rd.alloc((size_t)std::max<long long>(nr, 1)));
    if (nr > 0) {
        CHR3D_CUDA(cudaMemcpyAsync(jd.ri.p, ri, sizeof(int32_t) * (size_t)nr, cudaMemcpyHostToDevice, s));
        CHR3D_CUDA(cudaMemcpyAsync(jd.rj.p, rj, sizeof(int32_t) * (size_t)nr, cudaMemcpyHostToDevice, s));
        CHR3D_CUDA(cudaMemcpyAsync(jd.rd.p, rd, sizeof(double) * (size_t)nr, cudaMemcpyHostToDevice, s));
    }
    return 0;
}

static int ensemble_create(const chr3d_job* jobs, int njobs, const chr3d_sa_params* p, const chr3d_stage* program,
                           int nstages, bool want_forces, std::vector<DevBuf<float>>* fbufs, chr3d_ensemble** out) {
    if (!jobs || njobs < 1 || !p || !program || nstages < 1 || !out) { set_error("ensemble_create: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    Trace tr("create");
    std::unique_ptr<chr3d_ensemble> e(new chr3d_ensemble());
    make_dev_params(*p, e->P);
    e->nstages = nstages;
    std::vector<DevStage> hs(nstages);
    for (int k = 0; k < nstages; k++) make_dev_stage(*p, program[k], hs[k]);
    CHR3D_CUDA(e->prog.alloc(nstages));
    CHR3D_CUDA(cudaMemcpy(e->prog.p, hs.data(), sizeof(DevStage) * nstages, cudaMemcpyHostToDevice));
    std::vector<TrajDesc> tds;
    if (fbufs) { std::vector<DevBuf<float>> t(njobs); fbufs->swap(t); }
    // The two images of a model (INIT(+1) .. SAVE, INIT(-1) .. PICK; C3D:1605-1727) do not interact before PICK.  When few
    // trajectories share the GPU, running them side by side on smaller clusters beats running them one after the other on
    // large ones (20 models of 455 beads: 2 x 7.2 us per pair of evaluations on 6-CTA clusters, 11.7 us on 3-CTA clusters).
    int i_save = -1, i_pick = -1, n_init = 0, n_save = 0, n_pick = 0;
    for (int k = 0; k < nstages; k++) {
        if (program[k].op == CHR3D_OP_INIT) n_init++;
        if (program[k].op == CHR3D_OP_SAVE) { n_save++; i_save = k; }
        if (program[k].op == CHR3D_OP_PICK) { n_pick++; i_pick = k; }
    }
    const bool split_ok = !want_forces && !dev_getenv("CHR3D_NO_SPLIT") && n_init == 2 && n_save == 1 && n_pick == 1 && program[0].op == CHR3D_OP_INIT &&
                          i_save + 1 < i_pick && program[i_save + 1].op == CHR3D_OP_INIT;
    int total_traj = 0;                       // the launch shape of every job is chosen for the whole batch sharing the GPU
    for (int jn = 0; jn < njobs; jn++) total_traj += std::max(jobs[jn].n_models, 0);
    for (int jn = 0; jn < njobs; jn++) {
        const chr3d_job& jb = jobs[jn];
        if (jb.L < 2 || jb.n_models < 1) { set_error("job %d: need L>=2 and n_models>=1", jn); return CHR3D_ERR_ARG; }
        std::unique_ptr<JobDev> jd(new JobDev());
        jd->L = jb.L; jd->n_models = jb.n_models; jd->seed = jb.seed ? jb.seed : p->seed;
        const int N = jb.L, pitch = target_pitch(N);
        CHR3D_CUDA(jd->T.alloc((size_t)N * pitch));
        if (jb.IF) {
            DevBuf<double> dIF; DevBuf<long long> dT;
            if (int rc = if_to_tenths_certified(jb.IF, N, jb.alpha, jb.K, dIF, dT, nullptr, 0)) return rc;
            tr.lap("tenths");
            if (int rc = dev_tenths_to_targets(dT.p, N, jb.min_sep, p->vdw_min_sep, jd->T.p, 0)) return rc;
            // rows are at most the cells with j - i >= min_sep: room for all of them, so that counting and filling is one pass
            long long n = 0, room = 0;
            for (int i = 0; i < N; i++) room += N - std::min(N, std::max(0, i + jb.min_sep));
            room = std::max<long long>(room, 1);
            CHR3D_CUDA(jd->ri.alloc((size_t)room)); CHR3D_CUDA(jd->rj.alloc((size_t)room)); CHR3D_CUDA(jd->rd.alloc((size_t)room));
            if (int rc = dev_tenths_to_list(dT.p, N, jb.min_sep, jd->ri.p, jd->rj.p, jd->rd.p, room, &n, 0)) return rc;
            jd->nr = n;
        } else {
            if (jb.nr > 0 && (!jb.ri || !jb.rj || !jb.rd)) { set_error("job %d: restraint table pointers missing", jn); return CHR3D_ERR_ARG; }
            if (int rc = upload_list(*jd, jb.ri, jb.rj, jb.rd, jb.nr, 0, N, true, "chr3d_build_models", p->vdw_min_sep)) return rc;
            if (int rc = dev_list_to_targets(jd->ri.p, jd->rj.p, jd->rd.p, jb.nr, N, p->vdw_min_sep, jd->T.p, 0)) return rc;
        }
        CHR3D_CUDA(jd->Tt.alloc(sa_tiles_floats(N)));
        if (int rc = sa_pack_tiles(jd->T.p, N, pitch, jd->Tt.p, 0)) return rc;
        tr.lap("targets+list");
        e->rows_total += jd->nr * jb.n_models;
        const size_t nx = (size_t)jb.n_models * N * 3;
        CHR3D_CUDA(jd->x0.alloc(nx)); CHR3D_CUDA(jd->xout.alloc(nx)); CHR3D_CUDA(jd->scratch.alloc(2 * nx));
        CHR3D_CUDA(jd->xyz_d.alloc(nx)); CHR3D_CUDA(jd->e.alloc((size_t)jb.n_models * 4)); CHR3D_CUDA(jd->nev.alloc(jb.n_models));
        CHR3D_CUDA(jd->hand.alloc(jb.n_models)); CHR3D_CUDA(jd->sat.alloc(jb.n_models)); CHR3D_CUDA(jd->dev.alloc(jb.n_models));
        CHR3D_CUDA(cudaMemset(jd->e.p, 0, sizeof(double) * 4 * jb.n_models));
        if (jb.x0) {
            CHR3D_CUDA(cudaMemcpy(jd->xyz_d.p, jb.x0, sizeof(double) * nx, cudaMemcpyHostToDevice));
            d2f_kernel<<<(unsigned)((nx + 255) / 256), 256>>>(jd->xyz_d.p, jd->x0.p, (long long)nx);
            CHR3D_CUDA(cudaGetLastError());
        } else {
            if (int rc = dev_random_coil(jd->x0.p, N, jb.n_models, 0, p->bond_r0, jd->seed, 0)) return rc;
        }
        if (want_forces) CHR3D_CUDA((*fbufs)[jn].alloc(nx));
        tr.lap("alloc+coil");
        jd->has_x0 = jb.x0 != nullptr;
        if (int rc = sa_choose_config(N, total_traj, e->P.m, &jd->cfg)) {
            if (rc != CHR3D_ERR_SIZE) return rc;
            // too large for one thread-block cluster: the single-structure engine anneals the models one after the other
            if (int rc2 = chr3d_single_create_dense(jd->T.p, pitch, N, jd->seed, p, program, nstages, jd->nr, &jd->big)) return rc2;
            jd->cfg.cluster = 0; jd->cfg.threads = 0;
            e->jobs.push_back(std::move(jd));
            continue;
        }
        if (split_ok) {
            SaLaunchCfg c2;
            if (sa_choose_config(N, 2 * total_traj, e->P.m, &c2) == 0 && c2.cost < 1.9 * jd->cfg.cost) {
                jd->split = true; jd->cfg_img = c2;
                CHR3D_CUDA(jd->st_img.alloc(2 * nx)); CHR3D_CUDA(jd->esaved.alloc(jb.n_models)); CHR3D_CUDA(jd->nev_a.alloc(jb.n_models));
                CHR3D_CUDA(jd->nev_b.alloc(jb.n_models)); CHR3D_CUDA(jd->hand_tmp.alloc(2 * (size_t)jb.n_models));
            }
        }
        e->jobs.push_back(std::move(jd));
    }
    // Launch units: (job, phase).  Phase 1 holds whole trajectories of unsplit jobs and the two image trajectories of split
    // jobs; phase 2 (after all of phase 1) holds the trajectories that resume at PICK.  Inside a phase units are grouped by
    // launch shape, larger structures first (they finish last otherwise); one launch per group, groups of a phase run
    // concurrently on their own streams.
    {
        struct Unit { int job, phase; SaLaunchCfg cfg; int L; std::vector<TrajDesc> t; };
        std::vector<Unit> units;
        for (size_t jn = 0; jn < e->jobs.size(); jn++) {
            JobDev& jd = *e->jobs[jn];
            if (jd.big) continue;
            const int N = jd.L, pitch = target_pitch(N);
            auto base = [&](int m) {
                TrajDesc td; memset(&td, 0, sizeof td);
                td.T = jd.Tt.p; td.x0 = jd.x0.p + (size_t)m * N * 3; td.x_out = jd.xout.p + (size_t)m * N * 3;
                td.f_out = want_forces ? (*fbufs)[jn].p + (size_t)m * N * 3 : nullptr;
                td.scratch = jd.scratch.p + (size_t)m * N * 6; td.e_out = jd.e.p + 4 * m; td.nevals_out = jd.nev.p + m;
                td.hand_out = jd.hand.p + m; td.N = N; td.pitch = pitch; td.model = m; td.seed = jd.seed;
                return td;
            };
            Unit u1; u1.job = (int)jn; u1.phase = 1; u1.L = N; u1.cfg = jd.split ? jd.cfg_img : jd.cfg;
            Unit u2; u2.job = (int)jn; u2.phase = 2; u2.L = N; u2.cfg = jd.cfg;
            for (int m = 0; m < jd.n_models; m++) {
                if (!jd.split) { u1.t.push_back(base(m)); continue; }
                float* st = jd.st_img.p + (size_t)m * N * 6;
                TrajDesc a = base(m);          // image +1: INIT .. SAVE (x, v -> scratch, energy -> esaved)
                a.k0 = 0; a.k1 = i_save + 1; a.x_out = a.scratch; a.esaved_out = jd.esaved.p + m; a.nevals_out = jd.nev_a.p + m; a.hand_out = jd.hand_tmp.p + 2 * m;
                TrajDesc b = base(m);          // image -1: INIT .. last hot DYNAMICS (x, v -> st)
                b.k0 = i_save + 1; b.k1 = i_pick; b.x_out = st; b.state_out = st; b.nevals_out = jd.nev_b.p + m; b.hand_out = jd.hand_tmp.p + 2 * m + 1;
                TrajDesc c = base(m);          // PICK .. ENERGY, starting from image -1's end state
                c.k0 = i_pick; c.k1 = nstages; c.state_in = st; c.resume_image = program[i_save + 1].flag; c.esaved_in = jd.esaved.p + m;
                u1.t.push_back(a); u1.t.push_back(b); u2.t.push_back(c);
            }
            units.push_back(std::move(u1));
            if (jd.split) units.push_back(std::move(u2));
        }
        std::stable_sort(units.begin(), units.end(), [](const Unit& x, const Unit& y) {
            if (x.phase != y.phase) return x.phase < y.phase;
            if (x.cfg.cluster != y.cfg.cluster) return x.cfg.cluster > y.cfg.cluster;
            if (x.cfg.threads != y.cfg.threads) return x.cfg.threads > y.cfg.threads;
            return x.L > y.L;
        });
        for (const Unit& u : units) {
            const int pos = (int)tds.size();
            tds.insert(tds.end(), u.t.begin(), u.t.end());
            if (!e->groups.empty()) {
                LaunchGroup& g = e->groups.back();
                if (g.phase == u.phase && g.cfg.cluster == u.cfg.cluster && g.cfg.threads == u.cfg.threads) {
                    g.n_traj += (int)u.t.size(); g.cfg.smem = std::max(g.cfg.smem, u.cfg.smem); g.Nmax = std::max(g.Nmax, u.L); continue;
                }
            }
            LaunchGroup g; g.cfg = u.cfg; g.traj0 = pos; g.n_traj = (int)u.t.size(); g.Nmax = u.L; g.phase = u.phase;
            e->groups.push_back(g);
        }
    }
    if (e->groups.size() > 1 && !dev_getenv("CHR3D_SERIAL_GROUPS")) {      // CHR3D_SERIAL_GROUPS: development switch, one launch after the other
        e->gstreams.resize(e->groups.size()); e->gevents.resize(e->groups.size());
        for (size_t k = 0; k < e->groups.size(); k++) {
            CHR3D_CUDA(cudaStreamCreateWithFlags(&e->gstreams[k], cudaStreamNonBlocking));
            CHR3D_CUDA(cudaEventCreateWithFlags(&e->gevents[k], cudaEventDisableTiming));
        }
        CHR3D_CUDA(cudaEventCreateWithFlags(&e->evmid, cudaEventDisableTiming));
    }
    e->n_traj = (int)tds.size();
    CHR3D_CUDA(e->traj.alloc(std::max<size_t>(tds.size(), 1)));
    if (!tds.empty()) CHR3D_CUDA(cudaMemcpy(e->traj.p, tds.data(), sizeof(TrajDesc) * tds.size(), cudaMemcpyHostToDevice));
    CHR3D_CUDA(cudaEventCreate(&e->ev0)); CHR3D_CUDA(cudaEventCreate(&e->ev1));
    CHR3D_CUDA(cudaDeviceSynchronize());
    tr.lap("finish");
    *out = e.release();
    return 0;
}

CHR3D_API int chr3d_ensemble_create(const chr3d_job* jobs, int32_t njobs, const chr3d_sa_params* p,
                                    const chr3d_stage* program, int32_t nstages, chr3d_ensemble** out) {
    return ensemble_create(jobs, njobs, p, program, nstages, false, nullptr, out);
}

CHR3D_API int chr3d_ensemble_run(chr3d_ensemble* e, void* stream) {
    if (!e) { set_error("chr3d_ensemble_run: null handle"); return CHR3D_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    e->n_launches = 0;
    CHR3D_CUDA(cudaEventRecord(e->ev0, s));
    bool mid_recorded = false;
    for (size_t k = 0; k < e->groups.size(); k++) {
        const LaunchGroup& g = e->groups[k];
        cudaStream_t gs = e->gstreams.empty() ? s : e->gstreams[k];
        if (g.phase == 2 && !mid_recorded && gs != s) { CHR3D_CUDA(cudaEventRecord(e->evmid, s)); mid_recorded = true; }   // s has joined every phase-1 group
        if (gs != s) CHR3D_CUDA(cudaStreamWaitEvent(gs, g.phase == 2 ? e->evmid : e->ev0, 0));
        if (int rc = sa_launch(e->traj.p + g.traj0, g.n_traj, g.Nmax, e->P, e->prog.p, e->nstages, g.cfg, gs)) return rc;
        if (gs != s) { CHR3D_CUDA(cudaEventRecord(e->gevents[k], gs)); CHR3D_CUDA(cudaStreamWaitEvent(s, e->gevents[k], 0)); }
        e->n_launches++;
    }
    for (auto& jp : e->jobs) {
        JobDev& jd = *jp;
        if (!jd.big) continue;
        for (int m = 0; m < jd.n_models; m++) {
            const size_t off = (size_t)m * jd.L * 3;
            if (int rc = chr3d_single_run_model(jd.big, m, jd.has_x0 ? jd.x0.p + off : nullptr, s, jd.xout.p + off, jd.e.p + 4 * m, jd.nev.p + m, jd.hand.p + m)) return rc;
            e->n_launches++;
        }
    }
    CHR3D_CUDA(cudaEventRecord(e->ev1, s));
    e->ran = true;
    return 0;
}

CHR3D_API int chr3d_ensemble_fetch(chr3d_ensemble* e, chr3d_result* results) {
    if (!e || !results) { set_error("chr3d_ensemble_fetch: bad arguments"); return CHR3D_ERR_ARG; }
    Trace tr("fetch");
    CHR3D_CUDA(cudaDeviceSynchronize());
    tr.lap("wait_kernels");
    float ms = 0.f;
    if (e->ran) CHR3D_CUDA(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    for (size_t jn = 0; jn < e->jobs.size(); jn++) {
        JobDev& jd = *e->jobs[jn]; chr3d_result& r = results[jn];
        const size_t nx = (size_t)jd.n_models * jd.L * 3;
        f2d_kernel<<<(unsigned)((nx + 255) / 256), 256>>>(jd.xout.p, jd.xyz_d.p, (long long)nx);
        CHR3D_CUDA(cudaGetLastError());
        if (r.satisfied || r.sum_dev) {
            if (int rc = dev_assess(jd.xyz_d.p, jd.n_models, jd.L, jd.ri.p, jd.rj.p, jd.rd.p, jd.nr, jd.sat.p, jd.dev.p, 0)) return rc;
            if (r.satisfied) CHR3D_CUDA(cudaMemcpy(r.satisfied, jd.sat.p, sizeof(int32_t) * jd.n_models, cudaMemcpyDeviceToHost));
            if (r.sum_dev) CHR3D_CUDA(cudaMemcpy(r.sum_dev, jd.dev.p, sizeof(double) * jd.n_models, cudaMemcpyDeviceToHost));
        }
        if (r.xyz) CHR3D_CUDA(cudaMemcpy(r.xyz, jd.xyz_d.p, sizeof(double) * nx, cudaMemcpyDeviceToHost));
        if (r.energies) CHR3D_CUDA(cudaMemcpy(r.energies, jd.e.p, sizeof(double) * 4 * jd.n_models, cudaMemcpyDeviceToHost));
        if (r.force_evals) {
            CHR3D_CUDA(cudaMemcpy(r.force_evals, jd.nev.p, sizeof(long long) * jd.n_models, cudaMemcpyDeviceToHost));
            if (jd.split) {       // evaluations of the two image trajectories belong to the model as well
                std::vector<long long> a(jd.n_models), b(jd.n_models);
                CHR3D_CUDA(cudaMemcpy(a.data(), jd.nev_a.p, sizeof(long long) * jd.n_models, cudaMemcpyDeviceToHost));
                CHR3D_CUDA(cudaMemcpy(b.data(), jd.nev_b.p, sizeof(long long) * jd.n_models, cudaMemcpyDeviceToHost));
                for (int m = 0; m < jd.n_models; m++) r.force_evals[m] += a[m] + b[m];
            }
        }
        if (r.hand) CHR3D_CUDA(cudaMemcpy(r.hand, jd.hand.p, sizeof(int32_t) * jd.n_models, cudaMemcpyDeviceToHost));
        r.n_restraints = jd.nr; r.device_ms = ms;
        tr.lap("assess+d2h");
    }
    return 0;
}
CHR3D_API int chr3d_ensemble_info(const chr3d_ensemble* e, int32_t* cluster_size, int32_t* threads, int32_t* n_launches,
                                  int64_t* restraint_rows_total) {
    if (!e) { set_error("chr3d_ensemble_info: bad handle"); return CHR3D_ERR_ARG; }
    if (cluster_size) *cluster_size = e->groups.empty() ? 0 : e->groups[0].cfg.cluster;      // 0: only single-structure engine jobs
    if (threads) *threads = e->groups.empty() ? 0 : e->groups[0].cfg.threads;
    if (n_launches) *n_launches = (int)e->groups.size();
    if (restraint_rows_total) *restraint_rows_total = e->rows_total;
    return 0;
}
CHR3D_API int chr3d_ensemble_destroy(chr3d_ensemble* e) { delete e; return 0; }

CHR3D_API int chr3d_run_ensemble(const chr3d_job* jobs, int32_t njobs, const chr3d_sa_params* p,
                                 const chr3d_stage* program, int32_t nstages, chr3d_result* results) {
    chr3d_ensemble* e = nullptr;
    if (int rc = chr3d_ensemble_create(jobs, njobs, p, program, nstages, &e)) return rc;
    Trace tr("run_ensemble");
    int rc = chr3d_ensemble_run(e, nullptr);
    if (!rc) rc = chr3d_ensemble_fetch(e, results);
    tr.lap("run+fetch");
    chr3d_ensemble_destroy(e);
    tr.lap("destroy");
    return rc;
}

// ------------------------------ (2) energy + forces ---------------------------------------------
CHR3D_API int chr3d_energy_forces(const double* xyz, int32_t n_models, int32_t N, const int32_t* ri, const int32_t* rj,
                                  const double* rd, int64_t nr, const chr3d_sa_params* p, const chr3d_stage* weights,
                                  double* energies, double* forces) {
    if (!xyz || n_models < 1 || N < 2 || !p || !weights || !energies) { set_error("chr3d_energy_forces: bad arguments"); return CHR3D_ERR_ARG; }
    chr3d_job jb; memset(&jb, 0, sizeof jb);
    jb.L = N; jb.ri = ri; jb.rj = rj; jb.rd = rd; jb.nr = nr; jb.n_models = n_models; jb.x0 = xyz;
    chr3d_stage prog[2]; memset(prog, 0, sizeof prog);
    prog[0] = *weights; prog[0].op = CHR3D_OP_INIT; prog[0].flag = 1;
    prog[1] = *weights; prog[1].op = CHR3D_OP_ENERGY;
    std::vector<DevBuf<float>> fb;
    chr3d_ensemble* e = nullptr;
    if (int rc = ensemble_create(&jb, 1, p, prog, 2, forces != nullptr, &fb, &e)) return rc;
    std::unique_ptr<chr3d_ensemble> guard(e);
    if (int rc = chr3d_ensemble_run(e, nullptr)) return rc;
    chr3d_result r; memset(&r, 0, sizeof r); r.energies = energies;
    if (int rc = chr3d_ensemble_fetch(e, &r)) return rc;
    if (forces) {
        const size_t nx = (size_t)n_models * N * 3;
        std::vector<float> hf(nx);
        CHR3D_CUDA(cudaMemcpy(hf.data(), fb[0].p, sizeof(float) * nx, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < nx; k++) forces[k] = hf[k];
    }
    return 0;
}

// ------------------------------ (4) assessment / ranking / PDB ----------------------------------
CHR3D_API int chr3d_assess(const double* xyz, int32_t n_models, int32_t N, const int32_t* ri, const int32_t* rj,
                           const double* rd, int64_t nr, int32_t* satisfied, double* sum_dev) {
    if (!xyz || n_models < 1 || N < 1 || !satisfied || !sum_dev || (nr > 0 && (!ri || !rj || !rd))) { set_error("chr3d_assess: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    JobDev jd;
    if (int rc = upload_list(jd, ri, rj, rd, nr, 0, N, false, "chr3d_assess")) return rc;
    const size_t nx = (size_t)n_models * N * 3;
    CHR3D_CUDA(jd.xyz_d.alloc(nx)); CHR3D_CUDA(jd.sat.alloc(n_models)); CHR3D_CUDA(jd.dev.alloc(n_models));
    CHR3D_CUDA(cudaMemcpy(jd.xyz_d.p, xyz, sizeof(double) * nx, cudaMemcpyHostToDevice));
    if (int rc = dev_assess(jd.xyz_d.p, n_models, N, jd.ri.p, jd.rj.p, jd.rd.p, nr, jd.sat.p, jd.dev.p, 0)) return rc;
    CHR3D_CUDA(cudaMemcpy(satisfied, jd.sat.p, sizeof(int32_t) * n_models, cudaMemcpyDeviceToHost));
    CHR3D_CUDA(cudaMemcpy(sum_dev, jd.dev.p, sizeof(double) * n_models, cudaMemcpyDeviceToHost));
    return 0;
}
CHR3D_API int chr3d_assess_rows(const double* xyz, int32_t N, const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr,
                                int32_t* flag, double* dev, double* dist) {
    if (!xyz || N < 1 || nr < 0 || (nr > 0 && (!ri || !rj || !rd || !flag || !dev || !dist))) { set_error("chr3d_assess_rows: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    JobDev jd;
    if (int rc = upload_list(jd, ri, rj, rd, nr, 0, N, false, "chr3d_assess_rows")) return rc;
    DevBuf<int32_t> df; DevBuf<double> dd, ds;
    const size_t n1 = (size_t)std::max<long long>(nr, 1);
    CHR3D_CUDA(jd.xyz_d.alloc((size_t)N * 3)); CHR3D_CUDA(df.alloc(n1)); CHR3D_CUDA(dd.alloc(n1)); CHR3D_CUDA(ds.alloc(n1));
    CHR3D_CUDA(cudaMemcpy(jd.xyz_d.p, xyz, sizeof(double) * (size_t)N * 3, cudaMemcpyHostToDevice));
    if (int rc = dev_assess_rows(jd.xyz_d.p, N, jd.ri.p, jd.rj.p, jd.rd.p, nr, df.p, dd.p, ds.p, 0)) return rc;
    if (nr > 0) {
        CHR3D_CUDA(cudaMemcpy(flag, df.p, sizeof(int32_t) * (size_t)nr, cudaMemcpyDeviceToHost));
        CHR3D_CUDA(cudaMemcpy(dev, dd.p, sizeof(double) * (size_t)nr, cudaMemcpyDeviceToHost));
        CHR3D_CUDA(cudaMemcpy(dist, ds.p, sizeof(double) * (size_t)nr, cudaMemcpyDeviceToHost));
    }
    return 0;
}
CHR3D_API int chr3d_rank_models(const double* energies, int32_t n_models, int32_t* order) {
    if (!energies || !order || n_models < 1) { set_error("chr3d_rank_models: bad arguments"); return CHR3D_ERR_ARG; }
    std::vector<int32_t> idx(n_models);
    for (int i = 0; i < n_models; i++) idx[i] = i;
    // int() truncation of the NOE energy (C3D:617), ascending (C3D:823); ties by model index
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
        return (long long)energies[4 * a + CHR3D_E_NOE] < (long long)energies[4 * b + CHR3D_E_NOE]; });
    for (int i = 0; i < n_models; i++) order[i] = idx[i];
    return 0;
}
CHR3D_API int chr3d_write_pdb(const char* path, const double* xyz, int32_t N, const double* e4, const char* resnames,
                              int32_t with_conect) {
    if (!path || !xyz || N < 1) { set_error("chr3d_write_pdb: bad arguments"); return CHR3D_ERR_ARG; }
    std::string s; char b[160];
    if (e4) {   // the REMARK lines get_cns_energy() greps for (C3D:602-618)
        snprintf(b, sizeof b, "REMARK FILENAME=\"%s\"\n", path); s += b;
        snprintf(b, sizeof b, "REMARK overall = %.4f\n", e4[CHR3D_E_OVERALL]); s += b;
        snprintf(b, sizeof b, "REMARK bon = %.4f\n", e4[CHR3D_E_BON]); s += b;
        snprintf(b, sizeof b, "REMARK vdw = %.4f\n", e4[CHR3D_E_VDW]); s += b;
        snprintf(b, sizeof b, "REMARK noe = %.4f\n", e4[CHR3D_E_NOE]); s += b;
    }
    const size_t nref = strlen(REFSEQ);
    for (int i = 0; i < N; i++) {
        char rn[4] = {0, 0, 0, 0};
        if (resnames) memcpy(rn, resnames + 3 * (size_t)i, 3); else memcpy(rn, aa3(REFSEQ[(size_t)i % nref]), 3);
        // column layout of SURVEY App. B.5 (assembled by reindex_chain, C3D:854-856)
        snprintf(b, sizeof b, "ATOM  %5d  CA  %3s  %4d    %8.3f%8.3f%8.3f  1.00  0.00           C  \n",
                 (i + 1) % 100000, rn, (i + 1) % 10000, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        s += b;
    }
    if (with_conect) {
        s += "\n";                                                       // line left by sed s/END//g (C3D:818)
        for (int i = 1; i < N; i++) { snprintf(b, sizeof b, "CONECT%5d%5d\n", i, i + 1); s += b; }   // C3D:212
    }
    s += "END\n";
    return write_all(path, s);
}

// ------------------------------ (5) Spearman ----------------------------------------------------
CHR3D_API int chr3d_spearman(const double* IF, int32_t L, const double* xyz, int32_t n_models, int32_t range, double* rho) {
    if (!IF || !xyz || !rho || L < 1 || n_models < 1 || range < 0) { set_error("chr3d_spearman: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    DevBuf<double> dIF, dx;
    CHR3D_CUDA(dIF.alloc((size_t)L * L)); CHR3D_CUDA(dx.alloc((size_t)n_models * L * 3));
    CHR3D_CUDA(cudaMemcpy(dIF.p, IF, sizeof(double) * (size_t)L * L, cudaMemcpyHostToDevice));
    CHR3D_CUDA(cudaMemcpy(dx.p, xyz, sizeof(double) * (size_t)n_models * L * 3, cudaMemcpyHostToDevice));
    return dev_spearman(dIF.p, L, dx.p, n_models, range, rho, 0);
}

// chr3d_default_params plus the default-off bead-model stand-ins for what CNS's full-atom poly-GLY chain adds, as fitted on the 45
// shipped CNS models (profiles/r02_cns_pin.md): soft CA(i)-CA(i+2) lower bound, larger repel radius.
CHR3D_API void chr3d_cns_pin_params(chr3d_sa_params* p) {
    if (!p) return;
    chr3d_default_params(p);
    p->angle_d2 = 6.0; p->angle_k = 100.0; p->vdw_rmin = 5.5;
}

// Launch-shape options of the ensemble kernel (process-wide; unset = chosen by the cost model of sa_choose_config):
//   "sa_cluster" CTAs per trajectory (1..8), "sa_warps" warps per CTA (4..12), "sa_min_smem" bytes of shared memory a CTA asks
//   for at least (0 lets several CTAs share an SM), "sa_split" = "0": never run the two hot-MD images as trajectories of their own.
// value NULL or "" clears the option.  Options are read when an ensemble is created.
CHR3D_API int chr3d_set_option(const char* name, const char* value) {
    static const char* const known[][2] = {{"sa_cluster", "CHR3D_SA_CLUSTER"}, {"sa_warps", "CHR3D_SA_WARPS"}, {"sa_min_smem", "CHR3D_SA_MIN_SMEM"}, {"sa_split", "CHR3D_NO_SPLIT"},
                                           {"seqsum", "CHR3D_SEQSUM"}};
    if (!name) { set_error("chr3d_set_option: null name"); return CHR3D_ERR_ARG; }
    for (auto& k : known) {
        if (strcmp(name, k[0]) != 0) continue;
        std::lock_guard<std::mutex> lk(chr3d::g_opt_mu);
        auto& t = chr3d::opt_table();
        const bool clear = !value || !*value;
        if (strcmp(k[0], "sa_split") == 0) {          // stored under the name of the switch it replaces: present = no split
            if (clear || strcmp(value, "0") != 0) t.erase(k[1]); else t[k[1]] = "1";
        } else if (strcmp(k[0], "seqsum") == 0) {     // "serial" or cleared (= the parallel evaluation)
            if (clear) t.erase(k[1]);
            else if (strcmp(value, "serial") == 0) t[k[1]] = "serial";
            else { set_error("chr3d_set_option: seqsum wants \"serial\" or nothing, got '%s'", value); return CHR3D_ERR_ARG; }
        } else if (clear) t.erase(k[1]);
        else {
            char* end = nullptr; const long v = strtol(value, &end, 10);
            if (!end || *end || v < 0) { set_error("chr3d_set_option: %s wants a non-negative integer, got '%s'", name, value); return CHR3D_ERR_ARG; }
            t[k[1]] = value;
        }
        return 0;
    }
    set_error("chr3d_set_option: unknown option '%s' (sa_cluster, sa_warps, sa_min_smem, sa_split, seqsum)", name);
    return CHR3D_ERR_ARG;
}
