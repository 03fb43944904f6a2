// assess.cu — post-run scoring on the device, exact with respect to the Perl text pipeline.
//
//  * dev_assess   : count_satisfied_tbl_rows (C3D:447-485), sum_noe_dev (C3D:581-600) on the
//                   3-decimal PDB text coordinates and "%.3f" distances (calc_dist, C3D:716-729).
//  * dev_spearman : spearman_IF_pdb.pl (SP:32-71): Spearman rho between IF values and "%.3f"
//                   model distances over all ordered pairs |i-j| >= range, average ranks for ties
//                   (Statistics::RankCorrelation 0.1204, restated — SURVEY App. A.8).
//
// Everything that the Perl does through decimal text is done here in exact integer arithmetic
// (coordinates and distances in milli-Angstrom); the only floating point left is the handful of
// fp64 operations the Perl itself performs on those values, in the same order.
#include <cub/cub.cuh>
#include "common.h"
#include "seqsum.cuh"

namespace chr3d {

// exact round(x * 1000) with printf's semantics (round-half-even on the exact binary value)
__device__ __forceinline__ long long milli(double x) {
    double n = rint(x * 1000.0);
    double r = fma(x, 1000.0, -n);                 // exact residual
    if (r > 0.5) n += 1.0; else if (r < -0.5) n -= 1.0;
    else if (r == 0.5) { if (fmod(n, 2.0) != 0.0) n += 1.0; }
    else if (r == -0.5) { if (fmod(n, 2.0) != 0.0) n -= 1.0; }
    return (long long)n;
}
// "%.3f" of sqrt(D2 mA^2): n with (2n-1)^2 <= 4 D2 < (2n+1)^2
__device__ __forceinline__ long long round_sqrt(unsigned long long d2) {
    unsigned long long v = 4ull * d2;
    unsigned long long s = (unsigned long long)sqrt((double)v);
    while (s * s > v) s--;
    while ((s + 1) * (s + 1) <= v) s++;
    return (long long)((s + 1) >> 1);
}

__global__ void milli_kernel(const double* __restrict__ xyz, long long n, long long* __restrict__ out) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = milli(xyz[k]);
}

// one thread per (model, restraint): satisfied count (integer atomics) and the deviation term,
// stored at its (i,j) slot of a dense per-model matrix so that the sum can follow the key order
// of `sort keys %tbl_hash` (ascending i then j for the %3d-padded rows, C3D:458, 590).
__global__ void assess_terms_kernel(const long long* __restrict__ mxyz, int n_models, int N,
                                    const int32_t* __restrict__ ri, const int32_t* __restrict__ rj,
                                    const double* __restrict__ rd, long long nr,
                                    int* __restrict__ sat, double* __restrict__ dense, double* __restrict__ listterms) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int m = blockIdx.y;
    if (k >= nr) return;
    int i = ri[k], j = rj[k];
    const long long* a = mxyz + ((size_t)m * N + i) * 3; const long long* b = mxyz + ((size_t)m * N + j) * 3;
    long long dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    unsigned long long d2 = (unsigned long long)(dx * dx) + (unsigned long long)(dy * dy) + (unsigned long long)(dz * dz);
    double dist = (double)round_sqrt(d2) / 1000.0;       // the value Perl reads back from "%.3f"
    if (!(dist < 1000.0)) dist = 1000.0;                 // ssnoe_tbl_min_pdb_dist starts its minimum at 1000.0 (C3D:538, 546)
    const double D0 = rd[k], D1 = 0.0, D2 = 0.0, RELAX = 0.5;      // C3D:74, 352-354
    int c = 0;
    if (dist < (D0 + D2 + RELAX)) c++;                   // C3D:464
    if (dist < (D0 - D1 - RELAX)) c--;                   // C3D:469
    if (c) atomicAdd(&sat[m], c);
    double dev = 0.0;
    if (dist > (D0 + D2 + 0.2)) dev += dist - (D0 + D2); // C3D:592-594
    if (dist < (D0 - D1 - 0.2)) dev += (D0 - D1) - dist; // C3D:595-597
    if (dense) dense[((size_t)m * N + i) * N + j] = dev; else listterms[(size_t)m * nr + k] = dev;
}
// per-restraint rows of contact_violation.txt for one model (count_satisfied_tbl_rows, C3D:458-475):
// violation flag, deviation and the "%.3f" model distance
__global__ void assess_rows_kernel(const long long* __restrict__ mxyz, const int32_t* __restrict__ ri, const int32_t* __restrict__ rj,
                                   const double* __restrict__ rd, long long nr, int32_t* __restrict__ flag, double* __restrict__ devv,
                                   double* __restrict__ distv) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nr) return;
    const long long* a = mxyz + (size_t)ri[k] * 3; const long long* b = mxyz + (size_t)rj[k] * 3;
    long long dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    unsigned long long d2 = (unsigned long long)(dx * dx) + (unsigned long long)(dy * dy) + (unsigned long long)(dz * dz);
    double dist = (double)round_sqrt(d2) / 1000.0;
    if (!(dist < 1000.0)) dist = 1000.0;                     // C3D:538, 546
    const double D0 = rd[k], D1 = 0.0, D2 = 0.0, RELAX = 0.5;
    int vf = 1; double dev = dist - (D0 + D2);               // C3D:459, 463
    if (dist < (D0 + D2 + RELAX)) { vf = 0; dev = 0.0; }     // C3D:464-468
    if (dist < (D0 - D1 - RELAX)) { vf = 1; dev = -(D0 - D1 - dist); }   // C3D:469-473
    flag[k] = vf; devv[k] = dev; distv[k] = dist;
}
// large tables: fixed-shape tree sum (no Perl reference exists at that size)
__global__ void block_sum_kernel(const double* __restrict__ terms, long long per_model, double* __restrict__ out) {
    int m = blockIdx.x;
    const double* t = terms + (size_t)m * per_model;
    double s = 0.0;
    for (long long k = threadIdx.x; k < per_model; k += blockDim.x) s += t[k];
    __shared__ double sh[256];
    sh[threadIdx.x] = s; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[m] = sh[0];
}

// up to 999 beads the `sort keys %tbl_hash` order (string order of the %3d-padded rows) is the numeric (i, j) order, which the dense
// sequential sum reproduces; beyond that the Perl could not run anyway (C3D:93-94) and a fixed tree order is used
#define ASSESS_DENSE_MAX_N 999

int dev_assess(const double* d_xyz, int n_models, int N, const int32_t* d_ri, const int32_t* d_rj,
               const double* d_rd, long long nr, int32_t* d_sat, double* d_dev, cudaStream_t s) {
    DevBuf<long long> mx; DevBuf<double> terms;
    const long long nc = (long long)n_models * N * 3;
    CHR3D_CUDA(mx.alloc((size_t)nc));
    milli_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(d_xyz, nc, mx.p);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaMemsetAsync(d_sat, 0, sizeof(int32_t) * n_models, s));
    CHR3D_CUDA(cudaMemsetAsync(d_dev, 0, sizeof(double) * n_models, s));
    if (nr == 0) return 0;
    const bool dense = N <= ASSESS_DENSE_MAX_N;
    const long long per_model = dense ? (long long)N * N : nr;
    CHR3D_CUDA(terms.alloc((size_t)per_model * n_models));
    if (dense) CHR3D_CUDA(cudaMemsetAsync(terms.p, 0, sizeof(double) * (size_t)per_model * n_models, s));
    dim3 grid((unsigned)((nr + 255) / 256), n_models);
    assess_terms_kernel<<<grid, 256, 0, s>>>(mx.p, n_models, N, d_ri, d_rj, d_rd, nr, d_sat, dense ? terms.p : nullptr, dense ? nullptr : terms.p);
    CHR3D_CUDA(cudaGetLastError());
    if (dense) seq_sum_launch(terms.p, per_model, n_models, d_dev, s);
    else block_sum_kernel<<<n_models, 256, 0, s>>>(terms.p, per_model, d_dev);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int dev_assess_rows(const double* d_xyz, int N, const int32_t* d_ri, const int32_t* d_rj, const double* d_rd, long long nr,
                    int32_t* d_flag, double* d_dev, double* d_dist, cudaStream_t s) {
    DevBuf<long long> mx;
    CHR3D_CUDA(mx.alloc((size_t)N * 3));
    milli_kernel<<<(unsigned)((N * 3 + 255) / 256), 256, 0, s>>>(d_xyz, (long long)N * 3, mx.p);
    if (nr > 0) assess_rows_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(mx.p, d_ri, d_rj, d_rd, nr, d_flag, d_dev, d_dist);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Spearman
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_key(double v) {
    if (v == 0.0) v = 0.0;                            // -0 and +0 tie numerically
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// pair index -> (h, i) for the list of ordered pairs with |i-h| >= range in row-major order
__device__ __forceinline__ void pair_at(long long L, long long range, long long h, long long c, long long* i) {
    // row h keeps columns [0, h-range] and [h+range, L-1]
    if (range <= 0) { *i = c; return; }
    long long left = h - range + 1; if (left < 0) left = 0;
    *i = c < left ? c : (h + range + (c - left));
}
__global__ void row_offsets_kernel(int L, int range, long long* offs) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    long long s = 0;
    for (int h = 0; h < L; h++) {
        offs[h] = s;
        long long left = h - range + 1; if (left < 0) left = 0;
        long long right = (long long)L - (h + range); if (right < 0) right = 0;
        s += range <= 0 ? (long long)L : left + right;
    }
    offs[L] = s;
}
__global__ void if_keys_kernel(const double* __restrict__ IF, int L, int range, const long long* __restrict__ offs,
                               unsigned long long* __restrict__ keys) {
    int h = blockIdx.x;
    long long cnt = offs[h + 1] - offs[h];
    for (long long c = threadIdx.x; c < cnt; c += blockDim.x) {
        long long i; pair_at(L, range, h, c, &i);
        keys[offs[h] + c] = order_key(IF[(size_t)h * L + i]);        // SP:57-60
    }
}
__global__ void dist_keys_kernel(const long long* __restrict__ mx, int L, int range, const long long* __restrict__ offs,
                                 unsigned long long* __restrict__ keys) {
    int h = blockIdx.x;
    long long cnt = offs[h + 1] - offs[h];
    const long long* a = mx + (size_t)h * 3;
    for (long long c = threadIdx.x; c < cnt; c += blockDim.x) {
        long long i; pair_at(L, range, h, c, &i);
        const long long* b = mx + (size_t)i * 3;
        long long dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        unsigned long long d2 = (unsigned long long)(dx * dx) + (unsigned long long)(dy * dy) + (unsigned long long)(dz * dz);
        keys[offs[h] + c] = (unsigned long long)round_sqrt(d2);      // SP:45-47, in mA
    }
}
// twice the average 1-based rank of key among the sorted keys: lower_bound + upper_bound + 1
__global__ void rank2_kernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ sorted,
                             long long n, long long* __restrict__ rank2) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned long long v = keys[k];
    long long lo = 0, hi = n;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (sorted[mid] < v) lo = mid + 1; else hi = mid; }
    long long lb = lo; hi = n;
    while (lo < hi) { long long mid = (lo + hi) >> 1; if (sorted[mid] <= v) lo = mid + 1; else hi = mid; }
    rank2[k] = lb + lo + 1;
}
// exact moments of a = 2rx-(n+1), b = 2ry-(n+1); per-block partials in 128 bits
__global__ void moments_kernel(const long long* __restrict__ rx2, const long long* __restrict__ ry2, long long n,
                               unsigned long long* __restrict__ partial /* [blocks][3][2] */) {
    __int128 sxx = 0, syy = 0, sxy = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        __int128 a = rx2[k] - (n + 1), b = ry2[k] - (n + 1);
        sxx += a * a; syy += b * b; sxy += a * b;
    }
    __shared__ __int128 sh[3][256];
    sh[0][threadIdx.x] = sxx; sh[1][threadIdx.x] = syy; sh[2][threadIdx.x] = sxy;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) for (int q = 0; q < 3; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) for (int q = 0; q < 3; q++) {
        unsigned __int128 u = (unsigned __int128)sh[q][0];
        partial[((size_t)blockIdx.x * 3 + q) * 2] = (unsigned long long)u;
        partial[((size_t)blockIdx.x * 3 + q) * 2 + 1] = (unsigned long long)(u >> 64);
    }
}
__global__ void rho_kernel(const unsigned long long* __restrict__ partial, int blocks, double* __restrict__ rho) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    __int128 s[3] = {0, 0, 0};
    for (int b = 0; b < blocks; b++) for (int q = 0; q < 3; q++) {
        unsigned __int128 u = ((unsigned __int128)partial[((size_t)b * 3 + q) * 2 + 1] << 64) | partial[((size_t)b * 3 + q) * 2];
        s[q] += (__int128)u;
    }
    // sigma = S/4 exactly as the module accumulates (r - mean)^2 in fp64 (exact below 2^53)
    double sx = (double)s[0] / 4.0, sy = (double)s[1] / 4.0, sxy = (double)s[2] / 4.0;
    *rho = (sx * sy != 0.0) ? sxy / sqrt(sx * sy) : 0.0;
}

int dev_spearman(const double* d_IF, int L, const double* d_xyz, int n_models, int range, double* rho_host, cudaStream_t s) {
    if (range >= L) { for (int m = 0; m < n_models; m++) rho_host[m] = nan(""); return 0; }     // SP:38-41
    DevBuf<long long> offs, mx, rx2, ry2; DevBuf<unsigned long long> keys, sorted, partial; DevBuf<double> rho; DevBuf<unsigned char> tmp;
    CHR3D_CUDA(offs.alloc((size_t)L + 1));
    row_offsets_kernel<<<1, 32, 0, s>>>(L, range, offs.p);
    long long n = 0;
    CHR3D_CUDA(cudaMemcpyAsync(&n, offs.p + L, sizeof(long long), cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    if (n > 0x7fffffffLL) { set_error("spearman: %lld pairs exceed the sort limit of this build", n); return CHR3D_ERR_SIZE; }
    CHR3D_CUDA(keys.alloc((size_t)n)); CHR3D_CUDA(sorted.alloc((size_t)n)); CHR3D_CUDA(rx2.alloc((size_t)n)); CHR3D_CUDA(ry2.alloc((size_t)n));
    const int mblocks = 296;
    CHR3D_CUDA(partial.alloc((size_t)mblocks * 6)); CHR3D_CUDA(rho.alloc((size_t)n_models));
    const long long nc = (long long)n_models * L * 3;
    CHR3D_CUDA(mx.alloc((size_t)nc));
    milli_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(d_xyz, nc, mx.p);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys.p, sorted.p, (int)n, 0, 64, s);
    CHR3D_CUDA(tmp.alloc(tmp_bytes));
    // IF ranks (once for all models)
    if_keys_kernel<<<L, 128, 0, s>>>(d_IF, L, range, offs.p, keys.p);
    CHR3D_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, keys.p, sorted.p, (int)n, 0, 64, s));
    rank2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys.p, sorted.p, n, rx2.p);
    CHR3D_CUDA(cudaGetLastError());
    for (int m = 0; m < n_models; m++) {
        dist_keys_kernel<<<L, 128, 0, s>>>(mx.p + (size_t)m * L * 3, L, range, offs.p, keys.p);
        CHR3D_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, keys.p, sorted.p, (int)n, 0, 40, s));
        rank2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(keys.p, sorted.p, n, ry2.p);
        moments_kernel<<<mblocks, 256, 0, s>>>(rx2.p, ry2.p, n, partial.p);
        rho_kernel<<<1, 32, 0, s>>>(partial.p, mblocks, rho.p + m);
        CHR3D_CUDA(cudaGetLastError());
    }
    CHR3D_CUDA(cudaMemcpyAsync(rho_host, rho.p, sizeof(double) * n_models, cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// model-vs-model similarity (the scorer behind output_models/similarity.txt):
// Spearman rho between the upper-triangle distance vectors of two models, and the RMSE between them after
// scaling b's distances by mean(da)/mean(db).  Distances are fp64 sqrt((dx^2 + dy^2) + dz^2) without FMA
// contraction (the order a NumPy/Java scorer evaluates them in); ranks are exact (radix sort of the ordered
// fp64 keys, average ranks for ties), rho comes from exact 128-bit integer moments.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_of(long long k, int n, int& i, int& j) {     // k-th pair (i < j) in row-major order
    // row i starts at i*n - i(i+1)/2
    double b = 2.0 * n - 1.0;
    long long ii = (long long)floor((b - sqrt(b * b - 8.0 * (double)k)) * 0.5);
    if (ii < 0) ii = 0;
    while (ii > 0 && ii * n - ii * (ii + 1) / 2 > k) ii--;
    while ((ii + 1) * n - (ii + 1) * (ii + 2) / 2 <= k) ii++;
    i = (int)ii; j = (int)(k - (ii * n - ii * (ii + 1) / 2) + ii + 1);
}
__global__ void pdist_kernel(const double* __restrict__ x, int n, long long np, double* __restrict__ d, unsigned long long* __restrict__ keys) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    int i, j; pair_of(k, n, i, j);
    const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
    const double v = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    d[k] = v; keys[k] = order_key(v);
}
// fixed-order sums: per-block partials of (sum da, sum db) or of sum (da - sc*db)^2
__global__ void __launch_bounds__(256) sim_sums_kernel(const double* __restrict__ da, const double* __restrict__ db, long long np, const double* scale,
                                                       double* __restrict__ part /* [blocks][2] */) {
    double s0 = 0, s1 = 0;
    const double sc = scale ? *scale : 0.0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (long long)gridDim.x * blockDim.x) {
        if (scale) { const double e = da[k] - sc * db[k]; s0 += e * e; } else { s0 += da[k]; s1 += db[k]; }
    }
    __shared__ double sh[2][256];
    sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; } __syncthreads(); }
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = sh[0][0]; part[2 * blockIdx.x + 1] = sh[1][0]; }
}
__global__ void sim_finish_kernel(const double* part, int blocks, long long np, int stage, double* out /* [0] scale, [1] rmsd */) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double s0 = 0, s1 = 0;
    for (int b = 0; b < blocks; b++) { s0 += part[2 * b]; s1 += part[2 * b + 1]; }
    if (stage == 0) out[0] = (s0 / (double)np) / (s1 / (double)np); else out[1] = sqrt(s0 / (double)np);
}
int dev_model_similarity(const double* d_xa, const double* d_xb, int n, double* rho_host, double* rmsd_host, cudaStream_t s) {
    const long long np = (long long)n * (n - 1) / 2;
    if (np < 1) { *rho_host = nan(""); *rmsd_host = nan(""); return 0; }
    if (np > 0x7fffffffLL) { set_error("model similarity: %lld pairs exceed the sort limit of this build", np); return CHR3D_ERR_SIZE; }
    DevBuf<double> da, db, part, out; DevBuf<unsigned long long> keys, sorted, mom; DevBuf<long long> ra2, rb2; DevBuf<unsigned char> tmp;
    const int blocks = 296;
    CHR3D_CUDA(da.alloc((size_t)np)); CHR3D_CUDA(db.alloc((size_t)np)); CHR3D_CUDA(keys.alloc((size_t)np)); CHR3D_CUDA(sorted.alloc((size_t)np));
    CHR3D_CUDA(ra2.alloc((size_t)np)); CHR3D_CUDA(rb2.alloc((size_t)np)); CHR3D_CUDA(part.alloc((size_t)blocks * 2)); CHR3D_CUDA(out.alloc(3));
    CHR3D_CUDA(mom.alloc((size_t)blocks * 6));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys.p, sorted.p, (int)np, 0, 64, s);
    CHR3D_CUDA(tmp.alloc(tmp_bytes));
    const unsigned g = (unsigned)((np + 255) / 256);
    pdist_kernel<<<g, 256, 0, s>>>(d_xa, n, np, da.p, keys.p);
    CHR3D_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, keys.p, sorted.p, (int)np, 0, 64, s));
    rank2_kernel<<<g, 256, 0, s>>>(keys.p, sorted.p, np, ra2.p);
    pdist_kernel<<<g, 256, 0, s>>>(d_xb, n, np, db.p, keys.p);
    CHR3D_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, keys.p, sorted.p, (int)np, 0, 64, s));
    rank2_kernel<<<g, 256, 0, s>>>(keys.p, sorted.p, np, rb2.p);
    moments_kernel<<<blocks, 256, 0, s>>>(ra2.p, rb2.p, np, mom.p);
    rho_kernel<<<1, 32, 0, s>>>(mom.p, blocks, out.p + 2);
    sim_sums_kernel<<<blocks, 256, 0, s>>>(da.p, db.p, np, nullptr, part.p);
    sim_finish_kernel<<<1, 32, 0, s>>>(part.p, blocks, np, 0, out.p);
    sim_sums_kernel<<<blocks, 256, 0, s>>>(da.p, db.p, np, out.p, part.p);
    sim_finish_kernel<<<1, 32, 0, s>>>(part.p, blocks, np, 1, out.p);
    CHR3D_CUDA(cudaGetLastError());
    double h[3];
    CHR3D_CUDA(cudaMemcpyAsync(h, out.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    *rho_host = h[2]; *rmsd_host = h[1];
    return 0;
}

}  // namespace chr3d
