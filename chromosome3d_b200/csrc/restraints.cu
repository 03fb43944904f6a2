// restraints.cu — IF matrix -> distance restraints on the device.
//
// Restates IF2dist_new (C3D:110-162), dist2rr (C3D:181-206) and carr2tbl (C3D:340-362):
//   P = IF**alpha; mean = sum(P)/(L*L) (i-major sequential sum, diagonal included);
//   Q = P/mean; D = (Q==0) ? -1 : K/Q; text "%.1f"; keep j>i, j-i>=min_sep, value>0;
//   rows ordered by the *string* sort of "i j" (1-based).
// The device works in integer tenths of an Angstrom (the exact content of the "%.1f" text).
//
// Bit-exactness against the Perl:
//   * the sum is the reference's left-to-right sum (L <= SEQ_SUM_MAX_L), so the fp64 roundings of
//     C3D:132-139 are reproduced — evaluated in parallel without changing a bit (seqsum.cuh);
//     above that size (beyond the reference's own 663 bead cap, C3D:93-94) a fixed blocked order
//     is used;
//   * the decimal rounding is decided exactly with an FMA residual (no rint(10*x));
//   * CUDA's pow() is within 2 ulp of the correctly rounded result, glibc's within 1 ulp, so D may
//     differ from Perl's by a few ulp.  Every cell whose rounding would change under a relative
//     perturbation of CERT_EPS is flagged; the host re-derives flagged cells with libm pow in the
//     reference's order (expected frequency ~1e-10 per cell).  n_uncertain reports how many.
#include <math.h>
#include <algorithm>
#include <string>
#include <vector>
#include "common.h"
#include "seqsum.cuh"

namespace chr3d {

#define SEQ_SUM_MAX_L 8192
#define CERT_EPS 1e-13
#define MAX_FLAGGED CHR3D_MAX_FLAGGED

__global__ void pow_kernel(const double* __restrict__ IF, double* __restrict__ P, long long n, double alpha) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; k < n; k += stride) P[k] = pow(IF[k], alpha);
}

// large L: per-row sequential sums (one thread per row), then rows summed sequentially
__global__ void row_sum_kernel(const double* __restrict__ P, int L, double* __restrict__ rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const double* r = P + (size_t)i * L;
    double s = 0.0;
    for (int j = 0; j < L; j++) s += r[j];
    rows[i] = s;
}

// exact "%.1f" rounding of a positive double as integer tenths (round-half-even on exact ties)
__device__ __forceinline__ long long round_tenths(double D, int* uncertain) {
    if (D >= 4.0e15) { *uncertain = 0; return (long long)D * 10; }   // no fractional part representable
    double t = D * 10.0;
    double n = floor(t);
    // exact floor: D*10 - n evaluated with one rounding
    if (fma(D, 10.0, -n) < 0.0) n -= 1.0;
    else if (fma(D, 10.0, -(n + 1.0)) >= 0.0) n += 1.0;
    double r = fma(D, 10.0, -(n + 0.5));          // sign is exact
    long long ni = (long long)n;
    long long res = r > 0.0 ? ni + 1 : (r < 0.0 ? ni : ((ni & 1) ? ni + 1 : ni));
    // certification: would D(1 +- eps) round differently?
    double lo = D * (1.0 - CERT_EPS), hi = D * (1.0 + CERT_EPS);
    double rl = fma(lo, 10.0, -(n + 0.5)), rh = fma(hi, 10.0, -(n + 0.5));
    double fl = fma(lo, 10.0, -n), fh = fma(hi, 10.0, -(n + 1.0));
    *uncertain = ((rl > 0.0) != (rh > 0.0)) || r == 0.0 || fl < 0.0 || fh >= 0.0;
    return res;
}

__global__ void tenths_kernel(const double* __restrict__ P, long long n, double mean, double K,
                              long long* __restrict__ tenths, unsigned long long* n_unc, long long* flagged) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        double q = P[k] / mean;                   // C3D:142
        if (q == 0.0) { tenths[k] = -10; continue; }   // C3D:148-150 -> "-1.0"
        double D = K / q;                         // C3D:152
        int unc = 0;
        long long t = round_tenths(D, &unc);
        if (!(D == D) || isinf(D)) { unc = 1; t = 0; }
        tenths[k] = t;
        if (unc) {
            unsigned long long slot = atomicAdd(n_unc, 1ull);
            if (slot < MAX_FLAGGED) flagged[slot] = k;
        }
    }
}

int dev_if_to_tenths(const double* d_IF, int L, double alpha, double K, long long* d_tenths,
                     long long* n_uncertain_host, std::vector<long long>* flagged_host, cudaStream_t s) {
    const long long n = (long long)L * L;
    DevBuf<double> P, sum, rows;
    DevBuf<unsigned long long> nunc;
    DevBuf<long long> flagged;
    CHR3D_CUDA(P.alloc((size_t)n));
    CHR3D_CUDA(sum.alloc(1));
    CHR3D_CUDA(nunc.alloc(1));
    CHR3D_CUDA(flagged.alloc(MAX_FLAGGED));
    CHR3D_CUDA(cudaMemsetAsync(nunc.p, 0, sizeof(unsigned long long), s));
    int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    pow_kernel<<<blocks, 256, 0, s>>>(d_IF, P.p, n, alpha);
    CHR3D_CUDA(cudaGetLastError());
    if (L <= SEQ_SUM_MAX_L) {
        seq_sum_launch(P.p, n, 1, sum.p, s);
    } else {
        CHR3D_CUDA(rows.alloc((size_t)L));
        row_sum_kernel<<<(L + 127) / 128, 128, 0, s>>>(P.p, L, rows.p);
        seq_sum_launch(rows.p, L, 1, sum.p, s);
    }
    CHR3D_CUDA(cudaGetLastError());
    double total = 0.0;
    CHR3D_CUDA(cudaMemcpyAsync(&total, sum.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    double mean = total / ((double)L * (double)L);            // C3D:139
    tenths_kernel<<<blocks, 256, 0, s>>>(P.p, n, mean, K, d_tenths, nunc.p, flagged.p);
    CHR3D_CUDA(cudaGetLastError());
    unsigned long long nu = 0;
    CHR3D_CUDA(cudaMemcpyAsync(&nu, nunc.p, sizeof(nu), cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    if (n_uncertain_host) *n_uncertain_host = (long long)nu;
    if (flagged_host) {
        flagged_host->clear();
        if (nu > 0 && nu <= MAX_FLAGGED) {
            flagged_host->resize((size_t)nu);
            CHR3D_CUDA(cudaMemcpyAsync(flagged_host->data(), flagged.p, sizeof(long long) * (size_t)nu, cudaMemcpyDeviceToHost, s));
            CHR3D_CUDA(cudaStreamSynchronize(s));
        }
    }
    return 0;
}

// ---- the same derivation for a sparse matrix: only the non-zero cells, in i-major order (a CSR of the symmetric matrix).
// Zero cells contribute exactly 0 to the reference's sum (C3D:132-137: 0**alpha = 0 for alpha > 0) and get no restraint
// (C3D:148-150, 196), so summing the non-zero cells in the same order gives the same bits as the dense pass.
__global__ void row_sum_csr_kernel(const double* __restrict__ P, const long long* __restrict__ row_ptr, int L, double* __restrict__ rows) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    double s = 0.0;
    for (long long k = row_ptr[i]; k < row_ptr[i + 1]; k++) s += P[k];
    rows[i] = s;
}
// d_vals: n non-zero cell values in i-major order; d_row_ptr: L + 1 offsets.  Tenths per cell as dev_if_to_tenths gives them.
int dev_cells_to_tenths(const double* d_vals, long long n, const long long* d_row_ptr, int L, double alpha, double K, long long* d_tenths,
                        long long* n_uncertain_host, std::vector<long long>* flagged_host, cudaStream_t s) {
    DevBuf<double> P, sum, rows;
    DevBuf<unsigned long long> nunc;
    DevBuf<long long> flagged;
    CHR3D_CUDA(P.alloc((size_t)std::max<long long>(n, 1)));
    CHR3D_CUDA(sum.alloc(1)); CHR3D_CUDA(nunc.alloc(1)); CHR3D_CUDA(flagged.alloc(MAX_FLAGGED));
    CHR3D_CUDA(cudaMemsetAsync(nunc.p, 0, sizeof(unsigned long long), s));
    CHR3D_CUDA(cudaMemsetAsync(sum.p, 0, sizeof(double), s));
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 16));
    if (n > 0) pow_kernel<<<blocks, 256, 0, s>>>(d_vals, P.p, n, alpha);
    if (L <= SEQ_SUM_MAX_L) {
        if (n > 0) seq_sum_launch(P.p, n, 1, sum.p, s);
    } else {
        CHR3D_CUDA(rows.alloc((size_t)L));
        row_sum_csr_kernel<<<(L + 127) / 128, 128, 0, s>>>(P.p, d_row_ptr, L, rows.p);
        seq_sum_launch(rows.p, L, 1, sum.p, s);
    }
    CHR3D_CUDA(cudaGetLastError());
    double total = 0.0;
    CHR3D_CUDA(cudaMemcpyAsync(&total, sum.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    const double mean = total / ((double)L * (double)L);            // C3D:139
    if (n > 0) tenths_kernel<<<blocks, 256, 0, s>>>(P.p, n, mean, K, d_tenths, nunc.p, flagged.p);
    CHR3D_CUDA(cudaGetLastError());
    unsigned long long nu = 0;
    CHR3D_CUDA(cudaMemcpyAsync(&nu, nunc.p, sizeof(nu), cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    if (n_uncertain_host) *n_uncertain_host = (long long)nu;
    if (flagged_host) {
        flagged_host->clear();
        if (nu > 0 && nu <= MAX_FLAGGED) {
            flagged_host->resize((size_t)nu);
            CHR3D_CUDA(cudaMemcpyAsync(flagged_host->data(), flagged.p, sizeof(long long) * (size_t)nu, cudaMemcpyDeviceToHost, s));
            CHR3D_CUDA(cudaStreamSynchronize(s));
        }
    }
    return 0;
}

__global__ void targets_from_tenths_kernel(const long long* __restrict__ tenths, int L, int pitch, int min_sep,
                                           int vdw_min_sep, float* __restrict__ T) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= pitch) return;
    float v = -2.f;
    if (j < L) {
        int sep = abs(i - j);
        // the table keeps (i<j); the matrix is symmetric, so read the upper-triangle cell
        long long t = i < j ? tenths[(size_t)i * L + j] : tenths[(size_t)j * L + i];
        if (sep >= min_sep && t > 0) v = (float)((double)t / 10.0);
        else v = sep >= vdw_min_sep ? -1.f : -2.f;
    }
    T[(size_t)i * pitch + j] = v;
}
int dev_tenths_to_targets(const long long* d_tenths, int L, int min_sep, int vdw_min_sep, float* d_T, cudaStream_t s) {
    int pitch = target_pitch(L);
    dim3 grid((pitch + 127) / 128, L);
    targets_from_tenths_kernel<<<grid, 128, 0, s>>>(d_tenths, L, pitch, min_sep, vdw_min_sep, d_T);
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}

__global__ void targets_fill_kernel(int N, int pitch, int vdw_min_sep, float* __restrict__ T) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = blockIdx.y;
    if (j >= pitch) return;
    T[(size_t)i * pitch + j] = (j < N && abs(i - j) >= vdw_min_sep) ? -1.f : -2.f;
}
__global__ void targets_scatter_kernel(const int32_t* __restrict__ ri, const int32_t* __restrict__ rj,
                                       const double* __restrict__ rd, long long nr, int N, int pitch, float* __restrict__ T) {
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nr) return;
    int i = ri[k], j = rj[k];
    if (i < 0 || j < 0 || i >= N || j >= N || i == j) return;
    float d = (float)rd[k];
    T[(size_t)i * pitch + j] = d;
    T[(size_t)j * pitch + i] = d;
}
int dev_list_to_targets(const int32_t* d_ri, const int32_t* d_rj, const double* d_rd, long long nr, int N,
                        int vdw_min_sep, float* d_T, cudaStream_t s) {
    int pitch = target_pitch(N);
    dim3 grid((pitch + 127) / 128, N);
    targets_fill_kernel<<<grid, 128, 0, s>>>(N, pitch, vdw_min_sep, d_T);
    CHR3D_CUDA(cudaGetLastError());
    if (nr > 0) {
        targets_scatter_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(d_ri, d_rj, d_rd, nr, N, pitch, d_T);
        CHR3D_CUDA(cudaGetLastError());
    }
    return 0;
}

// ---- restraint list in the key order of C3D:203 ----------------------------------------------
// perm[a] = 0-based bead whose 1-based decimal string is the a-th smallest.  Rows are emitted for
// a ascending, b ascending with i=perm[a], j=perm[b], j-i>=min_sep, tenths>0.
__global__ void list_count_kernel(const long long* __restrict__ tenths, int L, int min_sep,
                                  const int* __restrict__ perm, unsigned long long* __restrict__ counts) {
    int a = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (a >= L) return;
    int i = perm[a];
    unsigned long long cnt = 0;
    for (int b0 = 0; b0 < L; b0 += 32) {
        int b = b0 + lane; bool ok = false;
        if (b < L) { int j = perm[b]; ok = (j - i >= min_sep) && tenths[(size_t)i * L + j] > 0; }
        cnt += __popc(__ballot_sync(0xffffffffu, ok));
    }
    if (lane == 0) counts[a] = cnt;
}
// exclusive prefix sum of the per-row counts, one warp (integers: any order gives the same offsets)
__global__ void list_scan_kernel(unsigned long long* counts, int L, unsigned long long* total) {
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    unsigned long long carry = 0;
    for (int a0 = 0; a0 < L; a0 += 32) {
        const int a = a0 + lane;
        const unsigned long long c = a < L ? counts[a] : 0ull;
        unsigned long long incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (a < L) counts[a] = carry + incl - c;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) *total = carry;
}
__global__ void list_fill_kernel(const long long* __restrict__ tenths, int L, int min_sep, const int* __restrict__ perm,
                                 const unsigned long long* __restrict__ offs, long long cap,
                                 int32_t* __restrict__ ri, int32_t* __restrict__ rj, double* __restrict__ rd) {
    int a = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (a >= L) return;
    int i = perm[a];
    unsigned long long base = offs[a];
    for (int b0 = 0; b0 < L; b0 += 32) {
        int b = b0 + lane; bool ok = false; int j = 0; long long t = 0;
        if (b < L) { j = perm[b]; if (j - i >= min_sep) { t = tenths[(size_t)i * L + j]; ok = t > 0; } }
        unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            unsigned long long pos = base + __popc(mask & ((1u << lane) - 1u));
            if ((long long)pos < cap) { ri[pos] = i; rj[pos] = j; rd[pos] = (double)t / 10.0; }
        }
        base += __popc(mask);
    }
}

int dev_tenths_to_list(const long long* d_tenths, int L, int min_sep, int32_t* d_ri, int32_t* d_rj,
                       double* d_rd, long long cap, long long* n_host, cudaStream_t s) {
    // decimal-string order of 1..L (index bookkeeping, done once on the host)
    std::vector<std::string> keys(L);
    std::vector<int> perm(L);
    for (int i = 0; i < L; i++) { keys[i] = std::to_string(i + 1); perm[i] = i; }
    std::sort(perm.begin(), perm.end(), [&](int x, int y) { return keys[x] < keys[y]; });
    DevBuf<int> dperm; DevBuf<unsigned long long> counts, total;
    CHR3D_CUDA(dperm.alloc(L)); CHR3D_CUDA(counts.alloc(L)); CHR3D_CUDA(total.alloc(1));
    CHR3D_CUDA(cudaMemcpyAsync(dperm.p, perm.data(), sizeof(int) * L, cudaMemcpyHostToDevice, s));
    const int wpb = 8;
    list_count_kernel<<<(L + wpb - 1) / wpb, wpb * 32, 0, s>>>(d_tenths, L, min_sep, dperm.p, counts.p);
    list_scan_kernel<<<1, 32, 0, s>>>(counts.p, L, total.p);
    CHR3D_CUDA(cudaGetLastError());
    unsigned long long n = 0;
    CHR3D_CUDA(cudaMemcpyAsync(&n, total.p, sizeof(n), cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    if (n_host) *n_host = (long long)n;
    if (cap > 0 && d_ri) {
        list_fill_kernel<<<(L + wpb - 1) / wpb, wpb * 32, 0, s>>>(d_tenths, L, min_sep, dperm.p, counts.p, cap, d_ri, d_rj, d_rd);
        CHR3D_CUDA(cudaGetLastError());
        CHR3D_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

}  // namespace chr3d

// The left-to-right sum as an entry point of its own (include/chr3d.h): host arrays in, the bits of the scalar loop out.
CHR3D_API int chr3d_seq_sum(const double* a, int64_t n_per, int32_t nseg, double* out) {
    using namespace chr3d;
    if (n_per < 0 || nseg < 0 || (nseg > 0 && !out) || (n_per > 0 && nseg > 0 && !a)) { set_error("chr3d_seq_sum: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    if (nseg == 0) return 0;
    DevBuf<double> d, r;
    const size_t n = (size_t)n_per * (size_t)nseg;
    CHR3D_CUDA(d.alloc(n > 0 ? n : 1)); CHR3D_CUDA(r.alloc((size_t)nseg));
    if (n > 0) CHR3D_CUDA(cudaMemcpy(d.p, a, sizeof(double) * n, cudaMemcpyHostToDevice));
    seq_sum_launch(d.p, n_per, nseg, r.p, 0);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaMemcpy(out, r.p, sizeof(double) * (size_t)nseg, cudaMemcpyDeviceToHost));
    return 0;
}
