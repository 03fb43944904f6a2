// dg.cu — distance-geometry embedding: the starting coordinates of an SA trajectory.
//
// Replaces the DG stage of dgsa.inp (C3D:1470-1526: `mmdg ... storebounds`, then per model `mmdg recallbounds ...` and
// a minimisation).  `mmdg` is a CNS built-in, so only the recipe is pinned by the reference; the algorithm is the textbook
// metric-matrix embedding at bead level (same definition as oracle/chr3d_oracle.c, "Distance-geometry embedding"):
//   bounds   U = L = target (restrained pairs), r0 (chain neighbours), [repel floor, inf) otherwise
//   smooth   U: Floyd-Warshall, one launch per pivot k (row/column k are fixed points of step k, so the parallel step
//            equals the sequential one);  L: one triangle pass, a thread per pair;  L = min(L, U) where the data are not metric
//   trial    D_ij = L_ij + u (U_ij - L_ij), u = Philox(pair, model; seed)
//   embed    metric matrix G, three largest positive eigenpairs by subspace iteration with Rayleigh-Ritz: one CTA per
//            model iterates in place (8 vectors in shared memory, G streamed from L2), a warp per inner product
// fp64 throughout, compiled with -fmad=false, every long sum in the fixed warp order of dg_dot32: results are
// bit-identical to the oracle's.  This is set-up work (O(N^3) once per job, O(N^2) per model and iteration), not the hot path.
#include <math.h>
#include "common.h"

namespace chr3d {

#define DG_P 8
#define DG_BIG 1.0e30
#define DG_STREAM 0x0D6000u
#define DG_THREADS 1024

__device__ inline void dg_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void dg_init_bounds_kernel(double* __restrict__ U, double* __restrict__ L, int N, double r0, double floor_d, int vdw_min_sep) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)N * N) return;
    const int i = (int)(k / N), j = (int)(k % N), sep = i > j ? i - j : j - i;
    U[k] = sep == 0 ? 0.0 : (sep == 1 ? r0 : DG_BIG);
    L[k] = sep == 0 ? 0.0 : (sep == 1 ? r0 : (sep >= vdw_min_sep ? floor_d : 0.0));
}
__global__ void dg_scatter_kernel(double* __restrict__ U, double* __restrict__ L, int N, const int32_t* __restrict__ ri, const int32_t* __restrict__ rj,
                                  const double* __restrict__ rd, long long nr) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nr) return;
    const size_t a = (size_t)ri[k] * N + rj[k], b = (size_t)rj[k] * N + ri[k];
    U[a] = U[b] = rd[k]; L[a] = L[b] = rd[k];
}
__global__ void dg_fw_step_kernel(double* __restrict__ U, int N, int k) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= N) return;
    const double v = U[(size_t)i * N + k] + U[(size_t)k * N + j];
    if (v < U[(size_t)i * N + j]) U[(size_t)i * N + j] = v;
}
// bad[0]: pairs whose lower bound had to be cut down to the upper one; bad[1]: pairs that keep some slack (L < U)
__global__ void dg_lower_kernel(const double* __restrict__ U, const double* __restrict__ L0, double* __restrict__ L, int N, unsigned long long* __restrict__ bad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= N || i == j) { if (j < N) L[(size_t)i * N + j] = 0.0; return; }
    double l = L0[(size_t)i * N + j];
    for (int k = 0; k < N; k++) {
        const double a = L0[(size_t)i * N + k] - U[(size_t)k * N + j], b = L0[(size_t)j * N + k] - U[(size_t)k * N + i];
        if (a > l) l = a;
        if (b > l) l = b;
    }
    const double u = U[(size_t)i * N + j];
    if (l > u) { l = u; if (i < j) atomicAdd(bad, 1ull); }
    if (l < u && i < j) atomicAdd(bad + 1, 1ull);
    L[(size_t)i * N + j] = l;
}
// squared trial distances of one model (symmetric)
__global__ void dg_trial_kernel(const double* __restrict__ U, const double* __restrict__ L, int N, unsigned long long seed, int model, double* __restrict__ D2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= N || j < i) return;
    double d = 0.0;
    if (i != j) {
        const unsigned long long pair = (unsigned long long)i * (unsigned long long)N + (unsigned long long)j;
        uint32_t r[4]; dg_philox((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)model, DG_STREAM, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double u = (r[0] + 0.5) * (1.0 / 4294967296.0);
        const double lo = L[(size_t)i * N + j], hi = U[(size_t)i * N + j];
        d = lo + u * (hi - lo);
    }
    D2[(size_t)i * N + j] = D2[(size_t)j * N + i] = d * d;
}
__global__ void dg_rowsum_kernel(const double* __restrict__ D2, int N, double* __restrict__ d0) {     // a thread per row, index order
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double s = 0.0;
    for (int j = 0; j < N; j++) s += D2[(size_t)i * N + j];
    d0[i] = s;
}
__global__ void dg_centre_kernel(double* __restrict__ d0, int N) {                                    // one thread: total in index order
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double total = 0.0;
    for (int i = 0; i < N; i++) total += d0[i];
    for (int i = 0; i < N; i++) d0[i] = d0[i] / N - total / (2.0 * (double)N * (double)N);
}
__global__ void dg_metric_kernel(const double* __restrict__ D2, const double* __restrict__ d0, int N, double* __restrict__ G) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j < N) G[(size_t)i * N + j] = 0.5 * (d0[i] + d0[j] - D2[(size_t)i * N + j]);
}

// inner product in the fixed warp order (all 32 lanes call; the result is valid in every lane)
__device__ __forceinline__ double dg_dot32(const double* a, int sa, const double* b, int sb, int n, int lane) {
    double p = 0.0;
    for (int i = lane; i < n; i += 32) { const double t = a[(size_t)i * sa] * b[(size_t)i * sb]; p = p + t; }
    for (int o = 16; o > 0; o >>= 1) {
        // lane l (l < o) adds the partial of lane l + o: the oracle's  p[l] = p[l] + p[l + o]
        const double q = __shfl_down_sync(0xffffffffu, p, o);
        p = p + q;
    }
    return __shfl_sync(0xffffffffu, p, 0);
}
__device__ void dg_jacobi(double* H, double* Q, int p) {      // one thread; same sweeps as the oracle's dg_jacobi
    for (int i = 0; i < p; i++) for (int j = 0; j < p; j++) Q[i * p + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0; for (int i = 0; i < p; i++) for (int j = i + 1; j < p; j++) off += H[i * p + j] * H[i * p + j];
        if (off == 0.0) break;
        for (int a = 0; a < p; a++) for (int b = a + 1; b < p; b++) {
            const double hab = H[a * p + b];
            if (hab == 0.0) continue;
            const double th = (H[b * p + b] - H[a * p + a]) / (2.0 * hab);
            const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < p; k++) { const double x = H[k * p + a], y = H[k * p + b]; H[k * p + a] = c * x - s * y; H[k * p + b] = s * x + c * y; }
            for (int k = 0; k < p; k++) { const double x = H[a * p + k], y = H[b * p + k]; H[a * p + k] = c * x - s * y; H[b * p + k] = s * x + c * y; }
            for (int k = 0; k < p; k++) { const double x = Q[k * p + a], y = Q[k * p + b]; Q[k * p + a] = c * x - s * y; Q[k * p + b] = s * x + c * y; }
        }
    }
}
// modified Gram-Schmidt of the P columns of V (shared memory), warp 0 only: the columns depend on each other in sequence
__device__ void dg_mgs(double* V, int N, int P, int lane) {
    for (int c = 0; c < P; c++) {
        for (int b = 0; b < c; b++) {
            const double dot = dg_dot32(V + b, P, V + c, P, N, lane);
            for (int i = lane; i < N; i += 32) V[(size_t)i * P + c] -= dot * V[(size_t)i * P + b];
            __syncwarp();
        }
        double nn = sqrt(dg_dot32(V + c, P, V + c, P, N, lane));
        if (nn < 1e-300) { for (int i = lane; i < N; i += 32) V[(size_t)i * P + c] = (i == c) ? 1.0 : 0.0; }
        else for (int i = lane; i < N; i += 32) V[(size_t)i * P + c] /= nn;
        __syncwarp();
    }
}
// one CTA per model: subspace iteration with Rayleigh-Ritz on G (N x N, global), coordinates out
__global__ void __launch_bounds__(DG_THREADS, 1) dg_embed_kernel(const double* __restrict__ Gall, int N, double* __restrict__ xyz, double* __restrict__ eig3) {
    extern __shared__ double dsm[];
    const int P = DG_P < N ? DG_P : N;
    double* V = dsm;                    // N x P
    double* W = V + (size_t)N * P;      // N x P
    __shared__ double H[DG_P * DG_P], Q[DG_P * DG_P], ritz[DG_P], prev[DG_P];
    __shared__ int stop;
    const double* G = Gall + (size_t)blockIdx.x * N * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = DG_THREADS / 32;
    for (int k = tid; k < N * P; k += DG_THREADS) {      // start vectors: fixed pseudo-random entries in (-0.5, 0.5), integer arithmetic only
        const int i = k / P, c = k % P;
        uint32_t r[4]; dg_philox((uint32_t)i, (uint32_t)c, 0u, DG_STREAM + 1u, 0x5eedu, 0u, r);
        V[k] = (r[0] + 0.5) * (1.0 / 4294967296.0) - 0.5;
    }
    if (tid < DG_P) prev[tid] = 0.0;
    if (tid == 0) stop = 0;
    __syncthreads();
    for (int it = 0; it < 500; it++) {
        if (warp == 0) dg_mgs(V, N, P, lane);
        __syncthreads();
        for (int k = warp; k < N * P; k += nwarp) {          // W = G V, a warp per entry
            const int i = k / P, c = k % P;
            const double s = dg_dot32(G + (size_t)i * N, 1, V + c, P, N, lane);
            if (lane == 0) W[k] = s;
        }
        __syncthreads();
        for (int k = warp; k < P * P; k += nwarp) {          // H = V^T W
            const int a = k / P, b = k % P;
            const double s = dg_dot32(V + a, P, W + b, P, N, lane);
            if (lane == 0) H[k] = s;
        }
        __syncthreads();
        if (tid == 0) {
            for (int a = 0; a < P; a++) for (int b = a + 1; b < P; b++) { const double m = 0.5 * (H[a * P + b] + H[b * P + a]); H[a * P + b] = H[b * P + a] = m; }
            dg_jacobi(H, Q, P);
            // converged when the three largest Ritz values (the ones the embedding uses) have settled
            double top[3] = {-1e300, -1e300, -1e300};
            for (int c = 0; c < P; c++) { ritz[c] = H[c * P + c]; double v = ritz[c]; for (int q = 0; q < 3; q++) if (v > top[q]) { const double t = top[q]; top[q] = v; v = t; } }
            double chg = 0.0, mag = 0.0;
            for (int q = 0; q < 3 && q < P; q++) { chg += fabs(top[q] - prev[q]); mag += fabs(top[q]); prev[q] = top[q]; }
            if (it > 2 && chg <= 1e-13 * mag) stop = 1;
        }
        __syncthreads();
        for (int i = tid; i < N; i += DG_THREADS) {          // Ritz vectors R = V Q (orthonormal) stay in W; next subspace V = (G V) Q
            double t[DG_P], r[DG_P];
            for (int c = 0; c < P; c++) { double s = 0.0, q = 0.0; for (int b = 0; b < P; b++) { s += W[(size_t)i * P + b] * Q[b * P + c]; q += V[(size_t)i * P + b] * Q[b * P + c]; } t[c] = s; r[c] = q; }
            for (int c = 0; c < P; c++) { V[(size_t)i * P + c] = t[c]; W[(size_t)i * P + c] = r[c]; }
        }
        __syncthreads();
        if (stop) break;
    }
    if (warp == 0) {      // the three largest positive Ritz values and their Ritz vectors (in W)
        bool used[DG_P];
        for (int c = 0; c < DG_P; c++) used[c] = false;
        for (int q = 0; q < 3; q++) {
            int best = -1;
            for (int c = 0; c < P; c++) if (!used[c] && (best < 0 || ritz[c] > ritz[best])) best = c;
            const double lam = best >= 0 ? ritz[best] : 0.0;
            if (best >= 0) used[best] = true;
            if (lane == 0 && eig3) eig3[3 * blockIdx.x + q] = lam;
            const double sc = lam > 0.0 ? sqrt(lam) : 0.0;
            // sign convention: the component of largest magnitude (first one on ties) is positive
            double vm = 0.0; int im = 0;
            if (best >= 0) {
                for (int i = lane; i < N; i += 32) { const double a = fabs(W[(size_t)i * P + best]); if (a > vm) { vm = a; im = i; } }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, vm, o); const int oi = __shfl_xor_sync(0xffffffffu, im, o);
                    if (ov > vm || (ov == vm && oi < im)) { vm = ov; im = oi; }
                }
            }
            const double sg = (best >= 0 && W[(size_t)im * P + best] < 0.0) ? -1.0 : 1.0;
            for (int i = lane; i < N; i += 32) xyz[((size_t)blockIdx.x * N + i) * 3 + q] = best >= 0 ? sg * sc * W[(size_t)i * P + best] : 0.0;
        }
    }
}

// bounds on the device; n_bad_host: pairs whose lower bound had to be cut down to the upper one (non-metric data)
int dev_dg_bounds(int N, const int32_t* d_ri, const int32_t* d_rj, const double* d_rd, long long nr, const chr3d_sa_params& p,
                  double* d_U, double* d_L, long long* n_bad_host, long long* n_slack_host, cudaStream_t s) {
    DevBuf<double> L0; DevBuf<unsigned long long> bad;
    const size_t nn = (size_t)N * N;
    CHR3D_CUDA(L0.alloc(nn)); CHR3D_CUDA(bad.alloc(2));
    CHR3D_CUDA(cudaMemsetAsync(bad.p, 0, 2 * sizeof(unsigned long long), s));
    dg_init_bounds_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(d_U, L0.p, N, p.bond_r0, p.vdw_rmin * p.cool_rad_finl, p.vdw_min_sep);
    if (nr > 0) dg_scatter_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(d_U, L0.p, N, d_ri, d_rj, d_rd, nr);
    const dim3 grid((unsigned)((N + 127) / 128), (unsigned)N);
    for (int k = 0; k < N; k++) dg_fw_step_kernel<<<grid, 128, 0, s>>>(d_U, N, k);
    dg_lower_kernel<<<grid, 128, 0, s>>>(d_U, L0.p, d_L, N, bad.p);
    CHR3D_CUDA(cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    CHR3D_CUDA(cudaMemcpyAsync(h, bad.p, sizeof h, cudaMemcpyDeviceToHost, s));
    CHR3D_CUDA(cudaStreamSynchronize(s));
    if (n_bad_host) *n_bad_host = (long long)h[0];
    if (n_slack_host) *n_slack_host = (long long)h[1];
    return 0;
}
// coordinates (n_models x N x 3, fp64, device) of the models first_model .. first_model + n_models - 1
int dev_dg_embed(int N, const double* d_U, const double* d_L, unsigned long long seed, int first_model, int n_models, double* d_xyz, double* d_eig3,
                 cudaStream_t s) {
    const int P = DG_P < N ? DG_P : N;
    const size_t smem = sizeof(double) * 2 * (size_t)N * P;
    if (smem > 200 * 1024) { set_error("dg embed: N=%d exceeds the shared-memory capacity of the embedding kernel (N <= 1600)", N); return CHR3D_ERR_SIZE; }
    CHR3D_CUDA(cudaFuncSetAttribute(dg_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const size_t nn = (size_t)N * N;
    const int chunk = std::max(1, std::min(n_models, (int)(((size_t)2 << 30) / (nn * sizeof(double) * 2))));     // <= 2 GB of scratch at a time
    DevBuf<double> D2, G, d0;
    CHR3D_CUDA(D2.alloc(nn)); CHR3D_CUDA(G.alloc(nn * chunk)); CHR3D_CUDA(d0.alloc((size_t)N));
    const dim3 grid((unsigned)((N + 127) / 128), (unsigned)N);
    for (int m0 = 0; m0 < n_models; m0 += chunk) {
        const int mc = std::min(chunk, n_models - m0);
        for (int m = 0; m < mc; m++) {
            dg_trial_kernel<<<grid, 128, 0, s>>>(d_U, d_L, N, seed, first_model + m0 + m, D2.p);
            dg_rowsum_kernel<<<(N + 127) / 128, 128, 0, s>>>(D2.p, N, d0.p);
            dg_centre_kernel<<<1, 32, 0, s>>>(d0.p, N);
            dg_metric_kernel<<<grid, 128, 0, s>>>(D2.p, d0.p, N, G.p + (size_t)m * nn);
        }
        dg_embed_kernel<<<mc, DG_THREADS, smem, s>>>(G.p, N, d_xyz + (size_t)m0 * N * 3, d_eig3 ? d_eig3 + 3 * (size_t)m0 : nullptr);
        CHR3D_CUDA(cudaGetLastError());
    }
    CHR3D_CUDA(cudaStreamSynchronize(s));
    return 0;
}

// ---- the rest of the DG stage (C3D:1516-1521): x *= dgscale, then `minimize lbfgs nstep=200 drop=1.` on the bounds-violation
// energy alone (`flags exclude * include dg`, C3D:1495).  Same definitions and the same operation order as the oracle's
// orc_dg_refine (bit-identical): E = scale * sum_{i<j} [max(0, d^2 - U^2)^2 + max(0, L^2 - d^2)^2], dgscale = least-squares
// factor onto the bound midpoints, L-BFGS of the SA minimiser in fp64.  One CTA per model; a warp per row of the bounds
// matrices (lanes stride the row, so U and L stream coalesced from L2), vectors in global scratch, the scalar logic of the
// minimiser replicated in every thread from block-uniform inner products.
struct DgRefineArgs {
    int N, nstep, m, max_ls;
    double scale, drop, c1, gtol, ftol;
    const double* U; const double* L;
    double* xyz; double* work; double* stats;
};
__device__ __forceinline__ double dg_fold32(double p) {      // lane 0 holds the oracle's p[0]
    for (int o = 16; o > 0; o >>= 1) { const double q = __shfl_down_sync(0xffffffffu, p, o); p = p + q; }
    return p;
}
__device__ __forceinline__ double dg_sum32(const double* a, int n, int lane) {
    double p = 0.0;
    for (int i = lane; i < n; i += 32) p = p + a[i];
    return __shfl_sync(0xffffffffu, dg_fold32(p), 0);
}
// block-wide: inner product in the fixed warp order, the value returned to every thread
__device__ double dg_block_dot(const double* a, const double* b, int n, double* slot) {
    __syncthreads();
    if (threadIdx.x < 32) { const double r = dg_dot32(a, 1, b, 1, n, threadIdx.x); if (threadIdx.x == 0) *slot = r; }
    __syncthreads();
    return *slot;
}
__device__ double dg_block_energy(int N, const double* __restrict__ U, const double* __restrict__ L, double scale, const double* x, double* g, double* er, double* slot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();
    for (int i = warp; i < N; i += nwarp) {
        double pe = 0.0, px = 0.0, py = 0.0, pz = 0.0;
        const double xi = x[3 * i], yi = x[3 * i + 1], zi = x[3 * i + 2];
        for (int j = lane; j < N; j += 32) {
            const double dx = xi - x[3 * j], dy = yi - x[3 * j + 1], dz = zi - x[3 * j + 2];
            const double d2 = (dx * dx + dy * dy) + dz * dz;
            const double u = U[(size_t)i * N + j], lo_b = L[(size_t)i * N + j];
            double hi = d2 - u * u; if (!(hi > 0.0)) hi = 0.0;
            double lo = lo_b * lo_b - d2; if (!(lo > 0.0)) lo = 0.0;
            const double w = hi - lo;
            pe = pe + (hi * hi + lo * lo);
            px = px + w * dx; py = py + w * dy; pz = pz + w * dz;
        }
        pe = dg_fold32(pe); px = dg_fold32(px); py = dg_fold32(py); pz = dg_fold32(pz);
        if (lane == 0) { er[i] = pe; g[3 * i] = (4.0 * scale) * px; g[3 * i + 1] = (4.0 * scale) * py; g[3 * i + 2] = (4.0 * scale) * pz; }
    }
    __syncthreads();
    if (warp == 0) { const double r = dg_sum32(er, N, lane); if (lane == 0) *slot = (0.5 * scale) * r; }
    __syncthreads();
    return *slot;
}
__global__ void __launch_bounds__(DG_THREADS, 1) dg_refine_kernel(const DgRefineArgs a) {
    const int N = a.N, n = 3 * N, m = a.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = DG_THREADS / 32;
    double* x = a.xyz + (size_t)blockIdx.x * n;
    double* w = a.work + (size_t)blockIdx.x * ((size_t)n * (4 + 2 * m) + 3 * (size_t)N);
    double *g = w, *gt = g + n, *xt = gt + n, *d = xt + n, *S = d + n, *Y = S + (size_t)n * m, *er = Y + (size_t)n * m, *num = er + N, *den = num + N;
    const double* __restrict__ U = a.U; const double* __restrict__ L = a.L;
    __shared__ double slot[4];
    double rho[CHR3D_LBFGS_MAX_M], al[CHR3D_LBFGS_MAX_M];
    for (int i = warp; i < N; i += nwarp) {                 // dgscale: sum m_ij d_ij / sum d_ij^2 over the pairs with a finite upper bound
        double pn = 0.0, pd = 0.0;
        for (int j = lane; j < N; j += 32) {
            const double dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
            const double d2 = (dx * dx + dy * dy) + dz * dz;
            const double mid = 0.5 * (U[(size_t)i * N + j] + L[(size_t)i * N + j]);
            if (i != j && mid < 0.5 * DG_BIG) { pn = pn + mid * sqrt(d2); pd = pd + d2; }
        }
        pn = dg_fold32(pn); pd = dg_fold32(pd);
        if (lane == 0) { num[i] = pn; den[i] = pd; }
    }
    __syncthreads();
    if (warp == 0) { const double sn = dg_sum32(num, N, lane), sd = dg_sum32(den, N, lane); if (lane == 0) slot[0] = sd > 0.0 ? sn / sd : 1.0; }
    __syncthreads();
    const double dgscale = slot[0];
    for (int i = tid; i < n; i += DG_THREADS) x[i] = x[i] * dgscale;
    double e = dg_block_energy(N, U, L, a.scale, x, g, er, &slot[1]), et = e;
    const double e_start = e;
    int cnt = 0, head = 0, iters = 0, evals = 1;
    for (int it = 0; it < a.nstep; it++) {
        const double gg = dg_block_dot(g, g, n, &slot[2]);
        if (sqrt(gg / n) < a.gtol) break;
        double t0;
        if (cnt == 0) {
            for (int i = tid; i < n; i += DG_THREADS) d[i] = -g[i];
            t0 = a.drop > 0 ? fmin(1.0, a.drop / gg) : fmin(1.0, 1.0 / sqrt(gg));
        } else {
            for (int i = tid; i < n; i += DG_THREADS) d[i] = g[i];
            for (int k = 0; k < cnt; k++) {
                const int idx = ((head - 1 - k) % m + m) % m;
                al[idx] = rho[idx] * dg_block_dot(S + (size_t)idx * n, d, n, &slot[2]);
                for (int i = tid; i < n; i += DG_THREADS) d[i] = d[i] - al[idx] * Y[(size_t)idx * n + i];
            }
            const int nw = ((head - 1) % m + m) % m;
            const double sy0 = dg_block_dot(S + (size_t)nw * n, Y + (size_t)nw * n, n, &slot[2]);
            const double yy0 = dg_block_dot(Y + (size_t)nw * n, Y + (size_t)nw * n, n, &slot[3]);
            const double gam = sy0 / yy0;
            for (int i = tid; i < n; i += DG_THREADS) d[i] = d[i] * gam;
            for (int k = cnt - 1; k >= 0; k--) {
                const int idx = ((head - 1 - k) % m + m) % m;
                const double be = rho[idx] * dg_block_dot(Y + (size_t)idx * n, d, n, &slot[2]);
                const double c = al[idx] - be;
                for (int i = tid; i < n; i += DG_THREADS) d[i] = d[i] + S[(size_t)idx * n + i] * c;
            }
            for (int i = tid; i < n; i += DG_THREADS) d[i] = -d[i];
            t0 = 1.0;
        }
        const double gd = dg_block_dot(g, d, n, &slot[3]);
        if (!(gd < 0)) { if (cnt == 0) break; cnt = 0; it--; continue; }
        double tt = t0; int ok = 0; const int cnt_before = cnt;
        for (int ls = 0; ls <= a.max_ls; ls++) {
            for (int i = tid; i < n; i += DG_THREADS) xt[i] = x[i] + tt * d[i];
            et = dg_block_energy(N, U, L, a.scale, xt, gt, er, &slot[1]); evals++;
            if (et < e && et <= e + a.c1 * tt * gd) { ok = 1; break; }
            tt = tt * 0.5;
        }
        iters++;
        if (!ok) { if (cnt == 0) break; cnt = 0; continue; }
        double* sk = S + (size_t)head * n; double* yk = Y + (size_t)head * n;
        for (int i = tid; i < n; i += DG_THREADS) { sk[i] = tt * d[i]; yk[i] = gt[i] - g[i]; }
        const double sy = dg_block_dot(sk, yk, n, &slot[2]), ss = dg_block_dot(sk, sk, n, &slot[3]), yy = dg_block_dot(yk, yk, n, &slot[2]);
        if (sy > 1e-10 * sqrt(ss * yy)) { rho[head] = 1.0 / sy; head = (head + 1) % m; if (cnt < m) cnt++; }
        const int small = cnt_before > 0 && (e - et) <= a.ftol * fmax(fabs(et), 1.0);
        __syncthreads();
        for (int i = tid; i < n; i += DG_THREADS) { x[i] = xt[i]; g[i] = gt[i]; }
        e = et;
        if (small) break;
    }
    const double g2 = dg_block_dot(g, g, n, &slot[3]);
    if (tid == 0 && a.stats) {
        double* st = a.stats + 6 * (size_t)blockIdx.x;
        st[0] = dgscale; st[1] = e_start; st[2] = e; st[3] = iters; st[4] = evals; st[5] = sqrt(g2 / n);
    }
}
// scale + minimise n_models embedded structures in place (d_xyz: n_models x N x 3); d_stats: 6 per model
int dev_dg_refine(int N, const double* d_U, const double* d_L, const chr3d_sa_params& p, double scale, int nstep, double drop, int n_models,
                  double* d_xyz, double* d_stats, cudaStream_t s) {
    if (p.lbfgs_m < 1 || p.lbfgs_m > CHR3D_LBFGS_MAX_M) { set_error("dg refine: lbfgs_m out of range"); return CHR3D_ERR_ARG; }
    DevBuf<double> work;
    const size_t per = (size_t)3 * N * (4 + 2 * p.lbfgs_m) + 3 * (size_t)N;
    CHR3D_CUDA(work.alloc(per * n_models));
    DgRefineArgs a;
    a.N = N; a.nstep = nstep; a.m = p.lbfgs_m; a.max_ls = p.lbfgs_max_ls; a.scale = scale; a.drop = drop; a.c1 = p.lbfgs_c1; a.gtol = p.lbfgs_gtol; a.ftol = p.lbfgs_ftol;
    a.U = d_U; a.L = d_L; a.xyz = d_xyz; a.work = work.p; a.stats = d_stats;
    dg_refine_kernel<<<n_models, DG_THREADS, 0, s>>>(a);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaStreamSynchronize(s));
    return 0;
}

}  // namespace chr3d

using namespace chr3d;

// Starting coordinates by distance-geometry embedding for n_models models of one restraint table (0-based ri < rj).
// xyz_out: n_models*N*3; eig3_out (optional): the three eigenvalues used, per model; n_inconsistent (optional): pairs whose
// bounds the triangle inequality contradicted.  Pass the result as chr3d_job.x0 to anneal from it instead of a random coil.
static int dg_stage(const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr, int32_t N, const chr3d_sa_params* p,
                    int32_t n_models, uint64_t seed, double* xyz_out, double* eig3_out, int64_t* n_inconsistent,
                    bool refine, double dg_scale, int nstep, double drop, double* stats_out) {
    if (N < 2 || n_models < 1 || !p || !xyz_out || nr < 0 || (nr > 0 && (!ri || !rj || !rd))) { set_error("chr3d_dg_embed / chr3d_dg_starts: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    for (int64_t k = 0; k < nr; k++) if (ri[k] < 0 || rj[k] < 0 || ri[k] >= N || rj[k] >= N || ri[k] == rj[k]) { set_error("chr3d_dg_embed: restraint %lld out of range", (long long)k); return CHR3D_ERR_ARG; }
    DevBuf<int32_t> di, dj; DevBuf<double> dd, U, L, x, e3;
    const size_t nn = (size_t)N * N;
    CHR3D_CUDA(di.alloc((size_t)std::max<int64_t>(nr, 1))); CHR3D_CUDA(dj.alloc((size_t)std::max<int64_t>(nr, 1))); CHR3D_CUDA(dd.alloc((size_t)std::max<int64_t>(nr, 1)));
    if (nr > 0) {
        CHR3D_CUDA(cudaMemcpy(di.p, ri, sizeof(int32_t) * (size_t)nr, cudaMemcpyHostToDevice));
        CHR3D_CUDA(cudaMemcpy(dj.p, rj, sizeof(int32_t) * (size_t)nr, cudaMemcpyHostToDevice));
        CHR3D_CUDA(cudaMemcpy(dd.p, rd, sizeof(double) * (size_t)nr, cudaMemcpyHostToDevice));
    }
    CHR3D_CUDA(U.alloc(nn)); CHR3D_CUDA(L.alloc(nn)); CHR3D_CUDA(x.alloc((size_t)n_models * N * 3)); CHR3D_CUDA(e3.alloc((size_t)n_models * 3));
    long long bad = 0, slack = 0;
    if (int rc = dev_dg_bounds(N, di.p, dj.p, dd.p, nr, *p, U.p, L.p, &bad, &slack, 0)) return rc;
    // no pair keeps any slack (the usual case for dense Hi-C tables: the smoothed bounds meet everywhere): every model draws the
    // same distance matrix, so one embedding serves all of them (bit-identical to embedding each)
    const int n_embed = slack == 0 ? 1 : n_models;
    if (int rc = dev_dg_embed(N, U.p, L.p, seed ? seed : p->seed, 0, n_embed, x.p, e3.p, 0)) return rc;
    DevBuf<double> st;
    CHR3D_CUDA(st.alloc((size_t)n_models * 6));
    if (refine) { if (int rc = dev_dg_refine(N, U.p, L.p, *p, dg_scale, nstep, drop, n_embed, x.p, st.p, 0)) return rc; }
    for (int m = n_embed; m < n_models; m++) {
        CHR3D_CUDA(cudaMemcpy(x.p + (size_t)m * N * 3, x.p, sizeof(double) * (size_t)N * 3, cudaMemcpyDeviceToDevice));
        CHR3D_CUDA(cudaMemcpy(e3.p + (size_t)m * 3, e3.p, sizeof(double) * 3, cudaMemcpyDeviceToDevice));
        if (refine) CHR3D_CUDA(cudaMemcpy(st.p + (size_t)m * 6, st.p, sizeof(double) * 6, cudaMemcpyDeviceToDevice));
    }
    CHR3D_CUDA(cudaMemcpy(xyz_out, x.p, sizeof(double) * (size_t)n_models * N * 3, cudaMemcpyDeviceToHost));
    if (eig3_out) CHR3D_CUDA(cudaMemcpy(eig3_out, e3.p, sizeof(double) * (size_t)n_models * 3, cudaMemcpyDeviceToHost));
    if (refine && stats_out) CHR3D_CUDA(cudaMemcpy(stats_out, st.p, sizeof(double) * (size_t)n_models * 6, cudaMemcpyDeviceToHost));
    if (n_inconsistent) *n_inconsistent = bad;
    return 0;
}
CHR3D_API int chr3d_dg_embed(const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr, int32_t N, const chr3d_sa_params* p,
                             int32_t n_models, uint64_t seed, double* xyz_out, double* eig3_out, int64_t* n_inconsistent) {
    return dg_stage(ri, rj, rd, nr, N, p, n_models, seed, xyz_out, eig3_out, n_inconsistent, false, 0.0, 0, 0.0, nullptr);
}
// The whole DG stage (C3D:1470-1526): bounds, embedding, x *= dgscale, nstep (md.dg.step = 200) L-BFGS iterations with drop (1.)
// on the bounds-violation energy scaled by dg_scale (md.dg.scale = 100).  stats_out (optional): 6 doubles per model
// {dgscale, E after scaling, E final, iterations, energy evaluations, rms gradient}.
CHR3D_API int chr3d_dg_starts(const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr, int32_t N, const chr3d_sa_params* p,
                              int32_t n_models, uint64_t seed, double dg_scale, int32_t nstep, double drop,
                              double* xyz_out, double* eig3_out, double* stats_out, int64_t* n_inconsistent) {
    if (!(dg_scale > 0) || nstep < 0) { set_error("chr3d_dg_starts: dg_scale must be positive and nstep >= 0"); return CHR3D_ERR_ARG; }
    return dg_stage(ri, rj, rd, nr, N, p, n_models, seed, xyz_out, eig3_out, n_inconsistent, true, dg_scale, nstep, drop, stats_out);
}
