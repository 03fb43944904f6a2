// sa_ensemble.cu — the small-N hot path: one thread-block cluster per SA trajectory.
//
// Replaces the CNS `dgsa.inp` run (C3D:254-289 launches it; protocol text C3D:1573-1829) for
// chromosomes up to ~1000 beads.  One launch anneals every model of every (matrix, alpha) job:
//
//   * a trajectory = one cluster of C CTAs (C = 1..8, chosen on the host from N and the number of
//     trajectories so that the 148 SMs are filled);  CTA `rank` owns a contiguous block of beads
//     ("rows"), integrates them, and after every update stores them into the coordinate array of
//     every CTA of the cluster through distributed shared memory;
//   * coordinates of all N beads live in shared memory (float4), the CTA's column block of the
//     dense target-distance matrix is staged into shared memory once with TMA bulk copies
//     (cp.async.bulk -> UBLKCP) when it fits, otherwise it is read through L2;
//   * force evaluation is the "gather" form: thread (row i, chunk q) walks j over its chunk with
//     x_j broadcast from shared memory and T[j][i] read conflict-free; no atomics, deterministic;
//   * NOE soft-square + repel are fused in that one pair loop, the bond term is added by the row
//     owner; energies are reduced with warp shuffles (fp64 across threads);
//   * the whole stage program (pre-minimisation, hot MD, cooling, L-BFGS, C3D:1573-1829) runs
//     inside the kernel; scalar reductions (energies, kinetic energy, L-BFGS inner products)
//     cross the cluster with one DSMEM exchange + one cluster barrier.
//
// L-BFGS is the vector-free formulation: all inner products of an iteration are reduced in one
// exchange and the two-loop recursion runs on the (2m+1)x(2m+1) Gram matrix.
#include <cooperative_groups.h>
#include <math.h>
#include "common.h"

namespace cg = cooperative_groups;

namespace chr3d {

#define SA_MAXC 8
#define SA_KTOT 64           // scalars per all-reduce (2 pair energies + owner columns)
#define SA_MAXB (2 * CHR3D_LBFGS_MAX_M + 1)

#define STREAM_COIL 0x00C011u
#define STREAM_MAXWELL 0x100000u

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (same definition as the oracle's)
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                           uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ inline void normal3(unsigned long long seed, uint32_t bead, uint32_t stream, uint32_t model, double z[3]) {
    uint32_t r[4];
    philox4x32(bead, stream, model, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double s = 1.0 / 4294967296.0;
    double u1 = (r[0] + 0.5) * s, u2 = (r[1] + 0.5) * s, u3 = (r[2] + 0.5) * s, u4 = (r[3] + 0.5) * s;
    double ra = sqrt(-2.0 * log(u1)), rb = sqrt(-2.0 * log(u3));
    const double twopi = 6.283185307179586476925286766559;
    z[0] = ra * cos(twopi * u2); z[1] = ra * sin(twopi * u2); z[2] = rb * cos(twopi * u4);
}

// seeded random-coil start: one thread per model walks the chain (N <= a few thousand)
__global__ void coil_kernel(float* x0, int N, int n_models, int first_model, double step, unsigned long long seed) {
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_models) return;
    float* x = x0 + (size_t)m * N * 3;
    double p[3] = {0, 0, 0}, c[3] = {0, 0, 0};
    for (int i = 0; i < N; i++) {
        if (i > 0) {
            double z[3]; normal3(seed, (uint32_t)i, STREAM_COIL, (uint32_t)(first_model + m), z);
            double nrm = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
            if (nrm < 1e-12) { z[0] = 1; z[1] = z[2] = 0; nrm = 1; }
            for (int k = 0; k < 3; k++) p[k] += step * z[k] / nrm;
        }
        for (int k = 0; k < 3; k++) c[k] += p[k];
    }
    // second pass: recompute and centre (keeps the kernel free of per-model scratch)
    p[0] = p[1] = p[2] = 0;
    for (int i = 0; i < N; i++) {
        if (i > 0) {
            double z[3]; normal3(seed, (uint32_t)i, STREAM_COIL, (uint32_t)(first_model + m), z);
            double nrm = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
            if (nrm < 1e-12) { z[0] = 1; z[1] = z[2] = 0; nrm = 1; }
            for (int k = 0; k < 3; k++) p[k] += step * z[k] / nrm;
        }
        for (int k = 0; k < 3; k++) x[3 * i + k] = (float)(p[k] - c[k] / N);
    }
}
int dev_random_coil(float* d_x0, int N, int n_models, int first_model, double step, unsigned long long seed,
                    cudaStream_t s) {
    coil_kernel<<<(n_models + 63) / 64, 64, 0, s>>>(d_x0, N, n_models, first_model, step, seed);
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// shared-memory layout (identical on host and device)
// ------------------------------------------------------------------------------------------------
struct SaShape {
    int N, C, nthreads, nwarps, rw_max, rows_pad, Q, m, nvec, stage_targets;
};
struct SaLayout {
    size_t xs, fpart, vec, P, ewarp, slots, gsum, B, delta, mbar, tblk, total;
};
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline void sa_shape(int N, int C, int nthreads, int m, int stage_targets, SaShape& s) {
    s.N = N; s.C = C; s.nthreads = nthreads; s.nwarps = nthreads / 32; s.m = m; s.stage_targets = stage_targets;
    int rw_total = (N + 31) / 32;
    s.rw_max = (rw_total + C - 1) / C;
    s.rows_pad = 32 * s.rw_max;
    s.Q = s.nwarps / s.rw_max; if (s.Q < 1) s.Q = 1;
    s.nvec = 3 + 2 * m;     // XB, G, PV, S[m], Y[m]
}
__host__ __device__ inline void sa_layout(const SaShape& s, SaLayout& L) {
    size_t o = 0;
    L.xs = o;    o = align_up(o + sizeof(float4) * (size_t)(s.N + 1), 16);
    L.fpart = o; o = align_up(o + sizeof(float4) * (size_t)s.Q * s.rows_pad, 16);
    L.vec = o;   o = align_up(o + sizeof(float) * (size_t)s.nvec * 3 * s.rows_pad, 16);
    L.P = o;     o = align_up(o + sizeof(float) * (size_t)(SA_KTOT - 2) * s.rows_pad, 16);
    L.ewarp = o; o = align_up(o + sizeof(double) * 2 * (size_t)s.nwarps, 16);
    L.slots = o; o = align_up(o + sizeof(double) * 2 * SA_MAXC * SA_KTOT, 16);
    L.gsum = o;  o = align_up(o + sizeof(double) * SA_KTOT, 16);
    L.B = o;     o = align_up(o + sizeof(double) * SA_MAXB * SA_MAXB, 16);
    L.delta = o; o = align_up(o + sizeof(double) * (SA_MAXB + 3), 16);
    L.mbar = o;  o = align_up(o + 16, 128);
    L.tblk = o;
    if (s.stage_targets) o += sizeof(float) * (size_t)s.N * s.rows_pad;
    L.total = o;
}

// ------------------------------------------------------------------------------------------------
// device context
// ------------------------------------------------------------------------------------------------
struct Ctx {
    // geometry
    int N, C, rank, tid, lane, warp, nwarps, nthreads;
    int r0, rows, rows_pad, Q;       // owned rows [r0, r0+rows)
    int pw_row, pw_q, pw_active;     // pair-loop role of this warp
    int m;
    // smem
    float4* xs; float4* fpart; float* vec; float* P; double* ewarp; double* slots; double* gsum; double* B;
    double* delta;
    const float* Tsrc; int tstride;  // T[j][r0+li] = Tsrc[j*tstride + li]
    int par;
    // owner state (valid when tid < rows_pad; row_ok when tid < rows)
    bool owner, row_ok; int row;
    float x[3], v[3], f[3];
    float ebond;
    long long nevals;
    DevParams prm;
};

__device__ __forceinline__ float* vecp(const Ctx& c, int v, int comp) { return c.vec + ((size_t)(v * 3 + comp)) * c.rows_pad; }
#define V_XB 0
#define V_G 1
#define V_PV 2
#define V_S(l) (3 + (l))
#define V_Y(l) (3 + c.m + (l))

__device__ __forceinline__ void csync(const Ctx& c) {
    if (c.C > 1) cg::this_cluster().sync(); else __syncthreads();
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// store the owner's coordinates into the coordinate array of every CTA of the cluster
__device__ __forceinline__ void publish_x(const Ctx& c) {
    if (!c.row_ok) return;
    float4 val = make_float4(c.x[0], c.x[1], c.x[2], 0.f);
    if (c.C > 1) {
        cg::cluster_group cl = cg::this_cluster();
        for (int r = 0; r < c.C; r++) cl.map_shared_rank(c.xs, r)[c.row] = val;
    } else {
        c.xs[c.row] = val;
    }
}

// All-reduce of K owner columns (P) + the two pair-energy columns (ewarp) across the cluster.
// Result: gsum[0]=E_noe, gsum[1]=E_vdw, gsum[2+k]=column k.  Every CTA sums the same partials in
// the same order, so all ranks take identical decisions afterwards.
__device__ void allreduce(Ctx& c, int K) {
    __syncthreads();
    const int total = K + 2;
    for (int k = c.warp; k < total; k += c.nwarps) {
        double s = 0;
        if (k < 2) { for (int w = c.lane; w < c.nwarps; w += 32) s += c.ewarp[w * 2 + k]; }
        else { const float* col = c.P + (size_t)(k - 2) * c.rows_pad; for (int t = c.lane; t < c.rows_pad; t += 32) s += (double)col[t]; }
        s = warp_sum(s);
        if (c.C > 1) {
            if (c.lane < c.C) {
                double* remote = cg::this_cluster().map_shared_rank(c.slots, c.lane);
                remote[((size_t)c.par * SA_MAXC + c.rank) * SA_KTOT + k] = s;
            }
        } else if (c.lane == 0) c.gsum[k] = s;
    }
    if (c.C > 1) {
        cg::this_cluster().sync();
        if (c.tid < total) {
            double s = 0;
            for (int r = 0; r < c.C; r++) s += c.slots[((size_t)c.par * SA_MAXC + r) * SA_KTOT + c.tid];
            c.gsum[c.tid] = s;
        }
        c.par ^= 1;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// energy + forces for the owned rows.  On return: c.f (owner), c.ebond (owner partial),
// c.ewarp[warp] = (E_noe, E_vdw) partials.  The caller must follow with allreduce().
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_default(const DevParams& p, float wnC, float wv4, float wv, float R2, float lo,
                                             float xi, float yi, float zi, float4 xj, float d,
                                             float& fx, float& fy, float& fz, float& en, float& ev) {
    float dx = xi - xj.x, dy = yi - xj.y, dz = zi - xj.z;
    float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    float r2c = fmaxf(r2, 1e-12f);
    float rinv = rsqrtf(r2c);
    float r = r2c * rinv;
    float delta = r - d;
    float h = fminf(fmaxf(delta, lo), p.rsw);
    float w = d > 0.f ? wnC : 0.f;            // w_noe * C
    float wh = w * h;
    en = fmaf(wh, 2.f * delta - h, en);
    float tr = fmaxf(R2 - r2, 0.f);
    tr = d > -1.5f ? tr : 0.f;
    ev = fmaf(wv * tr, tr, ev);
    float coef = fmaf(wv4, tr, -2.f * wh * rinv);   // 4 w_vdw t - (dE_noe/dr)/r
    fx = fmaf(coef, dx, fx); fy = fmaf(coef, dy, fy); fz = fmaf(coef, dz, fz);
}
__device__ __forceinline__ void pair_general(const DevParams& p, float wn, float wv4, float wv, float R2,
                                             float xi, float yi, float zi, float4 xj, float d,
                                             float& fx, float& fy, float& fz, float& en, float& ev) {
    float dx = xi - xj.x, dy = yi - xj.y, dz = zi - xj.z;
    float r2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    float r2c = fmaxf(r2, 1e-12f);
    float rinv = rsqrtf(r2c);
    float r = r2c * rinv;
    float e = 0.f, de = 0.f;
    if (d > 0.f) {
        float delta = r - d, u = fabsf(delta), sg = delta > 0.f ? 1.f : -1.f;
        bool soft = u > p.rsw && (delta > 0.f || p.lo_clamp > -1e30f);
        if (!soft) { e = p.C * u * u; de = 2.f * p.C * u * sg; }
        else { e = p.a + p.b / u + p.casym * u; de = (p.casym - p.b / (u * u)) * sg; }
    }
    en = fmaf(wn, e, en);
    float tr = fmaxf(R2 - r2, 0.f);
    tr = d > -1.5f ? tr : 0.f;
    ev = fmaf(wv * tr, tr, ev);
    float coef = fmaf(wv4, tr, -wn * de * rinv);
    fx = fmaf(coef, dx, fx); fy = fmaf(coef, dy, fy); fz = fmaf(coef, dz, fz);
}

__device__ void eval(Ctx& c, const DevStage& st, bool count) {
    csync(c);                                       // coordinates of every CTA are in place
    const DevParams& p = c.prm;
    float fx = 0.f, fy = 0.f, fz = 0.f, en = 0.f, ev = 0.f;
    if (c.pw_active) {
        const int li = 32 * c.pw_row + c.lane;
        int i = c.r0 + li; if (i >= c.N) i = c.N - 1;
        const float4 xi4 = c.xs[i];
        const float xi = xi4.x, yi = xi4.y, zi = xi4.z;
        const int j0 = (int)(((long long)c.N * c.pw_q) / c.Q), j1 = (int)(((long long)c.N * (c.pw_q + 1)) / c.Q);
        const float* tp = c.Tsrc + li;
        const int ts = c.tstride;
        const float wv = st.w_vdw, wv4 = 4.f * st.w_vdw, R2 = st.R2;
        if (!p.general_noe) {
            const float wnC = st.w_noe * p.C, lo = p.lo_clamp;
            int j = j0;
            for (; j + 4 <= j1; j += 4) {
                float4 a0 = c.xs[j], a1 = c.xs[j + 1], a2 = c.xs[j + 2], a3 = c.xs[j + 3];
                float d0 = tp[(size_t)j * ts], d1 = tp[(size_t)(j + 1) * ts], d2 = tp[(size_t)(j + 2) * ts], d3 = tp[(size_t)(j + 3) * ts];
                pair_default(p, wnC, wv4, wv, R2, lo, xi, yi, zi, a0, d0, fx, fy, fz, en, ev);
                pair_default(p, wnC, wv4, wv, R2, lo, xi, yi, zi, a1, d1, fx, fy, fz, en, ev);
                pair_default(p, wnC, wv4, wv, R2, lo, xi, yi, zi, a2, d2, fx, fy, fz, en, ev);
                pair_default(p, wnC, wv4, wv, R2, lo, xi, yi, zi, a3, d3, fx, fy, fz, en, ev);
            }
            for (; j < j1; j++) pair_default(p, wnC, wv4, wv, R2, lo, xi, yi, zi, c.xs[j], tp[(size_t)j * ts], fx, fy, fz, en, ev);
        } else {
            for (int j = j0; j < j1; j++) pair_general(p, st.w_noe, wv4, wv, R2, xi, yi, zi, c.xs[j], tp[(size_t)j * ts], fx, fy, fz, en, ev);
        }
        c.fpart[(size_t)c.pw_q * c.rows_pad + li] = make_float4(fx, fy, fz, 0.f);
    }
    // every unordered pair is visited from both of its rows
    double den = warp_sum(0.5 * (double)en), dev = warp_sum(0.5 * (double)ev);
    if (c.lane == 0) { c.ewarp[c.warp * 2] = den; c.ewarp[c.warp * 2 + 1] = dev; }
    __syncthreads();
    if (c.owner) {
        float gx = 0.f, gy = 0.f, gz = 0.f, eb = 0.f;
        if (c.row_ok) {
            for (int q = 0; q < c.Q; q++) { float4 t = c.fpart[(size_t)q * c.rows_pad + c.tid]; gx += t.x; gy += t.y; gz += t.z; }
            const float4 xi = c.xs[c.row];
            for (int s = -1; s <= 1; s += 2) {          // bonds (i-1,i) and (i,i+1)
                int nb = c.row + s;
                if (nb < 0 || nb >= c.N) continue;
                const float4 xn = c.xs[nb];
                float dx = xi.x - xn.x, dy = xi.y - xn.y, dz = xi.z - xn.z;
                float r2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-12f);
                float rinv = rsqrtf(r2), r = r2 * rinv, u = r - p.r0;
                float g = 2.f * st.w_bond * u * rinv;
                gx -= g * dx; gy -= g * dy; gz -= g * dz;
                if (s == 1) eb = st.w_bond * u * u;
            }
        }
        c.f[0] = gx; c.f[1] = gy; c.f[2] = gz; c.ebond = eb;
    }
    if (count) c.nevals++;
}

// ------------------------------------------------------------------------------------------------
// L-BFGS direction from the Gram matrix (vector-free two-loop recursion); one thread.
// basis index: S slot l -> l, Y slot l -> m+l, gradient -> 2m.  Writes delta[0..2m] such that
// p = -sum_l delta[l] b_l and returns g.p.
// ------------------------------------------------------------------------------------------------
__device__ double lbfgs_direction(double* B, double* delta, int m, int cnt, int head) {
    const int nb = 2 * m + 1, ig = 2 * m;
    double al[CHR3D_LBFGS_MAX_M];
    for (int l = 0; l < nb; l++) delta[l] = 0.0;
    delta[ig] = 1.0;
    for (int k = 0; k < cnt; k++) {
        int idx = ((head - 1 - k) % m + m) % m;
        double s = delta[ig] * B[ig * SA_MAXB + idx];
        for (int kk = 0; kk < cnt; kk++) { int l = ((head - 1 - kk) % m + m) % m;
            s += delta[l] * B[l * SA_MAXB + idx] + delta[m + l] * B[(m + l) * SA_MAXB + idx]; }
        al[idx] = s / B[idx * SA_MAXB + m + idx];
        delta[m + idx] -= al[idx];
    }
    int nw = ((head - 1) % m + m) % m;
    double gam = B[nw * SA_MAXB + m + nw] / B[(m + nw) * SA_MAXB + m + nw];
    for (int l = 0; l < nb; l++) delta[l] *= gam;
    for (int k = cnt - 1; k >= 0; k--) {
        int idx = ((head - 1 - k) % m + m) % m;
        double s = delta[ig] * B[ig * SA_MAXB + m + idx];
        for (int kk = 0; kk < cnt; kk++) { int l = ((head - 1 - kk) % m + m) % m;
            s += delta[l] * B[l * SA_MAXB + m + idx] + delta[m + l] * B[(m + l) * SA_MAXB + m + idx]; }
        double be = s / B[idx * SA_MAXB + m + idx];
        delta[idx] += al[idx] - be;
    }
    double gp = delta[ig] * B[ig * SA_MAXB + ig];
    for (int kk = 0; kk < cnt; kk++) { int l = ((head - 1 - kk) % m + m) % m;
        gp += delta[l] * B[l * SA_MAXB + ig] + delta[m + l] * B[(m + l) * SA_MAXB + ig]; }
    return -gp;
}

// column indices of the L-BFGS inner products inside P (column 0 is the bond energy)
#define DOT_L(l, w) (1 + 6 * (l) + (w))     // w: 0 s.S 1 s.Y 2 y.S 3 y.Y 4 g.S 5 g.Y
#define DOT_X(m, w) (1 + 6 * (m) + (w))     // w: 0 ss 1 sy 2 sg 3 yy 4 yg 5 gg

__device__ void op_minimize(Ctx& c, const DevStage& st) {
    const DevParams& p = c.prm;
    const int m = c.m, t = c.tid;
    const double n3 = 3.0 * c.N;
    eval(c, st, true);
    if (c.owner) {
        float gg = 0.f;
        for (int k = 0; k < 3; k++) { vecp(c, V_G, k)[t] = -c.f[k]; vecp(c, V_XB, k)[t] = c.x[k]; gg += c.f[k] * c.f[k]; }
        c.P[t] = c.ebond; c.P[c.rows_pad + t] = c.row_ok ? gg : 0.f;
    }
    allreduce(c, 2);
    double E = c.gsum[0] + c.gsum[1] + c.gsum[2], gg = c.gsum[3];
    int cnt = 0, head = 0, it = 0;
    double gp = 0.0;
    while (it < st.nstep) {
        if (sqrt(gg / n3) < (double)p.gtol) break;
        double t0 = 1.0;
        if (cnt == 0) {
            if (c.owner) for (int k = 0; k < 3; k++) vecp(c, V_PV, k)[t] = -vecp(c, V_G, k)[t];
            gp = -gg;
            t0 = st.drop > 0.f ? fmin(1.0, (double)st.drop / gg) : fmin(1.0, 1.0 / sqrt(gg));
        }
        if (!(gp < 0.0)) { if (cnt == 0) break; cnt = 0; continue; }
        double ts = t0, Et = 0.0; bool ok = false;
        float gn[3] = {0.f, 0.f, 0.f}, sv[3] = {0.f, 0.f, 0.f}, yv[3] = {0.f, 0.f, 0.f};
        for (int ls = 0; ls <= p.max_ls; ls++) {
            if (c.owner) {
                for (int k = 0; k < 3; k++) { sv[k] = (float)ts * vecp(c, V_PV, k)[t]; c.x[k] = vecp(c, V_XB, k)[t] + sv[k]; }
                publish_x(c);
            }
            eval(c, st, true);
            if (c.owner) {
                for (int k = 0; k < 3; k++) { gn[k] = -c.f[k]; yv[k] = gn[k] - vecp(c, V_G, k)[t]; }
                if (!c.row_ok) for (int k = 0; k < 3; k++) { gn[k] = sv[k] = yv[k] = 0.f; }
                const int rp = c.rows_pad;
                c.P[t] = c.ebond;
                for (int l = 0; l < m; l++) {
                    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (l != head) for (int k = 0; k < 3; k++) {
                        float sl = vecp(c, V_S(l), k)[t], yl = vecp(c, V_Y(l), k)[t];
                        a[0] = fmaf(sv[k], sl, a[0]); a[1] = fmaf(sv[k], yl, a[1]); a[2] = fmaf(yv[k], sl, a[2]);
                        a[3] = fmaf(yv[k], yl, a[3]); a[4] = fmaf(gn[k], sl, a[4]); a[5] = fmaf(gn[k], yl, a[5]);
                    }
                    for (int w = 0; w < 6; w++) c.P[(size_t)DOT_L(l, w) * rp + t] = a[w];
                }
                float b[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int k = 0; k < 3; k++) {
                    b[0] = fmaf(sv[k], sv[k], b[0]); b[1] = fmaf(sv[k], yv[k], b[1]); b[2] = fmaf(sv[k], gn[k], b[2]);
                    b[3] = fmaf(yv[k], yv[k], b[3]); b[4] = fmaf(yv[k], gn[k], b[4]); b[5] = fmaf(gn[k], gn[k], b[5]);
                }
                for (int w = 0; w < 6; w++) c.P[(size_t)DOT_X(m, w) * rp + t] = b[w];
            }
            allreduce(c, 1 + 6 * m + 6);
            Et = c.gsum[0] + c.gsum[1] + c.gsum[2];
            if (Et < E && Et <= E + (double)p.c1 * ts * gp) { ok = true; break; }
            ts *= 0.5;
        }
        it++;
        if (!ok) { if (cnt == 0) break; cnt = 0; continue; }
        const double* G = c.gsum + 2;       // owner columns
        double ss = G[DOT_X(m, 0)], sy = G[DOT_X(m, 1)], yy = G[DOT_X(m, 3)];
        bool keep = sy > 1e-10 * sqrt(ss * yy);
        if (c.owner) {
            for (int k = 0; k < 3; k++) {
                if (keep) { vecp(c, V_S(head), k)[t] = sv[k]; vecp(c, V_Y(head), k)[t] = yv[k]; }
                vecp(c, V_XB, k)[t] = c.x[k]; vecp(c, V_G, k)[t] = gn[k];
            }
        }
        const bool small = cnt > 0 && (E - Et) <= (double)p.ftol * fmax(fabs(Et), 1.0);
        E = Et; gg = G[DOT_X(m, 5)];
        if (small) break;
        if (keep) {
            if (t == 0) {
                double* B = c.B; const int ig = 2 * m, h = head;
                for (int l = 0; l < m; l++) {
                    if (l == h) continue;
                    B[h * SA_MAXB + l] = B[l * SA_MAXB + h] = G[DOT_L(l, 0)];
                    B[h * SA_MAXB + m + l] = B[(m + l) * SA_MAXB + h] = G[DOT_L(l, 1)];
                    B[(m + h) * SA_MAXB + l] = B[l * SA_MAXB + m + h] = G[DOT_L(l, 2)];
                    B[(m + h) * SA_MAXB + m + l] = B[(m + l) * SA_MAXB + m + h] = G[DOT_L(l, 3)];
                    B[ig * SA_MAXB + l] = B[l * SA_MAXB + ig] = G[DOT_L(l, 4)];
                    B[ig * SA_MAXB + m + l] = B[(m + l) * SA_MAXB + ig] = G[DOT_L(l, 5)];
                }
                B[h * SA_MAXB + h] = ss; B[h * SA_MAXB + m + h] = B[(m + h) * SA_MAXB + h] = sy;
                B[(m + h) * SA_MAXB + m + h] = yy;
                B[ig * SA_MAXB + h] = B[h * SA_MAXB + ig] = G[DOT_X(m, 2)];
                B[ig * SA_MAXB + m + h] = B[(m + h) * SA_MAXB + ig] = G[DOT_X(m, 4)];
                B[ig * SA_MAXB + ig] = gg;
            }
            head = (head + 1) % m; if (cnt < m) cnt++;
            if (t == 0) c.delta[SA_MAXB] = lbfgs_direction(c.B, c.delta, m, cnt, head);
            __syncthreads();
            gp = c.delta[SA_MAXB];
            if (c.owner) {
                for (int k = 0; k < 3; k++) {
                    float acc = (float)c.delta[2 * m] * vecp(c, V_G, k)[t];
                    for (int kk = 0; kk < cnt; kk++) {
                        int l = ((head - 1 - kk) % m + m) % m;
                        acc = fmaf((float)c.delta[l], vecp(c, V_S(l), k)[t], acc);
                        acc = fmaf((float)c.delta[m + l], vecp(c, V_Y(l), k)[t], acc);
                    }
                    vecp(c, V_PV, k)[t] = -acc;
                }
            }
            __syncthreads();   // delta is rewritten by thread 0 in the next iteration
        } else {
            cnt = 0;           // curvature condition failed: restart from steepest descent
        }
    }
    // leave x = last accepted point everywhere
    if (c.owner) { for (int k = 0; k < 3; k++) c.x[k] = vecp(c, V_XB, k)[t]; }
    csync(c);
    if (c.owner) publish_x(c);
}

__device__ void op_dynamics(Ctx& c, const DevStage& st) {
    const DevParams& p = c.prm;
    const int t = c.tid, rp = c.rows_pad;
    // cmremove=true
    if (c.owner) { for (int k = 0; k < 3; k++) c.P[(size_t)k * rp + t] = c.row_ok ? c.v[k] : 0.f; }
    if (c.lane == 0) { c.ewarp[c.warp * 2] = 0.0; c.ewarp[c.warp * 2 + 1] = 0.0; }
    allreduce(c, 3);
    if (c.owner && c.row_ok) for (int k = 0; k < 3; k++) c.v[k] -= (float)(c.gsum[2 + k] / c.N);
    eval(c, st, true);
    if (c.owner) {
        float ke = 0.f; if (c.row_ok) for (int k = 0; k < 3; k++) ke += c.v[k] * c.v[k];
        c.P[t] = c.ebond; c.P[rp + t] = 0.5f * p.mass * ke;
    }
    allreduce(c, 2);
    double ke = c.gsum[3];
    const double ndof = 3.0 * c.N - 3.0;
    const float h = 0.5f * st.dt_akma * p.inv_mass;
    for (int step = 0; step < st.nstep; step++) {
        double T = 2.0 * ke / ((double)p.kboltz * ndof), lam = 1.0;
        if (T > 0.0) {
            if (st.flag == CHR3D_COUPLE_TCOUP) lam = sqrt(fmax(0.0, 1.0 + 2.0 * (double)p.fbeta * (double)st.dt_ps * ((double)st.temp / T - 1.0)));
            else if (st.flag == CHR3D_COUPLE_VSCALE) lam = sqrt((double)st.temp / T);
        }
        if (c.owner && c.row_ok) {
            const float lf = (float)lam;
            for (int k = 0; k < 3; k++) { c.v[k] = fmaf(h, c.f[k], lf * c.v[k]); c.x[k] = fmaf(st.dt_akma, c.v[k], c.x[k]); }
            publish_x(c);
        }
        eval(c, st, true);
        if (c.owner) {
            float k2 = 0.f;
            if (c.row_ok) for (int k = 0; k < 3; k++) { c.v[k] = fmaf(h, c.f[k], c.v[k]); k2 += c.v[k] * c.v[k]; }
            c.P[t] = c.ebond; c.P[rp + t] = 0.5f * p.mass * k2;
        }
        allreduce(c, 2);
        ke = c.gsum[3];
    }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(512, 1)
sa_kernel(const TrajDesc* __restrict__ traj, DevParams prm, const DevStage* __restrict__ prog, int nstages,
          int stage_targets) {
    extern __shared__ __align__(128) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    Ctx c;
    c.C = (int)cluster.num_blocks();
    c.rank = (int)cluster.block_rank();
    const TrajDesc td = traj[blockIdx.x / c.C];
    c.N = td.N; c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    c.nthreads = blockDim.x; c.nwarps = c.nthreads >> 5; c.m = prm.m; c.prm = prm; c.par = 0; c.nevals = 0;

    SaShape sh; sa_shape(c.N, c.C, c.nthreads, c.m, stage_targets, sh);
    SaLayout L; sa_layout(sh, L);
    c.xs = (float4*)(smem + L.xs); c.fpart = (float4*)(smem + L.fpart); c.vec = (float*)(smem + L.vec);
    c.P = (float*)(smem + L.P); c.ewarp = (double*)(smem + L.ewarp); c.slots = (double*)(smem + L.slots);
    c.gsum = (double*)(smem + L.gsum); c.B = (double*)(smem + L.B); c.delta = (double*)(smem + L.delta);
    c.rows_pad = sh.rows_pad; c.Q = sh.Q;

    // row-warps of this rank: balanced contiguous split of ceil(N/32) row-warps over C ranks
    const int rw_total = (c.N + 31) / 32;
    const int rw0 = (int)(((long long)rw_total * c.rank) / c.C), rw1 = (int)(((long long)rw_total * (c.rank + 1)) / c.C);
    const int my_rw = rw1 - rw0;
    c.r0 = rw0 * 32;
    c.rows = min(c.N, rw1 * 32) - c.r0; if (c.rows < 0) c.rows = 0;
    // pair-loop role: warp -> (row-warp, chunk); Q chunks for every row-warp of this rank
    c.pw_row = my_rw > 0 ? c.warp % my_rw : 0;
    c.pw_q = my_rw > 0 ? c.warp / my_rw : 0;
    c.pw_active = (my_rw > 0 && c.pw_q < c.Q) ? 1 : 0;
    c.owner = c.tid < c.rows_pad;
    c.row_ok = c.tid < c.rows;
    c.row = c.r0 + c.tid;
    for (int k = 0; k < 3; k++) { c.x[k] = 0.f; c.v[k] = 0.f; c.f[k] = 0.f; }
    c.ebond = 0.f;

    // fpart rows of inactive chunks / padded rows must read as zero
    for (int i = c.tid; i < c.Q * c.rows_pad; i += c.nthreads) c.fpart[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = c.tid; i < 2 * SA_MAXC * SA_KTOT; i += c.nthreads) c.slots[i] = 0.0;
    for (int i = c.tid; i < SA_MAXB * SA_MAXB; i += c.nthreads) c.B[i] = 0.0;
    if (c.tid == 0) c.xs[c.N] = make_float4(0.f, 0.f, 0.f, 0.f);

    // targets: column block [r0, r0+rows_pad) of the symmetric matrix, i.e. T[j][r0+li]
    if (stage_targets && my_rw > 0) {
        float* tblk = (float*)(smem + L.tblk);
        uint64_t* mbar = (uint64_t*)(smem + L.mbar);
        const uint32_t row_bytes = (uint32_t)(my_rw * 32 * sizeof(float));
        if (c.tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (c.warp == 0) {
            if (c.lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(row_bytes * (uint32_t)c.N) : "memory");
            __syncwarp();
            for (int j = c.lane; j < c.N; j += 32) {
                const float* src = td.T + (size_t)j * td.pitch + c.r0;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tblk + (size_t)j * c.rows_pad)), "l"(src), "r"(row_bytes), "r"(smem_u32(mbar)) : "memory");
            }
        }
        // wait for phase 0
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(mbar)), "r"(0) : "memory");
        }
        // pad columns beyond this rank's rows (only when my_rw < rw_max) are never read
        c.Tsrc = tblk; c.tstride = c.rows_pad;
    } else {
        c.Tsrc = td.T + c.r0; c.tstride = td.pitch;
    }
    __syncthreads();
    csync(c);

    int cur_image = 1, hand = 1;
    double e_saved = 0.0;
    double e4[4] = {0, 0, 0, 0};
    for (int k = 0; k < nstages; k++) {
        const DevStage st = prog[k];
        switch (st.op) {
        case CHR3D_OP_INIT: {
            csync(c);
            if (c.owner && c.row_ok) {
                c.x[0] = td.x0[3 * c.row] * (float)st.flag; c.x[1] = td.x0[3 * c.row + 1]; c.x[2] = td.x0[3 * c.row + 2];
                c.v[0] = c.v[1] = c.v[2] = 0.f;
                publish_x(c);
            }
            cur_image = st.flag; hand = cur_image;
            break; }
        case CHR3D_OP_MINIMIZE: op_minimize(c, st); break;
        case CHR3D_OP_MAXWELL: {
            if (c.owner && c.row_ok) {
                double z[3]; normal3(td.seed, (uint32_t)c.row, STREAM_MAXWELL + (uint32_t)k, (uint32_t)td.model, z);
                double sd = sqrt((double)prm.kboltz * (double)st.temp / (double)prm.mass);
                for (int q = 0; q < 3; q++) c.v[q] = (float)(sd * z[q]);
            }
            break; }
        case CHR3D_OP_DYNAMICS: op_dynamics(c, st); break;
        case CHR3D_OP_SAVE: case CHR3D_OP_PICK: case CHR3D_OP_ENERGY: {
            eval(c, st, st.op != CHR3D_OP_ENERGY);
            if (c.owner) c.P[c.tid] = c.ebond;
            allreduce(c, 1);
            double e = c.gsum[0] + c.gsum[1] + c.gsum[2];
            if (st.op == CHR3D_OP_SAVE) {
                e_saved = e;
                if (c.owner && c.row_ok) for (int q = 0; q < 3; q++) { td.scratch[3 * c.row + q] = c.x[q]; td.scratch[3 * (c.N + c.row) + q] = c.v[q]; }
            } else if (st.op == CHR3D_OP_PICK) {
                if (e_saved > e) hand = cur_image;                 // C3D:1717-1721
                else {
                    hand = 1;
                    if (c.owner && c.row_ok) {
                        for (int q = 0; q < 3; q++) { c.x[q] = td.scratch[3 * c.row + q]; c.v[q] = td.scratch[3 * (c.N + c.row) + q]; }
                        publish_x(c);
                    }
                }
            } else {
                e4[CHR3D_E_OVERALL] = e; e4[CHR3D_E_BON] = c.gsum[2]; e4[CHR3D_E_VDW] = c.gsum[1]; e4[CHR3D_E_NOE] = c.gsum[0];
                if (td.f_out && c.owner && c.row_ok) for (int q = 0; q < 3; q++) td.f_out[3 * c.row + q] = c.f[q];
            }
            break; }
        case CHR3D_OP_CENTER: {
            if (c.owner) for (int q = 0; q < 3; q++) c.P[(size_t)q * c.rows_pad + c.tid] = c.row_ok ? c.x[q] : 0.f;
            if (c.lane == 0) { c.ewarp[c.warp * 2] = 0.0; c.ewarp[c.warp * 2 + 1] = 0.0; }
            allreduce(c, 3);
            if (c.owner && c.row_ok) { for (int q = 0; q < 3; q++) c.x[q] -= (float)(c.gsum[2 + q] / c.N); publish_x(c); }
            break; }
        default: break;
        }
    }
    if (c.owner && c.row_ok) for (int q = 0; q < 3; q++) td.x_out[3 * c.row + q] = c.x[q];
    if (c.rank == 0 && c.tid == 0) {
        for (int q = 0; q < 4; q++) td.e_out[q] = e4[q];
        *td.nevals_out = c.nevals; *td.hand_out = hand;
    }
    csync(c);   // no CTA may exit while peers can still write into its shared memory
}

// ------------------------------------------------------------------------------------------------
// host side: configuration choice and launch
// ------------------------------------------------------------------------------------------------
static size_t smem_for(int N, int C, int threads, int m, int stage) {
    SaShape sh; sa_shape(N, C, threads, m, stage, sh);
    SaLayout L; sa_layout(sh, L);
    return L.total;
}

int sa_choose_config(int N, int n_traj, int m, SaLaunchCfg* cfg) {
    int dev = 0, sms = 148, smem_max = 227 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int rw_total = (N + 31) / 32;
    double best = 1e300; SaLaunchCfg bc = {1, 128, 0, 0};
    for (int C = 1; C <= SA_MAXC; C++) {
        if (C > rw_total) break;
        int rw_max = (rw_total + C - 1) / C;
        // threads: enough chunks per row-warp to give every SM sub-partition several warps
        int Q = 1; while (rw_max * (Q + 1) * 32 <= 512 && (Q + 1) <= (N + 31) / 32 + 1) Q++;
        int threads = rw_max * Q * 32; if (threads < 64) threads = 64; if (threads > 512) continue;
        int stage = 1; size_t sm = smem_for(N, C, threads, m, 1);
        if (sm > (size_t)smem_max) { stage = 0; sm = smem_for(N, C, threads, m, 0); }
        if (sm > (size_t)smem_max) continue;
        // co-resident clusters (approximation: one CTA per SM for the big shapes, more for small ones)
        int cta_per_sm = (int)((size_t)smem_max / (sm + 1024)); if (cta_per_sm < 1) cta_per_sm = 1;
        int by_thr = 2048 / threads; if (cta_per_sm > by_thr) cta_per_sm = by_thr; if (cta_per_sm < 1) cta_per_sm = 1;
        int resident = (sms * cta_per_sm) / C; if (C > 1) resident = (int)(resident * 0.9); if (resident < 1) resident = 1;
        int waves = (n_traj + resident - 1) / resident;
        // cost model per force evaluation: pair work per thread + fixed barrier/reduction overhead
        double pair_cycles = (double)N / Q * 30.0 * ((double)(rw_max * Q) / 4.0 > 1.0 ? (double)(rw_max * Q) / 4.0 : 1.0);
        double over = 600.0 + (C > 1 ? 1200.0 : 0.0);
        double cost = waves * (pair_cycles + over);
        if (cost < best) { best = cost; bc.cluster = C; bc.threads = threads; bc.stage_targets = stage; bc.smem = sm; }
    }
    if (best >= 1e300) { set_error("no SA kernel configuration fits N=%d", N); return CHR3D_ERR_SIZE; }
    *cfg = bc;
    return 0;
}

int sa_launch(const TrajDesc* d_traj, int n_traj, int Nmax, const DevParams& P, const DevStage* d_prog, int nstages,
              const SaLaunchCfg& cfg, cudaStream_t s) {
    (void)Nmax;
    CHR3D_CUDA(cudaFuncSetAttribute(sa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(n_traj * cfg.cluster), 1, 1);
    lc.blockDim = dim3((unsigned)cfg.threads, 1, 1);
    lc.dynamicSmemBytes = cfg.smem;
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cfg.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int st = cfg.stage_targets;
    CHR3D_CUDA(cudaLaunchKernelEx(&lc, sa_kernel, d_traj, P, d_prog, nstages, st));
    return 0;
}

}  // namespace chr3d
