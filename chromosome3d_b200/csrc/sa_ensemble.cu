// sa_ensemble.cu — the small-N hot path: one thread-block cluster per SA trajectory.
//
// Replaces the CNS `dgsa.inp` run (C3D:254-289 launches it; protocol text C3D:1573-1829) for
// chromosomes up to ~1500 beads.  One launch anneals every model of every (matrix, alpha) job:
//
//   * a trajectory = one cluster of C CTAs of W warps (C, W chosen on the host from N, the number
//     of trajectories and the cluster occupancy of the device);
//   * the unordered pair matrix is cut into 32x32 micro-tiles (I <= J); tile t belongs to CTA
//     t mod C and there to warp (t / C) mod W, so CTAs and their schedulers carry equal shares.  A lane owns a 4-row x 8-column sub-block (32 pairs in registers),
//     every pair is evaluated ONCE and its force goes to both beads; row/column partial sums are
//     combined inside the warp by recursive-halving shuffles and stored (DSMEM) into the partial
//     buffer of the CTA that owns the bead block; the owner adds its nb+1 partials in fixed order.
//     No atomics anywhere: runs are bit-reproducible;
//   * target distances are tile-packed "lane-major" (each lane's 32 values = 8 float4 strided by
//     the warp), resident in shared memory when they fit, otherwise streamed coalesced from L2;
//   * coordinates of all beads live in the shared memory of every CTA; bead blocks are owned by
//     CTAs (contiguous ranges) whose threads integrate them and publish new coordinates to all
//     CTAs through distributed shared memory;
//   * NOE soft-square + repel are fused in the pair term, the bond term is added by the bead owner;
//     energies and all other scalars are reduced in fp64 in a fixed order;
//   * the whole stage program (pre-minimisation, hot MD, cooling, L-BFGS, C3D:1573-1829) runs
//     inside the kernel.  There is no cluster barrier on the per-evaluation path: partial forces, coordinates (published as
//     a = x + dt*h*f, b = dt*v, from which every CTA rebuilds x' = a + lambda*b itself once the kinetic-energy partials are in)
//     and the scalars of an all-reduce travel as st.async stores that credit an mbarrier of the receiving CTA, which waits only
//     for ITS inputs;
//   * the pair term runs on the packed fp32 pipe (FADD2 / FMUL2 / FFMA2): two columns per instruction;
//   * the host may give a trajectory a stage range and a saved state to resume from, so that the two hot-MD images of a model
//     run side by side (api.cu).
//
// L-BFGS is the vector-free formulation: all inner products of an iteration are reduced in one
// exchange and the two-loop recursion runs on the (2m+1)x(2m+1) Gram matrix.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>
#include <algorithm>
#include "common.h"

namespace cg = cooperative_groups;

namespace chr3d {

#define SA_MAXC 16
#define SA_KTOT 64           // scalars per all-reduce (owner columns)
#define SA_MAXW 20           // warps per CTA (kernel variants for <= 16 and <= 20 warps)
#define SA_TILE_FLOATS 1024  // 32x32 micro-tile
#define SA_GEO_BYTES 384     // SaGeo (first 128 bytes) and the trajectory descriptor (next 256) at the start of shared memory
#define SA_MAX_STAGES 160    // stage-program entries kept in shared memory (the dgsa.inp program has 113)
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

// seeded random-coil start (stand-in for the DG embedding): unit steps in parallel (the Philox / Box-Muller part),
// then one thread per model adds them up in chain order and centres — the same additions in the same order as the
// oracle's orc_random_coil, so the start coordinates are bit-identical to it
__global__ void coil_steps_kernel(double* __restrict__ d, int N, int n_models, int first_model, double step, unsigned long long seed) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)N * n_models) return;
    const int m = (int)(k / N), i = (int)(k % N);
    double z[3] = {0, 0, 0};
    if (i > 0) {
        normal3(seed, (uint32_t)i, STREAM_COIL, (uint32_t)(first_model + m), z);
        double nrm = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
        if (nrm < 1e-12) { z[0] = 1; z[1] = z[2] = 0; nrm = 1; }
        for (int q = 0; q < 3; q++) z[q] = step * z[q] / nrm;
    }
    for (int q = 0; q < 3; q++) d[3 * k + q] = z[q];
}
// prefix sums of the steps and the centring: sequential fp64 sums in index order (the oracle's random_coil), one WARP per
// (model, component) — the three components are independent chains.  The lanes stage the chain's steps through shared memory
// (a dependent global load per term is what made the one-thread version take 0.56 ms for 455 beads); lane 0 adds them in
// order, leaves every prefix p_i in place of its step and accumulates the sum of the prefixes in the same order as before;
// then all lanes write x_i = (float)(p_i - mean).  The same additions in the same order: bit-identical to the oracle.
#define COIL_CHUNK 1024
__global__ void __launch_bounds__(32) coil_kernel(float* x0, double* __restrict__ d, int N, int n_models) {
    __shared__ double buf[COIL_CHUNK];
    const int t = blockIdx.x, lane = threadIdx.x;
    const int m = t / 3, k = t - 3 * m;
    if (m >= n_models) return;
    float* x = x0 + (size_t)m * N * 3;
    double* dm = d + (size_t)m * N * 3;
    double p = 0, c = 0;
    for (int i0 = 0; i0 < N; i0 += COIL_CHUNK) {
        const int n = N - i0 < COIL_CHUNK ? N - i0 : COIL_CHUNK;
        for (int i = lane; i < n; i += 32) buf[i] = dm[3 * (size_t)(i0 + i) + k];
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < n; i++) {
                if (i0 + i > 0) p += buf[i];
                c += p;
                buf[i] = p;
            }
        }
        __syncwarp();
        for (int i = lane; i < n; i += 32) dm[3 * (size_t)(i0 + i) + k] = buf[i];
        __syncwarp();
    }
    c = __shfl_sync(0xffffffffu, c, 0);
    const double mean = c / N;
    for (int i = lane; i < N; i += 32) x[3 * (size_t)i + k] = (float)(dm[3 * (size_t)i + k] - mean);
}
int dev_random_coil(float* d_x0, int N, int n_models, int first_model, double step, unsigned long long seed,
                    cudaStream_t s) {
    DevBuf<double> d;
    CHR3D_CUDA(d.alloc((size_t)N * n_models * 3));
    const long long n = (long long)N * n_models;
    coil_steps_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d.p, N, n_models, first_model, step, seed);
    coil_kernel<<<3 * n_models, 32, 0, s>>>(d_x0, d.p, N, n_models);
    CHR3D_CUDA(cudaGetLastError());
    if (s != (cudaStream_t)0) CHR3D_CUDA(cudaStreamSynchronize(s));     // the scratch buffer is released on the default stream
    return 0;
}


// ------------------------------------------------------------------------------------------------
// tile-packed, lane-major targets.  Tile t = (I, J), I <= J, row-major over the upper triangle of
// the nb x nb block grid.  Inside a tile lane l = 4*rg + cg owns rows rg + 8r (r = 0..3) and columns
// cg + 4q (q = 0..7) — interleaved, so that the coordinate loads of a warp touch consecutive shared
// memory words (no bank conflicts); its value (r, q) sits in float4 number k = 2r + q/4 (component
// q%4) at Tt[t*1024 + (k*32 + l)*4 ..].  Values: >0 restraint distance, -1 none, -2 none and no repel
// (also: i >= j of diagonal tiles and padding).
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int sa_tile_index(int I, int J, int nb) { return I * nb - I * (I - 1) / 2 + (J - I); }
__host__ __device__ inline void sa_tile_coords(int t, int nb, int& I, int& J) {
    int i = 0;
    while (sa_tile_index(i + 1, i + 1, nb) <= t) i++;
    I = i; J = i + (t - sa_tile_index(i, i, nb));
}
__global__ void __launch_bounds__(256) sa_pack_tiles_kernel(const float* __restrict__ T, int N, int pitch, int nb, float* __restrict__ Tt) {
    const int t = blockIdx.x;
    int I, J; sa_tile_coords(t, nb, I, J);
    for (int e = threadIdx.x; e < SA_TILE_FLOATS; e += blockDim.x) {
        const int k = e >> 7, lane = (e >> 2) & 31, comp = e & 3;
        const int rg = lane >> 2, cg = lane & 3, r = k >> 1, c = (k & 1) * 4 + comp;
        const int i = 32 * I + rg + 8 * r, j = 32 * J + cg + 4 * c;
        Tt[(size_t)t * SA_TILE_FLOATS + e] = (i < N && j < N && i < j) ? T[(size_t)i * pitch + j] : -2.f;
    }
}
int sa_pack_tiles(const float* d_T, int N, int pitch, float* d_Tt, cudaStream_t s) {
    const int nb = (N + 31) / 32;
    sa_pack_tiles_kernel<<<nb * (nb + 1) / 2, 256, 0, s>>>(d_T, N, pitch, nb, d_Tt);
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}
size_t sa_tiles_floats(int N) { const size_t nb = (size_t)(N + 31) / 32; return nb * (nb + 1) / 2 * SA_TILE_FLOATS; }

// ------------------------------------------------------------------------------------------------
// shared-memory layout (identical on host and device)
// ------------------------------------------------------------------------------------------------
struct SaShape {
    int N, C, W, nthreads, nb, n32, ntiles, gwarps, rounds, own_max, rows_pad, m, nvec, res_rounds, nstages;
};
struct SaLayout {
    size_t prog, mbar, ab, part, vec, P, eslots, kslots, slots, gsum, B, delta, tij, own, tres, total;
};
__host__ __device__ inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__host__ __device__ inline void sa_layout(const SaShape& s, SaLayout& L) {
    size_t o = SA_GEO_BYTES;
    L.prog = o;   o = align_up(o + sizeof(DevStage) * (size_t)s.nstages, 16);
    L.mbar = o;   o = align_up(o + 32, 16);                            // four mbarriers: forces arrived, coordinates arrived, all-reduce slots (two parities)
    L.ab = o;     o = align_up(o + sizeof(float4) * 4 * (size_t)(s.n32 + 1), 16);   // [parity][a|b][n32+1]
    L.part = o;   o = align_up(o + sizeof(float4) * (size_t)s.own_max * (s.nb + 1) * 32, 16);
    L.vec = o;    o = align_up(o + sizeof(float) * (size_t)s.nvec * 3 * s.rows_pad, 16);
    L.P = o;      o = align_up(o + sizeof(float) * (size_t)(7 + 6 * s.m) * s.rows_pad, 16);   // columns: bond energy + 6m+6 inner products
    L.eslots = o; o = align_up(o + sizeof(double) * 2 * (size_t)s.gwarps, 16);
    L.kslots = o; o = align_up(o + 2 * sizeof(double) * (size_t)s.nb, 16);              // one 16-byte slot per bead block (st.async granularity)
    L.slots = o;  o = align_up(o + sizeof(double) * 2 * SA_MAXC * SA_KTOT, 16);
    L.gsum = o;   o = align_up(o + sizeof(double) * (SA_KTOT + 2), 16);
    L.B = o;      o = align_up(o + sizeof(double) * SA_MAXB * SA_MAXB, 16);
    L.delta = o;  o = align_up(o + sizeof(double) * (SA_MAXB + 3), 16);
    L.tij = o;    o = align_up(o + 2 * (size_t)s.rounds * s.W, 16);
    L.own = o;    o = align_up(o + 2 * (size_t)s.nb, 128);
    L.tres = o;
    o += sizeof(float) * SA_TILE_FLOATS * (size_t)s.res_rounds * s.W;
    L.total = o;
}
// res_rounds: how many of a warp's tiles stay resident in shared memory (decided from the dynamic smem size)
__host__ __device__ inline void sa_shape(int N, int C, int W, int m, int nstages, size_t smem_bytes, SaShape& s) {
    s.N = N; s.C = C; s.W = W; s.nthreads = 32 * W; s.m = m; s.nstages = nstages;
    s.nb = (N + 31) / 32; s.n32 = 32 * s.nb; s.ntiles = s.nb * (s.nb + 1) / 2; s.gwarps = C * W;
    s.rounds = ((s.ntiles + C - 1) / C + W - 1) / W;       // tiles of the busiest warp
    s.own_max = (s.nb + C - 1) / C; s.rows_pad = 32 * s.own_max;
    s.nvec = 3 + 2 * m;     // XB, G, PV, S[m], Y[m]
    s.res_rounds = 0;
    SaLayout L; sa_layout(s, L);
    if (smem_bytes > L.total) {
        size_t r = (smem_bytes - L.total) / (sizeof(float) * SA_TILE_FLOATS * (size_t)W);
        s.res_rounds = (int)(r < (size_t)s.rounds ? r : (size_t)s.rounds);
    }
}

extern __shared__ __align__(128) unsigned char sa_smem[];

// Cluster-uniform geometry, layout offsets and the stage program live at the start of shared memory (read with
// broadcast LDS when needed) instead of in per-thread registers: the pair loop needs the registers, and spilled
// per-thread copies cost local-memory round trips on the latency-critical control path.
struct SaGeo {
    int N, C, W, nthreads, nb, n32p, gwarps, own_max, rows_pad, m, res_rounds, rw0, nown, nstages;
    unsigned o_ab, o_kslots, o_part, o_vec, o_P, o_eslots, o_slots, o_gsum, o_B, o_delta, o_tij, o_own, o_tres, o_prog, o_mbar;
};
#define SGEO (*(const SaGeo*)sa_smem)
#define STD (*(const TrajDesc*)(sa_smem + 128))
static_assert(sizeof(SaGeo) <= 128 && sizeof(TrajDesc) <= SA_GEO_BYTES - 128, "SaGeo / TrajDesc outgrew their shared-memory slots");

// Development aid (-DSA_TIMING): SM-clock cycles spent in the phases of a force evaluation, summed over all
// evaluations, for the first owner warp and the last warp of cluster rank 0 of trajectory 0.
#ifdef SA_TIMING
__device__ unsigned long long sa_dbg[2][16];
#define TSTAMP(c, k) do { if ((c).tmark) { long long t_ = clock64(); (c).tacc[k] += t_ - (c).tlast; (c).tlast = t_; } } while (0)
#else
#define TSTAMP(c, k) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
// device context
// ------------------------------------------------------------------------------------------------
struct Ctx {
    // per-thread geometry (everything cluster-uniform is in SGEO)
    int rank, tid, lane, warp, gw, my_rounds;
    const float4* Tt;                // tile-packed targets in global memory
    int par;
    // coordinates of every bead are x = a + lam*b, (a, b) double-buffered by parity `buf` (see dyn_after_eval)
    int buf; float lam;
    unsigned phF, phX, phR;          // parities of the mbarriers (phR: bit p = parity of MB_R + p)
    // owner state (valid when tid < rows_pad; row_ok when the thread owns a real bead)
    bool owner, row_ok; int row;
    float x[3], v[3], f[3];
    float ebond;
    long long nevals;
#ifdef SA_TIMING
    bool tmark; long long tlast; long long tacc[10];
#endif
};
#define S_AB(c) ((float4*)(sa_smem + SGEO.o_ab))
#define S_KSLOTS(c) ((double*)(sa_smem + SGEO.o_kslots))
#define S_PART(c) ((float4*)(sa_smem + SGEO.o_part))
#define S_VEC(c) ((float*)(sa_smem + SGEO.o_vec))
#define S_P(c) ((float*)(sa_smem + SGEO.o_P))
#define S_ESLOTS(c) ((double*)(sa_smem + SGEO.o_eslots))
#define S_SLOTS(c) ((double*)(sa_smem + SGEO.o_slots))
#define S_GSUM(c) ((double*)(sa_smem + SGEO.o_gsum))
#define S_B(c) ((double*)(sa_smem + SGEO.o_B))
#define S_DELTA(c) ((double*)(sa_smem + SGEO.o_delta))
#define S_TIJ(c) ((uchar2*)(sa_smem + SGEO.o_tij))
#define S_OWN(c) ((uchar2*)(sa_smem + SGEO.o_own))
#define S_TRES(c) ((float4*)(sa_smem + SGEO.o_tres))
#define S_PROG(c) ((const DevStage*)(sa_smem + SGEO.o_prog))

__device__ __forceinline__ float* vecp(const Ctx& c, int v, int comp) { return S_VEC(c) + ((size_t)(v * 3 + comp)) * SGEO.rows_pad; }
#define V_XB 0
#define V_G 1
#define V_PV 2
#define V_S(l) (3 + (l))
#define V_Y(l) (3 + SGEO.m + (l))

__device__ __forceinline__ void csync(const Ctx& c) {
    if (SGEO.C > 1) cg::this_cluster().sync(); else __syncthreads();
}
__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Coordinates.  Every CTA holds (a, b) of all beads; x = a + lam*b with the cluster-uniform scalar lam.
// Publishing goes into the buffer of the other parity (readers of the current one may still be busy);
// once the coordinate phase has completed (await_coords), flip() makes it current.
__device__ __forceinline__ float4 ldx(const Ctx& c, int i) {
    const float4* A = S_AB(c) + (size_t)(2 * c.buf) * SGEO.n32p;
    const float4 a = A[i], b = A[SGEO.n32p + i];
    return make_float4(fmaf(c.lam, b.x, a.x), fmaf(c.lam, b.y, a.y), fmaf(c.lam, b.z, a.z), 0.f);
}
// ---- asynchronous cluster exchange ----------------------------------------------------------------------------
// Producers push 16-byte pieces into the shared memory of the consumer CTA with st.async, which also credits the
// bytes to an mbarrier of the consumer; the consumer declares how many bytes a phase brings (arrive.expect_tx by
// one thread) and waits for the phase.  No cluster-wide barrier and no producer-side fence: a CTA proceeds as soon
// as ITS inputs have landed.  Two barriers per CTA: MB_F (partial forces + pair energies), MB_X (coordinates +
// kinetic-energy partials).  Phases cannot mix: a producer can only start the next phase after it has consumed
// data that the consumer sent after completing the current one (see the header comment).
#define MB_F 0
#define MB_X 1
#define MB_R 2          // MB_R + parity: scalars of an all-reduce have arrived
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, int rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_async16(uint32_t dst_cluster, uint32_t mbar_cluster, float4 v) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(dst_cluster), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)),
                   "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ uint32_t mbar_addr(const Ctx& c, int which) { return smem_u32(sa_smem + SGEO.o_mbar + 8 * which); }
__device__ __forceinline__ void mbar_expect(const Ctx& c, int which, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr(c, which)), "r"(bytes) : "memory");
}
#ifdef SA_HANGDIAG
// Development aid (-DSA_HANGDIAG, scripts/hangdiag.sh): instead of trapping, the first wait that times out raises a global abort
// flag; every CTA then records where it stands (rank, evaluation count, the barrier it waits on, the raw words of its four
// mbarriers) and all waits return, so the kernel ends and the host can read the picture.
__device__ int sa_abort;
__device__ unsigned long long (*sa_diag)[8];       // 512 rows in pinned, mapped host memory: readable after the context has died
#endif
__device__ __forceinline__ void mbar_wait(const Ctx& c, int which, unsigned& parity) {
    const uint32_t bar = mbar_addr(c, which);
    uint32_t done = 0;
    const long long t0 = clock64();
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#ifdef SA_HANGDIAG
        if (!done) {
            if (clock64() - t0 > 1000000000ll) atomicExch(&sa_abort, 1);          // 0.5 s
            if (*(volatile int*)&sa_abort) {
                if (c.lane == 0 && blockIdx.x < 512 && atomicAdd((unsigned long long*)&sa_diag[blockIdx.x][7], 1ull) == 0ull) {
                    const unsigned long long* mb = (const unsigned long long*)(sa_smem + SGEO.o_mbar);
                    sa_diag[blockIdx.x][0] = ((unsigned long long)c.rank << 48) | ((unsigned long long)c.warp << 40) | ((unsigned long long)which << 32) | (unsigned long long)parity;
                    sa_diag[blockIdx.x][1] = (unsigned long long)c.nevals;
                    sa_diag[blockIdx.x][2] = ((unsigned long long)c.phF << 32) | ((unsigned long long)c.phX << 16) | (unsigned long long)c.phR;
                    for (int q = 0; q < 4; q++) sa_diag[blockIdx.x][3 + q] = ((const volatile unsigned long long*)mb)[q];
                    __threadfence_system();
                }
                const long long t1 = clock64();
                while (clock64() - t1 < 400000000ll) { }          // 0.2 s for everybody to record, then stop the kernel for good
                __trap();
            }
        }
#else
        if (!done && clock64() - t0 > 20000000000ll) __trap();        // ~10 s: a peer died; fail instead of hanging the GPU
#endif
    }
    parity ^= 1u;
}

// (The wait is the default acquire at CTA scope, as for TMA-fed pipelines: st.async data and its complete_tx reach the
// consumer's shared memory through the same path, and everything consumed after a wait lives in THIS CTA's shared memory.
// A cluster-scope acquire would make ptxas append CCTL.IVALL — an L1 invalidation — to every wait: measured 20 % of all
// stall samples of the kernel.)
__device__ __forceinline__ void st_async8(uint32_t dst_cluster, uint32_t mbar_cluster, double v) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(dst_cluster), "l"(__double_as_longlong(v)), "r"(mbar_cluster) : "memory");
}

// Publish (a, b) of the owned bead to every CTA of the cluster (buffer of the other parity).
__device__ __forceinline__ void publish_ab(const Ctx& c, float4 a, float4 b, bool with_b) {
    if (!c.row_ok) return;
    const uint32_t A = smem_u32(S_AB(c) + (size_t)(2 * (c.buf ^ 1)) * SGEO.n32p + c.row);
    const uint32_t bar = mbar_addr(c, MB_X);
    for (int r = 0; r < SGEO.C; r++) {
        const uint32_t rb = mapa_u32(bar, r), ra = mapa_u32(A, r);
        st_async16(ra, rb, a);
        if (with_b) st_async16(ra + 16u * (uint32_t)SGEO.n32p, rb, b);
    }
}
// x of the owner -> a (lam becomes 0 at the flip, b is ignored)
__device__ __forceinline__ void publish_x(const Ctx& c) {
    publish_ab(c, make_float4(c.x[0], c.x[1], c.x[2], 0.f), make_float4(0.f, 0.f, 0.f, 0.f), false);
}
__device__ __forceinline__ void flip(Ctx& c, float lam) { c.buf ^= 1; c.lam = lam; }
// every CTA waits until the coordinates of all N beads (16 or 32 bytes each, + `extra` bytes) have landed
__device__ __forceinline__ void await_coords(Ctx& c, bool with_b, uint32_t extra) {
    if (c.tid == 0) mbar_expect(c, MB_X, (uint32_t)SGEO.N * (with_b ? 32u : 16u) + extra);
    mbar_wait(c, MB_X, c.phX);
}
// publish_x + wait + flip: the usual way an op installs new coordinates
__device__ __forceinline__ void install_x(Ctx& c) { publish_x(c); await_coords(c, false, 0u); flip(c, 0.f); }

// All-reduce of K owner columns (P) across the cluster; with_e: also totals of the pair energies
// every warp of the cluster left in eslots (complete in every CTA after the barrier inside eval).
// Result: gsum[0]=E_noe, gsum[1]=E_vdw, gsum[2+k]=column k.  Every CTA sums the same partials in
// the same order, so all ranks take identical decisions afterwards.
__device__ void allreduce(Ctx& c, int K, bool with_e) {
    __syncthreads();
    const uint32_t bar = mbar_addr(c, MB_R + c.par);
    for (int k = c.warp; k < K; k += SGEO.W) {
        const float* col = S_P(c) + (size_t)k * SGEO.rows_pad;
        double s = 0;
        for (int t = c.lane; t < SGEO.rows_pad; t += 32) s += (double)col[t];
        s = warp_sum(s);
        if (SGEO.C > 1) {
            if (c.lane < SGEO.C)      // slot (parity, this rank, k) of CTA `lane`; the bytes are credited to its mbarrier of this parity
                st_async8(mapa_u32(smem_u32(S_SLOTS(c) + ((size_t)c.par * SA_MAXC + c.rank) * SA_KTOT + k), c.lane), mapa_u32(bar, c.lane), s);
        } else if (c.lane == 0) S_GSUM(c)[2 + k] = s;
    }
    if (with_e) {
        for (int e = 0; e < 2; e++) {
            if (c.warp != (SGEO.W - 1 - e) % SGEO.W) continue;
            double s = 0;
            for (int g = c.lane; g < SGEO.gwarps; g += 32) s += S_ESLOTS(c)[2 * g + e];
            s = warp_sum(s);
            if (c.lane == 0) S_GSUM(c)[e] = s;
        }
    }
    if (SGEO.C > 1) {
        // no cluster barrier: wait for the C*K scalars addressed to this CTA.  Consecutive all-reduces alternate between two
        // mbarriers (and two slot sets), so scalars of the next one can never complete the current phase; a rank cannot be two
        // all-reduces ahead, because the next one needs this rank's contribution.
        if (c.tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(8 * SGEO.C * K)) : "memory");
        {
            unsigned ph = (c.phR >> c.par) & 1u;
            mbar_wait(c, MB_R + c.par, ph);
            c.phR ^= 1u << c.par;
        }
        if (c.tid < K) {
            double s = 0;
            for (int r = 0; r < SGEO.C; r++) s += S_SLOTS(c)[((size_t)c.par * SA_MAXC + r) * SA_KTOT + c.tid];
            S_GSUM(c)[2 + c.tid] = s;
        }
        c.par ^= 1;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// pair term, symmetric: accumulates the force on i and on j and, if WANT_E, the energies.
// r^2 carries a 1e-12 floor (coincident beads), as the oracle's max(r^2, 1e-12).
// ------------------------------------------------------------------------------------------------
// r^-1/2 as one MUFU.RSQ (arguments here are >= 1e-12: no denormal handling needed)
__device__ __forceinline__ float rsqrt_fast(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct SaTerms { float wnC2m, rsw, wv, wv4, R2, wn, C, a, b, casym; int low_soft; };

// NEAR: the tile touches the diagonal band (J <= I + 1) or the padded last block column, the only places
// where cells carry the "no repel" sentinel -2.
template <bool WANT_E, bool GENERAL, bool NEAR>
__device__ __forceinline__ void pair_sym(const SaTerms& t, float xi, float yi, float zi, float xj, float yj, float zj, float d,
                                         float* fi, float* fj, float& en, float& ev) {
    const float dx = xi - xj, dy = yi - yj, dz = zi - zj;
    const float r2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, 1e-12f)));
    const float rinv = rsqrt_fast(r2);
    float tr = fmaxf(t.R2 - r2, 0.f);
    if (NEAR) tr = d > -1.5f ? tr : 0.f;
    float coef;
    if (!GENERAL) {
        const float delta = fmaf(r2, rinv, -d);                    // r - d
        const float h = fminf(delta, t.rsw);
        const float w2h = (d > 0.f ? t.wnC2m : 0.f) * h;           // -2 w C h
        if (WANT_E) en = fmaf(-0.5f * w2h, 2.f * delta - h, en);   // w C h (2 delta - h)
        coef = fmaf(t.wv4, tr, w2h * rinv);                        // 4 w_vdw t - (dE_noe/dr)/r
    } else {
        const float r = r2 * rinv, delta = r - d, u = fabsf(delta), sg = delta > 0.f ? 1.f : -1.f;
        const bool soft = u > t.rsw && (delta > 0.f || t.low_soft);
        const float uinv = 1.f / fmaxf(u, 1e-20f);
        const float e = soft ? t.a + t.b * uinv + t.casym * u : t.C * u * u;
        const float de = soft ? (t.casym - t.b * uinv * uinv) * sg : 2.f * t.C * u * sg;
        const float w = d > 0.f ? t.wn : 0.f;
        if (WANT_E) en = fmaf(w, e, en);
        coef = fmaf(t.wv4, tr, -w * de * rinv);
    }
    if (WANT_E) ev = fmaf(t.wv * tr, tr, ev);
    fi[0] = fmaf(coef, dx, fi[0]); fi[1] = fmaf(coef, dy, fi[1]); fi[2] = fmaf(coef, dz, fi[2]);
    fj[0] = fmaf(-coef, dx, fj[0]); fj[1] = fmaf(-coef, dy, fj[1]); fj[2] = fmaf(-coef, dz, fj[2]);
}

// The same pair term on Blackwell's packed fp32 pipe (FADD2 / FMUL2 / FFMA2, PTX *.f32x2): one call evaluates the pairs
// (row i, column j) and (row i, column j+4) — the two columns a lane holds side by side in a float4 of targets — in the
// two halves of 64-bit register pairs; the row coordinates ride as broadcast operands.  Each half computes exactly the
// scalar form's operations in the same order.  gj accumulates +coef*dx: the force on j is its negative.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
template <bool WANT_E, bool NEAR>
__device__ __forceinline__ void pair_sym2(const SaTerms& t, float xi, float yi, float zi, const float2 xj, const float2 yj, const float2 zj, const float2 d,
                                          float2* fi, float2* gj, float2& en, float2& ev) {
    const float2 dx = __fadd2_rn(bc(xi), f2(-xj.x, -xj.y)), dy = __fadd2_rn(bc(yi), f2(-yj.x, -yj.y)), dz = __fadd2_rn(bc(zi), f2(-zj.x, -zj.y));
    const float2 r2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __ffma2_rn(dz, dz, bc(1e-12f))));
    const float2 rinv = f2(rsqrt_fast(r2.x), rsqrt_fast(r2.y));
    float2 tr = __fadd2_rn(bc(t.R2), f2(-r2.x, -r2.y));
    tr = f2(fmaxf(tr.x, 0.f), fmaxf(tr.y, 0.f));
    if (NEAR) tr = f2(d.x > -1.5f ? tr.x : 0.f, d.y > -1.5f ? tr.y : 0.f);
    const float2 delta = __ffma2_rn(r2, rinv, f2(-d.x, -d.y));                         // r - d
    const float2 h = f2(fminf(delta.x, t.rsw), fminf(delta.y, t.rsw));
    const float2 has = f2(d.x > 0.f ? 1.f : 0.f, d.y > 0.f ? 1.f : 0.f);
    const float2 w2h = __fmul2_rn(__fmul2_rn(has, bc(t.wnC2m)), h);                    // -2 w C h
    if (WANT_E) {
        en = __ffma2_rn(__fmul2_rn(w2h, bc(-0.5f)), __ffma2_rn(delta, bc(2.f), f2(-h.x, -h.y)), en);   // w C h (2 delta - h)
        ev = __ffma2_rn(__fmul2_rn(tr, bc(t.wv)), tr, ev);
    }
    const float2 coef = __ffma2_rn(bc(t.wv4), tr, __fmul2_rn(w2h, rinv));             // 4 w_vdw t - (dE_noe/dr)/r
    fi[0] = __ffma2_rn(coef, dx, fi[0]); fi[1] = __ffma2_rn(coef, dy, fi[1]); fi[2] = __ffma2_rn(coef, dz, fi[2]);
    gj[0] = __ffma2_rn(coef, dx, gj[0]); gj[1] = __ffma2_rn(coef, dy, gj[1]); gj[2] = __ffma2_rn(coef, dz, gj[2]);
}

// ------------------------------------------------------------------------------------------------
// energy + forces.  On return: c.f (owner), c.ebond (owner partial); if WANT_E the pair-energy
// partials of every warp of the cluster are in eslots (pick them up with allreduce(.., true)).
// ------------------------------------------------------------------------------------------------
template <bool WANT_E, bool GENERAL>
__device__ __forceinline__ void eval_t(Ctx& c, const DevParams& prm, const DevStage& st, bool count) {
    const DevParams& p = prm;                       // precondition: coordinates installed (publish, await_coords, flip)
    SaTerms tm;
    tm.wnC2m = -2.f * st.w_noe * p.C; tm.rsw = p.rsw; tm.wv = st.w_vdw; tm.wv4 = 4.f * st.w_vdw; tm.R2 = st.R2;
    tm.wn = st.w_noe; tm.C = p.C; tm.a = p.a; tm.b = p.b; tm.casym = p.casym; tm.low_soft = p.lo_clamp > -1e30f ? 1 : 0;
    const int lane = c.lane, rg = lane >> 2, cg4 = lane & 3;
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8, b4 = lane & 16;
    float en = 0.f, ev = 0.f;
    TSTAMP(c, 0);                                   // everything since the last stamp: control, reductions
    for (int k = 0; k < c.my_rounds; k++) {
        const uchar2 ij = S_TIJ(c)[k * SGEO.W + c.warp];
        const int I = ij.x, J = ij.y;
        float4 d[8];
        if (k < SGEO.res_rounds) {
            const float4* tp = S_TRES(c) + ((size_t)(k * SGEO.W + c.warp) * 256 + lane);
#pragma unroll
            for (int q = 0; q < 8; q++) d[q] = tp[q * 32];
        } else {
            const float4* gp = c.Tt + ((size_t)(c.rank + SGEO.C * (c.warp + SGEO.W * k)) * 256 + lane);
#pragma unroll
            for (int q = 0; q < 8; q++) d[q] = __ldg(gp + q * 32);
        }
        float fi[4][3], fj[8][3];
        if constexpr (!GENERAL) {
            // packed pipe: the lane's columns as pairs (cg + 8p, cg + 8p + 4), p = 0..3 — adjacent components of its float4 targets
            float2 xj2[4], yj2[4], zj2[4], gj2[4][3], fi2[4][3];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 t0 = ldx(c, 32 * J + cg4 + 8 * q), t1 = ldx(c, 32 * J + cg4 + 8 * q + 4);
                xj2[q] = f2(t0.x, t1.x); yj2[q] = f2(t0.y, t1.y); zj2[q] = f2(t0.z, t1.z);
                gj2[q][0] = gj2[q][1] = gj2[q][2] = bc(0.f);
                fi2[q][0] = fi2[q][1] = fi2[q][2] = bc(0.f);
            }
            float2 en2 = bc(0.f), ev2 = bc(0.f);
            if (J <= I + 1 || J == SGEO.nb - 1) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const float4 xi = ldx(c, 32 * I + rg + 8 * r);
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const float4 d4 = d[2 * r + hh];
                        pair_sym2<WANT_E, true>(tm, xi.x, xi.y, xi.z, xj2[2 * hh], yj2[2 * hh], zj2[2 * hh], f2(d4.x, d4.y), fi2[r], gj2[2 * hh], en2, ev2);
                        pair_sym2<WANT_E, true>(tm, xi.x, xi.y, xi.z, xj2[2 * hh + 1], yj2[2 * hh + 1], zj2[2 * hh + 1], f2(d4.z, d4.w), fi2[r], gj2[2 * hh + 1], en2, ev2);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const float4 xi = ldx(c, 32 * I + rg + 8 * r);
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const float4 d4 = d[2 * r + hh];
                        pair_sym2<WANT_E, false>(tm, xi.x, xi.y, xi.z, xj2[2 * hh], yj2[2 * hh], zj2[2 * hh], f2(d4.x, d4.y), fi2[r], gj2[2 * hh], en2, ev2);
                        pair_sym2<WANT_E, false>(tm, xi.x, xi.y, xi.z, xj2[2 * hh + 1], yj2[2 * hh + 1], zj2[2 * hh + 1], f2(d4.z, d4.w), fi2[r], gj2[2 * hh + 1], en2, ev2);
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < 3; a++) {
#pragma unroll
                for (int r = 0; r < 4; r++) fi[r][a] = fi2[r][a].x + fi2[r][a].y;
#pragma unroll
                for (int q = 0; q < 4; q++) { fj[2 * q][a] = -gj2[q][a].x; fj[2 * q + 1][a] = -gj2[q][a].y; }
            }
            if (WANT_E) { en += en2.x + en2.y; ev += ev2.x + ev2.y; }
        } else {
            float xj[8], yj[8], zj[8];
    #pragma unroll
            for (int q = 0; q < 8; q++) { const float4 t = ldx(c, 32 * J + cg4 + 4 * q); xj[q] = t.x; yj[q] = t.y; zj[q] = t.z; }
    #pragma unroll
            for (int q = 0; q < 4; q++) fi[q][0] = fi[q][1] = fi[q][2] = 0.f;
    #pragma unroll
            for (int q = 0; q < 8; q++) fj[q][0] = fj[q][1] = fj[q][2] = 0.f;
            if (J <= I + 1 || J == SGEO.nb - 1) {
    #pragma unroll
                for (int r = 0; r < 4; r++) {
                    const float4 xi = ldx(c, 32 * I + rg + 8 * r);
    #pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const float4 d4 = d[2 * r + hh];
                        pair_sym<WANT_E, GENERAL, true>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 0], yj[4 * hh + 0], zj[4 * hh + 0], d4.x, fi[r], fj[4 * hh + 0], en, ev);
                        pair_sym<WANT_E, GENERAL, true>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 1], yj[4 * hh + 1], zj[4 * hh + 1], d4.y, fi[r], fj[4 * hh + 1], en, ev);
                        pair_sym<WANT_E, GENERAL, true>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 2], yj[4 * hh + 2], zj[4 * hh + 2], d4.z, fi[r], fj[4 * hh + 2], en, ev);
                        pair_sym<WANT_E, GENERAL, true>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 3], yj[4 * hh + 3], zj[4 * hh + 3], d4.w, fi[r], fj[4 * hh + 3], en, ev);
                    }
                }
            } else {
    #pragma unroll
                for (int r = 0; r < 4; r++) {
                    const float4 xi = ldx(c, 32 * I + rg + 8 * r);
    #pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const float4 d4 = d[2 * r + hh];
                        pair_sym<WANT_E, GENERAL, false>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 0], yj[4 * hh + 0], zj[4 * hh + 0], d4.x, fi[r], fj[4 * hh + 0], en, ev);
                        pair_sym<WANT_E, GENERAL, false>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 1], yj[4 * hh + 1], zj[4 * hh + 1], d4.y, fi[r], fj[4 * hh + 1], en, ev);
                        pair_sym<WANT_E, GENERAL, false>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 2], yj[4 * hh + 2], zj[4 * hh + 2], d4.z, fi[r], fj[4 * hh + 2], en, ev);
                        pair_sym<WANT_E, GENERAL, false>(tm, xi.x, xi.y, xi.z, xj[4 * hh + 3], yj[4 * hh + 3], zj[4 * hh + 3], d4.w, fi[r], fj[4 * hh + 3], en, ev);
                    }
                }
            }
        }
        // rows: the 4 lanes that differ in bits 0-1 (column groups) share rows -> recursive halving
        float rowf[3], colf[3];
        {
            float k0[2][3];
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float lo = fi[q][a], hi = fi[q + 2][a];
                    k0[q][a] = (b0 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b0 ? lo : hi, 1);
                }
#pragma unroll
            for (int a = 0; a < 3; a++) rowf[a] = (b1 ? k0[1][a] : k0[0][a]) + __shfl_xor_sync(0xffffffffu, b1 ? k0[0][a] : k0[1][a], 2);
        }
        // columns: the 8 lanes that differ in bits 2-4 (row groups) share columns
        {
            float k1[4][3], k2[2][3];
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float lo = fj[q][a], hi = fj[q + 4][a];
                    k1[q][a] = (b2 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b2 ? lo : hi, 4);
                }
#pragma unroll
            for (int q = 0; q < 2; q++)
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float lo = k1[q][a], hi = k1[q + 2][a];
                    k2[q][a] = (b3 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b3 ? lo : hi, 8);
                }
#pragma unroll
            for (int a = 0; a < 3; a++) colf[a] = (b4 ? k2[1][a] : k2[0][a]) + __shfl_xor_sync(0xffffffffu, b4 ? k2[0][a] : k2[1][a], 16);
        }
        const int rsel = rg + 8 * ((b0 ? 2 : 0) + (b1 ? 1 : 0));        // row of this tile held by the lane
        const int csel = cg4 + 4 * ((b2 ? 4 : 0) + (b3 ? 2 : 0) + (b4 ? 1 : 0));
        const uchar2 oi = S_OWN(c)[I], oj = S_OWN(c)[J];
        const int slot_i = J, slot_j = (I == J) ? SGEO.nb : I;
        const uint32_t dst_i = smem_u32(S_PART(c) + ((size_t)(oi.y * (SGEO.nb + 1) + slot_i) * 32 + rsel));
        const uint32_t dst_j = smem_u32(S_PART(c) + ((size_t)(oj.y * (SGEO.nb + 1) + slot_j) * 32 + csel));
        const uint32_t barF = mbar_addr(c, MB_F);
        st_async16(mapa_u32(dst_i, oi.x), mapa_u32(barF, oi.x), make_float4(rowf[0], rowf[1], rowf[2], 0.f));
        st_async16(mapa_u32(dst_j, oj.x), mapa_u32(barF, oj.x), make_float4(colf[0], colf[1], colf[2], 0.f));
    }
    if (WANT_E) {       // pair-energy partial of this warp -> slot gw of every CTA (16 bytes: two doubles)
        const double den = warp_sum((double)en), dev = warp_sum((double)ev);
        if (lane < SGEO.C) {
            const uint32_t dst = smem_u32(S_ESLOTS(c) + 2 * c.gw);
            st_async16(mapa_u32(dst, lane), mapa_u32(mbar_addr(c, MB_F), lane),
                       make_float4(__int_as_float(__double2loint(den)), __int_as_float(__double2hiint(den)),
                                   __int_as_float(__double2loint(dev)), __int_as_float(__double2hiint(dev))));
        }
    }
    TSTAMP(c, 1);                                   // pair phase (own tiles)
    // this CTA's inputs of the force phase: nb+1 partial vectors (512 B) per owned block, and the energy partials of every warp
    if (c.tid == 0) mbar_expect(c, MB_F, (uint32_t)SGEO.nown * (uint32_t)(SGEO.nb + 1) * 512u + (WANT_E ? 16u * (uint32_t)SGEO.gwarps : 0u));
    mbar_wait(c, MB_F, c.phF);
    // Every thread of the CTA must have seen this phase complete before any of them can move on: a parity wait only tells the
    // current phase from the previous one.  Without this barrier a warp that is late to poll (its tiles need not feed this
    // CTA's own beads, so nothing here waits for it) can find the mbarrier a full evaluation further — same parity again — and
    // wait for a phase that needs its own, never-sent partial forces: a cluster-wide deadlock.  The same barrier covers MB_X
    // (nobody can post the next coordinate phase before passing here).
    __syncthreads();
    TSTAMP(c, 2);                                   // wait for the partial forces
    float gx = 0.f, gy = 0.f, gz = 0.f, eb = 0.f;
    if (c.row_ok) {
        const float4* pp = S_PART(c) + ((size_t)(c.tid >> 5) * (SGEO.nb + 1)) * 32 + lane;
        int s = 0;
        for (; s + 4 <= SGEO.nb + 1; s += 4) {         // fixed summation order, loads issued four at a time
            const float4 t0 = pp[s * 32], t1 = pp[(s + 1) * 32], t2 = pp[(s + 2) * 32], t3 = pp[(s + 3) * 32];
            gx += t0.x; gy += t0.y; gz += t0.z; gx += t1.x; gy += t1.y; gz += t1.z;
            gx += t2.x; gy += t2.y; gz += t2.z; gx += t3.x; gy += t3.y; gz += t3.z;
        }
        for (; s <= SGEO.nb; s++) { const float4 t = pp[s * 32]; gx += t.x; gy += t.y; gz += t.z; }
        const float4 xi = ldx(c, c.row);
        for (int s = -1; s <= 1; s += 2) {          // bonds (i-1,i) and (i,i+1)
            const int nbr = c.row + s;
            if (nbr < 0 || nbr >= SGEO.N) continue;
            const float4 xn = ldx(c, nbr);
            const float dx = xi.x - xn.x, dy = xi.y - xn.y, dz = xi.z - xn.z;
            const float r2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-12f);
            const float rinv = rsqrtf(r2), r = r2 * rinv, u = r - p.r0;
            const float g = 2.f * st.w_bond * u * rinv;
            gx -= g * dx; gy -= g * dy; gz -= g * dz;
            if (s == 1) eb = st.w_bond * u * u;
        }
        if (p.ang_d2 > 0.f) {                       // soft lower bound on CA(i)-CA(i+2): the stand-in for CNS's angle term (C3D:1641)
            const float w_ang = st.w_bond * p.ang_ratio;
            for (int s = -2; s <= 2; s += 4) {
                const int nbr = c.row + s;
                if (nbr < 0 || nbr >= SGEO.N) continue;
                const float4 xn = ldx(c, nbr);
                const float dx = xi.x - xn.x, dy = xi.y - xn.y, dz = xi.z - xn.z;
                const float r2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-12f);
                const float rinv = rsqrtf(r2), r = r2 * rinv, u = fmaxf(p.ang_d2 - r, 0.f);
                const float g = -2.f * w_ang * u * rinv;
                gx -= g * dx; gy -= g * dy; gz -= g * dz;
                if (s == 2) eb += w_ang * u * u;
            }
        }
    }
    c.f[0] = gx; c.f[1] = gy; c.f[2] = gz; c.ebond = eb;
    if (count) c.nevals++;
    TSTAMP(c, 3);                                   // gather + bond
}

// ------------------------------------------------------------------------------------------------
// L-BFGS direction from the Gram matrix (vector-free two-loop recursion); one thread.
// basis index: S slot l -> l, Y slot l -> m+l, gradient -> 2m.  Writes delta[0..2m] such that
// p = -sum_l delta[l] b_l and returns g.p.
// ------------------------------------------------------------------------------------------------
__device__ double lbfgs_direction(double* B, double* delta, int m, int cnt, int head) {
    const int nb = 2 * m + 1, ig = 2 * m;
    double al[CHR3D_LBFGS_MAX_M];
    int ring[CHR3D_LBFGS_MAX_M];                 // history slots, newest first (no integer modulo inside the loops)
    { int l = head - 1; if (l < 0) l += m; for (int k = 0; k < cnt; k++) { ring[k] = l; l = l == 0 ? m - 1 : l - 1; } }
    for (int l = 0; l < nb; l++) delta[l] = 0.0;
    delta[ig] = 1.0;
    for (int k = 0; k < cnt; k++) {
        const int idx = ring[k];
        double s = delta[ig] * B[ig * SA_MAXB + idx];
        for (int kk = 0; kk < cnt; kk++) { const int l = ring[kk];
            s += delta[l] * B[l * SA_MAXB + idx] + delta[m + l] * B[(m + l) * SA_MAXB + idx]; }
        al[idx] = s / B[idx * SA_MAXB + m + idx];
        delta[m + idx] -= al[idx];
    }
    const int nw = ring[0];
    double gam = B[nw * SA_MAXB + m + nw] / B[(m + nw) * SA_MAXB + m + nw];
    for (int l = 0; l < nb; l++) delta[l] *= gam;
    for (int k = cnt - 1; k >= 0; k--) {
        const int idx = ring[k];
        double s = delta[ig] * B[ig * SA_MAXB + m + idx];
        for (int kk = 0; kk < cnt; kk++) { const int l = ring[kk];
            s += delta[l] * B[l * SA_MAXB + m + idx] + delta[m + l] * B[(m + l) * SA_MAXB + m + idx]; }
        double be = s / B[idx * SA_MAXB + m + idx];
        delta[idx] += al[idx] - be;
    }
    double gp = delta[ig] * B[ig * SA_MAXB + ig];
    for (int kk = 0; kk < cnt; kk++) { const int l = ring[kk];
        gp += delta[l] * B[l * SA_MAXB + ig] + delta[m + l] * B[(m + l) * SA_MAXB + ig]; }
    return -gp;
}

// column indices of the L-BFGS inner products inside P (column 0 is the bond energy)
#define DOT_L(l, w) (1 + 6 * (l) + (w))     // w: 0 s.S 1 s.Y 2 y.S 3 y.Y 4 g.S 5 g.Y
#define DOT_X(m, w) (1 + 6 * (m) + (w))     // w: 0 ss 1 sy 2 sg 3 yy 4 yg 5 gg

// Every op that needs forces is written as "prepare coordinates -> evaluate -> digest", so that the
// kernel has exactly one evaluation site per energy flavour (the pair loop is the bulk of the code).
struct OpState {
    int phase, step, cnt, head, it, ls;
    bool want_e, count;
    double E, gg, gp, ts, ke;
};

// ---- L-BFGS (definition shared with the oracle's sim_minimize) ---------------------------------
__device__ __forceinline__ void min_set_trial(Ctx& c, const OpState& os) {
    if (c.owner) {
        const int t = c.tid;
        for (int k = 0; k < 3; k++) c.x[k] = vecp(c, V_XB, k)[t] + (float)os.ts * vecp(c, V_PV, k)[t];
    }
    install_x(c);
}
__device__ __forceinline__ void min_finish(Ctx& c) {       // leave x = last accepted point everywhere
    if (c.owner) { for (int k = 0; k < 3; k++) c.x[k] = vecp(c, V_XB, k)[c.tid]; }
    install_x(c);
}
// top of an iteration: choose the direction / first step, publish the trial point.  false = minimiser done.
__device__ bool min_top(Ctx& c, const DevParams& prm, const DevStage& st, OpState& os) {
    const DevParams& p = prm;
    const int t = c.tid;
    const double n3 = 3.0 * SGEO.N;
    while (true) {
        if (!(os.it < st.nstep)) break;
        if (sqrt(os.gg / n3) < (double)p.gtol) break;
        double t0 = 1.0;
        if (os.cnt == 0) {
            if (c.owner) for (int k = 0; k < 3; k++) vecp(c, V_PV, k)[t] = -vecp(c, V_G, k)[t];
            os.gp = -os.gg;
            t0 = st.drop > 0.f ? fmin(1.0, (double)st.drop / os.gg) : fmin(1.0, 1.0 / sqrt(os.gg));
        }
        if (!(os.gp < 0.0)) { if (os.cnt == 0) break; os.cnt = 0; continue; }
        os.ts = t0; os.ls = 0;
        min_set_trial(c, os);
        return true;
    }
    min_finish(c);
    return false;
}
__device__ bool min_after_eval(Ctx& c, const DevParams& prm, const DevStage& st, OpState& os) {
    const DevParams& p = prm;
    const int m = SGEO.m, t = c.tid, rp = SGEO.rows_pad;
    float* P = S_P(c);
    double* gsum = S_GSUM(c);
    if (os.phase == 0) {                     // evaluation at the starting point
        if (c.owner) {
            float gg = 0.f;
            for (int k = 0; k < 3; k++) { vecp(c, V_G, k)[t] = -c.f[k]; vecp(c, V_XB, k)[t] = c.x[k]; gg += c.f[k] * c.f[k]; }
            P[t] = c.ebond; P[rp + t] = c.row_ok ? gg : 0.f;
        }
        allreduce(c, 2, true);
        os.E = gsum[0] + gsum[1] + gsum[2]; os.gg = gsum[3];
        os.cnt = 0; os.head = 0; os.it = 0; os.gp = 0.0; os.phase = 1;
        return min_top(c, prm, st, os);
    }
    // evaluation at a trial point x = xb + ts p
    float gn[3] = {0.f, 0.f, 0.f}, sv[3] = {0.f, 0.f, 0.f}, yv[3] = {0.f, 0.f, 0.f};
    const int head = os.head;
    if (c.owner) {
        for (int k = 0; k < 3; k++) { sv[k] = (float)os.ts * vecp(c, V_PV, k)[t]; gn[k] = -c.f[k]; yv[k] = gn[k] - vecp(c, V_G, k)[t]; }
        if (!c.row_ok) for (int k = 0; k < 3; k++) { gn[k] = sv[k] = yv[k] = 0.f; }
        P[t] = c.ebond;
        for (int l = 0; l < m; l++) {
            float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (l != head) for (int k = 0; k < 3; k++) {
                float sl = vecp(c, V_S(l), k)[t], yl = vecp(c, V_Y(l), k)[t];
                a[0] = fmaf(sv[k], sl, a[0]); a[1] = fmaf(sv[k], yl, a[1]); a[2] = fmaf(yv[k], sl, a[2]);
                a[3] = fmaf(yv[k], yl, a[3]); a[4] = fmaf(gn[k], sl, a[4]); a[5] = fmaf(gn[k], yl, a[5]);
            }
            for (int w = 0; w < 6; w++) P[(size_t)DOT_L(l, w) * rp + t] = a[w];
        }
        float b[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < 3; k++) {
            b[0] = fmaf(sv[k], sv[k], b[0]); b[1] = fmaf(sv[k], yv[k], b[1]); b[2] = fmaf(sv[k], gn[k], b[2]);
            b[3] = fmaf(yv[k], yv[k], b[3]); b[4] = fmaf(yv[k], gn[k], b[4]); b[5] = fmaf(gn[k], gn[k], b[5]);
        }
        for (int w = 0; w < 6; w++) P[(size_t)DOT_X(m, w) * rp + t] = b[w];
    }
    allreduce(c, 1 + 6 * m + 6, true);
    const double Et = gsum[0] + gsum[1] + gsum[2];
    if (!(Et < os.E && Et <= os.E + (double)p.c1 * os.ts * os.gp)) {        // Armijo test failed
        os.ts *= 0.5; os.ls++;
        if (os.ls <= p.max_ls) { min_set_trial(c, os); return true; }
        os.it++;
        if (os.cnt == 0) { min_finish(c); return false; }
        os.cnt = 0;
        return min_top(c, prm, st, os);
    }
    os.it++;
    const double* G = gsum + 2;       // owner columns
    const double ss = G[DOT_X(m, 0)], sy = G[DOT_X(m, 1)], yy = G[DOT_X(m, 3)];
    const bool keep = sy > 1e-10 * sqrt(ss * yy);
    if (c.owner) {
        for (int k = 0; k < 3; k++) {
            if (keep) { vecp(c, V_S(head), k)[t] = sv[k]; vecp(c, V_Y(head), k)[t] = yv[k]; }
            vecp(c, V_XB, k)[t] = c.x[k]; vecp(c, V_G, k)[t] = gn[k];
        }
    }
    const bool small = os.cnt > 0 && (os.E - Et) <= (double)p.ftol * fmax(fabs(Et), 1.0);
    os.E = Et; os.gg = G[DOT_X(m, 5)];
    if (small) { min_finish(c); return false; }
    if (keep) {
        double* B = S_B(c); double* delta = S_DELTA(c);
        if (t == 0) {
            const int ig = 2 * m, h = head;
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
            B[ig * SA_MAXB + ig] = os.gg;
        }
        os.head = (head + 1) % m; if (os.cnt < m) os.cnt++;
        if (t == 0) delta[SA_MAXB] = lbfgs_direction(B, delta, m, os.cnt, os.head);
        __syncthreads();
        os.gp = delta[SA_MAXB];
        if (c.owner) {
            for (int k = 0; k < 3; k++) {
                float acc = (float)delta[2 * m] * vecp(c, V_G, k)[t];
                int l = os.head - 1; if (l < 0) l += m;
                for (int kk = 0; kk < os.cnt; kk++) {
                    acc = fmaf((float)delta[l], vecp(c, V_S(l), k)[t], acc);
                    acc = fmaf((float)delta[m + l], vecp(c, V_Y(l), k)[t], acc);
                    l = l == 0 ? m - 1 : l - 1;
                }
                vecp(c, V_PV, k)[t] = -acc;
            }
        }
        __syncthreads();   // delta is rewritten by thread 0 in the next iteration
    } else {
        os.cnt = 0;        // curvature condition failed: restart from steepest descent
    }
    return min_top(c, prm, st, os);
}

// ---- Verlet MD (definition shared with the oracle's sim_dynamics) ------------------------------
__device__ void dyn_begin(Ctx& c) {          // cmremove=true
    const int t = c.tid, rp = SGEO.rows_pad;
    float* P = S_P(c);
    if (c.owner) { for (int k = 0; k < 3; k++) P[(size_t)k * rp + t] = c.row_ok ? c.v[k] : 0.f; }
    allreduce(c, 3, false);
    if (c.row_ok) for (int k = 0; k < 3; k++) c.v[k] -= (float)(S_GSUM(c)[2 + k] / SGEO.N);
}
// One exchange per step besides the force exchange: owners publish a = x + dt*h*f, b = dt*v (v after
// the closing half kick) and their kinetic-energy partials; after the barrier every CTA knows the
// temperature, hence lam, and x' = a + lam*b = x + dt*(lam*v + h*f) for all beads (read lazily by ldx).
__device__ bool dyn_after_eval(Ctx& c, const DevParams& prm, const DevStage& st, OpState& os) {
    const DevParams& p = prm;
    const float h = 0.5f * st.dt_akma * p.inv_mass;
    float k2 = 0.f;
    if (c.row_ok) for (int k = 0; k < 3; k++) { if (os.phase) c.v[k] = fmaf(h, c.f[k], c.v[k]); k2 += c.v[k] * c.v[k]; }
    os.phase = 1;
    if (os.step >= st.nstep) return false;         // x, v, f of the owners are final
    os.step++;
    if (c.owner && c.tid < 32 * SGEO.nown) {        // kinetic-energy partial of this owner warp -> its 16-byte slot in every CTA
        const double kw = warp_sum((double)(0.5f * p.mass * k2));
        const int slot = SGEO.rw0 + (c.tid >> 5);    // = global bead-block index: nb slots in total
        if (c.lane < SGEO.C) {
            const uint32_t dst = smem_u32(S_KSLOTS(c) + 2 * slot);
            st_async16(mapa_u32(dst, c.lane), mapa_u32(mbar_addr(c, MB_X), c.lane),
                       make_float4(__int_as_float(__double2loint(kw)), __int_as_float(__double2hiint(kw)), 0.f, 0.f));
        }
    }
    const float dth = st.dt_akma * h;
    publish_ab(c, make_float4(fmaf(dth, c.f[0], c.x[0]), fmaf(dth, c.f[1], c.x[1]), fmaf(dth, c.f[2], c.x[2]), 0.f),
               make_float4(st.dt_akma * c.v[0], st.dt_akma * c.v[1], st.dt_akma * c.v[2], 0.f), true);
    TSTAMP(c, 4);                                   // kick, kinetic-energy partial, publish
    await_coords(c, true, 16u * (uint32_t)SGEO.nb);
    TSTAMP(c, 5);                                   // coordinate barrier
    double ke = 0.0;
    for (int q = c.lane; q < SGEO.nb; q += 32) ke += S_KSLOTS(c)[2 * q];               // usually one slot per lane,
    ke = warp_sum(ke);                                                   // same butterfly order in every warp of the cluster
    // lam = sqrt(1 + 2 fbeta dt (T0/T - 1)) (tcoupling) or sqrt(T0/T) (vscaling), T = 2 KE / (kB (3N-3)); the sum is fp64,
    // the few scalar operations after it are fp32 (they sit on the critical path of every step)
    const float kef = (float)ke;
    float lf = 1.f;
    if (kef > 0.f) {
        const float tr = st.temp * (p.kboltz * (3.f * (float)SGEO.N - 3.f)) / (2.f * kef);      // T0 / T
        if (st.flag == CHR3D_COUPLE_TCOUP) lf = sqrtf(fmaxf(0.f, fmaf(2.f * p.fbeta * st.dt_ps, tr - 1.f, 1.f)));
        else if (st.flag == CHR3D_COUPLE_VSCALE) lf = sqrtf(tr);
    }
    if (c.row_ok) {
        for (int k = 0; k < 3; k++) {
            const float a = fmaf(dth, c.f[k], c.x[k]), b = st.dt_akma * c.v[k];
            c.x[k] = fmaf(lf, b, a);                // exactly what ldx() gives every other thread
            c.v[k] = fmaf(h, c.f[k], lf * c.v[k]);
        }
    }
    flip(c, lf);
    TSTAMP(c, 6);                                   // temperature, lambda, owner update
    return true;
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <bool GENERAL, int MAXW>
__global__ void __launch_bounds__(32 * MAXW, 1)
sa_kernel(const TrajDesc* __restrict__ traj, const __grid_constant__ DevParams prm, const DevStage* __restrict__ prog, int nstages, int smem_bytes) {
    cg::cluster_group cluster = cg::this_cluster();
    Ctx c;
    c.rank = (int)cluster.block_rank();
    const int nclu = (int)cluster.num_blocks();
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    c.par = 0; c.nevals = 0; c.buf = 0; c.lam = 0.f;
    {   // trajectory descriptor, geometry and layout: set up once by thread 0 in shared memory
        if (c.tid == 0) {
            const TrajDesc td = traj[blockIdx.x / nclu];
            *(TrajDesc*)(sa_smem + 128) = td;
            SaShape sh; sa_shape(td.N, nclu, (int)(blockDim.x >> 5), prm.m, nstages, (size_t)smem_bytes, sh);
            SaLayout L; sa_layout(sh, L);
            SaGeo* g = (SaGeo*)sa_smem;
            g->N = sh.N; g->C = sh.C; g->W = sh.W; g->nthreads = sh.nthreads; g->nb = sh.nb; g->n32p = sh.n32 + 1; g->gwarps = sh.gwarps;
            g->own_max = sh.own_max; g->rows_pad = sh.rows_pad; g->m = sh.m; g->res_rounds = sh.res_rounds; g->nstages = nstages;
            g->rw0 = (int)(((long long)sh.nb * c.rank) / sh.C);
            g->nown = (int)(((long long)sh.nb * (c.rank + 1)) / sh.C) - g->rw0;
            g->o_prog = (unsigned)L.prog; g->o_mbar = (unsigned)L.mbar; g->o_ab = (unsigned)L.ab; g->o_kslots = (unsigned)L.kslots; g->o_part = (unsigned)L.part; g->o_vec = (unsigned)L.vec;
            g->o_P = (unsigned)L.P; g->o_eslots = (unsigned)L.eslots; g->o_slots = (unsigned)L.slots; g->o_gsum = (unsigned)L.gsum; g->o_B = (unsigned)L.B;
            g->o_delta = (unsigned)L.delta; g->o_tij = (unsigned)L.tij; g->o_own = (unsigned)L.own; g->o_tres = (unsigned)L.tres;
        }
        __syncthreads();
    }
    c.Tt = (const float4*)STD.T;
    c.phF = 0u; c.phX = 0u; c.phR = 0u;
    if (c.tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr(c, MB_F)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr(c, MB_X)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr(c, MB_R)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr(c, MB_R + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int ntiles = SGEO.nb * (SGEO.nb + 1) / 2;
    c.gw = c.rank * SGEO.W + c.warp;
    {   // tiles of this CTA: t = rank + C*lt, lt = 0..nloc-1; of this warp: lt = warp + W*k
        const int nloc = ntiles > c.rank ? (ntiles - 1 - c.rank) / SGEO.C + 1 : 0;
        c.my_rounds = nloc > c.warp ? (nloc - 1 - c.warp) / SGEO.W + 1 : 0;
    }
    // bead blocks of this rank: balanced contiguous split of the nb blocks over C ranks
    c.owner = c.tid < SGEO.rows_pad;
    c.row = 32 * SGEO.rw0 + c.tid;
    c.row_ok = c.tid < 32 * SGEO.nown && c.row < SGEO.N;
    for (int k = 0; k < 3; k++) { c.x[k] = 0.f; c.v[k] = 0.f; c.f[k] = 0.f; }
    c.ebond = 0.f;

    for (int i = c.tid; i < nstages * (int)(sizeof(DevStage) / 4); i += SGEO.nthreads) ((int*)(sa_smem + SGEO.o_prog))[i] = ((const int*)prog)[i];
    for (int i = c.tid; i < 4 * SGEO.n32p; i += SGEO.nthreads) S_AB(c)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = c.tid; i < 2 * SGEO.nb; i += SGEO.nthreads) S_KSLOTS(c)[i] = 0.0;
    for (int i = c.tid; i < 2 * SA_MAXC * SA_KTOT; i += SGEO.nthreads) S_SLOTS(c)[i] = 0.0;
    for (int i = c.tid; i < 2 * SGEO.gwarps; i += SGEO.nthreads) S_ESLOTS(c)[i] = 0.0;
    for (int i = c.tid; i < SA_MAXB * SA_MAXB; i += SGEO.nthreads) S_B(c)[i] = 0.0;
    for (int i = c.tid; i < (7 + 6 * SGEO.m) * SGEO.rows_pad; i += SGEO.nthreads) S_P(c)[i] = 0.f;
    for (int b = c.tid; b < SGEO.nb; b += SGEO.nthreads) {           // owner rank / local index of every bead block
        int r = 0;
        while ((int)(((long long)SGEO.nb * (r + 1)) / SGEO.C) <= b) r++;
        S_OWN(c)[b] = make_uchar2((unsigned char)r, (unsigned char)(b - (int)(((long long)SGEO.nb * r) / SGEO.C)));
    }
    if (c.lane == 0) for (int k = 0; k < c.my_rounds; k++) {   // this warp's tiles
        int I, J; sa_tile_coords(c.rank + SGEO.C * (c.warp + SGEO.W * k), SGEO.nb, I, J);
        S_TIJ(c)[k * SGEO.W + c.warp] = make_uchar2((unsigned char)I, (unsigned char)J);
    }
    for (int k = 0; k < c.my_rounds && k < SGEO.res_rounds; k++) {   // resident targets: each warp stages its own tiles (coalesced)
        const float4* gp = c.Tt + ((size_t)(c.rank + SGEO.C * (c.warp + SGEO.W * k)) * 256 + c.lane);
        float4* sp = S_TRES(c) + ((size_t)(k * SGEO.W + c.warp) * 256 + c.lane);
#pragma unroll
        for (int q = 0; q < 8; q++) sp[q * 32] = __ldg(gp + q * 32);
    }
    __syncthreads();
    csync(c);
#ifdef SA_TIMING
    c.tmark = blockIdx.x == 0 && c.lane == 0 && (c.warp == 0 || c.warp == SGEO.W - 1);
    c.tlast = clock64(); for (int q = 0; q < 10; q++) c.tacc[q] = 0;
#endif

    int cur_image = 1, hand = 1;
    double e_saved = 0.0;
    bool after_dyn = false;
    const int k_begin = STD.k1 > 0 ? STD.k0 : 0, k_end = STD.k1 > 0 ? STD.k1 : nstages;
    if (STD.state_in) {            // resume: the other image's end state becomes the current one (as after its last DYNAMICS op)
        if (c.row_ok) for (int q = 0; q < 3; q++) { c.x[q] = STD.state_in[3 * c.row + q]; c.v[q] = STD.state_in[3 * (SGEO.N + c.row) + q]; }
        cur_image = STD.resume_image; hand = cur_image;
        if (STD.esaved_in) e_saved = *STD.esaved_in;
        csync(c);
        install_x(c);
    }
    for (int k = k_begin; k < k_end; k++) {
        const DevStage& st = S_PROG(c)[k];
        OpState os;
        os.phase = 0; os.step = 0; os.cnt = 0; os.head = 0; os.it = 0; os.ls = 0; os.want_e = true; os.count = true;
        os.E = os.gg = os.gp = os.ts = os.ke = 0.0;
        bool more = false;
        // A DYNAMICS op ends with a force evaluation and no exchange after it (the last step publishes nothing).  If the next
        // op starts with an evaluation of its own (MINIMIZE, SAVE, PICK, ENERGY), a fast CTA could send the partial forces of
        // that evaluation into the force phase a slow peer has not completed yet: the bytes would be credited to the wrong
        // phase (deadlock, seen as intermittent hangs with co-resident CTAs) and the partial slots overwritten under the
        // peer's gather.  A cluster barrier separates the two phases (twice per trajectory of the dgsa.inp program).
        if (after_dyn && (st.op == CHR3D_OP_MINIMIZE || st.op == CHR3D_OP_SAVE || st.op == CHR3D_OP_PICK || st.op == CHR3D_OP_ENERGY)) csync(c);
        if (st.op != CHR3D_OP_MAXWELL) after_dyn = st.op == CHR3D_OP_DYNAMICS;      // every other op ends with an exchange
        switch (st.op) {
        case CHR3D_OP_INIT:
            if (c.row_ok) {
                c.x[0] = STD.x0[3 * c.row] * (float)st.flag; c.x[1] = STD.x0[3 * c.row + 1]; c.x[2] = STD.x0[3 * c.row + 2];
                c.v[0] = c.v[1] = c.v[2] = 0.f;
            }
            csync(c);                   // two coordinate phases without a force phase in between must not mix
            install_x(c);
            cur_image = st.flag; hand = cur_image;
            break;
        case CHR3D_OP_MAXWELL:
            if (c.row_ok) {
                double z[3]; normal3(STD.seed, (uint32_t)c.row, STREAM_MAXWELL + (uint32_t)k, (uint32_t)STD.model, z);
                double sd = sqrt((double)prm.kboltz * (double)st.temp / (double)prm.mass);
                for (int q = 0; q < 3; q++) c.v[q] = (float)(sd * z[q]);
            }
            break;
        case CHR3D_OP_CENTER:
            if (c.owner) for (int q = 0; q < 3; q++) S_P(c)[(size_t)q * SGEO.rows_pad + c.tid] = c.row_ok ? c.x[q] : 0.f;
            allreduce(c, 3, false);
            if (c.row_ok) { for (int q = 0; q < 3; q++) c.x[q] -= (float)(S_GSUM(c)[2 + q] / SGEO.N); }
            install_x(c);
            break;
        case CHR3D_OP_MINIMIZE: more = true; break;
        case CHR3D_OP_DYNAMICS: dyn_begin(c); os.want_e = false; more = true; break;
        case CHR3D_OP_SAVE: case CHR3D_OP_PICK: more = true; break;
        case CHR3D_OP_ENERGY: os.count = false; more = true; break;
        default: break;
        }
        while (more) {
            if (os.want_e) eval_t<true, GENERAL>(c, prm, st, os.count); else eval_t<false, GENERAL>(c, prm, st, os.count);
            if (st.op == CHR3D_OP_MINIMIZE) { more = min_after_eval(c, prm, st, os); TSTAMP(c, 7); }
            else if (st.op == CHR3D_OP_DYNAMICS) more = dyn_after_eval(c, prm, st, os);
            else {
                more = false;
                if (c.owner) S_P(c)[c.tid] = c.ebond;
                allreduce(c, 1, true);
                const double* gsum = S_GSUM(c);
                const double e = gsum[0] + gsum[1] + gsum[2];
                if (st.op == CHR3D_OP_SAVE) {
                    e_saved = e;
                    if (c.row_ok) for (int q = 0; q < 3; q++) { STD.scratch[3 * c.row + q] = c.x[q]; STD.scratch[3 * (SGEO.N + c.row) + q] = c.v[q]; }
                } else if (st.op == CHR3D_OP_PICK) {
                    if (e_saved > e) hand = cur_image;                 // C3D:1717-1721
                    else {
                        hand = 1;
                        if (c.row_ok) {
                            for (int q = 0; q < 3; q++) { c.x[q] = STD.scratch[3 * c.row + q]; c.v[q] = STD.scratch[3 * (SGEO.N + c.row) + q]; }
                        }
                        install_x(c);
                    }
                } else {
                    if (c.rank == 0 && c.tid == 0) {
                        STD.e_out[CHR3D_E_OVERALL] = e; STD.e_out[CHR3D_E_BON] = gsum[2]; STD.e_out[CHR3D_E_VDW] = gsum[1]; STD.e_out[CHR3D_E_NOE] = gsum[0];
                    }
                    if (STD.f_out && c.row_ok) for (int q = 0; q < 3; q++) STD.f_out[3 * c.row + q] = c.f[q];
                }
            }
        }
    }
    if (c.row_ok) for (int q = 0; q < 3; q++) STD.x_out[3 * c.row + q] = c.x[q];
    if (STD.state_out && c.row_ok) for (int q = 0; q < 3; q++) { STD.state_out[3 * c.row + q] = c.x[q]; STD.state_out[3 * (SGEO.N + c.row) + q] = c.v[q]; }
    if (c.rank == 0 && c.tid == 0) {
        *STD.nevals_out = c.nevals; *STD.hand_out = hand;
        if (STD.esaved_out) *STD.esaved_out = e_saved;
    }
#ifdef SA_TIMING
    if (c.tmark) for (int q = 0; q < 10; q++) sa_dbg[c.warp == 0 ? 0 : 1][q] = (unsigned long long)c.tacc[q];
    if (c.tmark) sa_dbg[c.warp == 0 ? 0 : 1][10] = (unsigned long long)c.nevals;
#endif
    csync(c);   // no CTA may exit while peers can still write into its shared memory
}

// ------------------------------------------------------------------------------------------------
// host side: configuration choice and launch
// ------------------------------------------------------------------------------------------------
// kernel variant for a CTA size: the register cap follows from the thread count (one CTA per SM) — 168 registers up to
// 12 warps (no spills in the pair loop), 128 up to 16, 96 up to 20
typedef void (*SaKernelFn)(const TrajDesc*, DevParams, const DevStage*, int, int);
static SaKernelFn sa_pick_kernel(bool general, int threads) {
    if (threads <= 384) return general ? sa_kernel<true, 12> : sa_kernel<false, 12>;
    if (threads <= 512) return general ? sa_kernel<true, 16> : sa_kernel<false, 16>;
    return general ? sa_kernel<true, 20> : sa_kernel<false, 20>;
}
static size_t smem_fixed(int N, int C, int W, int m) {
    SaShape sh; sa_shape(N, C, W, m, SA_MAX_STAGES, 0, sh);
    SaLayout L; sa_layout(sh, L);
    return L.total;
}
static int max_active_clusters(int C, int threads, size_t smem) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(C * 148), 1, 1); lc.blockDim = dim3((unsigned)threads, 1, 1); lc.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = (unsigned)C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, sa_pick_kernel(false, threads), &lc) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// busiest scheduler of the busiest CTA: warps w = s, s+4, ... share scheduler s; warp w owns ceil((nloc - w) / W) tiles
static int tiles_per_scheduler(int ntiles, int C, int W) {
    const int nloc = (ntiles + C - 1) / C;
    int worst = 0;
    for (int sch = 0; sch < 4; sch++) {
        int t = 0;
        for (int w = sch; w < W; w += 4) if (nloc > w) t += (nloc - 1 - w) / W + 1;
        worst = t > worst ? t : worst;
    }
    return worst;
}

int sa_choose_config(int N, int n_traj, int m, SaLaunchCfg* cfg) {
    // the choice depends only on (N, n_traj, m) and the device: cached, the occupancy queries are slow
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int>, SaLaunchCfg> cache;
    int dev = 0, smem_max = 227 * 1024;
    cudaGetDevice(&dev);
    const int force_c = dev_getenv("CHR3D_SA_CLUSTER") ? atoi(dev_getenv("CHR3D_SA_CLUSTER")) : 0;
    const int force_w = dev_getenv("CHR3D_SA_WARPS") ? atoi(dev_getenv("CHR3D_SA_WARPS")) : 0;
    const auto key = std::make_tuple(dev, N, n_traj, m);
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end() && !force_c && !force_w) { *cfg = it->second; return 0; }
    }
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int smem_hw = smem_max;
    if (dev_getenv("CHR3D_SA_SMEM_CAP")) smem_max = std::min(smem_max, atoi(dev_getenv("CHR3D_SA_SMEM_CAP")));   // development: force >1 CTA per SM
    for (int th : {384, 512, 640}) {
        if (cudaFuncSetAttribute(sa_pick_kernel(false, th), cudaFuncAttributeMaxDynamicSharedMemorySize, smem_hw) != cudaSuccess) cudaGetLastError();
        if (cudaFuncSetAttribute(sa_pick_kernel(false, th), cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) cudaGetLastError();
    }
    const int nb = (N + 31) / 32, ntiles = nb * (nb + 1) / 2;
    // One CTA per SM, enforced by asking for more than half of the SM's shared memory: batches run faster this way (all 45
    // inputs x 20 models: 193 ms against 208 ms), because the shapes the cost model then picks leave fewer, larger launch
    // groups.  (Co-resident CTAs used to hang intermittently; that was a phase race in the exchange protocol, fixed in
    // sa_kernel — see "after_dyn" — and documented in DESIGN.md.)  CHR3D_SA_MIN_SMEM=0 lifts the floor.
    const size_t min_smem = dev_getenv("CHR3D_SA_MIN_SMEM") ? (size_t)atol(dev_getenv("CHR3D_SA_MIN_SMEM")) : (size_t)117 * 1024;
    double best = 1e300; SaLaunchCfg bc = {1, 128, 0, 0, 0.0};
    const int Cs[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16};     // > 8: non-portable cluster sizes, two CTAs of different trajectories per SM
    for (int C : Cs) {
        if (force_c && C != force_c) continue;
        if (C > nb && C > 1) continue;
        // Non-portable cluster sizes (> 8) only on request (CHR3D_SA_CLUSTER): with one CTA per SM (see min_smem) they cannot
        // hold 20 trajectories in one wave, and they have far less test mileage than the portable sizes.
        if (C > 8 && !force_c) continue;
        for (int W = 4; W <= SA_MAXW; W += 2) {
            if (force_w && W != force_w) continue;
            if (!force_w && (W & 3)) continue;
            const int own_max = (nb + C - 1) / C;
            if (own_max > W) continue;                           // every owned bead needs its own thread
            const size_t fixed = smem_fixed(N, C, W, m);
            if (fixed > (size_t)smem_max) continue;
            const int rounds = ((ntiles + C - 1) / C + W - 1) / W;
            // resident targets when they fit, otherwise as many rounds as fit (the rest streams from L2)
            size_t want = fixed + (size_t)rounds * W * SA_TILE_FLOATS * sizeof(float);
            size_t smem = want <= (size_t)smem_max ? want : (size_t)smem_max;
            smem = std::max(smem, std::min(min_smem, (size_t)smem_max));
            int res_rounds = (int)((smem - fixed) / ((size_t)W * SA_TILE_FLOATS * sizeof(float)));
            if (res_rounds > rounds) res_rounds = rounds;
            int resident = max_active_clusters(C, 32 * W, smem);
            if (resident < 1 && smem > fixed) { smem = fixed; res_rounds = 0; resident = max_active_clusters(C, 32 * W, smem); }
            if (resident < 1) continue;
            const int waves = (n_traj + resident - 1) / resident;
            // cost model per force evaluation (SM cycles), fitted to B200 measurements of the packed-math kernel
            // (chr1_500kb: 6x12 warps 14.1k cycles, 3x12 23.0k, 3x8 23.1k, 2x12 29.2k per evaluation):
            // pair phase = tiles of the busiest scheduler x ~1330 issue slots, stretched when few warps hide latency;
            // fixed part = owner phases, reductions, exchanges (grows with the warps that meet at them)
            const double wps = W / 4.0;
            const double spill = W > 16 ? 1.6 : (W > 12 ? 1.35 : 1.0);     // register cap of the wider variants: the pair loop spills
            const double tile_cycles = spill * 1330.0 * (1.0 + 1.0 / wps) + (rounds > res_rounds ? 300.0 * (rounds - res_rounds) / rounds : 0.0);
            const double pair_cycles = tiles_per_scheduler(ntiles, C, W) * tile_cycles;
            const double over = 3000.0 + 100.0 * W + (C > 1 ? 1000.0 : 0.0);
            const double cost = waves * (pair_cycles + over);
            if (cost < best) { best = cost; bc.cluster = C; bc.threads = 32 * W; bc.stage_targets = res_rounds; bc.smem = smem; bc.cost = cost; }
        }
    }
    if (best >= 1e300) { set_error("no SA kernel configuration fits N=%d (use the single-structure engine)", N); return CHR3D_ERR_SIZE; }
    *cfg = bc;
    if (!force_c && !force_w) { std::lock_guard<std::mutex> g(mu); cache[key] = bc; }
    return 0;
}

int sa_launch(const TrajDesc* d_traj, int n_traj, int Nmax, const DevParams& P, const DevStage* d_prog, int nstages,
              const SaLaunchCfg& cfg, cudaStream_t s) {
    (void)Nmax;
    if (nstages > SA_MAX_STAGES) { set_error("stage program has %d entries, the SA kernel keeps at most %d", nstages, SA_MAX_STAGES); return CHR3D_ERR_SIZE; }
    const bool general = P.general_noe || P.lo_clamp > -1e30f;
    SaKernelFn kern = sa_pick_kernel(general, cfg.threads);
    CHR3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    if (cfg.cluster > 8) CHR3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(n_traj * cfg.cluster), 1, 1);
    lc.blockDim = dim3((unsigned)cfg.threads, 1, 1);
    lc.dynamicSmemBytes = cfg.smem;
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cfg.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    int sm = (int)cfg.smem;
    CHR3D_CUDA(cudaLaunchKernelEx(&lc, kern, d_traj, P, d_prog, nstages, sm));
    return 0;
}

}  // namespace chr3d

#ifdef SA_HANGDIAG
extern "C" __attribute__((visibility("default"))) unsigned long long* chr3d_debug_sa_hang_init(void) {
    unsigned long long* h = nullptr; void* d = nullptr;
    if (cudaHostAlloc((void**)&h, sizeof(unsigned long long) * 4096, cudaHostAllocMapped) != cudaSuccess) return nullptr;
    memset(h, 0, sizeof(unsigned long long) * 4096);
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return nullptr;
    if (cudaMemcpyToSymbol(chr3d::sa_diag, &d, sizeof d) != cudaSuccess) return nullptr;
    return h;
}
#endif
#ifdef SA_TIMING
extern "C" __attribute__((visibility("default"))) int chr3d_debug_sa_timing(unsigned long long* out32) {
    return cudaMemcpyFromSymbol(out32, chr3d::sa_dbg, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -3;
}
#endif
