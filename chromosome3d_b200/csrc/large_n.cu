// large_n.cu — the large-N hot path: one structure of thousands of beads per GPU (or per rank).
//
// Replaces the same CNS dgsa.inp run (C3D:254-289, protocol C3D:1573-1829) for chromosomes far
// beyond the reference's 663-bead cap (C3D:93-94): synthetic 50 kb (N=4,985) and 10 kb (N=24,925)
// resolutions of BASELINE.json.
//
// Data layout in HBM
//   * targets: upper triangle of the dense target matrix in 128x128 fp32 tiles, tile-packed in
//     row-major tile order (I <= J), each tile one contiguous 64 KB block -> one TMA bulk copy
//     (cp.async.bulk, SASS UBLKCP) per tile.  4 B per unordered pair: the algorithmic minimum.
//     Sentinels: -1 no restraint, -2 no restraint and no repel (|i-j| < vdw_min_sep, lower part of
//     diagonal tiles, padding).
//   * coordinates / forces / L-BFGS vectors: float4 arrays of N.
// Force evaluation = ln_tile_kernel (persistent CTAs, contiguous tile ranges, three-stage TMA ring,
// symmetric tile evaluation: every unordered pair is evaluated once and its force goes to both beads)
//   - a warp owns 8 rows of the tile, a lane the 4 columns {lane, lane+32, lane+64, lane+96} (stored as one float4 per row
//     in the tile-packed layout, so target and coordinate reads are conflict-free): 32 pairs per thread per tile in registers;
//   - row forces stay in registers while consecutive tiles share the tile row, column forces are
//     reduced across the 16 warps through shared memory once per tile;
//   - partial forces are written per tile / per row segment and summed in a fixed order by
//     ln_reduce_kernel: no atomics, bit-reproducible.
// Multi-GPU: rank g owns a contiguous range of the tile list (and only that part of the targets);
// after the local reduction forces (+ energies) are summed with one NCCL all-reduce per force
// evaluation over NVLink and every rank integrates all N beads redundantly, so no coordinate
// exchange is needed (N*16 B per evaluation either way).
// MD and L-BFGS run as small vector kernels with all decisions taken on the device (no host
// round trip per step); the host only polls a `done` flag every few dozen evaluations.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <memory>
#include <vector>
#include <cooperative_groups.h>
#include "common.h"
#ifdef CHR3D_WITH_NCCL
#include <dlfcn.h>
#include <nccl.h>
// NCCL is bound at run time, when a multi-rank structure is first created: dlopen picks the libnccl.so.2 the process
// already has (PyTorch brings its own), otherwise the system's.  Linking it at build time would pin whichever copy the
// loader finds first and break a later `import torch` in the same process.
namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    bool ok = false;
};
NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy;
        }
    }
    return api;
}
}  // namespace
#define ncclGetUniqueId nccl_api().GetUniqueId
#define ncclCommInitRank nccl_api().CommInitRank
#define ncclAllReduce nccl_api().AllReduce
#define ncclCommDestroy nccl_api().CommDestroy
#endif

namespace chr3d {

#define LN_TB 128
#define LN_THREADS 512
#ifndef LN_BOUND_THREADS
#define LN_BOUND_THREADS LN_THREADS      // register cap of the tile kernel = 65536 / this (development: -DLN_BOUND_THREADS=640 ... probes the footprint)
#endif
#define LN_TILE_FLOATS (LN_TB * LN_TB)
#define LN_MAXB (2 * CHR3D_LBFGS_MAX_M + 1)
#define LN_KMAX 64

struct LnTerms { float wnC, wnC2m, lo, rsw, wv, wv4, R2; };
// the O(N) terms added by the bead's owner: harmonic bond (i, i+1) and, when d2 > 0, the soft lower bound on the CA(i)-CA(i+2)
// distance that stands in for CNS's angle term (C3D:1641; chr3d_sa_params.angle_d2 / angle_k), same `weights *` factor
struct LnBonded { float w_bond, r0, w_ang, d2; };

// one symmetric pair: accumulates the force on i (fi) and on j (fj) and, if WANT_E, the energies.
// r^2 carries a 1e-12 floor folded into the FMA chain (coincident beads), as the oracle's max(r^2, 1e-12).
// r^-1/2 as one MUFU.RSQ (arguments are >= 1e-12: no denormal handling needed)
__device__ __forceinline__ float ln_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// NEAR: the tile touches the diagonal band (J <= I + 1), where cells can carry the "no repel" sentinel -2
// (bonded neighbours, the diagonal, the lower half of diagonal tiles).  Padding rows/columns of the last
// tile row/column carry it too, so those tiles also take the NEAR path.
//
// The pair term runs on Blackwell's packed fp32 pipe (FADD2 / FMUL2 / FFMA2, PTX *.f32x2): one call evaluates the pairs
// (row a, column j) and (row a+1, column j) in the two halves of 64-bit register pairs.  The kernel is issue-bound with
// scalar FP32 (about 22 instructions per pair against 4 B of HBM traffic); packing halves the FMA-class instruction count.
// Scalar operands (column coordinates, weights) ride as broadcast operands of the packed instructions.  Each half computes
// exactly what the scalar form computed (same operations, same order), so results do not depend on the packing.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
template <bool WANT_E, bool LOW_SOFT, bool NEAR>
__device__ __forceinline__ void pair_sym2(const LnTerms& t, const float2 xi, const float2 yi, const float2 zi, float xj, float yj, float zj, const float2 d,
                                          float2& fix, float2& fiy, float2& fiz, float2& gjx, float2& gjy, float2& gjz, float2& en, float2& ev) {
    const float2 dx = __fadd2_rn(xi, bc(-xj)), dy = __fadd2_rn(yi, bc(-yj)), dz = __fadd2_rn(zi, bc(-zj));
    const float2 r2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, __ffma2_rn(dz, dz, bc(1e-12f))));
    const float2 rinv = f2(ln_rsqrt(r2.x), ln_rsqrt(r2.y));
    const float2 delta = __ffma2_rn(r2, rinv, f2(-d.x, -d.y));                          // r - d
    float2 h = f2(fminf(delta.x, t.rsw), fminf(delta.y, t.rsw));
    if (LOW_SOFT) h = f2(fmaxf(h.x, t.lo), fmaxf(h.y, t.lo));
    const float2 has = f2(d.x > 0.f ? 1.f : 0.f, d.y > 0.f ? 1.f : 0.f);                // FSET.BF
    float2 tr = __fadd2_rn(bc(t.R2), f2(-r2.x, -r2.y));
    tr = f2(fmaxf(tr.x, 0.f), fmaxf(tr.y, 0.f));
    if (NEAR) tr = f2(d.x > -1.5f ? tr.x : 0.f, d.y > -1.5f ? tr.y : 0.f);
    const float2 w2h = __fmul2_rn(__fmul2_rn(has, bc(t.wnC2m)), h);                     // -2 w C h
    if (WANT_E) {
        en = __ffma2_rn(__fmul2_rn(w2h, bc(-0.5f)), __ffma2_rn(delta, bc(2.f), f2(-h.x, -h.y)), en);   // w C h (2 delta - h)
        ev = __ffma2_rn(__fmul2_rn(tr, bc(t.wv)), tr, ev);
    }
    const float2 coef = __ffma2_rn(bc(t.wv4), tr, __fmul2_rn(w2h, rinv));              // 4 w_vdw t - (dE_noe/dr)/r
    fix = __ffma2_rn(coef, dx, fix); fiy = __ffma2_rn(coef, dy, fiy); fiz = __ffma2_rn(coef, dz, fiz);
    gjx = __ffma2_rn(coef, dx, gjx); gjy = __ffma2_rn(coef, dy, gjy); gjz = __ffma2_rn(coef, dz, gjz);   // force on j = -(gj.x + gj.y)
}

__host__ __device__ inline long long tile_index(long long I, long long J, long long nT) { return I * nT - I * (I - 1) / 2 + (J - I); }
__host__ __device__ inline void tile_coords(long long t, long long nT, int& I, int& J) {
    // largest I with I*nT - I(I-1)/2 <= t
    double a = (double)(2 * nT + 1);
    long long i = (long long)floor((a - sqrt(a * a - 8.0 * (double)t)) * 0.5);
    if (i < 0) i = 0;
    while (i > 0 && tile_index(i, i, nT) > t) i--;
    while (tile_index(i + 1, i + 1, nT) <= t) i++;
    I = (int)i; J = (int)(i + (t - tile_index(i, i, nT)));
}

// Position k of a tile-packed tile -> (row, column) inside the tile.  Thread (warp w, lane l) of ln_tile_kernel owns rows
// 8w..8w+7 and columns l+32c (c = 0..3); its 32 cells are 8 float4, number q = 2*rp + ch (row pair rp, column half ch) at
// float offset ((8w + q)*32 + l)*4, holding {(2rp, 2ch), (2rp+1, 2ch), (2rp, 2ch+1), (2rp+1, 2ch+1)}: two row-packed f32x2
// operands per 128-bit shared-memory load, warp loads contiguous (conflict-free).
__host__ __device__ inline void ln_tile_cell(int k, int& r, int& c) {
    const int e = k & 3, l = (k >> 2) & 31, q = (k >> 7) & 7, w = k >> 10;
    r = 8 * w + 2 * (q >> 1) + (e & 1);
    c = 32 * (2 * (q & 1) + (e >> 1)) + l;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// A wait that cannot complete is a bug in the pipeline protocol: trap after ~10 s instead of hanging the GPU.
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 20000000000ll) __trap();
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    if (!done) mbar_wait_slow(smem_addr(bar), parity);
}

struct LnTileArgs {
    const float* Dt;          // tile-packed targets of this rank: tiles [t0, t1)
    long long t0, t1, nT;
    int N;
    const float4* x;          // coordinates are x + lam * xb2 (lazy: the integrator publishes (a, b) and the scalar lam)
    const float4* xb2;        // may be nullptr (then lam is ignored)
    const float* lam;         // device scalar
    const int* done;          // may be nullptr; *done != 0: the minimiser has finished, skip the evaluation
    float* colpart;           // [(t - t0)][3][128]
    float4* rowpart;          // [cta][maxseg][128]
    double* epart;            // [cta][2]
    const long long* cstart;  // [ctas + 1] first tile of every CTA (cost-balanced partition of [t0, t1), ln_plan)
    int maxseg, stream_l2;
    int dbg;                  // development probe (-DCHR3D_DEV + CHR3D_LN_DEBUG): 1 = stream tiles without the pair term, 2 = pair term without tile traffic
    LnTerms terms;
};

// dynamic smem: LN_STAGES x 64 KB tile buffers + colp[3][16][128] + mbarriers (one per stage + "colp free").
// Three stages keep two tiles (128 KB per SM, 19 MB chip-wide) in flight behind the one being evaluated: with two stages only
// one 64 KB copy per SM is outstanding, which at the loaded HBM latency caps the kernel near 60 % of the copy bandwidth.
#define LN_STAGES 3
#define LN_BAR_OFF (LN_STAGES * LN_TILE_FLOATS * 4)       // mbarriers (one per stage + "colp free")
#define LN_RING_OFF (LN_BAR_OFF + 32)                    // LnRing: the pipeline state that survives from one evaluation to the next
#define LN_COLP_OFF (LN_BAR_OFF + 64)                    // colp[3][16][128]
#define LN_CBUF_OFF (LN_COLP_OFF + 3 * 16 * LN_TB * 4)   // column coordinates (a, b) of the current and the next tile
#define LN_SMEM (LN_CBUF_OFF + 2 * 2 * LN_TB * 16)
#define LN_SCRATCH_BYTES (LN_SMEM - LN_COLP_OFF)         // colp + cbuf: free between two tile phases (scratch of the step phases)

#ifdef CHR3D_DEV
#define LN_DBG(a) ((a).dbg)
#else
#define LN_DBG(a) 0
#endif

// Everything the tile phase of one CTA needs, in SHARED memory: the pair loop wants the whole register budget, so nothing of
// this is carried in registers — it is read where it is used (there is no L1 behind 229 KB of shared memory: a spill in the
// pair loop would be an L2 round trip, a shared-memory read is not).
//
// Ring state: a CTA's n tiles are walked as a periodic sequence — position g (counted over all evaluations of a launch) is
// tile g % n of the range, stage g % LN_STAGES.  After position g has been evaluated the copy of position g + LN_STAGES is
// issued, so at the end of an evaluation the first tiles of the next one are already in flight (the targets do not depend on
// the coordinates): a persistent kernel pays the pipeline prologue once per launch, not once per evaluation.
struct LnTileRun {
    const float* Dt;          // first tile of this CTA
    float* colpart;           // column partials of this CTA's tiles: [k][3][128]
    float4* rowpart;          // row partials of this CTA's segments: [seg][128]
    double* epart;            // [2]
    const float4* X; const float4* XB; float lam;      // coordinates X + lam * XB
    int n, I0, J0, nT, N;     // tiles, tile coordinates of the first one
    int more, dbg;            // more: another evaluation follows in this launch (keep the ring running across the boundary)
    int ring_tail;            // tiles of the next evaluation requested before this tile phase ends (the others at its start)
    int stream_l2;            // 1: the targets of this rank exceed what L2 can keep between evaluations -> the TMA copies carry an
                              // evict-first L2 policy, so that the stream does not push the partial forces and the vectors out of L2
    LnTerms terms;
    unsigned long long l2pol; // L2 cache policy of the TMA copies (created once per launch from stream_l2)
    unsigned int pending;     // stages with a copy outstanding
    unsigned int pos;         // position of the next tile in the periodic sequence (tiles consumed since the launch started)
    // bookkeeping of the walk, per tile, written one tile ahead by thread 0 and read by everybody at the top of the tile
    // (double-buffered by tile parity; the per-tile __syncthreads orders the two): nothing of it is carried in registers
    struct Step { int I, J, Jn, stage, tpar, cpar, near, newrow, seg, last; } step[2];
};
#define LN_RUN(field) (((volatile LnTileRun*)run)->field)

__device__ __forceinline__ void ln_ring_issue(LnTileRun* run, unsigned char* smem, int k, int stage) {      // one elected thread: 64 KB bulk copy of the CTA's tile k
    float* dbuf = (float*)smem; uint64_t* bar = (uint64_t*)(smem + LN_BAR_OFF);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar[stage])), "r"(LN_TILE_FLOATS * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_addr(dbuf + (size_t)stage * LN_TILE_FLOATS)), "l"(LN_RUN(Dt) + (size_t)k * LN_TILE_FLOATS),
                   "r"(LN_TILE_FLOATS * 4), "r"(smem_addr(&bar[stage])), "l"(LN_RUN(l2pol)) : "memory");
    LN_RUN(pending) = LN_RUN(pending) | (1u << stage);
}
// once per launch (all threads; run->Dt, n, dbg are set): mbarriers, ring state, the first LN_STAGES copies.
// periodic: the ring runs across evaluations (persistent kernel); otherwise only tiles of one evaluation are issued.
__device__ __forceinline__ void ln_ring_init(LnTileRun* run, unsigned char* smem, bool periodic) {
    uint64_t* bar = (uint64_t*)(smem + LN_BAR_OFF);
    if (threadIdx.x == 0) {
        for (int q = 0; q < LN_STAGES; q++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar[q])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar[LN_STAGES])), "r"(3 * LN_TB / 32));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        run->pending = 0; run->pos = 0;
        {
            unsigned long long pol;
#ifdef LN_NO_L2HINT
            if (false) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#else
            if (run->stream_l2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
            else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
            run->l2pol = pol;
        }
        const int n = run->n;
        if (n > 0 && run->dbg != 2) for (int q = 0; q < LN_STAGES; q++) if (periodic || q < n) ln_ring_issue(run, smem, q % n, q);
    }
    __syncthreads();
}
// before the kernel exits (all threads): wait for the copies still in flight
__device__ __forceinline__ void ln_ring_drain(LnTileRun* run, unsigned char* smem) {
    uint64_t* bar = (uint64_t*)(smem + LN_BAR_OFF);
    __syncthreads();
    // stage q is consumed next at the first position >= pos that maps to it; its parity there is (position / LN_STAGES) & 1
    if (threadIdx.x == 0) for (int q = 0; q < LN_STAGES; q++) if (run->pending >> q & 1u) {
        unsigned int g = run->pos; while (g % LN_STAGES != (unsigned)q) g++;
        mbar_wait(&bar[q], (g / LN_STAGES) & 1u);
    }
    __syncthreads();
}

// The first LN_STAGES tiles of the NEXT evaluation (one thread).  The persistent kernel calls this from the step phase, once the
// partial forces are on their way to the peers: issued at the end of the tile phase instead, these 28 MB (chip-wide) of bulk
// copies would sit in front of the latency-critical loads of the gather.
__device__ __forceinline__ void ln_ring_prefetch(LnTileRun* run, unsigned char* smem) {
    const int n = run->n;
    if (n <= 0 || run->dbg == 2) return;
    for (int q = 0; q < LN_STAGES; q++) {
        const unsigned int g = run->pos + (unsigned)q;
        if (!(run->pending >> (g % LN_STAGES) & 1u)) ln_ring_issue(run, smem, q % n, (int)(g % LN_STAGES));
    }
}

// One force evaluation of this CTA's tile range at the coordinates X + lam * XB (all in *run).
// tm: the pair-term constants, handed over by value from the KERNEL PARAMETER (constant bank -> uniform registers: they do not
// occupy vector registers of the pair loop)
template <bool WANT_E, bool LOW_SOFT>
__device__ __forceinline__ void ln_tile_phase(LnTileRun* run, unsigned char* smem, const LnTerms tm) {
#ifdef CHR3D_DEV
    const int dbg = run->dbg;
#else
    constexpr int dbg = 0;
#endif
    float* dbuf = (float*)smem;
    float* colp = (float*)(smem + LN_COLP_OFF);
    uint64_t* bar = (uint64_t*)(smem + LN_BAR_OFF);         // [0, LN_STAGES): tile landed; [LN_STAGES]: colp drained by the reducer warps
    float4* cbuf = (float4*)(smem + LN_CBUF_OFF);       // [stage][a|b][128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = run->n;
    const float lam = run->lam;
    // column coordinates of tile column Jn -> cbuf[stg] with cp.async (L2 latency hidden behind the previous tile's pairs)
    auto prefetch_cols = [&](int Jn, int stg) {
        if (tid < 2 * LN_TB) {
            const int col = tid & (LN_TB - 1), which = tid >> 7, N = LN_RUN(N);
            int j = Jn * LN_TB + col; if (j >= N) j = N - 1;
            const float4* src = (which ? LN_RUN(XB) : LN_RUN(X)) + j;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(cbuf + (stg * 2 + which) * LN_TB + col)), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // tile k of this evaluation = position pos + k of the periodic sequence: stage (pos + k) % LN_STAGES, whose mbarrier
    // completes its ((pos + k) / LN_STAGES)-th phase when the tile has landed; the "colp free" barrier completes one phase per
    // tile (consumed by the next tile, the last one at the end of the evaluation)
    auto make_step = [&](int k, int I, int J, int seg, int newrow) {         // thread 0
        LnTileRun::Step st;
        const unsigned int g = run->pos + (unsigned)k;
        st.I = I; st.J = J; st.stage = (int)(g % LN_STAGES); st.tpar = (int)((g / LN_STAGES) & 1u); st.cpar = (int)((g + 1u) & 1u);
        int Jn = J + 1; if (Jn == run->nT) Jn = I + 1;
        st.Jn = Jn; st.near = (J <= I + 1 || (J + 1) * LN_TB > run->N) ? 1 : 0; st.newrow = newrow; st.seg = seg; st.last = (J + 1 == run->nT || k + 1 == n) ? 1 : 0;
        run->step[k & 1] = st;
    };
    if (tid == 0 && n > 0) {
        make_step(0, run->I0, run->J0, 0, 1);
        if (dbg != 2) for (int q = 0; q < LN_STAGES && q < n; q++) {      // top the ring up (tiles a shortened tail left out, or a deferred ring)
            const unsigned int g = run->pos + (unsigned)q;
            if (!(run->pending >> (g % LN_STAGES) & 1u)) ln_ring_issue(run, smem, q, (int)(g % LN_STAGES));
        }
    }
    __syncthreads();
    if (n > 0) prefetch_cols(LN_RUN(step[0].J), (int)(LN_RUN(pos) & 1u));
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    float2 xi[4], yi[4], zi[4], fix[4], fiy[4], fiz[4];          // row pairs (8w + 2rp, 8w + 2rp + 1)
    float2 en = bc(0.f), ev = bc(0.f);

    auto flush_rows = [&](int seg) {
        // reduce the 8 row accumulators over the 32 lanes (columns) of this warp
        float4* rp_out = LN_RUN(rowpart) + (size_t)seg * LN_TB + warp * 8;
#pragma unroll
        for (int r = 0; r < 8; r++) {
            float vx = (r & 1) ? fix[r >> 1].y : fix[r >> 1].x, vy = (r & 1) ? fiy[r >> 1].y : fiy[r >> 1].x, vz = (r & 1) ? fiz[r >> 1].y : fiz[r >> 1].x;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { vx += __shfl_xor_sync(0xffffffffu, vx, o); vy += __shfl_xor_sync(0xffffffffu, vy, o); vz += __shfl_xor_sync(0xffffffffu, vz, o); }
            if (lane == r) rp_out[r] = make_float4(vx, vy, vz, 0.f);
        }
    };

    for (int k = 0; k < n; k++) {
        const volatile LnTileRun::Step* sp = &((volatile LnTileRun*)run)->step[k & 1];
        const int cstage = (int)((LN_RUN(pos) + (unsigned)k) & 1u);
        if (sp->newrow) {
            const float4* X = LN_RUN(X); const float4* XB = LN_RUN(XB); const int N = LN_RUN(N), I = sp->I;
#pragma unroll
            for (int rp = 0; rp < 4; rp++) {
                int i0 = I * LN_TB + warp * 8 + 2 * rp, i1 = i0 + 1;
                if (i0 >= N) i0 = N - 1;
                if (i1 >= N) i1 = N - 1;
                const float4 p0 = __ldcg(X + i0), q0 = __ldcg(XB + i0), p1 = __ldcg(X + i1), q1 = __ldcg(XB + i1);
                xi[rp] = f2(fmaf(lam, q0.x, p0.x), fmaf(lam, q1.x, p1.x)); yi[rp] = f2(fmaf(lam, q0.y, p0.y), fmaf(lam, q1.y, p1.y));
                zi[rp] = f2(fmaf(lam, q0.z, p0.z), fmaf(lam, q1.z, p1.z)); fix[rp] = fiy[rp] = fiz[rp] = bc(0.f);
            }
        }
        if (k + 1 < n) {
            prefetch_cols(sp->Jn, cstage ^ 1);
            if (tid == 0) {           // the next tile's bookkeeping, one tile ahead
                int I = sp->I, J = sp->J + 1, nr = 0, seg = sp->seg;
                if (J == LN_RUN(nT)) { I++; J = I; nr = 1; seg++; }
                make_step(k + 1, I, J, seg, nr);
            }
        }
        const int stage = sp->stage;
        if (dbg != 2) mbar_wait(&bar[stage], (uint32_t)sp->tpar);
        const float4* D4 = reinterpret_cast<const float4*>(dbuf + (size_t)stage * LN_TILE_FLOATS) + (warp * 8) * 32 + lane;   // see ln_tile_cell
        // padding beads (index >= N) exist only in the last tile row / column; their cells are -2 like the diagonal band's
        const bool near = sp->near != 0;
        // the lane's four columns in two halves (columns 32*(2ch) + lane and 32*(2ch+1) + lane): only two columns' coordinates
        // and force accumulators are live at a time, which is what keeps the pair loop inside the register budget
#ifdef LN_UNROLL_CH
#pragma unroll
#else
#pragma unroll 1          // (unrolled, the two halves overlap a little better, but in the persistent kernel ptxas then spills accumulators)
#endif
        for (int ch = 0; ch < 2; ch++) {
            float xj[2], yj[2], zj[2];
            float2 gjx[2], gjy[2], gjz[2];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const float4 p = cbuf[(cstage * 2) * LN_TB + (2 * ch + c) * 32 + lane], q = cbuf[(cstage * 2 + 1) * LN_TB + (2 * ch + c) * 32 + lane];
                xj[c] = fmaf(lam, q.x, p.x); yj[c] = fmaf(lam, q.y, p.y); zj[c] = fmaf(lam, q.z, p.z);
                gjx[c] = gjy[c] = gjz[c] = bc(0.f);
            }
            if (dbg == 1) {
            } else if (near) {
#pragma unroll
                for (int rp = 0; rp < 4; rp++) {
                    const float4 d4 = D4[(2 * rp + ch) * 32];
                    pair_sym2<WANT_E, LOW_SOFT, true>(tm, xi[rp], yi[rp], zi[rp], xj[0], yj[0], zj[0], f2(d4.x, d4.y), fix[rp], fiy[rp], fiz[rp], gjx[0], gjy[0], gjz[0], en, ev);
                    pair_sym2<WANT_E, LOW_SOFT, true>(tm, xi[rp], yi[rp], zi[rp], xj[1], yj[1], zj[1], f2(d4.z, d4.w), fix[rp], fiy[rp], fiz[rp], gjx[1], gjy[1], gjz[1], en, ev);
                }
            } else {
#pragma unroll
                for (int rp = 0; rp < 4; rp++) {
                    const float4 d4 = D4[(2 * rp + ch) * 32];
                    pair_sym2<WANT_E, LOW_SOFT, false>(tm, xi[rp], yi[rp], zi[rp], xj[0], yj[0], zj[0], f2(d4.x, d4.y), fix[rp], fiy[rp], fiz[rp], gjx[0], gjy[0], gjz[0], en, ev);
                    pair_sym2<WANT_E, LOW_SOFT, false>(tm, xi[rp], yi[rp], zi[rp], xj[1], yj[1], zj[1], f2(d4.z, d4.w), fix[rp], fiy[rp], fiz[rp], gjx[1], gjy[1], gjz[1], en, ev);
                }
            }
            // column partials: [comp][warp][128], this lane's 4 consecutive positions; the single buffer is free once the reducer
            // warps of the previous tile have drained it (they arrived long ago: this wait does not block in practice)
            if (ch == 0 && k > 0) mbar_wait(&bar[LN_STAGES], (uint32_t)sp->cpar);
            *reinterpret_cast<float2*>(colp + (0 * 16 + warp) * LN_TB + lane * 4 + 2 * ch) = make_float2(-(gjx[0].x + gjx[0].y), -(gjx[1].x + gjx[1].y));
            *reinterpret_cast<float2*>(colp + (1 * 16 + warp) * LN_TB + lane * 4 + 2 * ch) = make_float2(-(gjy[0].x + gjy[0].y), -(gjy[1].x + gjy[1].y));
            *reinterpret_cast<float2*>(colp + (2 * 16 + warp) * LN_TB + lane * 4 + 2 * ch) = make_float2(-(gjz[0].x + gjz[0].y), -(gjz[1].x + gjz[1].y));
        }
        if (sp->last) flush_rows(sp->seg);          // (before the barrier: behind it thread 0 may already rewrite this bookkeeping slot)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                 // tile buffer `stage` is free, column partials are complete, next column coordinates and bookkeeping are in place
        if (tid == 0 && dbg != 2) {      // next position of the periodic tile sequence that wants this stage
            LN_RUN(pending) = LN_RUN(pending) & ~(1u << stage);
            const int nx = k + LN_STAGES;
            if (nx < n) ln_ring_issue(run, smem, nx, stage);
            else if (LN_RUN(more) && nx - n < LN_RUN(ring_tail)) ln_ring_issue(run, smem, nx % n, stage);
        }
        if (tid < 3 * LN_TB) {
            const int comp = tid / LN_TB, col = tid % LN_TB;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 16; w++) s += colp[(comp * 16 + w) * LN_TB + col];
            LN_RUN(colpart)[((size_t)k * 3 + comp) * LN_TB + col] = s;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bar[LN_STAGES])) : "memory");
        }
    }
    // the last tile's reducer arrivals complete a phase nobody waited for: consume it, so that the next evaluation starts clean
    if (n > 0) mbar_wait(&bar[LN_STAGES], (LN_RUN(pos) + (unsigned)n - 1u) & 1u);
    // block energy (the column-partial buffer is free now)
    if (WANT_E) {
        double den = (double)(en.x + en.y), dev = (double)(ev.x + ev.y);
        double* esh = reinterpret_cast<double*>(colp);
        for (int o = 16; o > 0; o >>= 1) { den += __shfl_xor_sync(0xffffffffu, den, o); dev += __shfl_xor_sync(0xffffffffu, dev, o); }
        __syncthreads();
        if (lane == 0) { esh[warp] = den; esh[16 + warp] = dev; }
        __syncthreads();
        if (tid == 0) {
            double s0 = 0, s1 = 0;
            for (int w = 0; w < 16; w++) { s0 += esh[w]; s1 += esh[16 + w]; }
            double* ep = LN_RUN(epart); ep[0] = s0; ep[1] = s1;
        }
    }
    __syncthreads();
    if (tid == 0) run->pos = run->pos + (unsigned)n;
    __syncthreads();
}

// fills the static part of a CTA's LnTileRun (one thread)
__device__ __forceinline__ void ln_run_setup(LnTileRun* run, const float* Dt, long long t0, const long long* cstart, long long nT, int N, float* colpart, float4* rowpart, double* epart,
                                             int maxseg, int cta) {
    const long long c0 = cstart[cta], c1 = cstart[cta + 1];
    run->Dt = Dt + (size_t)(c0 - t0) * LN_TILE_FLOATS; run->colpart = colpart + (size_t)(c0 - t0) * 3 * LN_TB;
    run->rowpart = rowpart + (size_t)cta * maxseg * LN_TB; run->epart = epart + 2 * cta;
    run->n = (int)(c1 - c0); run->nT = (int)nT; run->N = N;
    int I = 0, J = 0; if (c1 > c0) tile_coords(c0, nT, I, J);
    run->I0 = I; run->J0 = J;
}

template <bool WANT_E, bool LOW_SOFT>
__global__ void __launch_bounds__(LN_BOUND_THREADS, 1) ln_tile_kernel(LnTileArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ LnTileRun run;
    asm volatile("griddepcontrol.wait;" ::: "memory");       // programmatic dependent launch: the previous kernel's results are complete from here on
    if (a.done && *a.done) return;
    if (threadIdx.x == 0) {
        ln_run_setup(&run, a.Dt, a.t0, a.cstart, a.nT, a.N, a.colpart, a.rowpart, a.epart, a.maxseg, blockIdx.x);
        run.X = a.x; run.XB = a.xb2 ? a.xb2 : a.x; run.lam = a.xb2 ? *a.lam : 0.f; run.more = 0; run.dbg = LN_DBG(a); run.terms = a.terms; run.stream_l2 = a.stream_l2; run.ring_tail = LN_STAGES;
    }
    __syncthreads();
    ln_ring_init(&run, smem, false);
    ln_tile_phase<WANT_E, LOW_SOFT>(&run, smem, a.terms);
}

// ------------------------------------------------------------------------------------------------
// fixed-order reduction of the partial forces; rank 0 adds the bond term
// ------------------------------------------------------------------------------------------------
// partial forces of this rank -> force on bead (B*128 + tid): column partials of the tiles (I, B) this rank owns,
// then the row partials of the CTA segments that walked tile row B (static CSR list built on the host), fixed order
struct LnGather {
    const float* colpart; const float4* rowpart; const int* rs_off; const int* rs_idx;
    long long t0, t1, nT;
};
#define LN_GROUPS 4                     // the step kernels run LN_GROUPS x LN_TB threads: the column partials of a bead block
#define LN_STEP_THREADS (LN_GROUPS * LN_TB)   // are summed by LN_GROUPS thread groups (I = grp, grp + LN_GROUPS, ...) in parallel
__device__ __forceinline__ void ln_gather_part(const LnGather& g, int B, int tid, int grp, float& fx, float& fy, float& fz) {
    // the tiles (I, B), I = 0..B, have increasing list positions, so the ones this rank owns are a contiguous I range
    int lo = 0, hi = B + 1;
    while (lo < hi && tile_index(lo, B, g.nT) < g.t0) lo++;
    while (hi > lo && tile_index(hi - 1, B, g.nT) >= g.t1) hi--;
    const float* base = g.colpart + ((tid & 31) * 4 + (tid >> 5));      // column -> position 4*lane + c inside a partial
    float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f}, az[4] = {0.f, 0.f, 0.f, 0.f};
    int I = lo + grp;
    for (; I + 3 * LN_GROUPS < hi; I += 4 * LN_GROUPS) {               // four independent loads in flight per component
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float* cp = base + (size_t)(tile_index(I + u * LN_GROUPS, B, g.nT) - g.t0) * 3 * LN_TB;
            ax[u] += cp[0]; ay[u] += cp[LN_TB]; az[u] += cp[2 * LN_TB];
        }
    }
    for (; I < hi; I += LN_GROUPS) {
        const float* cp = base + (size_t)(tile_index(I, B, g.nT) - g.t0) * 3 * LN_TB;
        ax[0] += cp[0]; ay[0] += cp[LN_TB]; az[0] += cp[2 * LN_TB];
    }
    fx = (ax[0] + ax[1]) + (ax[2] + ax[3]); fy = (ay[0] + ay[1]) + (ay[2] + ay[3]); fz = (az[0] + az[1]) + (az[2] + az[3]);
    if (grp == 0) for (int k = g.rs_off[B]; k < g.rs_off[B + 1]; k++) {
        const float4 v = g.rowpart[(size_t)g.rs_idx[k] * LN_TB + tid];
        fx += v.x; fy += v.y; fz += v.z;
    }
}
// all LN_STEP_THREADS threads call; the result is valid in group 0 (fixed summation order: group 0, 1, 2, 3)
__device__ __forceinline__ void ln_gather(const LnGather& g, int B, int tid, int grp, float& fx, float& fy, float& fz) {
    __shared__ float gsh[LN_GROUPS - 1][3][LN_TB];
    ln_gather_part(g, B, tid, grp, fx, fy, fz);
    if (grp > 0) { gsh[grp - 1][0][tid] = fx; gsh[grp - 1][1][tid] = fy; gsh[grp - 1][2][tid] = fz; }
    __syncthreads();
    if (grp == 0) for (int q = 0; q < LN_GROUPS - 1; q++) { fx += gsh[q][0][tid]; fy += gsh[q][1][tid]; fz += gsh[q][2][tid]; }
}
// bond term of bead i at lazily defined coordinates x = a + lam*b
__device__ __forceinline__ float4 ln_ldx(const float4* a, const float4* b, float lam, int i) {
    const float4 p = a[i], q = b[i];
    return make_float4(fmaf(lam, q.x, p.x), fmaf(lam, q.y, p.y), fmaf(lam, q.z, p.z), 0.f);
}
// bond (+ angle stand-in) terms of bead i at coordinates a + lam*b; LD: how a float4 is loaded (plain, or around L1)
template <typename LD>
__device__ __forceinline__ void ln_bonded(LD ld, const float4* a, const float4* b, float lam, int i, int N, const LnBonded& bd,
                                          float& fx, float& fy, float& fz, float& eb) {
    const float4 pi = ld(a + i), qi = ld(b + i);
    const float xi = fmaf(lam, qi.x, pi.x), yi = fmaf(lam, qi.y, pi.y), zi = fmaf(lam, qi.z, pi.z);
#pragma unroll 1
    for (int s = -1; s <= 1; s += 2) {
        const int nb = i + s; if (nb < 0 || nb >= N) continue;
        const float4 pn = ld(a + nb), qn = ld(b + nb);
        const float dx = xi - fmaf(lam, qn.x, pn.x), dy = yi - fmaf(lam, qn.y, pn.y), dz = zi - fmaf(lam, qn.z, pn.z);
        const float r2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-12f), rinv = ln_rsqrt(r2), r = r2 * rinv, u = r - bd.r0;
        const float g = 2.f * bd.w_bond * u * rinv;
        fx -= g * dx; fy -= g * dy; fz -= g * dz;
        if (s == 1) eb = bd.w_bond * u * u;
    }
    if (bd.w_ang > 0.f)
#pragma unroll 1
    for (int s = -2; s <= 2; s += 4) {
        const int nb = i + s; if (nb < 0 || nb >= N) continue;
        const float4 pn = ld(a + nb), qn = ld(b + nb);
        const float dx = xi - fmaf(lam, qn.x, pn.x), dy = yi - fmaf(lam, qn.y, pn.y), dz = zi - fmaf(lam, qn.z, pn.z);
        const float r2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-12f), rinv = ln_rsqrt(r2), r = r2 * rinv, u = fmaxf(bd.d2 - r, 0.f);
        const float g = -2.f * bd.w_ang * u * rinv;          // dE/dr / r of w (d2 - r)^2
        fx -= g * dx; fy -= g * dy; fz -= g * dz;
        if (s == 2) eb += bd.w_ang * u * u;
    }
}
struct LnLdPlain { __device__ __forceinline__ float4 operator()(const float4* p) const { return *p; } };
struct LnLdCg { __device__ __forceinline__ float4 operator()(const float4* p) const { return __ldcg(p); } };
__device__ __forceinline__ void ln_bond(const float4* a, const float4* b, float lam, int i, int N, const LnBonded& bd, float& fx, float& fy, float& fz, float& eb) {
    ln_bonded(LnLdPlain(), a, b, lam, i, N, bd, fx, fy, fz, eb);
}
__device__ __forceinline__ double ln_block_sum(double v, double* sh, int tid, bool lead) {   // sum over the LN_TB lead threads; every thread of the block calls
    __syncthreads();
    if (lead) sh[tid] = v;
    __syncthreads();
    for (int o = LN_TB / 2; o > 0; o >>= 1) { if (lead && tid < o) sh[tid] += sh[tid + o]; __syncthreads(); }
    return sh[0];
}

struct LnReduceArgs {
    LnGather g; const double* epart;
    int N, ncta, add_bond;
    const float4* x; const float4* xb2; const float* lam;     // coordinates x + lam*xb2 (xb2 may be nullptr)
    float4* f;                           // f has N + 1 entries; entry N is unused padding
    double* escal;                       // [3] e_noe, e_vdw, e_bond of this rank
    LnBonded bd;
};
__global__ void __launch_bounds__(LN_STEP_THREADS) ln_reduce_kernel(LnReduceArgs a) {
    const int B = blockIdx.x, tid = threadIdx.x & (LN_TB - 1), grp = threadIdx.x >> 7, i = B * LN_TB + tid;
    const bool lead = grp == 0;
    float fx, fy, fz, eb = 0.f;
    ln_gather(a.g, B, tid, grp, fx, fy, fz);
    if (lead && a.add_bond && i < a.N) ln_bond(a.x, a.xb2 ? a.xb2 : a.x, a.xb2 ? *a.lam : 0.f, i, a.N, a.bd, fx, fy, fz, eb);
    if (lead && i < a.N) a.f[i] = make_float4(fx, fy, fz, 0.f);
    __shared__ double sh[LN_TB];
    const double ebs = ln_block_sum((double)eb, sh, tid, lead);
    if (threadIdx.x == 0) a.escal[3 + B] = ebs;          // per-block bond partials follow the 3 totals
}
// totals of this rank: e_noe, e_vdw (tile kernel partials), e_bond (reduce kernel partials), fixed order
__global__ void ln_energy_kernel(const double* epart, int ncta, double* escal, int nblk) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double e0 = 0, e1 = 0, e2 = 0;
    for (int c = 0; c < ncta; c++) { e0 += epart[2 * c]; e1 += epart[2 * c + 1]; }
    for (int b = 0; b < nblk; b++) e2 += escal[3 + b];
    escal[0] = e0; escal[1] = e1; escal[2] = e2;
}

// ------------------------------------------------------------------------------------------------
// device-resident control state and the vector kernels of MD / L-BFGS
// ------------------------------------------------------------------------------------------------
struct LnState {
    double E, Enoe, Evdw, Ebond, gg, gp, tstep, ke, lam;
    double B[LN_MAXB * LN_MAXB]; double delta[LN_MAXB];
    double e_saved;
    int cnt, head, it, ls, done, accepted, keep, have_E, hand, cur_image, nsteps, fresh;
    long long nevals;
    float lamf, tstepf, zerof, pad_;      // device scalars read by the tile kernel: x = a + lamf*b (MD), x = xb + tstepf*p (L-BFGS)
    unsigned int revals, pad3_;           // evaluations since the tile plan was made (targets of the per-block completion counters)
    unsigned int counter, epoch;          // "last block done" counter of the legacy step kernels; evaluation counter of the peer exchange
    double eloc[2];                       // pair energies of this rank (persistent kernel)
    unsigned long long tprof[8];          // phase clock of the persistent kernel (ns, CTA 0): tile, barrier 1, step, barrier 2, update
};

struct LnVec {     // all float4[N]
    float4 *x, *xb, *g, *p, *v, *f, *S, *Y;   // S, Y: m * N
    int N, m;
};

__device__ __forceinline__ float dot3(const float4& a, const float4& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }

// layout of the inner products of (s, y, g') with the L-BFGS basis
#define LDOT_L(l, w) (6 * (l) + (w))
#define LDOT_X(m, w) (6 * (m) + (w))
// vector-free L-BFGS direction on the Gram matrix (same recursion as sa_ensemble.cu)
__device__ double ln_direction(double* B, double* delta, int m, int cnt, int head) {
    const int nb = 2 * m + 1, ig = 2 * m;
    double al[CHR3D_LBFGS_MAX_M];
    for (int l = 0; l < nb; l++) delta[l] = 0.0;
    delta[ig] = 1.0;
    for (int k = 0; k < cnt; k++) {
        int idx = ((head - 1 - k) % m + m) % m;
        double s = delta[ig] * B[ig * LN_MAXB + idx];
        for (int kk = 0; kk < cnt; kk++) { int l = ((head - 1 - kk) % m + m) % m;
            s += delta[l] * B[l * LN_MAXB + idx] + delta[m + l] * B[(m + l) * LN_MAXB + idx]; }
        al[idx] = s / B[idx * LN_MAXB + m + idx];
        delta[m + idx] -= al[idx];
    }
    int nw = ((head - 1) % m + m) % m;
    double gam = B[nw * LN_MAXB + m + nw] / B[(m + nw) * LN_MAXB + m + nw];
    for (int l = 0; l < nb; l++) delta[l] *= gam;
    for (int k = cnt - 1; k >= 0; k--) {
        int idx = ((head - 1 - k) % m + m) % m;
        double s = delta[ig] * B[ig * LN_MAXB + m + idx];
        for (int kk = 0; kk < cnt; kk++) { int l = ((head - 1 - kk) % m + m) % m;
            s += delta[l] * B[l * LN_MAXB + m + idx] + delta[m + l] * B[(m + l) * LN_MAXB + m + idx]; }
        double be = s / B[idx * LN_MAXB + m + idx];
        delta[idx] += al[idx] - be;
    }
    double gp = delta[ig] * B[ig * LN_MAXB + ig];
    for (int kk = 0; kk < cnt; kk++) { int l = ((head - 1 - kk) % m + m) % m;
        gp += delta[l] * B[l * LN_MAXB + ig] + delta[m + l] * B[(m + l) * LN_MAXB + ig]; }
    return -gp;
}

struct LnMinCfg { int m, max_ls, nstep, nparts; double c1, gtol, ftol, drop; int N; };

// one thread: sums the partial scalars in a fixed order and takes the line-search / history decision
// (G: the 6m+6 inner products, escal_all: e_noe, e_vdw, e_bond of the evaluation just done)
__device__ void ln_min_decide(LnState* st, const double* G, const double* escal_all, LnMinCfg c, int init) {
    const int m = c.m;
    const double Et = escal_all[0] + escal_all[1] + escal_all[2];
    st->nevals++;
    const double n3 = 3.0 * c.N;
    if (init) {             // first evaluation of a MINIMIZE op: x = xb, g' = g
        st->E = Et; st->Enoe = escal_all[0]; st->Evdw = escal_all[1]; st->Ebond = escal_all[2];
        st->gg = G[LDOT_X(m, 5)]; st->cnt = 0; st->head = 0; st->it = 0; st->ls = 0; st->accepted = 2;   // 2: adopt g, steepest descent
        st->gp = -st->gg;
        if (sqrt(st->gg / n3) < c.gtol || c.nstep <= 0) { st->done = 1; return; }
        st->tstep = c.drop > 0 ? fmin(1.0, c.drop / st->gg) : fmin(1.0, 1.0 / sqrt(st->gg));
        return;
    }
    st->accepted = 0;
    const double ts = st->tstep;
    if (Et < st->E && Et <= st->E + c.c1 * ts * st->gp) {
        // accept
        const double ss = G[LDOT_X(m, 0)], sy = G[LDOT_X(m, 1)], yy = G[LDOT_X(m, 3)];
        const bool keep = sy > 1e-10 * sqrt(ss * yy);
        const bool small = st->cnt > 0 && (st->E - Et) <= c.ftol * fmax(fabs(Et), 1.0);
        st->E = Et; st->Enoe = escal_all[0]; st->Evdw = escal_all[1]; st->Ebond = escal_all[2];
        st->gg = G[LDOT_X(m, 5)];
        st->it++; st->ls = 0;
        st->keep = keep ? 1 : 0; st->accepted = 1;
        if (keep) {
            double* B = st->B; const int ig = 2 * m, h = st->head;
            for (int l = 0; l < m; l++) {
                if (l == h) continue;
                B[h * LN_MAXB + l] = B[l * LN_MAXB + h] = G[LDOT_L(l, 0)];
                B[h * LN_MAXB + m + l] = B[(m + l) * LN_MAXB + h] = G[LDOT_L(l, 1)];
                B[(m + h) * LN_MAXB + l] = B[l * LN_MAXB + m + h] = G[LDOT_L(l, 2)];
                B[(m + h) * LN_MAXB + m + l] = B[(m + l) * LN_MAXB + m + h] = G[LDOT_L(l, 3)];
                B[ig * LN_MAXB + l] = B[l * LN_MAXB + ig] = G[LDOT_L(l, 4)];
                B[ig * LN_MAXB + m + l] = B[(m + l) * LN_MAXB + ig] = G[LDOT_L(l, 5)];
            }
            B[h * LN_MAXB + h] = ss; B[h * LN_MAXB + m + h] = B[(m + h) * LN_MAXB + h] = sy; B[(m + h) * LN_MAXB + m + h] = yy;
            B[ig * LN_MAXB + h] = B[h * LN_MAXB + ig] = G[LDOT_X(m, 2)];
            B[ig * LN_MAXB + m + h] = B[(m + h) * LN_MAXB + ig] = G[LDOT_X(m, 4)];
            B[ig * LN_MAXB + ig] = st->gg;
            st->head = (h + 1) % m; if (st->cnt < m) st->cnt++;
            st->gp = ln_direction(st->B, st->delta, m, st->cnt, st->head);
            st->tstep = 1.0;
        } else {
            st->cnt = 0;
        }
        if (small || st->it >= c.nstep || sqrt(st->gg / n3) < c.gtol) { st->done = 1; return; }
        if (st->cnt == 0 || !(st->gp < 0.0)) {      // steepest descent restart
            st->cnt = 0; st->gp = -st->gg;
            st->tstep = c.drop > 0 ? fmin(1.0, c.drop / st->gg) : fmin(1.0, 1.0 / sqrt(st->gg));
            st->accepted = 3;                         // 3: accepted, next direction = -g
        }
    } else {
        st->ls++;
        if (st->ls > c.max_ls) {
            st->it++; st->ls = 0;
            if (st->cnt == 0 || st->it >= c.nstep) { st->done = 1; st->accepted = 4; return; }   // 4: restore xb
            st->cnt = 0; st->gp = -st->gg;
            st->tstep = c.drop > 0 ? fmin(1.0, c.drop / st->gg) : fmin(1.0, 1.0 / sqrt(st->gg));
            st->accepted = 5;                         // 5: rejected, restart from -g at xb
        } else {
            st->tstep = ts * 0.5;
        }
    }
}
// applies the decision to the vectors
__global__ void ln_min_update_kernel(LnVec v, const LnState* st) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int acc = st->accepted;
    if (acc == 0) return;
    const int m = v.m, cnt = st->cnt, head = st->head;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.N; i += gridDim.x * blockDim.x) {
        if (acc == 4) { v.x[i] = v.xb[i]; continue; }
        if (acc == 5) { float4 g = v.g[i]; v.p[i] = make_float4(-g.x, -g.y, -g.z, 0.f); continue; }
        float4 x = v.x[i], f = v.f[i];
        float4 gn = make_float4(-f.x, -f.y, -f.z, 0.f);
        if (acc == 1 && st->keep) {
            const int h = ((head - 1) % m + m) % m;      // slot just filled
            float4 xb = v.xb[i], g = v.g[i];
            v.S[(size_t)h * v.N + i] = make_float4(x.x - xb.x, x.y - xb.y, x.z - xb.z, 0.f);
            v.Y[(size_t)h * v.N + i] = make_float4(gn.x - g.x, gn.y - g.y, gn.z - g.z, 0.f);
        }
        v.xb[i] = x; v.g[i] = gn;
        if (acc == 2 || acc == 3 || cnt == 0) { v.p[i] = make_float4(-gn.x, -gn.y, -gn.z, 0.f); continue; }
        float ax = (float)st->delta[2 * m] * gn.x, ay = (float)st->delta[2 * m] * gn.y, az = (float)st->delta[2 * m] * gn.z;
        for (int kk = 0; kk < cnt; kk++) {
            int l = ((head - 1 - kk) % m + m) % m;
            float ds = (float)st->delta[l], dy = (float)st->delta[m + l];
            float4 sl = v.S[(size_t)l * v.N + i], yl = v.Y[(size_t)l * v.N + i];
            ax = fmaf(ds, sl.x, fmaf(dy, yl.x, ax)); ay = fmaf(ds, sl.y, fmaf(dy, yl.y, ay)); az = fmaf(ds, sl.z, fmaf(dy, yl.z, az));
        }
        v.p[i] = make_float4(-ax, -ay, -az, 0.f);
    }
}

// ---- MD --------------------------------------------------------------------------------------
struct LnMdCfg { float dt_akma, h; double dt_ps, temp, fbeta, kboltz, mass; int flag, N, nparts, count; };
// v += h f (skipped for the initial evaluation), kinetic-energy partials
__global__ void __launch_bounds__(256) ln_md_post_kernel(LnVec v, LnMdCfg c, int kick, double* part) {
    float k2 = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.N; i += gridDim.x * blockDim.x) {
        float4 vv = v.v[i];
        if (kick) { float4 f = v.f[i]; vv.x = fmaf(c.h, f.x, vv.x); vv.y = fmaf(c.h, f.y, vv.y); vv.z = fmaf(c.h, f.z, vv.z); v.v[i] = vv; }
        k2 += dot3(vv, vv);
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = (double)k2; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * LN_KMAX] = sh[0];
}
__global__ void ln_md_control_kernel(LnState* st, const double* part, const double* escal, LnMdCfg c) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double k2 = 0; for (int b = 0; b < c.nparts; b++) k2 += part[(size_t)b * LN_KMAX];
    const double ke = 0.5 * c.mass * k2, ndof = 3.0 * c.N - 3.0;
    const double T = 2.0 * ke / (c.kboltz * ndof);
    double lam = 1.0;
    if (T > 0.0) {
        if (c.flag == CHR3D_COUPLE_TCOUP) lam = sqrt(fmax(0.0, 1.0 + 2.0 * c.fbeta * c.dt_ps * (c.temp / T - 1.0)));
        else if (c.flag == CHR3D_COUPLE_VSCALE) lam = sqrt(c.temp / T);
    }
    st->ke = ke; st->lam = lam; st->nevals += c.count;
    st->E = escal[0] + escal[1] + escal[2]; st->Enoe = escal[0]; st->Evdw = escal[1]; st->Ebond = escal[2];
}

// ------------------------------------------------------------------------------------------------
// Fused per-evaluation kernels.  A force evaluation is two launches: ln_tile_kernel (pair tiles, lazy
// coordinates) and one of the step kernels below, which gathers the partial forces of a bead block, adds
// the bond term, advances the integrator / collects the L-BFGS inner products, and lets the last block
// to finish ("last block done" counter) take the scalar decisions, so that no host round trip and no
// extra launch is needed.  With several ranks the gather is preceded by the force exchange (see ln_eval).
// ------------------------------------------------------------------------------------------------
struct LnEnergySrc {        // where the energies of the evaluation come from
    const double* epart; int ncta;      // local: per-CTA (e_noe, e_vdw) of the tile kernel
    const double* escal;                // != nullptr: already summed over ranks [e_noe, e_vdw, e_bond]
};
// true in exactly one block: the last one to arrive; it sees every other block's writes
__device__ __forceinline__ bool ln_last_block(unsigned int* counter, int nblk) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) { const unsigned int t = atomicAdd(counter, 1u); is_last = (t == (unsigned)nblk - 1u); }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last != 0;
}


// ---- force exchange over NVLink peer memory, fused into the persistent kernel (ln_persist.cuh) --------------------
// Every rank owns a communication buffer (cudaMalloc + CUDA IPC) that its peers map:
//   recv [2 parities][world][N + 1] float4   slot g = partial forces of rank g (no bond term); .w carries the EPOCH of the
//                                            evaluation the value belongs to
//   es   [2 parities][world][2] {double, epoch}   pair energies of rank g
// The CTA that handles bead block B gathers this rank's partial force on its 128 beads and PUSHES it into slot `rank` of every
// rank's recv as 16-byte remote stores (fire-and-forget: no fence, no separate flag — a 16-byte aligned store arrives as one
// piece, so the epoch in .w says that x, y, z next to it are the new ones; the same assumption NCCL's LL/LL128 protocols make
// on NVLink).  The receiver polls the slots of the contributing ranks until they show the current epoch and adds them in rank
// order (identical on every rank, so the redundant integration stays bit-identical).  Parity double-buffering makes overwriting
// safe: a rank cannot run two evaluations ahead of a peer, because each evaluation needs that peer's current data.
#define LN_MAXWORLD 8
struct LnXchg {
    int world, rank, nblk, n4;                 // n4 = N + 1
    float4* recv[LN_MAXWORLD]; double2* es[LN_MAXWORLD];      // es element: {value, epoch (as the bits of the second double)}
    int* err;                                  // host-visible word, set when a wait times out (a peer died): the host reports it
};
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct LnMdStepArgs {
    LnGather g; const float4* fsum;     // fsum != nullptr: total forces (bond included) already in fsum (NCCL path)
    const float4* a; const float4* b; const float* lam;     // coordinates of the evaluation just done: a + lam*b
    float4 *x, *v, *f, *a_next, *b_next;
    LnEnergySrc es;
    LnState* st; double* blockpart; int nblk, N;
    LnBonded bd; float h, dt_akma; int first, want_e;
    LnMdCfg md;
};
// Verlet step (definition of the oracle's sim_dynamics):  x = a + lam*b are the coordinates the forces were just
// computed at; v1 = lam*v + h*f_old is the velocity that led there; v2 = v1 + h*f closes the step.  Published for
// the next step: a' = x + dt*h*f, b' = dt*v2; the last block turns the kinetic energy of v2 into lam'.
__global__ void __launch_bounds__(LN_STEP_THREADS) ln_md_step_kernel(LnMdStepArgs p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int B = blockIdx.x, tid = threadIdx.x & (LN_TB - 1), grp = threadIdx.x >> 7, i = B * LN_TB + tid;
    const bool lead = grp == 0;
    __shared__ double sh[LN_TB];
    const float lam = *p.lam;
    float fx = 0.f, fy = 0.f, fz = 0.f, eb = 0.f, k2 = 0.f;
    if (p.fsum) { if (lead && i < p.N) { const float4 t = p.fsum[i]; fx = t.x; fy = t.y; fz = t.z; } }
    else {
        ln_gather(p.g, B, tid, grp, fx, fy, fz);
        if (lead && i < p.N) ln_bond(p.a, p.b, lam, i, p.N, p.bd, fx, fy, fz, eb);
    }
    if (lead && i < p.N) {
        const float4 xi = ln_ldx(p.a, p.b, lam, i);
        float4 vv = p.v[i];
        if (!p.first) {
            const float4 fo = p.f[i];
            vv.x = fmaf(p.h, fo.x, lam * vv.x); vv.y = fmaf(p.h, fo.y, lam * vv.y); vv.z = fmaf(p.h, fo.z, lam * vv.z);   // v1
            vv.x = fmaf(p.h, fx, vv.x); vv.y = fmaf(p.h, fy, vv.y); vv.z = fmaf(p.h, fz, vv.z);                          // v2
        }
        k2 = dot3(vv, vv);
        const float dth = p.dt_akma * p.h;
        p.x[i] = xi; p.v[i] = vv; p.f[i] = make_float4(fx, fy, fz, 0.f);
        p.a_next[i] = make_float4(fmaf(dth, fx, xi.x), fmaf(dth, fy, xi.y), fmaf(dth, fz, xi.z), 0.f);
        p.b_next[i] = make_float4(p.dt_akma * vv.x, p.dt_akma * vv.y, p.dt_akma * vv.z, 0.f);
    }
    const double kb = ln_block_sum((double)k2, sh, tid, lead), ebb = ln_block_sum((double)eb, sh, tid, lead);
    if (threadIdx.x == 0) { p.blockpart[(size_t)B * LN_KMAX] = kb; p.blockpart[(size_t)B * LN_KMAX + 1] = ebb; }
    if (!ln_last_block(&p.st->counter, p.nblk)) return;
    // the last block sums the per-block partials with all its threads (fixed tree order)
    __shared__ double red[4][LN_STEP_THREADS];
    {
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int q = threadIdx.x; q < p.nblk; q += LN_STEP_THREADS) { a0 += p.blockpart[(size_t)q * LN_KMAX]; a1 += p.blockpart[(size_t)q * LN_KMAX + 1]; }
        if (p.want_e && !p.es.escal) for (int q = threadIdx.x; q < p.es.ncta; q += LN_STEP_THREADS) { a2 += p.es.epart[2 * q]; a3 += p.es.epart[2 * q + 1]; }
        red[0][threadIdx.x] = a0; red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2; red[3][threadIdx.x] = a3;
        __syncthreads();
        for (int o = LN_STEP_THREADS / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) for (int c = 0; c < 4; c++) red[c][threadIdx.x] += red[c][threadIdx.x + o];
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        LnState* st = p.st;
        double k2s = red[0][0], e2 = red[1][0];
        const double ke = 0.5 * p.md.mass * k2s, ndof = 3.0 * p.N - 3.0;
        const double T = 2.0 * ke / (p.md.kboltz * ndof);
        double l2 = 1.0;
        if (T > 0.0) {
            if (p.md.flag == CHR3D_COUPLE_TCOUP) l2 = sqrt(fmax(0.0, 1.0 + 2.0 * p.md.fbeta * p.md.dt_ps * (p.md.temp / T - 1.0)));
            else if (p.md.flag == CHR3D_COUPLE_VSCALE) l2 = sqrt(p.md.temp / T);
        }
        st->ke = ke; st->lam = l2; st->lamf = (float)l2; st->nevals += p.md.count;
        if (p.want_e) {
            double e0 = 0, e1 = 0;
            if (p.es.escal) { e0 = p.es.escal[0]; e1 = p.es.escal[1]; e2 = p.es.escal[2]; }
            else {
                e0 = red[2][0]; e1 = red[3][0];
            }
            st->Enoe = e0; st->Evdw = e1; st->Ebond = e2; st->E = e0 + e1 + e2;
        }
        st->counter = 0;
    }
}

struct LnMinStepArgs {
    LnGather g; const float4* fsum;
    LnVec v; const float* lam; int init;         // init: evaluation at the starting point (x = v.x, nothing pending)
    LnEnergySrc es;
    LnState* st; double* blockpart; int nblk;
    LnBonded bd;
    LnMinCfg cfg;
};
// L-BFGS trial x = xb + t*p (t = *lam): forces, the 6m+6 inner products of (s, y, g') with the history, and in the
// last block the line-search / history decision of ln_min_decide.  ln_min_update_kernel then applies it.
__global__ void __launch_bounds__(LN_STEP_THREADS) ln_min_step_kernel(LnMinStepArgs p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (p.st->done) { if (blockIdx.x == 0 && threadIdx.x == 0) p.st->accepted = 0; return; }   // later update launches are no-ops
    const int B = blockIdx.x, tid = threadIdx.x & (LN_TB - 1), grp = threadIdx.x >> 7, i = B * LN_TB + tid, lane = tid & 31, warp = tid >> 5;
    const bool lead = grp == 0;
    const int m = p.v.m, K = 6 * m + 6, N = p.v.N, head = p.st->head;
    __shared__ double sh[LN_TB];
    __shared__ float shk[LN_TB / 32][LN_KMAX];
    __shared__ double G[LN_KMAX];
    const float lam = p.init ? 0.f : *p.lam;
    const float4* A = p.init ? p.v.x : p.v.xb;
    float fx = 0.f, fy = 0.f, fz = 0.f, eb = 0.f;
    if (p.fsum) { if (lead && i < N) { const float4 t = p.fsum[i]; fx = t.x; fy = t.y; fz = t.z; } }
    else {
        ln_gather(p.g, B, tid, grp, fx, fy, fz);
        if (lead && i < N) ln_bond(A, p.v.p, lam, i, N, p.bd, fx, fy, fz, eb);
    }
    float acc[LN_KMAX];
#pragma unroll 1
    for (int k = 0; k < K; k++) acc[k] = 0.f;
    if (lead && i < N) {
        const float4 x = ln_ldx(A, p.v.p, lam, i), xb = p.v.xb[i], g = p.v.g[i];
        const float4 s = make_float4(x.x - xb.x, x.y - xb.y, x.z - xb.z, 0.f);
        const float4 gn = make_float4(-fx, -fy, -fz, 0.f);
        const float4 y = make_float4(gn.x - g.x, gn.y - g.y, gn.z - g.z, 0.f);
        p.v.x[i] = x; p.v.f[i] = make_float4(fx, fy, fz, 0.f);
        for (int l = 0; l < m; l++) {
            if (l == head) continue;
            const float4 sl = p.v.S[(size_t)l * N + i], yl = p.v.Y[(size_t)l * N + i];
            acc[LDOT_L(l, 0)] = dot3(s, sl); acc[LDOT_L(l, 1)] = dot3(s, yl); acc[LDOT_L(l, 2)] = dot3(y, sl);
            acc[LDOT_L(l, 3)] = dot3(y, yl); acc[LDOT_L(l, 4)] = dot3(gn, sl); acc[LDOT_L(l, 5)] = dot3(gn, yl);
        }
        acc[LDOT_X(m, 0)] = dot3(s, s); acc[LDOT_X(m, 1)] = dot3(s, y); acc[LDOT_X(m, 2)] = dot3(s, gn);
        acc[LDOT_X(m, 3)] = dot3(y, y); acc[LDOT_X(m, 4)] = dot3(y, gn); acc[LDOT_X(m, 5)] = dot3(gn, gn);
    }
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        float s = acc[k];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lead && lane == 0) shk[warp][k] = s;
    }
    const double ebb = ln_block_sum((double)eb, sh, tid, lead);  // (its barriers also publish shk)
    if (lead && tid < K) {
        double s = 0;
        for (int w = 0; w < LN_TB / 32; w++) s += (double)shk[w][tid];
        p.blockpart[(size_t)B * LN_KMAX + tid] = s;
    }
    if (threadIdx.x == 0) p.blockpart[(size_t)B * LN_KMAX + K] = ebb;
    if (!ln_last_block(&p.st->counter, p.nblk)) return;
    {   // per-block partials -> totals with all threads of the last block: 8 slices of blocks per scalar, fixed order
        __shared__ double red2[LN_STEP_THREADS / LN_KMAX][LN_KMAX];
        const int k = threadIdx.x & (LN_KMAX - 1), sl = threadIdx.x / LN_KMAX;
        double s = 0;
        if (k <= K) for (int q = sl; q < p.nblk; q += LN_STEP_THREADS / LN_KMAX) s += p.blockpart[(size_t)q * LN_KMAX + k];
        red2[sl][k] = s;
        __syncthreads();
        if ((int)threadIdx.x <= K) { double t = 0; for (int q = 0; q < LN_STEP_THREADS / LN_KMAX; q++) t += red2[q][threadIdx.x]; G[threadIdx.x] = t; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double e3[3] = {0, 0, G[K]};
        if (p.es.escal) { e3[0] = p.es.escal[0]; e3[1] = p.es.escal[1]; e3[2] = p.es.escal[2]; }
        else {
            for (int q = 0; q < p.es.ncta; q++) { e3[0] += p.es.epart[2 * q]; e3[1] += p.es.epart[2 * q + 1]; }
        }
        ln_min_decide(p.st, G, e3, p.cfg, p.init);
        p.st->tstepf = (float)p.st->tstep;
        p.st->counter = 0;
    }
}

#include "ln_persist.cuh"
#include "ln_sparse.cuh"

// sums of a float4 array (COM velocity / geometric centre): partials then subtraction
__global__ void __launch_bounds__(256) ln_sum3_kernel(const float4* a, int N, double* part) {
    double sx = 0, sy = 0, sz = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) { float4 q = a[i]; sx += q.x; sy += q.y; sz += q.z; }
    __shared__ double sh[3][256];
    sh[0][threadIdx.x] = sx; sh[1][threadIdx.x] = sy; sh[2][threadIdx.x] = sz; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) for (int q = 0; q < 3; q++) sh[q][threadIdx.x] += sh[q][threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) for (int q = 0; q < 3; q++) part[(size_t)blockIdx.x * LN_KMAX + q] = sh[q][0];
}
__global__ void ln_sub_mean_kernel(float4* a, int N, const double* part, int nparts) {
    __shared__ float mean[3];
    if (threadIdx.x == 0) for (int q = 0; q < 3; q++) { double s = 0; for (int b = 0; b < nparts; b++) s += part[(size_t)b * LN_KMAX + q]; mean[q] = (float)(s / N); }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) { float4 q = a[i]; q.x -= mean[0]; q.y -= mean[1]; q.z -= mean[2]; a[i] = q; }
}
__global__ void ln_init_kernel(LnVec v, const float* x0, int flag) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.N; i += gridDim.x * blockDim.x) {
        v.x[i] = make_float4(x0[3 * i] * (float)flag, x0[3 * i + 1], x0[3 * i + 2], 0.f);
        v.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
__global__ void ln_copy_xb_kernel(LnVec v) {    // xb = x (start of a MINIMIZE op)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.N; i += gridDim.x * blockDim.x) v.xb[i] = v.x[i];
}
__global__ void ln_state_reset_kernel(LnState* st, int full) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    st->done = 0; st->accepted = 0; st->ls = 0; st->it = 0; st->cnt = 0; st->head = 0; st->lam = 1.0;
    if (full) { st->nevals = 0; st->hand = 1; st->cur_image = 1; st->e_saved = 0; }
}

// Philox (same definition as the oracle / sa_ensemble.cu)
__device__ inline void ln_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__global__ void ln_maxwell_kernel(LnVec v, unsigned long long seed, uint32_t stream, uint32_t model, double sd) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.N; i += gridDim.x * blockDim.x) {
        uint32_t r[4]; ln_philox((uint32_t)i, stream, model, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double s = 1.0 / 4294967296.0, twopi = 6.283185307179586476925286766559;
        double u1 = (r[0] + 0.5) * s, u2 = (r[1] + 0.5) * s, u3 = (r[2] + 0.5) * s, u4 = (r[3] + 0.5) * s;
        double ra = sqrt(-2.0 * log(u1)), rb = sqrt(-2.0 * log(u3));
        v.v[i] = make_float4((float)(sd * ra * cos(twopi * u2)), (float)(sd * ra * sin(twopi * u2)), (float)(sd * rb * cos(twopi * u4)), 0.f);
    }
}
__global__ void ln_save_kernel(LnVec v, float4* sx, float4* sv, LnState* st) {
    if (blockIdx.x == 0 && threadIdx.x == 0) st->e_saved = st->E;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.N; i += gridDim.x * blockDim.x) { sx[i] = v.x[i]; sv[i] = v.v[i]; }
}
__global__ void ln_pick_kernel(LnVec v, const float4* sx, const float4* sv, LnState* st) {
    const bool keep_cur = st->e_saved > st->E;          // C3D:1717-1721
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < v.N; i += gridDim.x * blockDim.x) if (!keep_cur) { v.x[i] = sx[i]; v.v[i] = sv[i]; }
}
__global__ void ln_pick_state_kernel(LnState* st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    st->hand = st->e_saved > st->E ? st->cur_image : 1;
}
__global__ void ln_set_image_kernel(LnState* st, int image) { if (threadIdx.x == 0 && blockIdx.x == 0) { st->cur_image = image; st->hand = image; } }

// ------------------------------------------------------------------------------------------------
// target construction (large L: no certification, blocked sum order — beyond the reference's cap)
// ------------------------------------------------------------------------------------------------
__global__ void ln_row_pow_sum_kernel(const double* __restrict__ IF, int L, double alpha, double* __restrict__ rows) {
    // one warp per row, lane-strided partial sums combined in a fixed order
    int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x & 31;
    if (row >= L) return;
    double s = 0.0;
    for (int j = lane; j < L; j += 32) s += pow(IF[(size_t)row * L + j], alpha);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) rows[row] = s;
}
__global__ void ln_sum_rows_kernel(const double* rows, int L, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0; for (int i = 0; i < L; i++) s += rows[i];
    *out = s / ((double)L * (double)L);
}
__device__ __forceinline__ long long ln_round_tenths(double D) {
    if (D >= 4.0e15) return (long long)D * 10;
    double n = floor(D * 10.0);
    if (fma(D, 10.0, -n) < 0.0) n -= 1.0; else if (fma(D, 10.0, -(n + 1.0)) >= 0.0) n += 1.0;
    double r = fma(D, 10.0, -(n + 0.5));
    long long ni = (long long)n;
    return r > 0.0 ? ni + 1 : (r < 0.0 ? ni : ((ni & 1) ? ni + 1 : ni));
}
// tile-packed fp32 targets straight from IF (C3D:134-152 per cell, "%.1f" rounding, dist2rr filters)
__global__ void __launch_bounds__(256) ln_build_tiles_kernel(const double* __restrict__ IF, int L, double alpha, double K, const double* mean_p,
                                                              int min_sep, int vdw_min_sep, long long t0, long long t1, long long nT, float* __restrict__ Dt) {
    const long long t = t0 + blockIdx.x;
    if (t >= t1) return;
    int I, J; tile_coords(t, nT, I, J);
    const double mean = *mean_p;
    float* out = Dt + (size_t)(t - t0) * LN_TILE_FLOATS;
    for (int k = threadIdx.x; k < LN_TILE_FLOATS; k += blockDim.x) {
        int r, c; ln_tile_cell(k, r, c);
        const int i = I * LN_TB + r, j = J * LN_TB + c;
        float v = -2.f;
        if (i < L && j < L && j > i) {
            const int sep = j - i;
            v = sep >= vdw_min_sep ? -1.f : -2.f;
            if (sep >= min_sep) {
                double q = pow(IF[(size_t)i * L + j], alpha) / mean;
                if (q != 0.0) { long long tt = ln_round_tenths(K / q); if (tt > 0) v = (float)((double)tt / 10.0); }
            }
        }
        out[k] = v;
    }
}
// tile-packed targets from the certified `%.1f` tenths (L within the reference's range: bit-exact table)
__global__ void __launch_bounds__(256) ln_tiles_from_tenths_kernel(const long long* __restrict__ tenths, int L, int min_sep, int vdw_min_sep,
                                                                    long long t0, long long t1, long long nT, float* __restrict__ Dt) {
    const long long t = t0 + blockIdx.x;
    if (t >= t1) return;
    int I, J; tile_coords(t, nT, I, J);
    float* out = Dt + (size_t)(t - t0) * LN_TILE_FLOATS;
    for (int k = threadIdx.x; k < LN_TILE_FLOATS; k += blockDim.x) {
        int r, c; ln_tile_cell(k, r, c);
        const int i = I * LN_TB + r, j = J * LN_TB + c;
        float v = -2.f;
        if (i < L && j < L && j > i) {
            const int sep = j - i;
            v = sep >= vdw_min_sep ? -1.f : -2.f;
            const long long tt = tenths[(size_t)i * L + j];
            if (sep >= min_sep && tt > 0) v = (float)((double)tt / 10.0);
        }
        out[k] = v;
    }
}
// restraint rows held by this rank (cells with a positive target)
__global__ void __launch_bounds__(256) ln_count_rows_kernel(const float* __restrict__ Dt, long long n, unsigned long long* __restrict__ out) {
    unsigned long long c = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) c += Dt[k] > 0.f ? 1ull : 0ull;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}
// dense [N][pitch] fp32 targets (small-N layout) -> tile-packed; used by tests to cross-check both engines
__global__ void __launch_bounds__(256) ln_tiles_from_dense_kernel(const float* __restrict__ T, int N, int pitch, long long t0, long long t1, long long nT,
                                                                   float* __restrict__ Dt) {
    const long long t = t0 + blockIdx.x;
    if (t >= t1) return;
    int I, J; tile_coords(t, nT, I, J);
    float* out = Dt + (size_t)(t - t0) * LN_TILE_FLOATS;
    for (int k = threadIdx.x; k < LN_TILE_FLOATS; k += blockDim.x) {
        int r, c; ln_tile_cell(k, r, c);
        const int i = I * LN_TB + r, j = J * LN_TB + c;
        out[k] = (i < N && j < N && j > i) ? T[(size_t)i * pitch + j] : -2.f;
    }
}

// synthetic IF matrix of SURVEY §8(d) generated on the device (same definition as datasets.synthetic_if)
__global__ void ln_synth_if_kernel(double* __restrict__ IF, int N, unsigned long long seed, double A, double slope, double sigma, double zero_frac) {
    const long long n = (long long)N * N;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        int a = (int)(k / N), b = (int)(k % N);
        int i = a < b ? a : b, j = a < b ? b : a;
        unsigned long long lin = (unsigned long long)i * (unsigned long long)N + (unsigned long long)j;
        uint32_t r[4]; ln_philox((uint32_t)lin, (uint32_t)(lin >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        const double s = 1.0 / 4294967296.0, twopi = 6.283185307179586476925286766559;
        double u1 = (r[0] + 0.5) * s, u2 = (r[1] + 0.5) * s, u3 = (r[2] + 0.5) * s;
        double z = sqrt(-2.0 * log(u1)) * cos(twopi * u2);
        double v;
        if (i == j) v = 4.0 * A * exp(0.3 * z);
        else { v = A * pow((double)(j - i), slope) * exp(sigma * z); if (u3 < zero_frac) v = 0.0; }
        IF[k] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// host driver
// ------------------------------------------------------------------------------------------------
struct LnEngine;
// the engines one persistent launch serves: one (the normal case: one process per GPU), or the virtual ranks of a test group
struct LnGroup { std::vector<LnEngine*> es; DevBuf<LnCtx> ctxs; int grid = 0; bool dirty = true; };

struct LnEngine {
    int N = 0, nT = 0, rank = 0, world = 1, m = 5, ncta = 0, maxseg = 0, nblk = 0, nparts = 0, sms = 148, stream_l2 = 0;
    long long n_tiles = 0, t0 = 0, t1 = 0, pairs_total = 0; std::vector<long long> rank_cut;
    DevBuf<float> Dt, colpart, x0; DevBuf<float4> rowpart, vx, vxb, vg, vp, vv, vf, vS, vY, sx, sv, va[2], vb[2], vfr;
    DevBuf<int> rs_off, rs_idx, blk_off, blk_idx, tb_off, tb_idx, need; DevBuf<unsigned int> ready; DevBuf<double2> kslot; DevBuf<double> epart, escal, part, blockpart; DevBuf<LnState> st; DevBuf<unsigned int> bar;
    DevBuf<long long> cstart; DevBuf<unsigned long long> cta_ns, trace; std::vector<long long> h_cstart;
    chr3d_sa_params prm; unsigned long long seed = 0; int model = 0;     // model: Philox stream of the coil / Maxwell draws
    cudaStream_t stream = nullptr;
    long long n_launches = 0;
#ifdef CHR3D_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif
    // peer-memory exchange (world > 1): this rank's buffer, the mapped peer buffers
    void* xbuf = nullptr; size_t xbytes = 0; void* peer[LN_MAXWORLD] = {nullptr}; bool p2p = false, peers_local = false;
    int* xerr = nullptr;                       // pinned host word the kernel writes before it traps on a wait that cannot complete
    bool l2_window_set = false;
    // sparse force provider (ln_sparse.cuh): symmetric CSR of the restraint table, hashed cell grid for the repel term
    bool sparse = false; long long sp_nnz = 0; int sp_hbits = 0, sp_blocks = 0; float sp_inv_cs = 0.f;
    DevBuf<long long> sp_row_ptr, sp_nreb; DevBuf<int> sp_col, sp_cell_start, sp_cell_count, sp_cell_fill, sp_order, sp_rebuild; DevBuf<float> sp_dist;
    DevBuf<float4> sp_xs, sp_xref; DevBuf<double> sp_epart; DevBuf<unsigned int> sp_done; DevBuf<int> sp_bsum; DevBuf<int4> sp_ent;
    bool legacy = false;                       // multi-launch evaluation (NCCL all-reduce between ranks): the baseline path
    LnGroup self; LnGroup* group = nullptr;    // group: the launch group this engine currently belongs to (&self unless virtual ranks)
    static size_t up16(size_t v) { return (v + 15) / 16 * 16; }
    size_t x_off_es() const { return sizeof(float4) * 2 * (size_t)world * (N + 1); }
    size_t x_total() const { return x_off_es() + sizeof(double2) * 2 * 2 * (size_t)world; }
    LnVec vec() { LnVec v; v.x = vx.p; v.xb = vxb.p; v.g = vg.p; v.p = vp.p; v.v = vv.p; v.f = vf.p; v.S = vS.p; v.Y = vY.p; v.N = N; v.m = m; return v; }
    bool persistent() const { return !sparse && !legacy && (world == 1 || p2p); }
};

// Cost model of the tile walk, in units of one far tile: a tile that touches the diagonal band or the padded last column takes
// the slower path (sentinel selects), and every tile row a CTA starts costs a flush of the row accumulators plus the exposed
// latency of the row-coordinate loads.  Fitted to per-CTA phase times on B200 (scripts/ln_cta_fit.py over chr3d_single_profile_ctas:
// near/far = 0.99 .. 1.05, row/far = 0.79 .. 0.93 at N = 24,925 / 4,985; an equal-count split leaves the CTAs of the short bottom
// rows 12 % behind, this one brings max/mean of the per-CTA times to 1.015).
#define LN_COST_NEAR 1.0
#define LN_COST_ROW 0.9
static inline double ln_tile_cost(int I, int J, int nT, int N) {
    const bool near = J <= I + 1 || (long long)(J + 1) * LN_TB > N;
    return (near ? LN_COST_NEAR : 1.0) + (J == I ? LN_COST_ROW : 0.0);
}
// cuts the tile range [lo, hi) into `parts` contiguous pieces of equal modelled cost (every piece non-empty when hi - lo >= parts)
static std::vector<long long> ln_cost_split(long long lo, long long hi, int parts, int nT, int N) {
    std::vector<long long> cut(parts + 1, hi);
    cut[0] = lo;
    if (hi <= lo) return cut;
    double total = 0.0;
    { int I, J; tile_coords(lo, nT, I, J);
      for (long long t = lo; t < hi; t++) { total += ln_tile_cost(I, J, nT, N); if (++J == nT) { I++; J = I; } } }
    int I, J; tile_coords(lo, nT, I, J);
    double acc = 0.0; int k = 1;
    for (long long t = lo; t < hi && k < parts; t++) {
        acc += ln_tile_cost(I, J, nT, N);
        if (++J == nT) { I++; J = I; }
        // piece k-1 ends after tile t once its share is reached; leave at least one tile for each remaining piece
        while (k < parts && (acc >= total * k / parts || hi - (t + 1) <= parts - k)) { cut[k] = std::max(t + 1, cut[k - 1] + 1); k++; }
    }
    for (int q = 1; q <= parts; q++) cut[q] = std::min(hi, std::max(cut[q], cut[q - 1] + (q < parts ? 1 : 0)));
    cut[parts] = hi;
    return cut;
}

// tile range -> CTA ranges (cost-balanced): row-partial slots, the static CSR list "which (CTA, segment) row partials belong to
// tile row B" (mirrors the walk of ln_tile_phase), and the bead blocks every CTA gathers / integrates (balanced by gather length)
static int ln_plan(LnEngine& e, int ncta) {
    const long long ntl = e.t1 - e.t0;
    e.ncta = (int)std::max<long long>(1, std::min<long long>(ncta, std::max<long long>(ntl, 1)));
    e.h_cstart = ln_cost_split(e.t0, e.t1, e.ncta, e.nT, e.N);
    std::vector<std::vector<int>> rows(e.nT);
    int maxseg = 1;
    for (int pass = 0; pass < 2; pass++) {         // pass 0: segments per CTA -> maxseg; pass 1: the list
        for (int c = 0; c < e.ncta; c++) {
            int curI = -1, seg = -1;
            for (long long t = e.h_cstart[c]; t < e.h_cstart[c + 1]; t++) {
                int I, J; tile_coords(t, e.nT, I, J);
                if (I != curI) { curI = I; seg++; if (pass) rows[I].push_back(c * e.maxseg + seg); }
            }
            if (!pass) maxseg = std::max(maxseg, seg + 1);
        }
        if (!pass) e.maxseg = maxseg;
    }
    CHR3D_CUDA(e.rowpart.alloc((size_t)e.ncta * e.maxseg * LN_TB));
    CHR3D_CUDA(e.epart.alloc((size_t)e.ncta * 2));
    CHR3D_CUDA(cudaMemset(e.epart.p, 0, sizeof(double) * 2 * e.ncta));
    CHR3D_CUDA(e.cstart.alloc(e.h_cstart.size()));
    CHR3D_CUDA(cudaMemcpy(e.cstart.p, e.h_cstart.data(), sizeof(long long) * e.h_cstart.size(), cudaMemcpyHostToDevice));
    CHR3D_CUDA(e.cta_ns.alloc(e.ncta));
    CHR3D_CUDA(cudaMemset(e.cta_ns.p, 0, sizeof(unsigned long long) * e.ncta));
    std::vector<int> off(e.nT + 1, 0), idx;
    for (int B = 0; B < e.nT; B++) { off[B + 1] = off[B] + (int)rows[B].size(); idx.insert(idx.end(), rows[B].begin(), rows[B].end()); }
    CHR3D_CUDA(e.rs_off.alloc(off.size())); CHR3D_CUDA(e.rs_idx.alloc(std::max<size_t>(idx.size(), 1)));
    CHR3D_CUDA(cudaMemcpy(e.rs_off.p, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice));
    if (!idx.empty()) CHR3D_CUDA(cudaMemcpy(e.rs_idx.p, idx.data(), sizeof(int) * idx.size(), cudaMemcpyHostToDevice));
    {   // bead blocks -> CTAs, longest gather first onto the least loaded CTA (LPT); cost = owned tiles of the block's column + row partials + a constant
        int I0, J0, I1, J1; tile_coords(e.t0, e.nT, I0, J0);
        if (e.t1 >= e.n_tiles) { I1 = e.nT; J1 = e.nT; } else tile_coords(e.t1, e.nT, I1, J1);
        std::vector<std::pair<int, int>> cost(e.nblk);
        for (int B = 0; B < e.nblk; B++) {
            int lo = (B >= J0) ? I0 : I0 + 1, hi = (B < J1) ? I1 + 1 : I1;
            hi = std::min(hi, B + 1);
            cost[B] = std::make_pair(std::max(0, hi - lo) + (int)rows[B].size() + 8, B);
        }
        std::sort(cost.begin(), cost.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first != b.first ? a.first > b.first : a.second < b.second; });
        std::vector<std::vector<int>> mine(e.ncta); std::vector<long long> load(e.ncta, 0);
        for (auto& cb : cost) {
            int best = 0; for (int c = 1; c < e.ncta; c++) if (load[c] < load[best]) best = c;
            mine[best].push_back(cb.second); load[best] += cb.first;
        }
        std::vector<int> boff(e.ncta + 1, 0), bidx;
        for (int c = 0; c < e.ncta; c++) { std::sort(mine[c].begin(), mine[c].end()); boff[c + 1] = boff[c] + (int)mine[c].size(); bidx.insert(bidx.end(), mine[c].begin(), mine[c].end()); }
        CHR3D_CUDA(e.blk_off.alloc(boff.size())); CHR3D_CUDA(e.blk_idx.alloc(std::max<size_t>(bidx.size(), 1)));
        CHR3D_CUDA(cudaMemcpy(e.blk_off.p, boff.data(), sizeof(int) * boff.size(), cudaMemcpyHostToDevice));
        if (!bidx.empty()) CHR3D_CUDA(cudaMemcpy(e.blk_idx.p, bidx.data(), sizeof(int) * bidx.size(), cudaMemcpyHostToDevice));
    }
    {   // bead blocks every CTA's tiles feed (rows and columns), and how many CTAs feed each block: the completion counters
        // of the barrier-free evaluations count per CTA, from zero at this plan
        std::vector<int> toff(e.ncta + 1, 0), tidx, need(e.nblk, 0);
        for (int c = 0; c < e.ncta; c++) {
            std::vector<char> seen(e.nblk, 0);
            for (long long t = e.h_cstart[c]; t < e.h_cstart[c + 1]; t++) { int I, J; tile_coords(t, e.nT, I, J); seen[I] = 1; seen[J] = 1; }
            for (int B = 0; B < e.nblk; B++) if (seen[B]) { tidx.push_back(B); need[B]++; }
            toff[c + 1] = (int)tidx.size();
        }
        CHR3D_CUDA(e.tb_off.alloc(toff.size())); CHR3D_CUDA(e.tb_idx.alloc(std::max<size_t>(tidx.size(), 1))); CHR3D_CUDA(e.need.alloc(e.nblk));
        CHR3D_CUDA(cudaMemcpy(e.tb_off.p, toff.data(), sizeof(int) * toff.size(), cudaMemcpyHostToDevice));
        if (!tidx.empty()) CHR3D_CUDA(cudaMemcpy(e.tb_idx.p, tidx.data(), sizeof(int) * tidx.size(), cudaMemcpyHostToDevice));
        CHR3D_CUDA(cudaMemcpy(e.need.p, need.data(), sizeof(int) * e.nblk, cudaMemcpyHostToDevice));
        CHR3D_CUDA(e.ready.alloc(e.nblk)); CHR3D_CUDA(cudaMemset(e.ready.p, 0, sizeof(unsigned int) * e.nblk));
        CHR3D_CUDA(e.kslot.alloc((size_t)4 * e.nblk)); CHR3D_CUDA(cudaMemset(e.kslot.p, 0, sizeof(double2) * 4 * e.nblk));
        if (e.st.p) CHR3D_CUDA(cudaMemset(&e.st.p->revals, 0, sizeof(unsigned int)));
    }
    if (e.group) e.group->dirty = true;
    return 0;
}

// state, vectors and scratch every force provider needs
static int ln_alloc_vectors(LnEngine& e) {
    const int N = e.N;
    e.nT = (N + LN_TB - 1) / LN_TB; e.nblk = e.nT;
    e.nparts = std::min(64, std::max(1, (N + 255) / 256));
    e.self.es.assign(1, &e); e.group = &e.self;
    int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&e.sms, cudaDevAttrMultiProcessorCount, dev);
    CHR3D_CUDA(e.st.alloc(1));
    CHR3D_CUDA(cudaMemset(e.st.p, 0, sizeof(LnState)));
    CHR3D_CUDA(e.escal.alloc((size_t)3 + e.nblk));
    CHR3D_CUDA(e.part.alloc((size_t)e.nparts * LN_KMAX));
    const size_t n4 = (size_t)N + 1;
    CHR3D_CUDA(e.vx.alloc(n4)); CHR3D_CUDA(e.vxb.alloc(n4)); CHR3D_CUDA(e.vg.alloc(n4)); CHR3D_CUDA(e.vp.alloc(n4));
    CHR3D_CUDA(e.vv.alloc(n4)); CHR3D_CUDA(e.vf.alloc(n4)); CHR3D_CUDA(e.sx.alloc(n4)); CHR3D_CUDA(e.sv.alloc(n4));
    CHR3D_CUDA(e.vS.alloc((size_t)e.m * N)); CHR3D_CUDA(e.vY.alloc((size_t)e.m * N));
    CHR3D_CUDA(e.x0.alloc((size_t)N * 3));
    for (int q = 0; q < 2; q++) { CHR3D_CUDA(e.va[q].alloc(n4)); CHR3D_CUDA(e.vb[q].alloc(n4)); }
    CHR3D_CUDA(e.vfr.alloc(n4)); CHR3D_CUDA(e.blockpart.alloc((size_t)e.nblk * LN_KMAX));
    CHR3D_CUDA(cudaMemset(e.vf.p, 0, sizeof(float4) * n4));
    CHR3D_CUDA(cudaMemset(e.vp.p, 0, sizeof(float4) * n4));
    CHR3D_CUDA(cudaMallocHost((void**)&e.xerr, sizeof(int))); *e.xerr = 0;
    return 0;
}
static int ln_alloc(LnEngine& e) {
    const int N = e.N;
    if (int rc = ln_alloc_vectors(e)) return rc;
    e.n_tiles = (long long)e.nT * (e.nT + 1) / 2;
    {   // rank boundaries: contiguous tile ranges of equal modelled cost (every rank derives the same cuts)
        std::vector<long long> cut = ln_cost_split(0, e.n_tiles, e.world, e.nT, e.N);
        e.t0 = cut[e.rank]; e.t1 = cut[e.rank + 1]; e.rank_cut = cut;
    }
    int dev = 0; cudaGetDevice(&dev);
    const long long ntl = e.t1 - e.t0;
    {   // targets that cannot stay in L2 from one evaluation to the next are streamed with an evict-first policy
        int l2 = 0; cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
        e.stream_l2 = (double)ntl * LN_TILE_FLOATS * 4 > 0.6 * (double)l2 ? 1 : 0;
    }
    CHR3D_CUDA(e.Dt.alloc((size_t)std::max<long long>(ntl, 1) * LN_TILE_FLOATS));
    CHR3D_CUDA(e.colpart.alloc((size_t)std::max<long long>(ntl, 1) * 3 * LN_TB));
    if (int rc = ln_plan(e, e.sms)) return rc;
    CHR3D_CUDA(e.bar.alloc(64));
    CHR3D_CUDA(cudaMemset(e.bar.p, 0, sizeof(unsigned int) * 64));
#ifdef CHR3D_DEV
    CHR3D_CUDA(e.trace.alloc(2 * 512 * 2)); CHR3D_CUDA(cudaMemset(e.trace.p, 0, sizeof(unsigned long long) * 2 * 512 * 2));
#endif
    CHR3D_CUDA(cudaFuncSetAttribute(ln_tile_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    CHR3D_CUDA(cudaFuncSetAttribute(ln_tile_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    CHR3D_CUDA(cudaFuncSetAttribute(ln_tile_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    CHR3D_CUDA(cudaFuncSetAttribute(ln_tile_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    CHR3D_CUDA(cudaFuncSetAttribute(ln_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    CHR3D_CUDA(cudaFuncSetAttribute(ln_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LN_SMEM));
    return 0;
}

static void ln_terms(const chr3d_sa_params& p, const chr3d_stage& s, LnTerms& t) {
    t.wnC = (float)(s.w_noe * p.noe_sqconst); t.wnC2m = -2.f * t.wnC; t.rsw = (float)p.noe_rswitch;
    t.lo = p.noe_low_soft ? -(float)p.noe_rswitch : -INFINITY;
    t.wv = (float)s.w_vdw; t.wv4 = 4.f * (float)s.w_vdw;
    double r = s.repel * p.vdw_rmin; t.R2 = (float)(r * r);
}
static LnBonded ln_bonded_terms(const chr3d_sa_params& p, const chr3d_stage& s) {
    LnBonded b; b.w_bond = (float)s.w_bond; b.r0 = (float)p.bond_r0; b.d2 = (float)p.angle_d2;
    b.w_ang = (p.angle_d2 > 0 && p.bond_k > 0) ? (float)(s.w_bond * (p.angle_k / p.bond_k)) : 0.f;
    return b;
}
static int ln_dbg() {
    static const int dbg = dev_getenv("CHR3D_LN_DEBUG") ? atoi(dev_getenv("CHR3D_LN_DEBUG")) : 0;
    return dbg;
}

// ---- the persistent path: one cooperative launch per op ---------------------------------------------------------------
static void ln_fill_ctx(LnEngine& e, int cta0, LnCtx& c) {
    memset(&c, 0, sizeof c);
    c.Dt = e.Dt.p; c.t0 = e.t0; c.t1 = e.t1; c.nT = e.nT; c.N = e.N; c.nblk = e.nblk; c.maxseg = e.maxseg; c.cta0 = cta0; c.ncta = e.ncta;
    c.colpart = e.colpart.p; c.rowpart = e.rowpart.p; c.epart = e.epart.p; c.rs_off = e.rs_off.p; c.rs_idx = e.rs_idx.p;
    c.trace = e.trace.p;
    c.tb_off = e.tb_off.p; c.tb_idx = e.tb_idx.p; c.need = e.need.p; c.ready = e.ready.p; c.kslot = e.kslot.p;
    c.stream_l2 = e.stream_l2; c.cstart = e.cstart.p; c.blk_off = e.blk_off.p; c.blk_idx = e.blk_idx.p; c.cta_ns = e.cta_ns.p;
    tile_coords(e.t0, e.nT, c.gI0, c.gJ0);
    if (e.t1 >= e.n_tiles) { c.gI1 = e.nT; c.gJ1 = e.nT; } else tile_coords(e.t1, e.nT, c.gI1, c.gJ1);
    c.v = e.vec(); for (int q = 0; q < 2; q++) { c.va[q] = e.va[q].p; c.vb[q] = e.vb[q].p; }
    c.st = e.st.p; c.blockpart = e.blockpart.p; c.bar = e.bar.p;
    LnXchg& x = c.xc;
    x.world = e.p2p ? e.world : 1; x.rank = e.rank; x.nblk = e.nblk; x.n4 = e.N + 1; x.err = e.xerr;
    for (int g = 0; g < e.world && g < LN_MAXWORLD; g++) {
        int I, J; tile_coords(e.rank_cut[g], e.nT, I, J); c.bmin[g] = I;
        if (e.p2p) {
            char* base = (char*)e.peer[g];
            x.recv[g] = (float4*)base; x.es[g] = (double2*)(base + e.x_off_es());
        }
    }
}
static int ln_group_sync(LnGroup& g) {        // (re)build the device copy of the rank contexts
    if (!g.dirty) return 0;
    std::vector<LnCtx> h(g.es.size());
    int cta0 = 0;
    for (size_t k = 0; k < g.es.size(); k++) { ln_fill_ctx(*g.es[k], cta0, h[k]); cta0 += g.es[k]->ncta; }
    g.grid = cta0;
    CHR3D_CUDA(g.ctxs.alloc(h.size()));
    CHR3D_CUDA(cudaMemcpy(g.ctxs.p, h.data(), sizeof(LnCtx) * h.size(), cudaMemcpyHostToDevice));
    g.dirty = false;
    return 0;
}
static int ln_launch_persist(LnGroup& g, LnOp& op, cudaStream_t stream) {
    if (int rc = ln_group_sync(g)) return rc;
    LnEngine& e0 = *g.es[0];
    if (dev_getenv("CHR3D_LN_L2_PERSIST") && !e0.l2_window_set) {      // development: pin the column partials in L2 (access policy window)
        int dev = 0, maxp = 0, maxw = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev); cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        const size_t bytes = std::min<size_t>(std::min<size_t>((size_t)maxw, (size_t)maxp), sizeof(float) * (size_t)std::max<long long>(e0.t1 - e0.t0, 1) * 3 * LN_TB);
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes);
        cudaStreamAttrValue av; memset(&av, 0, sizeof av);
        av.accessPolicyWindow.base_ptr = e0.colpart.p; av.accessPolicyWindow.num_bytes = bytes; av.accessPolicyWindow.hitRatio = 1.0f;
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting; av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
        e0.l2_window_set = true;
    }
    op.dbg = ln_dbg();
    const LnCtx* ctxs = g.ctxs.p; int nvr = (int)g.es.size();
    void* args[3] = {(void*)&ctxs, (void*)&nvr, (void*)&op};
    const void* fn = e0.prm.noe_low_soft ? (const void*)ln_persist_kernel<true> : (const void*)ln_persist_kernel<false>;
    // cooperative: all CTAs are co-resident (they wait on one another at the grid barriers and on the peers' flags)
    CHR3D_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)g.grid), dim3(LN_P_THREADS), args, (size_t)LN_SMEM, stream));
    for (LnEngine* e : g.es) e->n_launches++;
    return 0;
}
static void ln_op_common(LnEngine& e, const chr3d_stage& s, LnOp& op) {
    memset(&op, 0, sizeof op);
    ln_terms(e.prm, s, op.terms); op.bd = ln_bonded_terms(e.prm, s); op.count_eval = 1;
    op.flow = 1;
    if (dev_getenv("CHR3D_LN_FLOW")) op.flow = atoi(dev_getenv("CHR3D_LN_FLOW"));
    op.ring_tail = LN_STAGES;
    if (dev_getenv("CHR3D_LN_RING_TAIL")) op.ring_tail = atoi(dev_getenv("CHR3D_LN_RING_TAIL"));
    op.defer_ring = 0;
    if (dev_getenv("CHR3D_LN_DEFER_RING")) op.defer_ring = atoi(dev_getenv("CHR3D_LN_DEFER_RING"));
}
static void ln_md_cfg(LnEngine& e, const chr3d_stage& s, LnMdCfg& c) {
    c.dt_akma = (float)(s.dt / CHR3D_TIMFAC); c.h = (float)(0.5 * (s.dt / CHR3D_TIMFAC) / e.prm.mass); c.dt_ps = s.dt; c.temp = s.temp;
    c.fbeta = e.prm.fbeta; c.kboltz = CHR3D_KBOLTZ; c.mass = e.prm.mass; c.flag = s.flag; c.N = e.N; c.nparts = e.nparts; c.count = 1;
}
static void ln_min_cfg(LnEngine& e, const chr3d_stage& s, LnMinCfg& c) {
    c.m = e.m; c.max_ls = e.prm.lbfgs_max_ls; c.nstep = s.nstep; c.nparts = e.nparts; c.c1 = e.prm.lbfgs_c1; c.gtol = e.prm.lbfgs_gtol;
    c.ftol = e.prm.lbfgs_ftol; c.drop = s.drop; c.N = e.N;
}

// Launch with programmatic stream serialization (legacy multi-launch path): the kernel is set up while its predecessor on the
// stream still runs and waits at its griddepcontrol.wait until the predecessor has finished (nobody triggers early).
template <typename Kern, typename Arg>
static cudaError_t ln_launch_pdl(Kern kern, unsigned grid, unsigned block, size_t smem, cudaStream_t s, const Arg& arg) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(block); lc.dynamicSmemBytes = smem; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    return cudaLaunchKernelEx(&lc, kern, arg);
}
template <typename Kern, typename A1, typename A2>
static cudaError_t ln_launch_pdl2(Kern kern, unsigned grid, unsigned block, cudaStream_t s, const A1& a1, const A2& a2) {
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(block); lc.dynamicSmemBytes = 0; lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    return cudaLaunchKernelEx(&lc, kern, a1, a2);
}
static void ln_launch_tiles(LnEngine& e, LnTileArgs a, bool want_e) {
    a.dbg = ln_dbg();
    const bool soft = e.prm.noe_low_soft != 0;
    if (want_e) { if (soft) ln_launch_pdl(ln_tile_kernel<true, true>, e.ncta, LN_THREADS, LN_SMEM, e.stream, a); else ln_launch_pdl(ln_tile_kernel<true, false>, e.ncta, LN_THREADS, LN_SMEM, e.stream, a); }
    else { if (soft) ln_launch_pdl(ln_tile_kernel<false, true>, e.ncta, LN_THREADS, LN_SMEM, e.stream, a); else ln_launch_pdl(ln_tile_kernel<false, false>, e.ncta, LN_THREADS, LN_SMEM, e.stream, a); }
}

static LnGather ln_gather_args(LnEngine& e) {
    LnGather g; g.colpart = e.colpart.p; g.rowpart = e.rowpart.p; g.rs_off = e.rs_off.p; g.rs_idx = e.rs_idx.p; g.t0 = e.t0; g.t1 = e.t1; g.nT = e.nT;
    return g;
}
// pair tiles at the coordinates x + lam*xb2 (xb2 == nullptr: plain x)
static void ln_tiles(LnEngine& e, const chr3d_stage& s, const float4* x, const float4* xb2, const float* lam, const int* done, bool want_e) {
    LnTileArgs a; a.Dt = e.Dt.p; a.t0 = e.t0; a.t1 = e.t1; a.nT = e.nT; a.N = e.N; a.x = x; a.xb2 = xb2; a.lam = lam; a.done = done;
    a.colpart = e.colpart.p; a.rowpart = e.rowpart.p; a.epart = e.epart.p; a.cstart = e.cstart.p; a.maxseg = e.maxseg; a.stream_l2 = e.stream_l2; ln_terms(e.prm, s, a.terms);
    ln_launch_tiles(e, a, want_e);
    e.n_launches++;
}
// ---- sparse force provider: total forces into e.vfr, energies into e.escal[0..2] (ln_sparse.cuh) ------------------------------
static SpGrid sp_grid(LnEngine& e) {
    SpGrid g; g.inv_cs = e.sp_inv_cs; g.hbits = e.sp_hbits; g.cell_start = e.sp_cell_start.p; g.cell_count = e.sp_cell_count.p; g.cell_fill = e.sp_cell_fill.p;
    g.order = e.sp_order.p; g.bsum = e.sp_bsum.p; g.ent = e.sp_ent.p; g.xref = e.sp_xref.p; g.rebuild = e.sp_rebuild.p; g.n_rebuilds = e.sp_nreb.p;
    return g;
}
static int sp_eval(LnEngine& e, const chr3d_stage& s, const float4* a, const float4* b, const float* lam, bool force_rebuild) {
    SpEvalArgs A;
    A.a = a; A.b = b; A.lam = lam; A.N = e.N; A.force_rebuild = force_rebuild ? 1 : 0; A.xs = e.sp_xs.p; A.g = sp_grid(e);
    A.m.row_ptr = e.sp_row_ptr.p; A.m.col = e.sp_col.p; A.m.dist = e.sp_dist.p;
    ln_terms(e.prm, s, A.t); A.bd = ln_bonded_terms(e.prm, s); A.vdw_min_sep = e.prm.vdw_min_sep;
    A.fsum = e.vfr.p; A.epart = e.sp_epart.p; A.escal = e.escal.p; A.ticket = e.sp_done.p;
    void* args[] = {&A};
    const void* fn = e.prm.noe_low_soft ? (const void*)sp_eval_kernel<true> : (const void*)sp_eval_kernel<false>;
    CHR3D_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)e.sp_blocks), dim3(SP_THREADS), args, 0, e.stream));      // one launch per evaluation
    e.n_launches += 1;
    return 0;
}

// ---- legacy multi-launch path (the NCCL baseline between ranks) ---------------------------------------------------------
// several ranks: reduce this rank's partial forces (+ bond on rank 0) into vfr, sum vfr and the three energies over the
// ranks with NCCL; afterwards every rank holds the total force in vfr and the totals in escal[0..2]
static int ln_exchange(LnEngine& e, const chr3d_stage& s, const float4* x, const float4* xb2, const float* lam) {
    LnReduceArgs r; r.g = ln_gather_args(e); r.epart = e.epart.p; r.N = e.N; r.ncta = e.ncta; r.add_bond = e.rank == 0 ? 1 : 0;
    r.x = x; r.xb2 = xb2; r.lam = lam; r.f = e.vfr.p; r.escal = e.escal.p; r.bd = ln_bonded_terms(e.prm, s);
    ln_reduce_kernel<<<e.nblk, LN_STEP_THREADS, 0, e.stream>>>(r);
    ln_energy_kernel<<<1, 32, 0, e.stream>>>(e.epart.p, e.ncta, e.escal.p, e.nblk);
    e.n_launches += 2;
#ifdef CHR3D_WITH_NCCL
    if (e.world > 1 && e.comm) {
        if (ncclAllReduce(e.vfr.p, e.vfr.p, (size_t)e.N * 4, ncclFloat, ncclSum, e.comm, e.stream) != ncclSuccess ||
            ncclAllReduce(e.escal.p, e.escal.p, 3, ncclDouble, ncclSum, e.comm, e.stream) != ncclSuccess) {
            set_error("ncclAllReduce failed"); return CHR3D_ERR_CUDA;
        }
        e.n_launches += 2;
    }
#endif
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}
// one energy + force evaluation at explicit coordinates e.vx, forces into e.vf, totals in escal[0..2]
static int ln_eval_legacy(LnEngine& e, const chr3d_stage& s, bool want_e = true) {
    if (e.sparse) {
        if (int rc = sp_eval(e, s, e.vx.p, nullptr, nullptr, true)) return rc;
        CHR3D_CUDA(cudaMemcpyAsync(e.vf.p, e.vfr.p, sizeof(float4) * (size_t)e.N, cudaMemcpyDeviceToDevice, e.stream));
        return 0;
    }
    ln_tiles(e, s, e.vx.p, nullptr, nullptr, nullptr, want_e);
    LnReduceArgs r; r.g = ln_gather_args(e); r.epart = e.epart.p; r.N = e.N; r.ncta = e.ncta; r.add_bond = e.rank == 0 ? 1 : 0;
    r.x = e.vx.p; r.xb2 = nullptr; r.lam = nullptr; r.f = e.vf.p; r.escal = e.escal.p; r.bd = ln_bonded_terms(e.prm, s);
    ln_reduce_kernel<<<e.nblk, LN_STEP_THREADS, 0, e.stream>>>(r);
    ln_energy_kernel<<<1, 32, 0, e.stream>>>(e.epart.p, e.ncta, e.escal.p, e.nblk);
    e.n_launches += 2;
#ifdef CHR3D_WITH_NCCL
    if (e.world > 1 && e.comm) {
        if (ncclAllReduce(e.vf.p, e.vf.p, (size_t)e.N * 4, ncclFloat, ncclSum, e.comm, e.stream) != ncclSuccess ||
            ncclAllReduce(e.escal.p, e.escal.p, 3, ncclDouble, ncclSum, e.comm, e.stream) != ncclSuccess) {
            set_error("ncclAllReduce failed"); return CHR3D_ERR_CUDA;
        }
    }
#endif
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}

static int ln_poll_done(LnEngine& e, int* done) {
    int h = 0;
    CHR3D_CUDA(cudaMemcpyAsync(&h, &e.st.p->done, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
    CHR3D_CUDA(cudaStreamSynchronize(e.stream));
    *done = h;
    return 0;
}

// L-BFGS, legacy: per trial three launches (tiles at xb + t*p, ln_min_step_kernel, ln_min_update_kernel); the host polls
// the `done` flag every 32 trials, launches after `done` return immediately
static int ln_minimize_legacy(LnEngine& e, const chr3d_stage& s) {
    const int vb = std::min(296, (e.N + 255) / 256);
    LnVec v = e.vec();
    const bool multi = e.world > 1 || e.sparse;          // forces arrive complete in vfr, energies in escal
    LnMinCfg c; ln_min_cfg(e, s, c);
    LnMinStepArgs m; m.g = ln_gather_args(e); m.fsum = multi ? e.vfr.p : nullptr; m.v = v; m.lam = &e.st.p->tstepf; m.init = 1;
    m.es.epart = e.epart.p; m.es.ncta = e.ncta; m.es.escal = multi ? e.escal.p : nullptr; m.st = e.st.p; m.blockpart = e.blockpart.p; m.nblk = e.nblk;
    m.bd = ln_bonded_terms(e.prm, s); m.cfg = c;
    if (e.sparse) { if (int rc = sp_eval(e, s, e.vx.p, nullptr, nullptr, true)) return rc; }
    else {
        ln_tiles(e, s, e.vx.p, nullptr, nullptr, nullptr, true);
        if (multi) if (int rc = ln_exchange(e, s, e.vx.p, nullptr, nullptr)) return rc;
    }
    ln_min_step_kernel<<<e.nblk, LN_STEP_THREADS, 0, e.stream>>>(m);
    ln_min_update_kernel<<<vb, 256, 0, e.stream>>>(v, e.st.p);
    e.n_launches += 2;
    m.init = 0;
    const long long max_trials = (long long)s.nstep * (e.prm.lbfgs_max_ls + 2) + 2;
    const int poll_every = 32;
    for (long long trial = 0; trial < max_trials;) {
        for (int k = 0; k < poll_every; k++, trial++) {
            if (e.sparse) { if (int rc = sp_eval(e, s, e.vxb.p, e.vp.p, &e.st.p->tstepf, false)) return rc; }
            else {
                ln_tiles(e, s, e.vxb.p, e.vp.p, &e.st.p->tstepf, &e.st.p->done, true);
                if (multi) if (int rc = ln_exchange(e, s, e.vxb.p, e.vp.p, &e.st.p->tstepf)) return rc;
            }
            ln_launch_pdl(ln_min_step_kernel, e.nblk, LN_STEP_THREADS, 0, e.stream, m);
            ln_launch_pdl2(ln_min_update_kernel, vb, 256, e.stream, v, (const LnState*)e.st.p);
            e.n_launches += 2;
        }
        int done = 0;
        if (int rc = ln_poll_done(e, &done)) return rc;
        if (done) break;
    }
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}
// Verlet MD, legacy: per step two launches (tiles at a + lam*b, ln_md_step_kernel); (a, b) double-buffered by parity
static int ln_dynamics_legacy(LnEngine& e, const chr3d_stage& s) {
    const bool multi = e.world > 1 || e.sparse;
    LnMdCfg c; ln_md_cfg(e, s, c);
    LnMdStepArgs p; p.g = ln_gather_args(e); p.fsum = multi ? e.vfr.p : nullptr; p.x = e.vx.p; p.v = e.vv.p; p.f = e.vf.p;
    p.es.epart = e.epart.p; p.es.ncta = e.ncta; p.es.escal = multi ? e.escal.p : nullptr; p.st = e.st.p; p.blockpart = e.blockpart.p; p.nblk = e.nblk; p.N = e.N;
    p.bd = ln_bonded_terms(e.prm, s); p.h = c.h; p.dt_akma = c.dt_akma; p.md = c;
    // evaluation at the starting point: explicit x, nothing pending
    if (e.sparse) { if (int rc = sp_eval(e, s, e.vx.p, nullptr, nullptr, true)) return rc; }
    else {
        ln_tiles(e, s, e.vx.p, nullptr, nullptr, nullptr, false);
        if (multi) if (int rc = ln_exchange(e, s, e.vx.p, nullptr, nullptr)) return rc;
    }
    p.a = e.vx.p; p.b = e.vx.p; p.lam = &e.st.p->zerof; p.a_next = e.va[0].p; p.b_next = e.vb[0].p; p.first = 1; p.want_e = 0;
    ln_launch_pdl(ln_md_step_kernel, e.nblk, LN_STEP_THREADS, 0, e.stream, p);
    e.n_launches++;
    int par = 0;
    for (int step = 0; step < s.nstep; step++) {
        const bool want_e = step == s.nstep - 1;                               // energies only matter after the last step
        if (e.sparse) { if (int rc = sp_eval(e, s, e.va[par].p, e.vb[par].p, &e.st.p->lamf, false)) return rc; }
        else {
            ln_tiles(e, s, e.va[par].p, e.vb[par].p, &e.st.p->lamf, nullptr, want_e);
            if (multi) if (int rc = ln_exchange(e, s, e.va[par].p, e.vb[par].p, &e.st.p->lamf)) return rc;
        }
        p.a = e.va[par].p; p.b = e.vb[par].p; p.lam = &e.st.p->lamf; p.a_next = e.va[par ^ 1].p; p.b_next = e.vb[par ^ 1].p; p.first = 0; p.want_e = want_e ? 1 : 0;
        ln_launch_pdl(ln_md_step_kernel, e.nblk, LN_STEP_THREADS, 0, e.stream, p);
        e.n_launches++;
        par ^= 1;
    }
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}
static int ln_energy_only_legacy(LnEngine& e, const chr3d_stage& s, bool count) {
    if (int rc = ln_eval_legacy(e, s)) return rc;
    LnMdCfg c; memset(&c, 0, sizeof c); c.mass = e.prm.mass; c.kboltz = CHR3D_KBOLTZ; c.N = e.N; c.nparts = e.nparts; c.flag = 0; c.temp = 0; c.count = count ? 1 : 0;
    LnVec v = e.vec();
    ln_md_post_kernel<<<e.nparts, 256, 0, e.stream>>>(v, c, 0, e.part.p);
    ln_md_control_kernel<<<1, 32, 0, e.stream>>>(e.st.p, e.part.p, e.escal.p, c);     // fills E terms (and counts the evaluation)
    e.n_launches += 2;
    return 0;
}

// ---- ops on a launch group ---------------------------------------------------------------------------------------------
static int ln_minimize(LnGroup& g, const chr3d_stage& s) {
    for (LnEngine* ep : g.es) {
        LnEngine& e = *ep;
        const int vb = std::min(296, (e.N + 255) / 256);
        LnVec v = e.vec();
        ln_state_reset_kernel<<<1, 32, 0, e.stream>>>(e.st.p, 0);
        ln_copy_xb_kernel<<<vb, 256, 0, e.stream>>>(v);
        // g := 0 so that the initial "y" is -f and g.g comes out of the same inner products; p := 0 keeps x + 0*p finite
        CHR3D_CUDA(cudaMemsetAsync(e.vg.p, 0, sizeof(float4) * (size_t)(e.N + 1), e.stream));
        CHR3D_CUDA(cudaMemsetAsync(e.vp.p, 0, sizeof(float4) * (size_t)(e.N + 1), e.stream));
        e.n_launches += 4;
    }
    LnEngine& e0 = *g.es[0];
    if (!e0.persistent()) return ln_minimize_legacy(e0, s);
    LnOp op; ln_op_common(e0, s, op); op.kind = LN_OP_MIN; ln_min_cfg(e0, s, op.min);
    op.max_trials = (long long)s.nstep * (e0.prm.lbfgs_max_ls + 2) + 2;
    return ln_launch_persist(g, op, e0.stream);          // leaves its loop on the device once the minimiser is done
}
static int ln_dynamics(LnGroup& g, const chr3d_stage& s) {
    for (LnEngine* ep : g.es) {
        LnEngine& e = *ep;
        const int vb = std::min(296, (e.N + 255) / 256);
        ln_sum3_kernel<<<e.nparts, 256, 0, e.stream>>>(e.vv.p, e.N, e.part.p);             // cmremove
        ln_sub_mean_kernel<<<vb, 256, 0, e.stream>>>(e.vv.p, e.N, e.part.p, e.nparts);
        e.n_launches += 2;
    }
    LnEngine& e0 = *g.es[0];
    if (!e0.persistent()) return ln_dynamics_legacy(e0, s);
    LnOp op; ln_op_common(e0, s, op); op.kind = LN_OP_MD; op.nstep = s.nstep; ln_md_cfg(e0, s, op.md);
    return ln_launch_persist(g, op, e0.stream);
}
static int ln_energy_only(LnGroup& g, const chr3d_stage& s, bool count) {
    LnEngine& e0 = *g.es[0];
    if (!e0.persistent()) return ln_energy_only_legacy(e0, s, count);
    LnOp op; ln_op_common(e0, s, op); op.kind = LN_OP_EVAL; op.count_eval = count ? 1 : 0;
    return ln_launch_persist(g, op, e0.stream);
}

static int ln_run_program(LnGroup& g, const chr3d_stage* prog, int nst) {
    for (LnEngine* ep : g.es) { ln_state_reset_kernel<<<1, 32, 0, ep->stream>>>(ep->st.p, 1); }
    for (int k = 0; k < nst; k++) {
        const chr3d_stage& s = prog[k];
        switch (s.op) {
        case CHR3D_OP_MINIMIZE: if (int rc = ln_minimize(g, s)) return rc; break;
        case CHR3D_OP_DYNAMICS: if (int rc = ln_dynamics(g, s)) return rc; break;
        case CHR3D_OP_SAVE: case CHR3D_OP_PICK: if (int rc = ln_energy_only(g, s, true)) return rc; break;
        case CHR3D_OP_ENERGY: if (int rc = ln_energy_only(g, s, false)) return rc; break;
        default: break;
        }
        for (LnEngine* ep : g.es) {
            LnEngine& e = *ep;
            const int vb = std::min(296, (e.N + 255) / 256);
            LnVec v = e.vec();
            switch (s.op) {
            case CHR3D_OP_INIT:
                ln_init_kernel<<<vb, 256, 0, e.stream>>>(v, e.x0.p, s.flag);
                ln_set_image_kernel<<<1, 32, 0, e.stream>>>(e.st.p, s.flag);
                break;
            case CHR3D_OP_MAXWELL:
                ln_maxwell_kernel<<<vb, 256, 0, e.stream>>>(v, e.seed, 0x100000u + (uint32_t)k, (uint32_t)e.model, sqrt(CHR3D_KBOLTZ * s.temp / e.prm.mass));
                break;
            case CHR3D_OP_SAVE:
                ln_save_kernel<<<vb, 256, 0, e.stream>>>(v, e.sx.p, e.sv.p, e.st.p);
                break;
            case CHR3D_OP_PICK:
                ln_pick_kernel<<<vb, 256, 0, e.stream>>>(v, e.sx.p, e.sv.p, e.st.p);
                ln_pick_state_kernel<<<1, 32, 0, e.stream>>>(e.st.p);
                break;
            case CHR3D_OP_CENTER:
                ln_sum3_kernel<<<e.nparts, 256, 0, e.stream>>>(e.vx.p, e.N, e.part.p);
                ln_sub_mean_kernel<<<vb, 256, 0, e.stream>>>(e.vx.p, e.N, e.part.p, e.nparts);
                break;
            default: break;
            }
        }
        CHR3D_CUDA(cudaGetLastError());
    }
    return 0;
}

// a wait inside the persistent kernel gave up (a peer died or stalled): report it instead of returning stale results
static int ln_check_comm(LnEngine& e, const char* where) {
    if (e.xerr && *e.xerr) {
        set_error("%s: the persistent kernel gave up waiting (%s, code %d): a peer rank died or stalled; the results are invalid", where,
                  *e.xerr == 1 ? "peer flag" : "grid barrier", *e.xerr);
        return CHR3D_ERR_CUDA;
    }
    return 0;
}

}  // namespace chr3d

using namespace chr3d;

struct chr3d_single { LnEngine e; cudaEvent_t ev0 = nullptr, ev1 = nullptr; std::vector<chr3d_stage> prog; LnGroup vgroup; };

// job: IF (host, N*N) or synthetic (IF == NULL, synth_seed != 0); x0 optional.
CHR3D_API int chr3d_single_create(const double* IF, int32_t N, double alpha, double K, int32_t min_sep, uint64_t synth_seed,
                                  const double* x0, uint64_t seed, const chr3d_sa_params* p, const chr3d_stage* program, int32_t nstages,
                                  int32_t rank, int32_t world, const void* nccl_id, chr3d_single** out) {
    if (N < 2 || !p || !program || nstages < 1 || !out || rank < 0 || world < 1 || rank >= world) { set_error("chr3d_single_create: bad arguments"); return CHR3D_ERR_ARG; }
    if (!IF && !synth_seed) { set_error("chr3d_single_create: need an IF matrix or a synthetic seed"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    std::unique_ptr<chr3d_single> h(new chr3d_single());
    LnEngine& e = h->e;
    e.N = N; e.rank = rank; e.world = world; e.prm = *p; e.seed = seed ? seed : p->seed;
    e.m = std::max(1, std::min((int)p->lbfgs_m, CHR3D_LBFGS_MAX_M));
    h->prog.assign(program, program + nstages);
    {
        const long long nT = (N + LN_TB - 1) / LN_TB, n_tiles = nT * (nT + 1) / 2;
        if (world > LN_MAXWORLD || world > n_tiles) { set_error("chr3d_single_create: world %d exceeds %d ranks or the %lld tiles of the structure", world, LN_MAXWORLD, n_tiles); return CHR3D_ERR_ARG; }
    }
    if (world > 1) {
        // nccl_id == NULL: no NCCL communicator — the ranks exchange over peer memory once chr3d_single_comm_connect (or
        // chr3d_single_group_connect) has run; until then forces/energies stay this rank's partial sums and
        // chr3d_single_run refuses to integrate
        if (nccl_id) {
#ifdef CHR3D_WITH_NCCL
            if (!nccl_api().ok) { set_error("libnccl.so.2 could not be loaded"); return CHR3D_ERR_ARG; }
            ncclUniqueId id; memcpy(&id, nccl_id, sizeof id);
            if (ncclCommInitRank(&e.comm, world, id, rank) != ncclSuccess) { set_error("ncclCommInitRank failed"); return CHR3D_ERR_CUDA; }
#else
            set_error("libchr3d was built without NCCL"); return CHR3D_ERR_ARG;
#endif
        }
    }
    if (int rc = ln_alloc(e)) return rc;
    if (world > 1) {
        e.xbytes = e.x_total();
        CHR3D_CUDA(cudaMalloc(&e.xbuf, e.xbytes));           // plain cudaMalloc: exportable through CUDA IPC
        CHR3D_CUDA(cudaMemset(e.xbuf, 0, e.xbytes));
    }
    {   // targets
        DevBuf<double> dIF, rows, mean; DevBuf<long long> dT;
        const size_t n2 = (size_t)N * N;
        const long long ntl = e.t1 - e.t0;
        if (IF && N <= CHR3D_EXACT_TABLE_MAX_L) {
            // within the reference's reach: the certified, bit-exact table of restraints.cu (sequential mean, libm-checked rounding)
            if (int rc = if_to_tenths_certified(IF, N, alpha, K, dIF, dT, nullptr, 0)) return rc;
            if (ntl > 0) ln_tiles_from_tenths_kernel<<<(unsigned)ntl, 256>>>(dT.p, N, min_sep, p->vdw_min_sep, e.t0, e.t1, e.nT, e.Dt.p);
        } else {
            CHR3D_CUDA(dIF.alloc(n2)); CHR3D_CUDA(rows.alloc(N)); CHR3D_CUDA(mean.alloc(1));
            if (IF) CHR3D_CUDA(cudaMemcpy(dIF.p, IF, sizeof(double) * n2, cudaMemcpyHostToDevice));
            else { ln_synth_if_kernel<<<148 * 8, 256>>>(dIF.p, N, synth_seed, 6.8e4, -0.95, 0.6, 5e-4); CHR3D_CUDA(cudaGetLastError()); }
            ln_row_pow_sum_kernel<<<(N + 7) / 8, 256>>>(dIF.p, N, alpha, rows.p);
            ln_sum_rows_kernel<<<1, 32>>>(rows.p, N, mean.p);
            if (ntl > 0) ln_build_tiles_kernel<<<(unsigned)ntl, 256>>>(dIF.p, N, alpha, K, mean.p, min_sep, p->vdw_min_sep, e.t0, e.t1, e.nT, e.Dt.p);
        }
        CHR3D_CUDA(cudaGetLastError());
        DevBuf<unsigned long long> cnt; CHR3D_CUDA(cnt.alloc(1)); CHR3D_CUDA(cudaMemset(cnt.p, 0, sizeof(unsigned long long)));
        if (ntl > 0) ln_count_rows_kernel<<<148 * 4, 256>>>(e.Dt.p, ntl * (long long)LN_TILE_FLOATS, cnt.p);
        unsigned long long hc = 0;
        CHR3D_CUDA(cudaMemcpy(&hc, cnt.p, sizeof hc, cudaMemcpyDeviceToHost));
        e.pairs_total = (long long)hc;          // restraint rows whose tile this rank owns
    }
    if (x0) {
        std::vector<float> hx((size_t)N * 3);
        for (size_t k = 0; k < hx.size(); k++) hx[k] = (float)x0[k];
        CHR3D_CUDA(cudaMemcpy(e.x0.p, hx.data(), sizeof(float) * hx.size(), cudaMemcpyHostToDevice));
    } else {
        if (int rc = dev_random_coil(e.x0.p, N, 1, 0, p->bond_r0, e.seed, 0)) return rc;
    }
    CHR3D_CUDA(cudaEventCreate(&h->ev0)); CHR3D_CUDA(cudaEventCreate(&h->ev1));
    CHR3D_CUDA(cudaDeviceSynchronize());
    *out = h.release();
    return 0;
}

// ---- internal: the ensemble API hands structures that exceed the cluster kernel's capacity to this engine ----------
__global__ void ln_pack_result_kernel(const float4* __restrict__ x, int N, const LnState* st, float* __restrict__ xout, double* e4, long long* nevals, int* hand) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) { const float4 q = x[i]; xout[3 * i] = q.x; xout[3 * i + 1] = q.y; xout[3 * i + 2] = q.z; }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        e4[CHR3D_E_OVERALL] = st->E; e4[CHR3D_E_BON] = st->Ebond; e4[CHR3D_E_VDW] = st->Evdw; e4[CHR3D_E_NOE] = st->Enoe;
        *nevals = st->nevals; *hand = st->hand;
    }
}
// targets from a dense [N][pitch] device matrix (restraint-list jobs); one GPU
int chr3d_single_create_dense(const float* d_T, int pitch, int32_t N, uint64_t seed, const chr3d_sa_params* p, const chr3d_stage* program,
                              int32_t nstages, long long n_rows, chr3d_single** out) {
    std::unique_ptr<chr3d_single> h(new chr3d_single());
    LnEngine& e = h->e;
    e.N = N; e.rank = 0; e.world = 1; e.prm = *p; e.seed = seed ? seed : p->seed;
    e.m = std::max(1, std::min((int)p->lbfgs_m, CHR3D_LBFGS_MAX_M));
    h->prog.assign(program, program + nstages);
    if (int rc = ln_alloc(e)) return rc;
    ln_tiles_from_dense_kernel<<<(unsigned)e.n_tiles, 256>>>(d_T, N, pitch, e.t0, e.t1, e.nT, e.Dt.p);
    CHR3D_CUDA(cudaGetLastError());
    e.pairs_total = n_rows;
    CHR3D_CUDA(cudaEventCreate(&h->ev0)); CHR3D_CUDA(cudaEventCreate(&h->ev1));
    CHR3D_CUDA(cudaDeviceSynchronize());
    *out = h.release();
    return 0;
}
// one model: start coordinates (d_x0 == nullptr: seeded random coil of `model`), the stage program, results packed on the device
int chr3d_single_run_model(chr3d_single* h, int model, const float* d_x0, void* stream, float* d_xout, double* d_e4, long long* d_nevals, int* d_hand) {
    LnEngine& e = h->e;
    e.stream = (cudaStream_t)stream; e.model = model;
    if (d_x0) CHR3D_CUDA(cudaMemcpyAsync(e.x0.p, d_x0, sizeof(float) * 3 * (size_t)e.N, cudaMemcpyDeviceToDevice, e.stream));
    else { if (int rc = dev_random_coil(e.x0.p, e.N, 1, model, e.prm.bond_r0, e.seed, e.stream)) return rc; }
    if (int rc = ln_run_program(e.self, h->prog.data(), (int)h->prog.size())) return rc;
    ln_pack_result_kernel<<<std::min(296, (e.N + 255) / 256), 256, 0, e.stream>>>(e.vx.p, e.N, e.st.p, d_xout, d_e4, d_nevals, d_hand);
    CHR3D_CUDA(cudaGetLastError());
    return 0;
}

// One structure from an explicit (sparse) restraint table: rows (ri < rj, 0-based, distance rd) as chr3d_if_to_restraints or a
// contact.tbl give them.  Forces come from the sparse provider (CSR restraint rows + cell-list repel, ln_sparse.cuh): memory
// traffic per evaluation is proportional to the number of rows, not to N^2.  One GPU.
CHR3D_API int chr3d_single_create_sparse(int32_t N, const int32_t* ri, const int32_t* rj, const double* rd, int64_t nr, const double* x0, uint64_t seed,
                                         const chr3d_sa_params* p, const chr3d_stage* program, int32_t nstages, chr3d_single** out) {
    if (N < 2 || nr < 0 || (nr > 0 && (!ri || !rj || !rd)) || !p || !program || nstages < 1 || !out) { set_error("chr3d_single_create_sparse: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    std::unique_ptr<chr3d_single> h(new chr3d_single());
    LnEngine& e = h->e;
    e.N = N; e.rank = 0; e.world = 1; e.prm = *p; e.seed = seed ? seed : p->seed; e.sparse = true;
    e.m = std::max(1, std::min((int)p->lbfgs_m, CHR3D_LBFGS_MAX_M));
    h->prog.assign(program, program + nstages);
    if (int rc = ln_alloc_vectors(e)) return rc;
    {   // symmetric CSR, columns ascending; rows are validated (0 <= i < j < N) and must not repeat
        std::vector<long long> ptr((size_t)N + 1, 0);
        for (int64_t k = 0; k < nr; k++) {
            if (ri[k] < 0 || rj[k] <= ri[k] || rj[k] >= N || !(rd[k] > 0)) { set_error("chr3d_single_create_sparse: row %lld is not 0 <= i < j < N with d > 0", (long long)k); return CHR3D_ERR_ARG; }
            ptr[ri[k] + 1]++; ptr[rj[k] + 1]++;
        }
        for (int i = 0; i < N; i++) ptr[i + 1] += ptr[i];
        const size_t nnz = (size_t)ptr[N];
        std::vector<std::pair<int, float>> ent(nnz);
        std::vector<long long> fill(ptr.begin(), ptr.end() - 1);
        for (int64_t k = 0; k < nr; k++) { ent[fill[ri[k]]++] = std::make_pair((int)rj[k], (float)rd[k]); ent[fill[rj[k]]++] = std::make_pair((int)ri[k], (float)rd[k]); }
        std::vector<int> col(std::max<size_t>(nnz, 1)); std::vector<float> dist(std::max<size_t>(nnz, 1));
        for (int i = 0; i < N; i++) {
            std::sort(ent.begin() + ptr[i], ent.begin() + ptr[i + 1]);
            for (long long q = ptr[i]; q < ptr[i + 1]; q++) {
                if (q > ptr[i] && ent[q].first == ent[q - 1].first) { set_error("chr3d_single_create_sparse: the pair (%d, %d) is listed twice", i, ent[q].first); return CHR3D_ERR_ARG; }
                col[q] = ent[q].first; dist[q] = ent[q].second;
            }
        }
        e.sp_nnz = (long long)nnz;
        CHR3D_CUDA(e.sp_row_ptr.alloc(ptr.size())); CHR3D_CUDA(e.sp_col.alloc(col.size())); CHR3D_CUDA(e.sp_dist.alloc(dist.size()));
        CHR3D_CUDA(cudaMemcpy(e.sp_row_ptr.p, ptr.data(), sizeof(long long) * ptr.size(), cudaMemcpyHostToDevice));
        CHR3D_CUDA(cudaMemcpy(e.sp_col.p, col.data(), sizeof(int) * col.size(), cudaMemcpyHostToDevice));
        CHR3D_CUDA(cudaMemcpy(e.sp_dist.p, dist.data(), sizeof(float) * dist.size(), cudaMemcpyHostToDevice));
    }
    {   // cell grid: edge = largest repel radius of the program + skin, 2N buckets (hashed: no bounding box)
        double rmax = 0.0;
        for (const chr3d_stage& s : h->prog) rmax = std::max(rmax, s.repel * p->vdw_rmin);
        if (!(rmax > 0.0)) rmax = p->vdw_rmin;
        e.sp_inv_cs = (float)(1.0 / (rmax + (double)SP_SKIN));
        e.sp_hbits = 4; while ((1 << e.sp_hbits) < 2 * N && e.sp_hbits < 17) e.sp_hbits++;
        const size_t H = (size_t)1 << e.sp_hbits;
        {   // one cooperative launch per evaluation: as many CTAs as are co-resident (at most eight per SM), no more than a warp per bead needs
            int bps = 0;
            CHR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, sp_eval_kernel<false>, SP_THREADS, 0));
            int bps1 = 0;
            CHR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps1, sp_eval_kernel<true>, SP_THREADS, 0));
            bps = std::max(1, std::min(std::min(bps, bps1), 8));
            e.sp_blocks = std::max(1, std::min(bps * e.sms, (N * 32 + SP_THREADS - 1) / SP_THREADS));
        }
        CHR3D_CUDA(e.sp_cell_start.alloc(H + 1)); CHR3D_CUDA(e.sp_cell_count.alloc(H)); CHR3D_CUDA(e.sp_cell_fill.alloc(H)); CHR3D_CUDA(e.sp_order.alloc(N));
        CHR3D_CUDA(e.sp_rebuild.alloc(1)); CHR3D_CUDA(e.sp_nreb.alloc(1)); CHR3D_CUDA(e.sp_done.alloc(1));
        CHR3D_CUDA(cudaMemset(e.sp_rebuild.p, 0, sizeof(int))); CHR3D_CUDA(cudaMemset(e.sp_nreb.p, 0, sizeof(long long))); CHR3D_CUDA(cudaMemset(e.sp_done.p, 0, sizeof(unsigned int)));
        CHR3D_CUDA(cudaMemset(e.sp_cell_start.p, 0, sizeof(int) * (H + 1))); CHR3D_CUDA(cudaMemset(e.sp_cell_count.p, 0, sizeof(int) * H)); CHR3D_CUDA(cudaMemset(e.sp_cell_fill.p, 0, sizeof(int) * H));
        CHR3D_CUDA(e.sp_xs.alloc(N)); CHR3D_CUDA(e.sp_xref.alloc(N)); CHR3D_CUDA(e.sp_epart.alloc((size_t)e.sp_blocks * 3)); CHR3D_CUDA(e.sp_bsum.alloc((size_t)e.sp_blocks)); CHR3D_CUDA(e.sp_ent.alloc(N));
        CHR3D_CUDA(cudaMemset(e.sp_xref.p, 0, sizeof(float4) * N));
    }
    e.pairs_total = nr;
    if (x0) {
        std::vector<float> hx((size_t)N * 3);
        for (size_t k = 0; k < hx.size(); k++) hx[k] = (float)x0[k];
        CHR3D_CUDA(cudaMemcpy(e.x0.p, hx.data(), sizeof(float) * hx.size(), cudaMemcpyHostToDevice));
    } else {
        if (int rc = dev_random_coil(e.x0.p, N, 1, 0, p->bond_r0, e.seed, 0)) return rc;
    }
    CHR3D_CUDA(cudaEventCreate(&h->ev0)); CHR3D_CUDA(cudaEventCreate(&h->ev1));
    CHR3D_CUDA(cudaDeviceSynchronize());
    *out = h.release();
    return 0;
}
// Host-only (no device needed): the tile-range split the engine uses for a structure of N beads on `world` ranks with
// `ctas` CTAs per rank.  rank_cut[world + 1]: tile t of the row-major upper-triangle list belongs to rank g when
// rank_cut[g] <= t < rank_cut[g + 1]; cta_cut[world * (ctas + 1)] (optional): the same per rank for its CTAs; cost[world]
// (optional): modelled cost of every rank's range (a far tile = 1).
CHR3D_API int chr3d_single_tile_split(int32_t N, int32_t world, int32_t ctas, int64_t* rank_cut, int64_t* cta_cut, double* cost) {
    if (N < 2 || world < 1 || ctas < 1 || !rank_cut) { set_error("chr3d_single_tile_split: bad arguments"); return CHR3D_ERR_ARG; }
    const int nT = (N + LN_TB - 1) / LN_TB;
    const long long n_tiles = (long long)nT * (nT + 1) / 2;
    const std::vector<long long> cut = ln_cost_split(0, n_tiles, world, nT, N);
    for (int g = 0; g <= world; g++) rank_cut[g] = cut[g];
    for (int g = 0; g < world; g++) {
        if (cta_cut) {
            const int nc = (int)std::max<long long>(1, std::min<long long>(ctas, std::max<long long>(cut[g + 1] - cut[g], 1)));
            const std::vector<long long> cc = ln_cost_split(cut[g], cut[g + 1], nc, nT, N);
            for (int c = 0; c <= ctas; c++) cta_cut[(size_t)g * (ctas + 1) + c] = c <= nc ? cc[c] : cut[g + 1];
        }
        if (cost) {
            double a = 0.0;
            if (cut[g + 1] > cut[g]) { int I, J; tile_coords(cut[g], nT, I, J); for (long long t = cut[g]; t < cut[g + 1]; t++) { a += ln_tile_cost(I, J, nT, N); if (++J == nT) { I++; J = I; } } }
            cost[g] = a;
        }
    }
    return 0;
}
// sparse provider: bin rebuilds so far, entries of the symmetric CSR, bytes the NOE term reads per evaluation
CHR3D_API int chr3d_single_sparse_info(chr3d_single* h, int64_t* n_rebuilds, int64_t* csr_entries, int64_t* noe_bytes_per_eval) {
    if (!h || !h->e.sparse) { set_error("chr3d_single_sparse_info: not a sparse structure"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e;
    CHR3D_CUDA(cudaStreamSynchronize(e.stream));
    long long nb = 0;
    CHR3D_CUDA(cudaMemcpy(&nb, e.sp_nreb.p, sizeof nb, cudaMemcpyDeviceToHost));
    if (n_rebuilds) *n_rebuilds = nb;
    if (csr_entries) *csr_entries = e.sp_nnz;
    if (noe_bytes_per_eval) *noe_bytes_per_eval = e.sp_nnz * 8 + (long long)(e.N + 1) * 8;
    return 0;
}

static bool ln_integrates(const std::vector<chr3d_stage>& prog) {
    for (const chr3d_stage& s : prog) if (s.op == CHR3D_OP_MINIMIZE || s.op == CHR3D_OP_DYNAMICS) return true;
    return false;
}

CHR3D_API int chr3d_single_run(chr3d_single* h, void* stream) {
    if (!h) { set_error("chr3d_single_run: null handle"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e;
    if (e.group != &e.self) { set_error("chr3d_single_run: this structure is a virtual rank of a group (use chr3d_single_group_run)"); return CHR3D_ERR_ARG; }
    bool connected = e.world == 1 || e.p2p;
#ifdef CHR3D_WITH_NCCL
    if (e.comm) connected = true;
#endif
    if (!connected && ln_integrates(h->prog)) {
        set_error("chr3d_single_run: a sharded structure without a communicator or a peer connection can only evaluate (INIT/ENERGY)"); return CHR3D_ERR_ARG;
    }
    e.legacy = (e.world > 1 && !e.p2p) || e.sparse;
    if (dev_getenv("CHR3D_LN_LEGACY")) e.legacy = true;
    if (e.legacy && e.world > 1 && !connected) {     // partial sums of a shard on one GPU (tests of the partition)
        e.legacy = true;
    }
    e.stream = (cudaStream_t)stream; e.n_launches = 0;
    CHR3D_CUDA(cudaEventRecord(h->ev0, e.stream));
    if (int rc = ln_run_program(e.self, h->prog.data(), (int)h->prog.size())) return rc;
    CHR3D_CUDA(cudaEventRecord(h->ev1, e.stream));
    return 0;
}

// Virtual ranks: `n` structures created in THIS process on THIS device with rank g of world n become one launch group — the
// persistent kernel hosts all of them in one cooperative grid (SMs / n CTAs each) and their "peer" buffers are plain device
// pointers.  This is the multi-rank protocol (tile-range shards, peer exchange, redundant integration) on a single GPU.
CHR3D_API int chr3d_single_group_connect(chr3d_single** hs, int32_t n) {
    if (!hs || n < 1 || n > LN_MAXWORLD) { set_error("chr3d_single_group_connect: bad arguments"); return CHR3D_ERR_ARG; }
    for (int g = 0; g < n; g++) if (!hs[g] || hs[g]->e.world != n || hs[g]->e.rank != g || (n > 1 && !hs[g]->e.xbuf)) { set_error("chr3d_single_group_connect: handle %d is not rank %d of %d", g, g, n); return CHR3D_ERR_ARG; }
    LnGroup& grp = hs[0]->vgroup;
    grp.es.clear(); grp.dirty = true;
    const int each = std::max(1, hs[0]->e.sms / n);
    for (int g = 0; g < n; g++) {
        LnEngine& e = hs[g]->e;
        for (int q = 0; q < n; q++) e.peer[q] = hs[q]->e.xbuf;
        e.p2p = n > 1; e.peers_local = true; e.legacy = false;
        e.group = &grp;
        if (int rc = ln_plan(e, each)) return rc;
        grp.es.push_back(&e);
    }
    return 0;
}
CHR3D_API int chr3d_single_group_run(chr3d_single** hs, int32_t n, void* stream) {
    if (!hs || n < 1 || !hs[0] || (int)hs[0]->vgroup.es.size() != n) { set_error("chr3d_single_group_run: the handles are not a connected group"); return CHR3D_ERR_ARG; }
    for (int g = 0; g < n; g++) { hs[g]->e.stream = (cudaStream_t)stream; hs[g]->e.n_launches = 0; }
    CHR3D_CUDA(cudaEventRecord(hs[0]->ev0, (cudaStream_t)stream));
    if (int rc = ln_run_program(hs[0]->vgroup, hs[0]->prog.data(), (int)hs[0]->prog.size())) return rc;
    CHR3D_CUDA(cudaEventRecord(hs[0]->ev1, (cudaStream_t)stream));
    for (int g = 1; g < n; g++) { CHR3D_CUDA(cudaEventRecord(hs[g]->ev0, (cudaStream_t)stream)); CHR3D_CUDA(cudaEventRecord(hs[g]->ev1, (cudaStream_t)stream)); }
    return 0;
}

// `reps` force evaluations at the start coordinates, timed on the device: the roofline measurement of ln_tile_kernel.
CHR3D_API int chr3d_single_eval_bench(chr3d_single* h, void* stream, int32_t reps, int32_t want_energy, double* tile_ms, double* eval_ms) {
    if (!h || reps < 1) { set_error("chr3d_single_eval_bench: bad arguments"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e; e.stream = (cudaStream_t)stream;
    if (e.group != &e.self || e.sparse) { set_error("chr3d_single_eval_bench: not available for virtual ranks / sparse structures"); return CHR3D_ERR_ARG; }
    e.legacy = e.world > 1 && !e.p2p;
    if (dev_getenv("CHR3D_LN_LEGACY")) e.legacy = true;
    LnVec v = e.vec();
    ln_init_kernel<<<std::min(296, (e.N + 255) / 256), 256, 0, e.stream>>>(v, e.x0.p, 1);
    chr3d_stage s = h->prog.back();
    cudaEvent_t a, b; CHR3D_CUDA(cudaEventCreate(&a)); CHR3D_CUDA(cudaEventCreate(&b));
    LnTileArgs ta; ta.Dt = e.Dt.p; ta.t0 = e.t0; ta.t1 = e.t1; ta.nT = e.nT; ta.N = e.N; ta.x = e.vx.p; ta.xb2 = nullptr; ta.lam = nullptr; ta.done = nullptr;
    ta.colpart = e.colpart.p; ta.rowpart = e.rowpart.p; ta.epart = e.epart.p; ta.cstart = e.cstart.p; ta.maxseg = e.maxseg; ta.stream_l2 = e.stream_l2; ln_terms(e.prm, s, ta.terms);
    for (int w = 0; w < 3; w++) ln_launch_tiles(e, ta, want_energy != 0);
    CHR3D_CUDA(cudaEventRecord(a, e.stream));
    for (int r = 0; r < reps; r++) ln_launch_tiles(e, ta, want_energy != 0);
    CHR3D_CUDA(cudaEventRecord(b, e.stream));
    CHR3D_CUDA(cudaEventSynchronize(b));
    float ms = 0.f; CHR3D_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (tile_ms) *tile_ms = ms / reps;
    // a whole force evaluation as the integrator issues it: `reps` MD steps with dt = 0 (the state does not move)
    chr3d_stage md = s; md.op = CHR3D_OP_DYNAMICS; md.dt = 0.0; md.temp = 0.0; md.flag = 0; md.nstep = 2;
    CHR3D_CUDA(cudaMemsetAsync(e.vv.p, 0, sizeof(float4) * (size_t)(e.N + 1), e.stream));
    if (int rc = ln_dynamics(e.self, md)) return rc;
    md.nstep = reps;
    CHR3D_CUDA(cudaEventRecord(a, e.stream));
    if (int rc = ln_dynamics(e.self, md)) return rc;
    CHR3D_CUDA(cudaEventRecord(b, e.stream));
    CHR3D_CUDA(cudaEventSynchronize(b));
    CHR3D_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (eval_ms) *eval_ms = ms / (reps + 1);     // the stage's starting-point evaluation is one more
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ln_check_comm(e, "chr3d_single_eval_bench");
}

CHR3D_API int chr3d_single_fetch(chr3d_single* h, double* xyz, double* energies4, double* forces, int64_t* force_evals, int32_t* hand,
                                 double* device_ms, int64_t* n_launches, int64_t* restraint_pairs) {
    if (!h) { set_error("chr3d_single_fetch: null handle"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e;
    {
        const cudaError_t se = cudaStreamSynchronize(e.stream);
        if (int rc = ln_check_comm(e, "chr3d_single_fetch")) return rc;
        CHR3D_CUDA(se);
    }
    LnState st;
    CHR3D_CUDA(cudaMemcpy(&st, e.st.p, sizeof st, cudaMemcpyDeviceToHost));
    std::vector<float4> buf((size_t)e.N);
    if (xyz) {
        CHR3D_CUDA(cudaMemcpy(buf.data(), e.vx.p, sizeof(float4) * e.N, cudaMemcpyDeviceToHost));
        for (int i = 0; i < e.N; i++) { xyz[3 * i] = buf[i].x; xyz[3 * i + 1] = buf[i].y; xyz[3 * i + 2] = buf[i].z; }
    }
    if (forces) {
        CHR3D_CUDA(cudaMemcpy(buf.data(), e.vf.p, sizeof(float4) * e.N, cudaMemcpyDeviceToHost));
        for (int i = 0; i < e.N; i++) { forces[3 * i] = buf[i].x; forces[3 * i + 1] = buf[i].y; forces[3 * i + 2] = buf[i].z; }
    }
    if (energies4) { energies4[CHR3D_E_OVERALL] = st.E; energies4[CHR3D_E_BON] = st.Ebond; energies4[CHR3D_E_VDW] = st.Evdw; energies4[CHR3D_E_NOE] = st.Enoe; }
    if (force_evals) *force_evals = st.nevals;
    if (hand) *hand = st.hand;
    if (device_ms) { float ms = 0.f; if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) { ms = 0.f; cudaGetLastError(); } *device_ms = ms; }
    if (n_launches) *n_launches = e.n_launches;
    if (restraint_pairs) *restraint_pairs = e.pairs_total;
    return 0;
}

// Phase clock of the persistent kernel since the last call (milliseconds, as seen by CTA 0 of this rank), 8 entries: [0] tile
// phases, [1] first grid barrier (+ energy publication), [2] step phases (one rank: gather + integration; several ranks: summation
// + integration), [3] second barrier (+ decisions), [4] L-BFGS update + third barrier, [5] push to the peers, [6] wait for
// the peers' partial forces, [7] gather of this rank's partials (several ranks).  Reading resets the clock.
CHR3D_API int chr3d_single_profile(chr3d_single* h, double* ms8) {
    double* ms5 = ms8;
    if (!h || !ms5) { set_error("chr3d_single_profile: bad arguments"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e;
    CHR3D_CUDA(cudaStreamSynchronize(e.stream));
    unsigned long long t[8];
    CHR3D_CUDA(cudaMemcpy(t, &e.st.p->tprof[0], sizeof t, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 8; q++) ms5[q] = 1e-6 * (double)t[q];
    CHR3D_CUDA(cudaMemset(&e.st.p->tprof[0], 0, sizeof t));
    return 0;
}

// Per-CTA view of the tile walk: tiles, "near" tiles (slow path), tile rows started, and the time every CTA spent in tile phases
// of the persistent kernel since the last call (ms; reading resets it).  Arrays of `cap` entries; *n = CTAs of this rank.
CHR3D_API int chr3d_single_profile_ctas(chr3d_single* h, int32_t cap, int32_t* n, int32_t* tiles, int32_t* near_tiles, int32_t* rows, double* tile_ms) {
    if (!h || !n) { set_error("chr3d_single_profile_ctas: bad arguments"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e;
    *n = e.ncta;
    CHR3D_CUDA(cudaStreamSynchronize(e.stream));
    std::vector<unsigned long long> ns(e.ncta);
    CHR3D_CUDA(cudaMemcpy(ns.data(), e.cta_ns.p, sizeof(unsigned long long) * e.ncta, cudaMemcpyDeviceToHost));
    CHR3D_CUDA(cudaMemset(e.cta_ns.p, 0, sizeof(unsigned long long) * e.ncta));
    for (int c = 0; c < e.ncta && c < cap; c++) {
        int nt = 0, nn = 0, nr = 0, curI = -1;
        for (long long t = e.h_cstart[c]; t < e.h_cstart[c + 1]; t++) {
            int I, J; tile_coords(t, e.nT, I, J);
            nt++; if (J <= I + 1 || (long long)(J + 1) * LN_TB > e.N) nn++;
            if (I != curI) { curI = I; nr++; }
        }
        if (tiles) tiles[c] = nt; if (near_tiles) near_tiles[c] = nn; if (rows) rows[c] = nr; if (tile_ms) tile_ms[c] = 1e-6 * (double)ns[c];
    }
    return 0;
}

// development (-DCHR3D_DEV builds): raw event trace of two CTAs of the last persistent launch, 2 x 512 x (event, ns)
CHR3D_API int chr3d_single_trace(chr3d_single* h, unsigned long long* out2048) {
    if (!h || !out2048 || !h->e.trace.p) { set_error("chr3d_single_trace: not a development build"); return CHR3D_ERR_ARG; }
    CHR3D_CUDA(cudaStreamSynchronize(h->e.stream));
    CHR3D_CUDA(cudaMemcpy(out2048, h->e.trace.p, sizeof(unsigned long long) * 2048, cudaMemcpyDeviceToHost));
    return 0;
}

CHR3D_API int chr3d_single_info(const chr3d_single* h, int64_t* n_tiles, int64_t* tiles_this_rank, int32_t* ctas, int64_t* target_bytes) {
    if (!h) { set_error("chr3d_single_info: null handle"); return CHR3D_ERR_ARG; }
    if (n_tiles) *n_tiles = h->e.n_tiles;
    if (tiles_this_rank) *tiles_this_rank = h->e.t1 - h->e.t0;
    if (ctas) *ctas = h->e.ncta;
    if (target_bytes) *target_bytes = (h->e.t1 - h->e.t0) * (long long)LN_TILE_FLOATS * 4;
    return 0;
}

// Test aid: this rank's targets as a dense row-major N*N fp32 matrix (cells outside the rank's tiles: NaN; the lower triangle
// mirrors the upper one).  Sentinels as stored: -1 no restraint, -2 no restraint and no repel.
__global__ void __launch_bounds__(256) ln_tiles_to_dense_kernel(const float* __restrict__ Dt, int N, long long t0, long long t1, long long nT, float* __restrict__ T) {
    const long long t = t0 + blockIdx.x;
    if (t >= t1) return;
    int I, J; tile_coords(t, nT, I, J);
    const float* in = Dt + (size_t)(t - t0) * LN_TILE_FLOATS;
    for (int k = threadIdx.x; k < LN_TILE_FLOATS; k += blockDim.x) {
        int r, c; ln_tile_cell(k, r, c);
        const int i = I * LN_TB + r, j = J * LN_TB + c;
        if (i < N && j < N && j > i) { T[(size_t)i * N + j] = in[k]; T[(size_t)j * N + i] = in[k]; }
    }
}
CHR3D_API int chr3d_single_targets(chr3d_single* h, float* out) {
    if (!h || !out) { set_error("chr3d_single_targets: bad arguments"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e;
    const size_t n2 = (size_t)e.N * e.N;
    DevBuf<float> d; CHR3D_CUDA(d.alloc(n2));
    CHR3D_CUDA(cudaMemset(d.p, 0xff, sizeof(float) * n2));          // NaN
    const long long ntl = e.t1 - e.t0;
    if (ntl > 0) ln_tiles_to_dense_kernel<<<(unsigned)ntl, 256>>>(e.Dt.p, e.N, e.t0, e.t1, e.nT, d.p);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaMemcpy(out, d.p, sizeof(float) * n2, cudaMemcpyDeviceToHost));
    return 0;
}

CHR3D_API int chr3d_single_destroy(chr3d_single* h) {
    if (!h) return 0;
#ifdef CHR3D_WITH_NCCL
    if (h->e.comm) ncclCommDestroy(h->e.comm);
#endif
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    cudaDeviceSynchronize();
    if (!h->e.peers_local) for (int g = 0; g < h->e.world && g < LN_MAXWORLD; g++) if (h->e.peer[g] && g != h->e.rank) cudaIpcCloseMemHandle(h->e.peer[g]);
    if (h->e.xbuf) cudaFree(h->e.xbuf);
    if (h->e.xerr) cudaFreeHost(h->e.xerr);
    cudaGetLastError();
    delete h;
    return 0;
}

// Peer-memory exchange set-up (world > 1): every rank exports the IPC handle of its communication buffer
// (64 bytes), the caller gathers the handles of all ranks in rank order (e.g. torch.distributed.all_gather_object)
// and hands them to chr3d_single_comm_connect on every rank.  After that the force exchange of every evaluation
// is fused into the persistent kernel (remote stores + flags over NVLink); without it the engine uses NCCL all-reduce.
CHR3D_API int chr3d_single_comm_handle(chr3d_single* h, void* out64) {
    if (!h || !out64 || !h->e.xbuf) { set_error("chr3d_single_comm_handle: no communication buffer (world == 1?)"); return CHR3D_ERR_ARG; }
    cudaIpcMemHandle_t mh;
    CHR3D_CUDA(cudaIpcGetMemHandle(&mh, h->e.xbuf));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out64, &mh, 64);
    return 0;
}
CHR3D_API int chr3d_single_comm_connect(chr3d_single* h, const void* handles, int32_t world) {
    if (!h || !handles || world != h->e.world || !h->e.xbuf) { set_error("chr3d_single_comm_connect: bad arguments"); return CHR3D_ERR_ARG; }
    LnEngine& e = h->e;
    for (int g = 0; g < world; g++) {
        if (g == e.rank) { e.peer[g] = e.xbuf; continue; }
        cudaIpcMemHandle_t mh; memcpy(&mh, (const char*)handles + 64 * (size_t)g, 64);
        CHR3D_CUDA(cudaIpcOpenMemHandle(&e.peer[g], mh, cudaIpcMemLazyEnablePeerAccess));
    }
    e.p2p = true; e.self.dirty = true;
    return 0;
}
CHR3D_API int chr3d_single_comm_error(chr3d_single* h, int32_t* err) {
    if (!h || !err) { set_error("chr3d_single_comm_error: bad arguments"); return CHR3D_ERR_ARG; }
    *err = h->e.xerr ? *h->e.xerr : 0;
    return 0;
}

CHR3D_API int chr3d_nccl_unique_id(void* out128) {
#ifdef CHR3D_WITH_NCCL
    if (!nccl_api().ok) { set_error("libnccl.so.2 could not be loaded"); return CHR3D_ERR_ARG; }
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) { set_error("ncclGetUniqueId failed"); return CHR3D_ERR_CUDA; }
    memcpy(out128, &id, sizeof id);
    return 0;
#else
    (void)out128; set_error("libchr3d was built without NCCL"); return CHR3D_ERR_ARG;
#endif
}

// synthetic IF matrix on the host side of the ABI (small sizes, for tests): generated on the device, copied back
CHR3D_API int chr3d_synthetic_if(int32_t N, uint64_t seed, double* out) {
    if (N < 1 || !out) { set_error("chr3d_synthetic_if: bad arguments"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    DevBuf<double> d; CHR3D_CUDA(d.alloc((size_t)N * N));
    ln_synth_if_kernel<<<148 * 8, 256>>>(d.p, N, seed, 6.8e4, -0.95, 0.6, 5e-4);
    CHR3D_CUDA(cudaGetLastError());
    CHR3D_CUDA(cudaMemcpy(out, d.p, sizeof(double) * (size_t)N * N, cudaMemcpyDeviceToHost));
    return 0;
}
