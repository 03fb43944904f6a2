// ln_persist.cuh — the persistent evaluation kernel of the large-N engine (included by large_n.cu inside namespace chr3d).
//
// One cooperative launch per DYNAMICS / MINIMIZE / ENERGY op of the stage program (C3D:1573-1829) runs ALL force evaluations
// of that op: per evaluation
//     tile phase (TMA ring over this CTA's tile range, ln_tile_phase)
//  -> grid barrier
//  -> step phase: fixed-order gather of the partial forces of a bead block, exchange with the other ranks over NVLink peer
//     memory (remote stores + epoch flags), bond term, Verlet step / L-BFGS inner products
//  -> grid barrier whose last arriver takes the scalar decisions (thermostat factor, line search, history update)
// with no launch, no host round trip and no pipeline prologue in between: the TMA ring keeps running across the evaluation
// boundary (the targets do not depend on the coordinates).  The minimiser leaves its loop on the device when it has converged.
//
// Ranks: the tile list is cut into one contiguous range per rank (row blocks balanced by pair count); a rank
// holds partial forces on the beads of its tile rows and of every column right of them.  Every rank pushes the partial force
// of bead block B into slot `rank` of every peer's receive buffer (only blocks it contributes to), the receivers add the slots
// in rank order and integrate all beads redundantly: identical arithmetic on every rank, hence identical coordinates and no
// coordinate exchange (one NVLink hop per evaluation instead of two).
//
// A launch may host several VIRTUAL ranks (blockIdx ranges of one cooperative grid, peer pointers into the same device):
// that is how the multi-rank protocol is tested on a single GPU, where separately launched kernels that wait on one another
// would not be guaranteed to run concurrently.

#define LN_P_THREADS LN_THREADS
#define LN_P_GGRP 5                         // the gather of a bead block: 96 threads per [3][128] partial, 5 partials in flight
enum { LN_OP_MD = 1, LN_OP_MIN = 2, LN_OP_EVAL = 3 };

struct LnCtx {                              // one (virtual) rank
    const float* Dt; long long t0, t1, nT; int N, nblk, maxseg, cta0, ncta, stream_l2;
    float* colpart; float4* rowpart; double* epart;
    const int* rs_off; const int* rs_idx;
    const long long* cstart;                // [ncta + 1] first tile of every CTA (cost-balanced, ln_plan)
    const int* blk_off; const int* blk_idx; // bead blocks of every CTA (balanced by gather length)
    unsigned long long* cta_ns;             // [ncta] time spent in tile phases (ns)
    // barrier-free evaluations (the bulk of an MD op): which bead blocks this CTA's tiles feed, how many CTAs feed each block,
    // the per-block completion counters and the per-block {sum of v.v, epoch} / {bond energy, epoch} slots
    const int* tb_off; const int* tb_idx;   // blocks touched by every CTA (rows and columns of its tiles)
    const int* need;                        // [nblk] CTAs that touch block B
    unsigned int* ready;                    // [nblk] contributions finished, monotonic: evaluation E of this plan is complete at E * need[B]
    double2* kslot;                         // [2 parities][nblk][2]
    unsigned long long* trace;              // development trace buffer (or nullptr)
    int gI0, gJ0, gI1, gJ1;                 // tile coordinates of t0 and t1: the owned tiles of a column are a contiguous row range
    LnVec v; float4* va[2]; float4* vb[2];
    LnState* st; double* blockpart;
    unsigned int* bar;                      // [0] arrivals, [32] generation (separate 128-byte lines)
    LnXchg xc; int bmin[LN_MAXWORLD];       // bmin[g]: first bead block rank g contributes to (its first tile row)
};
struct LnOp {
    int kind, nstep, count_eval, dbg;
    int flow;               // 1: the evaluations in the middle of an MD op run without grid barriers (see "barrier-free evaluations")
    int ring_tail;          // how many tiles of the next evaluation are requested before the current tile phase ends (0 .. LN_STAGES);
                            // the rest is requested when the next tile phase starts
    int defer_ring;         // 1: the tile ring stops at the end of an evaluation and is restarted from the step phase, behind the push
                            // to the peers (several ranks: the copies would otherwise queue in front of the latency-critical gather);
                            // 0: the ring keeps running across the evaluation boundary (one rank: there is no idle time to hide it in)
    LnTerms terms; LnBonded bd;
    LnMdCfg md; LnMinCfg min; long long max_trials;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
    unsigned int v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 16-byte pieces of the peer exchange: one store / one load each (the epoch travels inside the piece)
__device__ __forceinline__ void st_v4_sys(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_v4_sys(const float4* p) {
    float4 v; asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_d2_sys(double2* p, double a, double b) {
    asm volatile("st.relaxed.sys.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ double2 ld_d2_sys(const double2* p) {
    double2 v; asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned long long ln_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// A wait that can never complete (a peer died, a protocol bug) must not hang the GPU: after ~20 s the reason goes to the
// host-visible error word and the kernel traps; chr3d_single_run / _fetch report it.
#define LN_WAIT_LIMIT 40000000000ll
__device__ __noinline__ void ln_fail(int* err, int code) {
    if (err) { *err = code; __threadfence_system(); }
    __trap();
}
__device__ __forceinline__ void ln_wait_flag_sys(const unsigned int* p, unsigned int epoch, int* err) {
    if ((int)(ld_acquire_sys(p) - epoch) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(p) - epoch) < 0) if (clock64() - t0 > LN_WAIT_LIMIT) ln_fail(err, 1);
}

// ---- grid barrier of one (virtual) rank ------------------------------------------------------------------------------
// arrive: returns true in exactly one CTA, the last to arrive, which sees every other CTA's writes and runs the serial section;
// the others wait inside for the release.  `gen` is the barrier generation (a register copy, equal in all CTAs).
__device__ __forceinline__ bool ln_bar_arrive(const LnCtx& c, unsigned int gen, int* s_flag) {
    __syncthreads();
    if (threadIdx.x == 0) {
        // release: everything this CTA wrote (ordered before by the barrier above) is visible to whoever acquires the counter;
        // acquire: the last arriver sees every other CTA's writes.  (An acquire at gpu scope also invalidates this SM's L1.)
        unsigned int old;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(c.bar) : "memory");
        const int last = old == (unsigned)c.ncta - 1u;
        if (!last) {
            // poll with relaxed loads (an acquire load drags an L1 invalidation behind every poll), acquire once at the end
            if (ld_relaxed_gpu(c.bar + 32) == gen) {
                const long long t0 = clock64();
                while (ld_relaxed_gpu(c.bar + 32) == gen) if (clock64() - t0 > LN_WAIT_LIMIT) ln_fail(c.xc.err, 2);
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        *s_flag = last;
    }
    __syncthreads();
    return *s_flag != 0;
}
__device__ __forceinline__ void ln_bar_release(const LnCtx& c, unsigned int gen) {      // the last arriver, after its serial section
    __syncthreads();
    if (threadIdx.x == 0) { *c.bar = 0u; st_release_gpu(c.bar + 32, gen + 1u); }     // (the release orders the counter reset before the new generation)
}

// Loop state lives in shared memory, not in registers: the tile phase is inlined into the kernel and needs the whole register
// budget (there is no L1 behind 229 KB of shared memory, a spill in the pair loop is an L2 round trip).
// tacc: phase clock of CTA 0 of a rank (ns, accumulated into LnState::tprof): 0 tile phase, 1 first barrier (+ energy
// publication), 2 step phase (one rank: gather + integration; several ranks: summation + integration only), 3 second barrier
// (+ decisions), 4 L-BFGS update + third barrier, 5 gather + push to the peers, 6 wait for the peers' partial forces
struct LnLoop { int step, par, stop, prefetch; unsigned int gen, epoch, revals; float lam; unsigned long long tlast, tacc[8], t0tile, ttile; };
__shared__ LnCtx lnp_c;
__shared__ LnOp lnp_op;
__shared__ LnLoop lnp_L;
__shared__ int lnp_flag;
// Development trace (-DCHR3D_DEV): raw (event, globaltimer) pairs of two CTAs of a rank, first LN_TRACE_MAX events of a launch
#ifdef CHR3D_DEV
#define LN_TRACE_MAX 512
#define LN_TRACE(ev) do { const int w_ = ((int)blockIdx.x - lnp_c.cta0 == 0) ? 0 : (((int)blockIdx.x - lnp_c.cta0 == lnp_c.ncta / 2) ? 1 : -1); \
    if (w_ >= 0 && threadIdx.x == 0 && lnp_c.trace) { const unsigned int n_ = lnp_trace_n; if (n_ < LN_TRACE_MAX) { \
        lnp_c.trace[((size_t)w_ * LN_TRACE_MAX + n_) * 2] = (unsigned long long)(ev); lnp_c.trace[((size_t)w_ * LN_TRACE_MAX + n_) * 2 + 1] = ln_globaltimer(); lnp_trace_n = n_ + 1; } } } while (0)
__shared__ unsigned int lnp_trace_n;
#else
#define LN_TRACE(ev) do { } while (0)
#endif
__device__ __forceinline__ void ln_tprof(volatile LnLoop* L, int slot, bool timing) {
    if (timing) { const unsigned long long t = ln_globaltimer(); L->tacc[slot] += t - L->tlast; L->tlast = t; }
}
// ---- scratch in the dynamic shared memory: the column-partial buffer of the tile phase is free between tile phases ------
struct LnScratch {
    union {
        float gred[LN_P_GGRP][3 * LN_TB];                   // gather: one [3][128] partial sum per thread group
        float prod[6 * CHR3D_LBFGS_MAX_M + 6][LN_TB];       // L-BFGS: per-bead products (also the scratch of the serial sections)
    };
    double dsum[32];                                        // block sums
    double G[LN_KMAX];                                      // L-BFGS: reduced inner products
};
static_assert(sizeof(LnScratch) <= LN_SCRATCH_BYTES, "scratch must fit the column-partial + column-coordinate buffers");
static_assert(6 * CHR3D_LBFGS_MAX_M + 6 <= LN_KMAX && 2 * LN_P_THREADS * 8 <= (6 * CHR3D_LBFGS_MAX_M + 6) * LN_TB * 4, "serial-section scratch");

// ---- barrier-free evaluations -------------------------------------------------------------------------------------------------
// Inside an MD op only the first evaluation (nothing to integrate yet) and the last one (energies) need the whole grid in step.
// The evaluations in between run as a dataflow:
//   * a CTA that has finished its tile phase adds 1 (release) to the counter of every bead block its tiles feed;
//   * the CTA that gathers block B waits until the counter shows that all feeders are through (acquire): no grid barrier
//     between tile phase and gather;
//   * after integrating block B its CTA publishes {sum of v.v over the block, epoch} (and the bond energy) in a 16-byte slot;
//   * before its next tile phase every CTA waits for the slots of ALL blocks and adds them up itself, in the fixed order the
//     barrier path uses (lane-strided sums of one warp, then the butterfly): every CTA gets the same thermostat factor with no
//     second grid barrier, and because that factor needs every block, no CTA can run ahead far enough to overwrite anything
//     that is still being read (partials, coordinate buffers, slots of the other parity).
__device__ __forceinline__ void red_release_gpu(unsigned int* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void ln_signal_tiles_done(const LnCtx& c) {           // all threads, after the tile phase
    __syncthreads();
    const int cta = blockIdx.x - c.cta0;
    for (int k = c.tb_off[cta] + (int)threadIdx.x; k < c.tb_off[cta + 1]; k += LN_P_THREADS) red_release_gpu(c.ready + c.tb_idx[k]);
}
__device__ __forceinline__ void ln_wait_ready(const LnCtx& c, int B, unsigned int revals) {     // one thread
    const unsigned int target = revals * (unsigned)c.need[B];
    if ((int)(ld_acquire_gpu(c.ready + B) - target) >= 0) return;
    const long long t0 = clock64();
    while ((int)(ld_relaxed_gpu(c.ready + B) - target) < 0) if (clock64() - t0 > LN_WAIT_LIMIT) ln_fail(c.xc.err, 3);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// one thread, after the block's vectors have been written (and a CTA barrier): publish the block's scalars of evaluation `epoch`
__device__ __forceinline__ void ln_publish_block(const LnCtx& c, int B, unsigned int epoch, double k2, double eb) {
    __threadfence();
    double2* s = c.kslot + ((size_t)(epoch & 1u) * c.nblk + B) * 2;
    const double ep = __longlong_as_double((long long)epoch);
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(s), "d"(k2), "d"(ep) : "memory");
    asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(s + 1), "d"(eb), "d"(ep) : "memory");
}
// warp 0 of every CTA: wait for the slots of all blocks of evaluation `epoch`, thermostat factor from their sum (as ln_decide_md)
__device__ __forceinline__ float ln_lambda_from_slots(const LnCtx& c, const LnOp& op, unsigned int epoch, double* ke_out) {
    const int lane = threadIdx.x & 31;
    double a0 = 0;
    for (int q = lane; q < c.nblk; q += 32) {
        const double2* s = c.kslot + ((size_t)(epoch & 1u) * c.nblk + q) * 2;
        double2 v;
        asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(s) : "memory");
        if (__double_as_longlong(v.y) != (long long)epoch) {
            const long long t0 = clock64();
            do {
                asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(s) : "memory");
                if (clock64() - t0 > LN_WAIT_LIMIT) ln_fail(c.xc.err, 4);
            } while (__double_as_longlong(v.y) != (long long)epoch);
        }
        a0 += v.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");              // the coordinates the integrators wrote before their slots
    const double ke = 0.5 * op.md.mass * a0, ndof = 3.0 * c.N - 3.0;
    const double T = 2.0 * ke / (op.md.kboltz * ndof);
    double l2 = 1.0;
    if (T > 0.0) {
        if (op.md.flag == CHR3D_COUPLE_TCOUP) l2 = sqrt(fmax(0.0, 1.0 + 2.0 * op.md.fbeta * op.md.dt_ps * (op.md.temp / T - 1.0)));
        else if (op.md.flag == CHR3D_COUPLE_VSCALE) l2 = sqrt(op.md.temp / T);
    }
    if (ke_out) *ke_out = ke;
    return (float)l2;
}

// The bead blocks of this CTA and the row partials that belong to them, copied into shared memory once per launch: the step
// phase is a chain of dependent memory accesses, and every index it would otherwise fetch from global memory costs a round
// trip to L2 per evaluation.
#define LNP_MAXB 8
#define LNP_MAXRS 16
struct LnBlkPlan { int nb, cached; int B[LNP_MAXB], lo[LNP_MAXB], hi[LNP_MAXB], rs_cnt[LNP_MAXB], rs[LNP_MAXB][LNP_MAXRS]; };
__device__ __forceinline__ void ln_blk_bounds(const LnCtx& c, int B, int& lo, int& hi) {      // owned tiles of column B: rows [lo, hi)
    lo = (B >= c.gJ0) ? c.gI0 : c.gI0 + 1; hi = (B < c.gJ1) ? c.gI1 + 1 : c.gI1;
    if (hi > B + 1) hi = B + 1;
}
__device__ void ln_blk_plan_load(const LnCtx& c, LnBlkPlan* bp) {          // one thread
    const int cta = blockIdx.x - c.cta0;
    const int b0 = c.blk_off[cta], nb = c.blk_off[cta + 1] - b0;
    bp->nb = nb; bp->cached = nb <= LNP_MAXB ? 1 : 0;
    for (int k = 0; k < nb && k < LNP_MAXB; k++) {
        const int B = c.blk_idx[b0 + k];
        bp->B[k] = B; ln_blk_bounds(c, B, bp->lo[k], bp->hi[k]);
        const int r0 = c.rs_off[B], nr = c.rs_off[B + 1] - r0;
        bp->rs_cnt[k] = nr;
        if (nr > LNP_MAXRS) bp->cached = 0;
        for (int q = 0; q < nr && q < LNP_MAXRS; q++) bp->rs[k][q] = c.rs_idx[r0 + q];
    }
}
__shared__ LnBlkPlan lnp_bp;
__shared__ LnTileRun lnp_run;
extern __shared__ __align__(128) unsigned char lnp_smem[];

// This rank's partial force on the beads of bead block B.  The CTA works on `nslots` blocks at a time (1 or 2): slot s is
// the thread range [s*512/nslots, (s+1)*512/nslots), inside it groups of 96 threads each sum every G-th partial of the column
// (16-byte loads, sixteen in flight per thread: the gather is bound by dependent round trips to L2), the first 128
// threads of the slot ("lead" threads, lt = 0..127) then add the groups and the row partials in a fixed order.
// k = index of the block in the CTA's cached plan (-1: not cached, look the row partials up in global memory).
// All threads call; the result is valid in the lead threads of the slot.
// (Deliberately compact code — one rolled loop, 32-bit index arithmetic: the step phases run once per evaluation, straight
// out of a cold instruction cache, so their cost is largely the instruction fetch.)
__device__ __noinline__ void ln_gather4(const LnCtx& c, int B, int k, int slot, int nslots, LnScratch* sc, float& fx, float& fy, float& fz) {
    const int per = LN_P_THREADS / nslots, ts = threadIdx.x - slot * per, G = per / 96;
    int lo, hi;
    if (k >= 0) { lo = lnp_bp.lo[k]; hi = lnp_bp.hi[k]; } else ln_blk_bounds(c, B, lo, hi);
    if (B >= 0 && ts < G * 96) {
        const int q = ts % 96, grp = ts / 96, nT = (int)c.nT, t0 = (int)c.t0;
        float4 acc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        // tile (I, B) sits at I*nT - I(I-1)/2 + B - I in the tile list; the rank's partials start at tile t0 (96 float4 each)
        const float4* base = reinterpret_cast<const float4*>(c.colpart) + q;
#pragma unroll 1
        for (int I = lo + grp; I < hi; I += 16 * G) {
            float4 u[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int Ij = I + j * G;
                u[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (Ij < hi) u[j] = __ldcg(base + (size_t)(unsigned)((Ij * nT - ((Ij * (Ij - 1)) >> 1) + B - Ij - t0) * 96));
            }
#pragma unroll
            for (int j = 0; j < 16; j++) { acc[j & 3].x += u[j].x; acc[j & 3].y += u[j].y; acc[j & 3].z += u[j].z; acc[j & 3].w += u[j].w; }
        }
        float4 a0;
        a0.x = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x); a0.y = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
        a0.z = (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z); a0.w = (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w);
        *reinterpret_cast<float4*>(&sc->gred[slot * G + grp][q * 4]) = a0;
    }
    LN_TRACE(25);
    // the row partials of this block (the lead threads; issued behind the column partials, before anybody waits)
    float rx = 0.f, ry = 0.f, rz = 0.f;
    if (B >= 0 && ts < LN_TB) {
        const int r0 = k >= 0 ? 0 : c.rs_off[B], r1 = k >= 0 ? lnp_bp.rs_cnt[k] : c.rs_off[B + 1];
#pragma unroll 1
        for (int r = r0; r < r1; r += 8) {                       // eight independent loads in flight (a long tile row is cut into a dozen segments)
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r + j < r1) v[j] = __ldcg(c.rowpart + (size_t)(k >= 0 ? lnp_bp.rs[k][r + j] : c.rs_idx[r + j]) * LN_TB + ts);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) { rx += v[j].x; ry += v[j].y; rz += v[j].z; }
        }
    }
    LN_TRACE(26);
    __syncthreads();
    LN_TRACE(27);
    fx = fy = fz = 0.f;
    if (B >= 0 && ts < LN_TB) {
        const int pos = 4 * (ts & 31) + (ts >> 5);              // column -> position inside a partial (ln_tile_phase)
#pragma unroll 1
        for (int g = 0; g < G; g++) { fx += sc->gred[slot * G + g][pos]; fy += sc->gred[slot * G + g][LN_TB + pos]; fz += sc->gred[slot * G + g][2 * LN_TB + pos]; }
        fx += rx; fy += ry; fz += rz;
    }
    __syncthreads();                                           // gred may be overwritten by the next block
}

// Total force on the beads of this CTA's blocks: gather, exchange over the ranks.
// emit(active, B, i, lt, slot, fx, fy, fz) is called by ALL threads of the CTA for every round of blocks (block-wide reductions
// inside are allowed); a thread is a "lead" thread of bead i when lt >= 0 (lt = 0..127 = bead inside the block) and active.
// PAIR: with several ranks and more than one block, the CTA works on two blocks at a time (two halves of the CTA): every block
// is a chain of dependent memory latencies (partials, peers' slots, vectors), so two of them side by side take the time of one.
// Several ranks: pass 0 gathers and pushes every block (slot `rank` of every rank's receive buffer, the epoch in .w), pass 1
// polls the slots of the contributors of each block until they carry this epoch and adds them in rank order.  One rank: a
// single pass, the gathered force goes straight to emit.  (One call site each for the gather and for emit: compact code.)
// preload(active, i, lt) -> whatever emit needs from memory that does not depend on the forces (coordinates of the bead and of its
// chain neighbours, velocity, previous force ...): issued BEFORE the gather (one rank) or before the poll of the peers' slots
// (several ranks), so that those round trips to L2 overlap instead of queueing up behind one another.
struct LnNoPre {};
template <bool PAIR, typename Preload, typename Emit>
__device__ __forceinline__ void ln_forces_of_my_blocks(const LnCtx& c, LnScratch* sc, unsigned int epoch, volatile LnLoop* L, bool timing, Preload preload, Emit emit,
                                                       unsigned int flow_revals = 0u /* != 0: barrier-free evaluation number of this plan: wait for the feeders of each block */) {
    const int tid = threadIdx.x, cta = blockIdx.x - c.cta0;
    const LnXchg& x = c.xc;
    const int nb = lnp_bp.nb, cached = lnp_bp.cached;
    const bool multi = x.world > 1;
    const int parity = (int)(epoch & 1u);
    const size_t myslot = ((size_t)parity * x.world + x.rank) * x.n4;
    const int nslots = (PAIR && multi && nb > 1) ? 2 : 1, per = LN_P_THREADS / nslots, slot = tid / per, ts = tid - slot * per, lt = ts < LN_TB ? ts : -1;
#pragma unroll 1
    for (int pass = 0; pass < (multi ? 2 : 1); pass++) {
        // the partial forces are out (several ranks): now the first tiles of the next evaluation may start to stream in
        if (pass == 1 && tid == 0 && L->prefetch) ln_ring_prefetch(&lnp_run, lnp_smem);
#pragma unroll 1
        for (int k0 = 0; k0 < nb; k0 += nslots) {
            const int k = k0 + slot;
            const int B = k < nb ? (cached ? lnp_bp.B[k] : c.blk_idx[c.blk_off[cta] + k]) : -1;
            const int i = B * LN_TB + lt;
            float fx = 0.f, fy = 0.f, fz = 0.f;
            decltype(preload(false, 0, 0)) pre = decltype(preload(false, 0, 0))();
            if (pass == (multi ? 1 : 0)) pre = preload(B >= 0 && lt >= 0 && i < c.N, i, lt);
            if (pass == 0) {
                const int Bg = (B >= 0 && multi && B < c.bmin[x.rank]) ? -1 : B;          // nothing of this block in this rank's tiles
                LN_TRACE(10 + k0);
                if (flow_revals) { if (Bg >= 0 && ts == 0) ln_wait_ready(c, Bg, flow_revals); __syncthreads(); }
                LN_TRACE(20 + k0);
                ln_gather4(c, Bg, (cached && Bg >= 0) ? k : -1, slot, nslots, sc, fx, fy, fz);
                LN_TRACE(30 + k0);
                if (multi) {
                    ln_tprof(L, 7, timing);
                    if (Bg >= 0 && lt >= 0 && i < c.N) {
                        const float4 v = make_float4(fx, fy, fz, __uint_as_float(epoch));
#pragma unroll 1
                        for (int g = 0; g < x.world; g++) st_v4_sys(x.recv[g] + myslot + i, v);
                    }
                    LN_TRACE(40 + k0);
                    if (k0 + nslots >= nb) ln_tprof(L, 5, timing);
                    continue;
                }
            } else {
                if (B >= 0 && lt >= 0 && i < c.N) {
                    const float4* R = x.recv[x.rank] + (size_t)parity * x.world * x.n4 + i;
#pragma unroll 1
                    for (int g = 0; g < x.world; g++) if (B >= c.bmin[g]) {
                        float4 v = ld_v4_sys(R + (size_t)g * x.n4);
                        if (__float_as_uint(v.w) != epoch) {
                            const long long t0 = clock64();
                            do { v = ld_v4_sys(R + (size_t)g * x.n4); if (clock64() - t0 > LN_WAIT_LIMIT) ln_fail(x.err, 1); } while (__float_as_uint(v.w) != epoch);
                        }
                        fx += v.x; fy += v.y; fz += v.z;
                    }
                }
                __syncthreads();
                LN_TRACE(50 + k0);
                if (k0 + nslots >= nb) ln_tprof(L, 6, timing);
            }
            emit(B >= 0, B, i, lt, slot, fx, fy, fz, pre);
            LN_TRACE(60 + k0);
        }
    }
    if (!multi && tid == 0 && L->prefetch) ln_ring_prefetch(&lnp_run, lnp_smem);
}

// sum of `v` over the 128 lead threads (4 warps) of each slot, result in every thread of the slot; all threads call; fixed order
__device__ __forceinline__ double ln_lead_sum(double v, double* sh8, int lt, int slot) {
    if (lt >= 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((lt & 31) == 0) sh8[slot * 4 + (lt >> 5)] = v;
    }
    __syncthreads();
    const double s = (sh8[slot * 4] + sh8[slot * 4 + 1]) + (sh8[slot * 4 + 2] + sh8[slot * 4 + 3]);
    __syncthreads();
    return s;
}
// coordinates a + lam*b read around L1 (written by other CTAs of this launch)
__device__ __forceinline__ float4 ln_ldx_cg(const float4* a, const float4* b, float lam, int i) {
    const float4 p = __ldcg(a + i), q = __ldcg(b + i);
    return make_float4(fmaf(lam, q.x, p.x), fmaf(lam, q.y, p.y), fmaf(lam, q.z, p.z), 0.f);
}
__device__ __forceinline__ void ln_bond_cg(const float4* a, const float4* b, float lam, int i, int N, const LnBonded& bd, float& fx, float& fy, float& fz, float& eb) {
    ln_bonded(LnLdCg(), a, b, lam, i, N, bd, fx, fy, fz, eb);
}

// pair energies of this rank (sum of the per-CTA partials of the tile phase, fixed order) -> every rank; called by all threads
// of the last arriver of the first barrier of an evaluation that wants energies
__device__ __forceinline__ void ln_energy_publish(const LnCtx& c, LnScratch* sc, unsigned int epoch) {
    const int tid = threadIdx.x;
    double a0 = 0, a1 = 0;
    for (int q = tid; q < c.ncta; q += LN_P_THREADS) { a0 += __ldcg(c.epart + 2 * q); a1 += __ldcg(c.epart + 2 * q + 1); }
    double* red = reinterpret_cast<double*>(sc->prod);          // [2][512]
    red[tid] = a0; red[LN_P_THREADS + tid] = a1;
    __syncthreads();
    for (int o = LN_P_THREADS / 2; o > 0; o >>= 1) { if (tid < o) { red[tid] += red[tid + o]; red[LN_P_THREADS + tid] += red[LN_P_THREADS + tid + o]; } __syncthreads(); }
    if (tid == 0) {
        const LnXchg& x = c.xc;
        c.st->eloc[0] = red[0]; c.st->eloc[1] = red[LN_P_THREADS];
        if (x.world > 1) {
            const int parity = (int)(epoch & 1u);
            const double ep = __longlong_as_double((long long)epoch);
            for (int g = 0; g < x.world; g++) {
                double2* d = x.es[g] + ((size_t)parity * x.world + x.rank) * 2;
                st_d2_sys(d, red[0], ep); st_d2_sys(d + 1, red[LN_P_THREADS], ep);
            }
        }
    }
    __syncthreads();
}
// one thread: pair energies summed over the ranks in rank order (identical on every rank)
__device__ __forceinline__ void ln_energy_collect(const LnCtx& c, unsigned int epoch, double& e0, double& e1) {
    const LnXchg& x = c.xc;
    if (x.world <= 1) { e0 = c.st->eloc[0]; e1 = c.st->eloc[1]; return; }
    const int parity = (int)(epoch & 1u);
    e0 = e1 = 0.0;
    for (int g = 0; g < x.world; g++) {
        const double2* d = x.es[x.rank] + ((size_t)parity * x.world + g) * 2;
        for (int q = 0; q < 2; q++) {
            double2 v = ld_d2_sys(d + q);
            if (__double_as_longlong(v.y) != (long long)epoch) {
                const long long t0 = clock64();
                do { v = ld_d2_sys(d + q); if (clock64() - t0 > LN_WAIT_LIMIT) ln_fail(x.err, 1); } while (__double_as_longlong(v.y) != (long long)epoch);
            }
            if (q == 0) e0 += v.x; else e1 += v.x;
        }
    }
}

// ---- the step phases ----------------------------------------------------------------------------------------------------
// Verlet step (definition of the oracle's sim_dynamics, see ln_md_step_kernel)
__device__ __noinline__ void ln_step_md(const LnCtx& c, const LnOp& op, LnScratch* sc, const float4* A, const float4* Bv, float lam, int first,
                                        float4* a_next, float4* b_next, unsigned int epoch, volatile LnLoop* L, bool timing, unsigned int flow_revals) {
    const int tid = threadIdx.x;
    const float h = op.md.h, dt = op.md.dt_akma, dth = dt * h;
    struct MdPre { float4 p0, q0, pm, qm, pp, qp, v, f; };
    const bool fast_bond = !(op.bd.w_ang > 0.f);
    ln_forces_of_my_blocks<true>(c, sc, epoch, L, timing,
        [&](bool live, int i, int lt) {
            MdPre r;
            r.p0 = r.q0 = r.pm = r.qm = r.pp = r.qp = r.v = r.f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                r.p0 = __ldcg(A + i); r.q0 = __ldcg(Bv + i); r.v = __ldcg(c.v.v + i); r.f = __ldcg(c.v.f + i);
                if (i > 0) { r.pm = __ldcg(A + i - 1); r.qm = __ldcg(Bv + i - 1); }
                if (i + 1 < c.N) { r.pp = __ldcg(A + i + 1); r.qp = __ldcg(Bv + i + 1); }
            }
            return r;
        },
        [&](bool active, int B, int i, int lt, int slot, float fx, float fy, float fz, const MdPre& r) {
        float eb = 0.f, k2 = 0.f;
        if (active && lt >= 0 && i < c.N) {
            const float4 xi = make_float4(fmaf(lam, r.q0.x, r.p0.x), fmaf(lam, r.q0.y, r.p0.y), fmaf(lam, r.q0.z, r.p0.z), 0.f);
            if (fast_bond) {           // the bonds (i-1, i) and (i, i+1) from the preloaded neighbours: same arithmetic as ln_bonded
#pragma unroll
                for (int s = -1; s <= 1; s += 2) {
                    const int nb = i + s; if (nb < 0 || nb >= c.N) continue;
                    const float4 pn = s < 0 ? r.pm : r.pp, qn = s < 0 ? r.qm : r.qp;
                    const float dx = xi.x - fmaf(lam, qn.x, pn.x), dy = xi.y - fmaf(lam, qn.y, pn.y), dz = xi.z - fmaf(lam, qn.z, pn.z);
                    const float r2 = fmaxf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)), 1e-12f), rinv = ln_rsqrt(r2), rr = r2 * rinv, u = rr - op.bd.r0;
                    const float g = 2.f * op.bd.w_bond * u * rinv;
                    fx -= g * dx; fy -= g * dy; fz -= g * dz;
                    if (s == 1) eb = op.bd.w_bond * u * u;
                }
            } else ln_bond_cg(A, Bv, lam, i, c.N, op.bd, fx, fy, fz, eb);
            float4 vv = r.v;
            if (!first) {
                const float4 fo = r.f;
                vv.x = fmaf(h, fo.x, lam * vv.x); vv.y = fmaf(h, fo.y, lam * vv.y); vv.z = fmaf(h, fo.z, lam * vv.z);   // v1
                vv.x = fmaf(h, fx, vv.x); vv.y = fmaf(h, fy, vv.y); vv.z = fmaf(h, fz, vv.z);                          // v2
            }
            k2 = dot3(vv, vv);
            c.v.x[i] = xi; c.v.v[i] = vv; c.v.f[i] = make_float4(fx, fy, fz, 0.f);
            a_next[i] = make_float4(fmaf(dth, fx, xi.x), fmaf(dth, fy, xi.y), fmaf(dth, fz, xi.z), 0.f);
            b_next[i] = make_float4(dt * vv.x, dt * vv.y, dt * vv.z, 0.f);
        }
        const double kb = ln_lead_sum((double)k2, sc->dsum, lt, slot), ebb = ln_lead_sum((double)eb, sc->dsum, lt, slot);
        if (active && lt == 0) { c.blockpart[(size_t)B * LN_KMAX] = kb; c.blockpart[(size_t)B * LN_KMAX + 1] = ebb; ln_publish_block(c, B, epoch, kb, ebb); }
    }, flow_revals);
}
// serial section after ln_step_md (all threads of the last arriver)
__device__ __noinline__ void ln_decide_md(const LnCtx& c, const LnOp& op, LnScratch* sc, int want_e, unsigned int epoch) {
    const int tid = threadIdx.x;
    if (tid >= 32) return;                  // one warp: lane-strided partial sums, then the butterfly (fixed order, no block barrier)
    double a0 = 0, a1 = 0;
    for (int q = tid; q < c.nblk; q += 32) { a0 += __ldcg(c.blockpart + (size_t)q * LN_KMAX); a1 += __ldcg(c.blockpart + (size_t)q * LN_KMAX + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
    if (tid == 0) {
        LnState* st = c.st;
        const double ke = 0.5 * op.md.mass * a0, ndof = 3.0 * c.N - 3.0;
        const double T = 2.0 * ke / (op.md.kboltz * ndof);
        double l2 = 1.0;
        if (T > 0.0) {
            if (op.md.flag == CHR3D_COUPLE_TCOUP) l2 = sqrt(fmax(0.0, 1.0 + 2.0 * op.md.fbeta * op.md.dt_ps * (op.md.temp / T - 1.0)));
            else if (op.md.flag == CHR3D_COUPLE_VSCALE) l2 = sqrt(op.md.temp / T);
        }
        st->ke = ke; st->lam = l2; st->lamf = (float)l2; st->nevals += op.count_eval;
        if (want_e) {
            double e0, e1; ln_energy_collect(c, epoch, e0, e1);
            st->Enoe = e0; st->Evdw = e1; st->Ebond = a1; st->E = e0 + e1 + a1;
        }
    }
}

// L-BFGS trial x = xb + t*p (see ln_min_step_kernel): forces, the 6m+6 inner products of (s, y, g') with the history
__device__ __noinline__ void ln_step_min(const LnCtx& c, const LnOp& op, LnScratch* sc, float lam, int init, unsigned int epoch, volatile LnLoop* L, bool timing) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LnVec& v = c.v;
    const int m = v.m, K = 6 * m + 6, N = v.N;
    const int head = __ldcg(&c.st->head);
    const float4* A = init ? v.x : v.xb;
    ln_forces_of_my_blocks<false>(c, sc, epoch, L, timing, [&](bool, int, int) { return LnNoPre(); },
                                  [&](bool active, int B, int i, int lt, int slot, float fx, float fy, float fz, const LnNoPre&) {
        float eb = 0.f;
        if (tid < LN_TB) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f), y = s, gn = s;
            const bool live = active && i < N;
            if (live) {
                ln_bond_cg(A, v.p, lam, i, N, op.bd, fx, fy, fz, eb);
                const float4 x = ln_ldx_cg(A, v.p, lam, i), xb = v.xb[i], g = v.g[i];
                s = make_float4(x.x - xb.x, x.y - xb.y, x.z - xb.z, 0.f);
                gn = make_float4(-fx, -fy, -fz, 0.f);
                y = make_float4(gn.x - g.x, gn.y - g.y, gn.z - g.z, 0.f);
                v.x[i] = x; v.f[i] = make_float4(fx, fy, fz, 0.f);
            }
            for (int l = 0; l < m; l++) {
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, d4 = 0.f, d5 = 0.f;
                if (live && l != head) {
                    const float4 sl = v.S[(size_t)l * N + i], yl = v.Y[(size_t)l * N + i];
                    d0 = dot3(s, sl); d1 = dot3(s, yl); d2 = dot3(y, sl); d3 = dot3(y, yl); d4 = dot3(gn, sl); d5 = dot3(gn, yl);
                }
                sc->prod[LDOT_L(l, 0)][tid] = d0; sc->prod[LDOT_L(l, 1)][tid] = d1; sc->prod[LDOT_L(l, 2)][tid] = d2;
                sc->prod[LDOT_L(l, 3)][tid] = d3; sc->prod[LDOT_L(l, 4)][tid] = d4; sc->prod[LDOT_L(l, 5)][tid] = d5;
            }
            sc->prod[LDOT_X(m, 0)][tid] = dot3(s, s); sc->prod[LDOT_X(m, 1)][tid] = dot3(s, y); sc->prod[LDOT_X(m, 2)][tid] = dot3(s, gn);
            sc->prod[LDOT_X(m, 3)][tid] = dot3(y, y); sc->prod[LDOT_X(m, 4)][tid] = dot3(y, gn); sc->prod[LDOT_X(m, 5)][tid] = dot3(gn, gn);
        }
        const double ebb = ln_lead_sum((double)eb, sc->dsum, lt, slot);       // (its barriers also publish prod)
        for (int k = warp; k < K; k += LN_P_THREADS / 32) {           // a warp per inner product, fixed order
            double s = ((double)sc->prod[k][lane] + (double)sc->prod[k][lane + 32]) + ((double)sc->prod[k][lane + 64] + (double)sc->prod[k][lane + 96]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) c.blockpart[(size_t)B * LN_KMAX + k] = s;
        }
        if (tid == 0) c.blockpart[(size_t)B * LN_KMAX + K] = ebb;
        __syncthreads();
    });
}
__device__ __noinline__ void ln_decide_min(const LnCtx& c, const LnOp& op, LnScratch* sc, int init, unsigned int epoch) {
    const int tid = threadIdx.x;
    const int K = 6 * c.v.m + 6;
    {   // per-block partials -> totals: 8 slices of blocks per scalar, fixed order
        double* red2 = reinterpret_cast<double*>(sc->prod);          // [8][LN_KMAX]
        const int k = tid & (LN_KMAX - 1), sl = tid / LN_KMAX;
        double s = 0;
        if (k <= K) for (int q = sl; q < c.nblk; q += LN_P_THREADS / LN_KMAX) s += __ldcg(c.blockpart + (size_t)q * LN_KMAX + k);
        red2[sl * LN_KMAX + k] = s;
        __syncthreads();
        if (tid <= K) { double t = 0; for (int q = 0; q < LN_P_THREADS / LN_KMAX; q++) t += red2[q * LN_KMAX + tid]; sc->G[tid] = t; }
    }
    __syncthreads();
    if (tid == 0) {
        double e3[3] = {0, 0, sc->G[K]};
        ln_energy_collect(c, epoch, e3[0], e3[1]);
        ln_min_decide(c.st, sc->G, e3, op.min, init);
        c.st->tstepf = (float)c.st->tstep;
    }
}
// applies the decision to the vectors of this CTA's blocks (ln_min_update_kernel)
__device__ __noinline__ void ln_update_min(const LnCtx& c) {
    const LnState* st = c.st;
    const int acc = __ldcg(&st->accepted);
    if (acc == 0) return;
    const LnVec& v = c.v;
    const int m = v.m, cnt = __ldcg(&st->cnt), head = __ldcg(&st->head), keep = __ldcg(&st->keep);
    const int tid = threadIdx.x, cta = blockIdx.x - c.cta0;
    if (tid >= LN_TB) return;
    float dl[2 * CHR3D_LBFGS_MAX_M + 1];
#pragma unroll
    for (int l = 0; l < 2 * CHR3D_LBFGS_MAX_M + 1; l++) dl[l] = l < 2 * m + 1 ? (float)__ldcg(&st->delta[l]) : 0.f;
    for (int k = 0; k < lnp_bp.nb; k++) {
        const int B = lnp_bp.cached ? lnp_bp.B[k] : c.blk_idx[c.blk_off[cta] + k];
        const int i = B * LN_TB + tid;
        if (i >= v.N) continue;
        if (acc == 4) { v.x[i] = v.xb[i]; continue; }
        if (acc == 5) { const float4 g = v.g[i]; v.p[i] = make_float4(-g.x, -g.y, -g.z, 0.f); continue; }
        const float4 x = v.x[i], f = v.f[i];
        const float4 gn = make_float4(-f.x, -f.y, -f.z, 0.f);
        if (acc == 1 && keep) {
            const int h = ((head - 1) % m + m) % m;      // slot just filled
            const float4 xb = v.xb[i], g = v.g[i];
            v.S[(size_t)h * v.N + i] = make_float4(x.x - xb.x, x.y - xb.y, x.z - xb.z, 0.f);
            v.Y[(size_t)h * v.N + i] = make_float4(gn.x - g.x, gn.y - g.y, gn.z - g.z, 0.f);
        }
        v.xb[i] = x; v.g[i] = gn;
        if (acc == 2 || acc == 3 || cnt == 0) { v.p[i] = make_float4(-gn.x, -gn.y, -gn.z, 0.f); continue; }
        float ax = dl[2 * m] * gn.x, ay = dl[2 * m] * gn.y, az = dl[2 * m] * gn.z;
        for (int kk = 0; kk < cnt; kk++) {
            const int l = ((head - 1 - kk) % m + m) % m;
            const float ds = dl[l], dy = dl[m + l];
            const float4 sl = v.S[(size_t)l * v.N + i], yl = v.Y[(size_t)l * v.N + i];
            ax = fmaf(ds, sl.x, fmaf(dy, yl.x, ax)); ay = fmaf(ds, sl.y, fmaf(dy, yl.y, ay)); az = fmaf(ds, sl.z, fmaf(dy, yl.z, az));
        }
        v.p[i] = make_float4(-ax, -ay, -az, 0.f);
    }
}

// energies + forces at explicit coordinates (SAVE / PICK / ENERGY ops)
__device__ __noinline__ void ln_step_eval(const LnCtx& c, const LnOp& op, LnScratch* sc, unsigned int epoch, volatile LnLoop* L, bool timing) {
    const int tid = threadIdx.x;
    ln_forces_of_my_blocks<true>(c, sc, epoch, L, timing, [&](bool, int, int) { return LnNoPre(); },
                                 [&](bool active, int B, int i, int lt, int slot, float fx, float fy, float fz, const LnNoPre&) {
        float eb = 0.f;
        if (active && lt >= 0 && i < c.N) {
            ln_bond_cg(c.v.x, c.v.x, 0.f, i, c.N, op.bd, fx, fy, fz, eb);
            c.v.f[i] = make_float4(fx, fy, fz, 0.f);
        }
        const double ebb = ln_lead_sum((double)eb, sc->dsum, lt, slot);
        if (active && lt == 0) c.blockpart[(size_t)B * LN_KMAX] = ebb;
    });
}
__device__ __noinline__ void ln_decide_eval(const LnCtx& c, const LnOp& op, LnScratch* sc, unsigned int epoch) {
    const int tid = threadIdx.x;
    double a0 = 0;
    for (int q = tid; q < c.nblk; q += LN_P_THREADS) a0 += __ldcg(c.blockpart + (size_t)q * LN_KMAX);
    double* red = reinterpret_cast<double*>(sc->prod);
    red[tid] = a0;
    __syncthreads();
    for (int o = LN_P_THREADS / 2; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
    if (tid == 0) {
        LnState* st = c.st;
        double e0, e1; ln_energy_collect(c, epoch, e0, e1);
        st->Enoe = e0; st->Evdw = e1; st->Ebond = red[0]; st->E = e0 + e1 + red[0]; st->nevals += op.count_eval;
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------
// grid barrier; `serial` runs in the last arriver (all its threads) before the others are released
template <typename Serial>
__device__ __forceinline__ void ln_barrier(const LnCtx& c, volatile LnLoop* L, int* s_flag, Serial serial) {
    const unsigned int gen = L->gen;
    if (ln_bar_arrive(c, gen, s_flag)) { serial(); ln_bar_release(c, gen); }
    __syncthreads();
    if (threadIdx.x == 0) L->gen = gen + 1u;
    __syncthreads();
}

// The state of a persistent launch, in shared memory at file scope: the out-of-line phases below take NO arguments, so the
// kernel body keeps nothing alive in registers across the inlined tile phase (which needs all of them).
#define LNP_SC (reinterpret_cast<LnScratch*>(lnp_smem + LN_COLP_OFF))
#define LNP_TIMING (blockIdx.x == (unsigned)lnp_c.cta0 && threadIdx.x == 0)

// What follows a tile phase: first barrier, step phase, second barrier with the decisions, loop state.
__device__ __noinline__ void ln_md_rest() {
    const LnCtx& c = lnp_c; const LnOp& op = lnp_op; LnScratch* sc = LNP_SC; volatile LnLoop* L = &lnp_L; int* s_flag = &lnp_flag;
    const bool timing = LNP_TIMING;
    if (threadIdx.x == 0) L->ttile += ln_globaltimer() - L->t0tile;
    const int step = L->step, par = L->par;
    const int first = step < 0, want_e = !first && step == op.nstep - 1;
    ln_tprof(L, 0, timing);
    const unsigned int epoch = L->epoch + 1u;
    ln_barrier(c, L, s_flag, [&]() { if (want_e) ln_energy_publish(c, sc, epoch); });
    ln_tprof(L, 1, timing);
    {
        const float4* A = first ? c.v.x : c.va[par]; const float4* Bv = first ? c.v.x : c.vb[par];
        const float lam = first ? 0.f : L->lam;
        ln_step_md(c, op, sc, A, Bv, lam, first, first ? c.va[0] : c.va[par ^ 1], first ? c.vb[0] : c.vb[par ^ 1], epoch, L, timing, 0u);
    }
    ln_tprof(L, 2, timing);
    ln_barrier(c, L, s_flag, [&]() { ln_decide_md(c, op, sc, want_e, epoch); });
    if (threadIdx.x == 0) { L->epoch = epoch; L->step = step + 1; if (!first) L->par = par ^ 1; }
    __syncthreads();
    ln_tprof(L, 3, timing);
}
// the same for an evaluation in the middle of an MD op: no grid barrier (see "barrier-free evaluations")
__device__ __noinline__ void ln_md_rest_flow() {
    const LnCtx& c = lnp_c; const LnOp& op = lnp_op; LnScratch* sc = LNP_SC; volatile LnLoop* L = &lnp_L;
    const bool timing = LNP_TIMING;
    if (threadIdx.x == 0) L->ttile += ln_globaltimer() - L->t0tile;
    LN_TRACE(3);
    const int step = L->step, par = L->par;
    ln_tprof(L, 0, timing);
    const unsigned int epoch = L->epoch + 1u;
    ln_step_md(c, op, sc, c.va[par], c.vb[par], L->lam, 0, c.va[par ^ 1], c.vb[par ^ 1], epoch, L, timing, L->revals);
    if (threadIdx.x == 0) { L->epoch = epoch; L->step = step + 1; L->par = par ^ 1; if (timing) c.st->nevals += op.count_eval; }
    __syncthreads();
    ln_tprof(L, 2, timing);
}
__device__ __noinline__ void ln_min_rest() {
    const LnCtx& c = lnp_c; const LnOp& op = lnp_op; LnScratch* sc = LNP_SC; volatile LnLoop* L = &lnp_L; int* s_flag = &lnp_flag;
    const bool timing = LNP_TIMING;
    if (threadIdx.x == 0) L->ttile += ln_globaltimer() - L->t0tile;
    const int step = L->step, init = step < 0;
    ln_tprof(L, 0, timing);
    const unsigned int epoch = L->epoch + 1u;
    ln_barrier(c, L, s_flag, [&]() { ln_energy_publish(c, sc, epoch); });
    ln_tprof(L, 1, timing);
    ln_step_min(c, op, sc, init ? 0.f : __ldcg(&c.st->tstepf), init, epoch, L, timing);
    ln_tprof(L, 2, timing);
    ln_barrier(c, L, s_flag, [&]() { ln_decide_min(c, op, sc, init, epoch); });
    ln_tprof(L, 3, timing);
    ln_update_min(c);
    ln_barrier(c, L, s_flag, [&]() {});
    if (threadIdx.x == 0) { L->epoch = epoch; L->step = step + 1; L->stop = (step + 1 >= (int)op.max_trials || __ldcg(&c.st->done)) ? 1 : 0; }
    __syncthreads();
    ln_tprof(L, 4, timing);
}
__device__ __noinline__ void ln_eval_rest() {
    const LnCtx& c = lnp_c; const LnOp& op = lnp_op; LnScratch* sc = LNP_SC; volatile LnLoop* L = &lnp_L; int* s_flag = &lnp_flag;
    const bool timing = LNP_TIMING;
    if (threadIdx.x == 0) L->ttile += ln_globaltimer() - L->t0tile;
    ln_tprof(L, 0, timing);
    const unsigned int epoch = L->epoch + 1u;
    ln_barrier(c, L, s_flag, [&]() { ln_energy_publish(c, sc, epoch); });
    ln_tprof(L, 1, timing);
    ln_step_eval(c, op, sc, epoch, L, timing);
    ln_tprof(L, 2, timing);
    ln_barrier(c, L, s_flag, [&]() { ln_decide_eval(c, op, sc, epoch); });
    if (threadIdx.x == 0) L->epoch = epoch;
    __syncthreads();
    ln_tprof(L, 3, timing);
}
// coordinates of the next evaluation -> lnp_run, start of the tile clock (all threads call)
__device__ __noinline__ void ln_md_setup() {
    const LnCtx& c = lnp_c; volatile LnLoop* L = &lnp_L; LnTileRun* run = &lnp_run;
    const int step = L->step, first = step < 0;
    if (threadIdx.x < 32) {
        // every evaluation but the first continues from an integration: wait for the scalars of all bead blocks of that evaluation
        // (every CTA adds them up in the same order, so every CTA gets the same factor; no grid barrier needed for it)
        float lam = 0.f; double ke = 0.0;
        LN_TRACE(1);
        if (!first) lam = ln_lambda_from_slots(c, lnp_op, L->epoch, &ke);
        LN_TRACE(2);
        if (threadIdx.x == 0) {
            const int par = L->par;
            const int more = step + 1 < lnp_op.nstep ? 1 : 0;
            run->X = first ? c.v.x : c.va[par]; run->XB = first ? c.v.x : c.vb[par];
            run->lam = lam; L->lam = lam;
            run->more = lnp_op.defer_ring ? 0 : more; L->prefetch = lnp_op.defer_ring ? more : 0;
            L->revals = L->revals + 1u;
            if (!first && LNP_TIMING) { c.st->ke = ke; c.st->lam = (double)lam; c.st->lamf = lam; }
            L->t0tile = ln_globaltimer();
        }
    }
    __syncthreads();
}
__device__ __noinline__ void ln_min_setup() {
    if (threadIdx.x == 0) {
        const LnCtx& c = lnp_c; LnTileRun* run = &lnp_run;
        const int init = lnp_L.step < 0;
        run->X = init ? c.v.x : c.v.xb; run->XB = init ? c.v.x : c.v.p;
        run->lam = init ? 0.f : __ldcg(&c.st->tstepf);
        run->more = lnp_op.defer_ring ? 0 : 1; lnp_L.prefetch = lnp_op.defer_ring ? 1 : 0;
        lnp_L.revals = lnp_L.revals + 1u;
        lnp_L.t0tile = ln_globaltimer();
    }
    __syncthreads();
}
__device__ __noinline__ void ln_launch_setup(const LnCtx* __restrict__ ctxs, int nvr, const LnOp* op) {
    {   // which (virtual) rank does this CTA belong to
        int vr = 0;
        while (vr + 1 < nvr && (int)blockIdx.x >= ctxs[vr + 1].cta0) vr++;
        const int* src = reinterpret_cast<const int*>(ctxs + vr); int* dst = reinterpret_cast<int*>(&lnp_c);
        for (int k = threadIdx.x; k < (int)(sizeof(LnCtx) / 4); k += LN_P_THREADS) dst[k] = src[k];
        const int* so = reinterpret_cast<const int*>(op); int* sd = reinterpret_cast<int*>(&lnp_op);
        for (int k = threadIdx.x; k < (int)(sizeof(LnOp) / 4); k += LN_P_THREADS) sd[k] = so[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const LnCtx& c = lnp_c;
        ln_run_setup(&lnp_run, c.Dt, c.t0, c.cstart, c.nT, c.N, c.colpart, c.rowpart, c.epart, c.maxseg, blockIdx.x - c.cta0);
        lnp_run.terms = lnp_op.terms; lnp_run.dbg = lnp_op.dbg; lnp_run.stream_l2 = c.stream_l2; lnp_run.ring_tail = lnp_op.ring_tail;
        lnp_L.step = -1; lnp_L.par = 0; lnp_L.stop = 0;
        lnp_L.gen = ld_acquire_gpu(c.bar + 32);             // stable between launches
        lnp_L.epoch = c.st->epoch; lnp_L.revals = c.st->revals; lnp_L.lam = 0.f;
        for (int q = 0; q < 8; q++) lnp_L.tacc[q] = 0;
        lnp_L.tlast = ln_globaltimer(); lnp_L.ttile = 0;
    }
#ifdef CHR3D_DEV
    if (threadIdx.x == 0) lnp_trace_n = 0;
#endif
    if (threadIdx.x == 32) ln_blk_plan_load(lnp_c, &lnp_bp);
    __syncthreads();
    ln_ring_init(&lnp_run, lnp_smem, true);
}
__device__ __noinline__ void ln_launch_finish() {
    ln_ring_drain(&lnp_run, lnp_smem);
    if (threadIdx.x == 0) lnp_c.cta_ns[blockIdx.x - lnp_c.cta0] += lnp_L.ttile;
    if (LNP_TIMING) { lnp_c.st->epoch = lnp_L.epoch; lnp_c.st->revals = lnp_L.revals; for (int q = 0; q < 8; q++) lnp_c.st->tprof[q] += lnp_L.tacc[q]; }
}

template <bool LOW_SOFT>
__global__ void __launch_bounds__(LN_P_THREADS, 1) ln_persist_kernel(const LnCtx* __restrict__ ctxs, int nvr, const __grid_constant__ LnOp op) {
    ln_launch_setup(ctxs, nvr, &op);
    if (op.kind == LN_OP_MD) {
        // the first evaluation (nothing to integrate yet) with the grid barriers ...
        ln_md_setup();
        ln_tile_phase<false, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
        ln_signal_tiles_done(lnp_c);
        ln_md_rest();
        // ... the evaluations in the middle as a dataflow, no grid barrier ...
        while (((volatile LnLoop*)&lnp_L)->step < op.nstep - 1) {
            ln_md_setup();
            ln_tile_phase<false, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
            ln_signal_tiles_done(lnp_c);
            if (op.flow) ln_md_rest_flow(); else ln_md_rest();
        }
        // ... the last one with energies and the barriers again
        if (op.nstep > 0) {
            ln_md_setup();
            ln_tile_phase<true, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
            ln_signal_tiles_done(lnp_c);
            ln_md_rest();
        }
    } else if (op.kind == LN_OP_MIN) {
        while (!((volatile LnLoop*)&lnp_L)->stop) {
            ln_min_setup();
            ln_tile_phase<true, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
            ln_signal_tiles_done(lnp_c);
            ln_min_rest();
        }
    } else {
        if (threadIdx.x == 0) { lnp_run.X = lnp_c.v.x; lnp_run.XB = lnp_c.v.x; lnp_run.lam = 0.f; lnp_run.more = 0; lnp_L.prefetch = 0; lnp_L.revals = lnp_L.revals + 1u; lnp_L.t0tile = ln_globaltimer(); }
        __syncthreads();
        ln_tile_phase<true, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
        ln_signal_tiles_done(lnp_c);
        ln_eval_rest();
    }
    ln_launch_finish();
}
