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
    int gI0, gJ0, gI1, gJ1;                 // tile coordinates of t0 and t1: the owned tiles of a column are a contiguous row range
    LnVec v; float4* va[2]; float4* vb[2];
    LnState* st; double* blockpart;
    unsigned int* bar;                      // [0] arrivals, [32] generation (separate 128-byte lines)
    LnXchg xc; int bmin[LN_MAXWORLD];       // bmin[g]: first bead block rank g contributes to (its first tile row)
};
struct LnOp {
    int kind, nstep, count_eval, dbg;
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
struct LnLoop { int step, par, stop; unsigned int gen, epoch; unsigned long long tlast, tacc[8], t0tile, ttile; };
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

// this rank's partial force on bead (B*128 + tid), valid in threads tid < 128; all threads call.  Fixed summation order.
__device__ __forceinline__ void ln_gather4(const LnCtx& c, int B, LnScratch* sc, float& fx, float& fy, float& fz) {
    const int tid = threadIdx.x;
    int lo = (B >= c.gJ0) ? c.gI0 : c.gI0 + 1, hi = (B < c.gJ1) ? c.gI1 + 1 : c.gI1;
    if (hi > B + 1) hi = B + 1;
    if (tid < LN_P_GGRP * 96) {
        const int q = tid % 96, grp = tid / 96;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        const float4* base = reinterpret_cast<const float4*>(c.colpart) + q;
        int I = lo + grp;
#define LN_GLD(u, k) const float4 u = __ldcg(base + (size_t)(tile_index(I + (k) * LN_P_GGRP, B, c.nT) - c.t0) * 96)
#define LN_GADD(a, u) a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w
        for (; I + 15 * LN_P_GGRP < hi; I += 16 * LN_P_GGRP) {     // sixteen independent 16-byte loads in flight per thread: the gather is
            LN_GLD(u0, 0); LN_GLD(u1, 1); LN_GLD(u2, 2); LN_GLD(u3, 3); LN_GLD(u4, 4); LN_GLD(u5, 5); LN_GLD(u6, 6); LN_GLD(u7, 7);     // bound by dependent round trips to L2
            LN_GLD(u8, 8); LN_GLD(u9, 9); LN_GLD(u10, 10); LN_GLD(u11, 11); LN_GLD(u12, 12); LN_GLD(u13, 13); LN_GLD(u14, 14); LN_GLD(u15, 15);
            LN_GADD(a0, u0); LN_GADD(a1, u1); LN_GADD(a2, u2); LN_GADD(a3, u3); LN_GADD(a0, u4); LN_GADD(a1, u5); LN_GADD(a2, u6); LN_GADD(a3, u7);
            LN_GADD(a0, u8); LN_GADD(a1, u9); LN_GADD(a2, u10); LN_GADD(a3, u11); LN_GADD(a0, u12); LN_GADD(a1, u13); LN_GADD(a2, u14); LN_GADD(a3, u15);
        }
        for (; I + 7 * LN_P_GGRP < hi; I += 8 * LN_P_GGRP) {
            LN_GLD(u0, 0); LN_GLD(u1, 1); LN_GLD(u2, 2); LN_GLD(u3, 3); LN_GLD(u4, 4); LN_GLD(u5, 5); LN_GLD(u6, 6); LN_GLD(u7, 7);
            LN_GADD(a0, u0); LN_GADD(a1, u1); LN_GADD(a2, u2); LN_GADD(a3, u3); LN_GADD(a0, u4); LN_GADD(a1, u5); LN_GADD(a2, u6); LN_GADD(a3, u7);
        }
        for (; I + 3 * LN_P_GGRP < hi; I += 4 * LN_P_GGRP) {
            LN_GLD(u0, 0); LN_GLD(u1, 1); LN_GLD(u2, 2); LN_GLD(u3, 3);
            LN_GADD(a0, u0); LN_GADD(a1, u1); LN_GADD(a2, u2); LN_GADD(a3, u3);
        }
        for (; I < hi; I += LN_P_GGRP) { LN_GLD(u0, 0); LN_GADD(a0, u0); }
#undef LN_GLD
#undef LN_GADD
        a0.x = (a0.x + a1.x) + (a2.x + a3.x); a0.y = (a0.y + a1.y) + (a2.y + a3.y); a0.z = (a0.z + a1.z) + (a2.z + a3.z); a0.w = (a0.w + a1.w) + (a2.w + a3.w);
        *reinterpret_cast<float4*>(&sc->gred[grp][q * 4]) = a0;
    }
    __syncthreads();
    fx = fy = fz = 0.f;
    if (tid < LN_TB) {
        const int pos = 4 * (tid & 31) + (tid >> 5);            // column -> position inside a partial (ln_tile_phase)
#pragma unroll
        for (int g = 0; g < LN_P_GGRP; g++) { fx += sc->gred[g][pos]; fy += sc->gred[g][LN_TB + pos]; fz += sc->gred[g][2 * LN_TB + pos]; }
        for (int k = c.rs_off[B]; k < c.rs_off[B + 1]; k++) {
            const float4 v = __ldcg(c.rowpart + (size_t)c.rs_idx[k] * LN_TB + tid);
            fx += v.x; fy += v.y; fz += v.z;
        }
    }
    __syncthreads();                                           // gred may be overwritten by the next block
}

// total force on the beads of this CTA's blocks: gather, exchange over the ranks.  `emit(B, i, fx, fy, fz)` is called by the
// 128 lead threads of every block (i = bead index, may be >= N) with all threads of the CTA participating (block-wide
// reductions inside emit are allowed).
template <typename Emit>
__device__ __forceinline__ void ln_forces_of_my_blocks(const LnCtx& c, LnScratch* sc, unsigned int epoch, volatile LnLoop* L, bool timing, Emit emit) {
    const int tid = threadIdx.x, cta = blockIdx.x - c.cta0;
    const LnXchg& x = c.xc;
    if (x.world <= 1) {
        for (int bk = c.blk_off[cta], B = 0; bk < c.blk_off[cta + 1] && ((B = c.blk_idx[bk]), true); bk++) {
            float fx, fy, fz;
            ln_gather4(c, B, sc, fx, fy, fz);
            emit(B, B * LN_TB + tid, fx, fy, fz);
        }
        return;
    }
    const int parity = (int)(epoch & 1u);
    const size_t slot = ((size_t)parity * x.world + x.rank) * x.n4;
    // push: this rank's partial of every block it contributes to -> slot `rank` of every rank (self included), the epoch in .w
    for (int bk = c.blk_off[cta], B = 0; bk < c.blk_off[cta + 1] && ((B = c.blk_idx[bk]), true); bk++) {
        if (B < c.bmin[x.rank]) continue;
        float fx, fy, fz;
        ln_gather4(c, B, sc, fx, fy, fz);
        const int i = B * LN_TB + tid;
        if (tid < LN_TB && i < c.N) {
            const float4 v = make_float4(fx, fy, fz, __uint_as_float(epoch));
            for (int g = 0; g < x.world; g++) st_v4_sys(x.recv[g] + slot + i, v);
        }
    }
    ln_tprof(L, 5, timing);
    // poll the slots of the contributors of each block until they carry this epoch, add them in rank order
    for (int bk = c.blk_off[cta], B = 0; bk < c.blk_off[cta + 1] && ((B = c.blk_idx[bk]), true); bk++) {
        const int i = B * LN_TB + tid;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        if (tid < LN_TB && i < c.N) {
            const float4* R = x.recv[x.rank] + (size_t)parity * x.world * x.n4 + i;
            for (int g = 0; g < x.world; g++) if (B >= c.bmin[g]) {
                float4 v = ld_v4_sys(R + (size_t)g * x.n4);
                if (__float_as_uint(v.w) != epoch) {
                    const long long t0 = clock64();
                    do { v = ld_v4_sys(R + (size_t)g * x.n4); if (clock64() - t0 > LN_WAIT_LIMIT) ln_fail(x.err, 1); } while (__float_as_uint(v.w) != epoch);
                }
                fx += v.x; fy += v.y; fz += v.z;
            }
        }
        if (bk + 1 == c.blk_off[cta + 1]) ln_tprof(L, 6, timing);
        emit(B, i, fx, fy, fz);
    }
}

// sum of `v` over the 128 lead threads (4 warps), result in every thread; all threads call; fixed order
__device__ __forceinline__ double ln_lead_sum(double v, double* sh4) {
    const int tid = threadIdx.x;
    if (tid < LN_TB) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) sh4[tid >> 5] = v;
    }
    __syncthreads();
    const double s = (sh4[0] + sh4[1]) + (sh4[2] + sh4[3]);
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
                                        float4* a_next, float4* b_next, unsigned int epoch, volatile LnLoop* L, bool timing) {
    const int tid = threadIdx.x;
    const float h = op.md.h, dt = op.md.dt_akma, dth = dt * h;
    ln_forces_of_my_blocks(c, sc, epoch, L, timing, [&](int B, int i, float fx, float fy, float fz) {
        float eb = 0.f, k2 = 0.f;
        if (tid < LN_TB && i < c.N) {
            ln_bond_cg(A, Bv, lam, i, c.N, op.bd, fx, fy, fz, eb);
            const float4 xi = ln_ldx_cg(A, Bv, lam, i);
            float4 vv = c.v.v[i];
            if (!first) {
                const float4 fo = c.v.f[i];
                vv.x = fmaf(h, fo.x, lam * vv.x); vv.y = fmaf(h, fo.y, lam * vv.y); vv.z = fmaf(h, fo.z, lam * vv.z);   // v1
                vv.x = fmaf(h, fx, vv.x); vv.y = fmaf(h, fy, vv.y); vv.z = fmaf(h, fz, vv.z);                          // v2
            }
            k2 = dot3(vv, vv);
            c.v.x[i] = xi; c.v.v[i] = vv; c.v.f[i] = make_float4(fx, fy, fz, 0.f);
            a_next[i] = make_float4(fmaf(dth, fx, xi.x), fmaf(dth, fy, xi.y), fmaf(dth, fz, xi.z), 0.f);
            b_next[i] = make_float4(dt * vv.x, dt * vv.y, dt * vv.z, 0.f);
        }
        const double kb = ln_lead_sum((double)k2, sc->dsum), ebb = ln_lead_sum((double)eb, sc->dsum);
        if (tid == 0) { c.blockpart[(size_t)B * LN_KMAX] = kb; c.blockpart[(size_t)B * LN_KMAX + 1] = ebb; }
    });
}
// serial section after ln_step_md (all threads of the last arriver)
__device__ __noinline__ void ln_decide_md(const LnCtx& c, const LnOp& op, LnScratch* sc, int want_e, unsigned int epoch) {
    const int tid = threadIdx.x;
    double a0 = 0, a1 = 0;
    for (int q = tid; q < c.nblk; q += LN_P_THREADS) { a0 += __ldcg(c.blockpart + (size_t)q * LN_KMAX); a1 += __ldcg(c.blockpart + (size_t)q * LN_KMAX + 1); }
    double* red = reinterpret_cast<double*>(sc->prod);
    red[tid] = a0; red[LN_P_THREADS + tid] = a1;
    __syncthreads();
    for (int o = LN_P_THREADS / 2; o > 0; o >>= 1) { if (tid < o) { red[tid] += red[tid + o]; red[LN_P_THREADS + tid] += red[LN_P_THREADS + tid + o]; } __syncthreads(); }
    if (tid == 0) {
        LnState* st = c.st;
        const double ke = 0.5 * op.md.mass * red[0], ndof = 3.0 * c.N - 3.0;
        const double T = 2.0 * ke / (op.md.kboltz * ndof);
        double l2 = 1.0;
        if (T > 0.0) {
            if (op.md.flag == CHR3D_COUPLE_TCOUP) l2 = sqrt(fmax(0.0, 1.0 + 2.0 * op.md.fbeta * op.md.dt_ps * (op.md.temp / T - 1.0)));
            else if (op.md.flag == CHR3D_COUPLE_VSCALE) l2 = sqrt(op.md.temp / T);
        }
        st->ke = ke; st->lam = l2; st->lamf = (float)l2; st->nevals += op.count_eval;
        if (want_e) {
            double e0, e1; ln_energy_collect(c, epoch, e0, e1);
            st->Enoe = e0; st->Evdw = e1; st->Ebond = red[LN_P_THREADS]; st->E = e0 + e1 + red[LN_P_THREADS];
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
    ln_forces_of_my_blocks(c, sc, epoch, L, timing, [&](int B, int i, float fx, float fy, float fz) {
        float eb = 0.f;
        if (tid < LN_TB) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f), y = s, gn = s;
            const bool live = i < N;
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
        const double ebb = ln_lead_sum((double)eb, sc->dsum);       // (its barriers also publish prod)
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
    for (int bk = c.blk_off[cta], B = 0; bk < c.blk_off[cta + 1] && ((B = c.blk_idx[bk]), true); bk++) {
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
    ln_forces_of_my_blocks(c, sc, epoch, L, timing, [&](int B, int i, float fx, float fy, float fz) {
        float eb = 0.f;
        if (tid < LN_TB && i < c.N) {
            ln_bond_cg(c.v.x, c.v.x, 0.f, i, c.N, op.bd, fx, fy, fz, eb);
            c.v.f[i] = make_float4(fx, fy, fz, 0.f);
        }
        const double ebb = ln_lead_sum((double)eb, sc->dsum);
        if (tid == 0) c.blockpart[(size_t)B * LN_KMAX] = ebb;
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
__shared__ LnCtx lnp_c;
__shared__ LnOp lnp_op;
__shared__ LnTileRun lnp_run;
__shared__ LnLoop lnp_L;
__shared__ int lnp_flag;
extern __shared__ __align__(128) unsigned char lnp_smem[];
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
        const float lam = first ? 0.f : __ldcg(&c.st->lamf);
        ln_step_md(c, op, sc, A, Bv, lam, first, first ? c.va[0] : c.va[par ^ 1], first ? c.vb[0] : c.vb[par ^ 1], epoch, L, timing);
    }
    ln_tprof(L, 2, timing);
    ln_barrier(c, L, s_flag, [&]() { ln_decide_md(c, op, sc, want_e, epoch); });
    if (threadIdx.x == 0) { L->epoch = epoch; L->step = step + 1; if (!first) L->par = par ^ 1; }
    __syncthreads();
    ln_tprof(L, 3, timing);
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
    if (threadIdx.x == 0) {
        const LnCtx& c = lnp_c; volatile LnLoop* L = &lnp_L; LnTileRun* run = &lnp_run;
        const int step = L->step, par = L->par, first = step < 0;
        run->X = first ? c.v.x : c.va[par]; run->XB = first ? c.v.x : c.vb[par];
        run->lam = first ? 0.f : __ldcg(&c.st->lamf); run->more = step + 1 < lnp_op.nstep;
        L->t0tile = ln_globaltimer();
    }
    __syncthreads();
}
__device__ __noinline__ void ln_min_setup() {
    if (threadIdx.x == 0) {
        const LnCtx& c = lnp_c; LnTileRun* run = &lnp_run;
        const int init = lnp_L.step < 0;
        run->X = init ? c.v.x : c.v.xb; run->XB = init ? c.v.x : c.v.p;
        run->lam = init ? 0.f : __ldcg(&c.st->tstepf); run->more = 1;
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
        lnp_run.terms = lnp_op.terms; lnp_run.dbg = lnp_op.dbg; lnp_run.stream_l2 = c.stream_l2;
        lnp_L.step = -1; lnp_L.par = 0; lnp_L.stop = 0;
        lnp_L.gen = ld_acquire_gpu(c.bar + 32);             // stable between launches
        lnp_L.epoch = c.st->epoch;
        for (int q = 0; q < 8; q++) lnp_L.tacc[q] = 0;
        lnp_L.tlast = ln_globaltimer(); lnp_L.ttile = 0;
    }
    __syncthreads();
    ln_ring_init(&lnp_run, lnp_smem, true);
}
__device__ __noinline__ void ln_launch_finish() {
    ln_ring_drain(&lnp_run, lnp_smem);
    if (threadIdx.x == 0) lnp_c.cta_ns[blockIdx.x - lnp_c.cta0] += lnp_L.ttile;
    if (LNP_TIMING) { lnp_c.st->epoch = lnp_L.epoch; for (int q = 0; q < 8; q++) lnp_c.st->tprof[q] += lnp_L.tacc[q]; }
}

template <bool LOW_SOFT>
__global__ void __launch_bounds__(LN_P_THREADS, 1) ln_persist_kernel(const LnCtx* __restrict__ ctxs, int nvr, const __grid_constant__ LnOp op) {
    ln_launch_setup(ctxs, nvr, &op);
    if (op.kind == LN_OP_MD) {
        // every evaluation but the last one without energies
        while (((volatile LnLoop*)&lnp_L)->step < 0 || ((volatile LnLoop*)&lnp_L)->step < op.nstep - 1) {
            ln_md_setup();
            ln_tile_phase<false, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
            ln_md_rest();
        }
        if (op.nstep > 0) {
            ln_md_setup();
            ln_tile_phase<true, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
            ln_md_rest();
        }
    } else if (op.kind == LN_OP_MIN) {
        while (!((volatile LnLoop*)&lnp_L)->stop) {
            ln_min_setup();
            ln_tile_phase<true, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
            ln_min_rest();
        }
    } else {
        if (threadIdx.x == 0) { lnp_run.X = lnp_c.v.x; lnp_run.XB = lnp_c.v.x; lnp_run.lam = 0.f; lnp_run.more = 0; lnp_L.t0tile = ln_globaltimer(); }
        __syncthreads();
        ln_tile_phase<true, LOW_SOFT>(&lnp_run, lnp_smem, op.terms);
        ln_eval_rest();
    }
    ln_launch_finish();
}
