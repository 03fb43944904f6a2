// seqsum.cuh — exact left-to-right fp64 summation (reproduces the rounding sequence of a scalar
// CPU loop, which the Perl's `$sum += ...` loops are).  One block per segment.  Two kernels with the same result:
//   seq_sum_exact_kernel   the sum evaluated in parallel, bit for bit (integer prefix scans inside a binade, seqsum_core.h)
//   seq_sum_staged_kernel  the serial chain: the block stages chunks through shared memory with coalesced loads while thread 0
//                          adds them in order (4.4 ns per term; kept as the cross-check and behind chr3d_set_option("seqsum", "serial"))
#pragma once
#include <limits.h>
#include "seqsum_core.h"

namespace chr3d {

#define SEQSUM_CHUNK 2048

static __global__ void __launch_bounds__(256) seq_sum_staged_kernel(const double* __restrict__ src_all, long long n_per,
                                                                    double* __restrict__ out) {
    __shared__ double buf[2][SEQSUM_CHUNK];
    const double* src = src_all + (size_t)blockIdx.x * n_per;
    const long long nchunks = (n_per + SEQSUM_CHUNK - 1) / SEQSUM_CHUNK;
    double s = 0.0;
    // chunk 0
    for (int k = threadIdx.x; k < SEQSUM_CHUNK; k += blockDim.x) buf[0][k] = (long long)k < n_per ? src[k] : 0.0;
    __syncthreads();
    for (long long c = 0; c < nchunks; c++) {
        const int cur = (int)(c & 1);
        if (threadIdx.x >= 32) {                 // warps 1.. prefetch the next chunk
            if (c + 1 < nchunks) {
                const long long base = (c + 1) * SEQSUM_CHUNK;
                for (int k = threadIdx.x - 32; k < SEQSUM_CHUNK; k += blockDim.x - 32)
                    buf[cur ^ 1][k] = base + k < n_per ? src[base + k] : 0.0;
            }
        } else if (threadIdx.x == 0) {           // exact order; padding zeros do not change the sum
            const double* b = buf[cur];
            long long left = n_per - c * SEQSUM_CHUNK; if (left > SEQSUM_CHUNK) left = SEQSUM_CHUNK;
            int k = 0;
            for (; k + 8 <= (int)left; k += 8) {
                double a0 = b[k], a1 = b[k + 1], a2 = b[k + 2], a3 = b[k + 3], a4 = b[k + 4], a5 = b[k + 5], a6 = b[k + 6], a7 = b[k + 7];
                s += a0; s += a1; s += a2; s += a3; s += a4; s += a5; s += a6; s += a7;
            }
            for (; k < (int)left; k++) s += b[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// ---- the same sum, in parallel ------------------------------------------------------------------------------------------
// Per pass the block looks at a window of SEQSUM_THREADS x SEQSUM_PER_THREAD terms behind the current position: every thread
// turns its terms into integer contributions in units of the running sum's last place (seqsum_contrib_fp: three fp64 additions
// per term; seqsum_contrib states the same thing in integer shifts), a block-wide prefix scan gives the sum after every term,
// and the first term that cannot be folded in (a tie, a carry out of the binade, a negative / non-finite term, a term of a
// quarter of the running sum or more, or no valid frame yet) ends the pass: everything before it is added at once, that term
// with a real fp64 add by the thread that owns it.  Real data (positive, unremarkable) needs one short pass per doubling of the
// sum and one per full window after that: 36 passes for the 207,025 cells of chr1_500kb at alpha = 1.1 (0.92 ms as a serial
// chain; 0.24 ms with the integer contributions, less with the three additions: profiles/r02j_launches_e2e.md,
// r02k_e2e_ab.json).  Inputs that stop all the time (denormal sums, dyadic values that tie at every step) would pay a pass per
// term, so a pass that folded in fewer than 32 terms doubles the serial run that follows its stop (up to 4,096 terms; any
// longer pass resets it to the one stopping term): such inputs degrade to the serial chain, not below it.
#define SEQSUM_THREADS 1024
#define SEQSUM_PER_THREAD 8
#define SEQSUM_WINDOW (SEQSUM_THREADS * SEQSUM_PER_THREAD)
#define SEQSUM_MAX_RUN 4096

static __global__ void __launch_bounds__(SEQSUM_THREADS) seq_sum_exact_kernel(const double* __restrict__ src_all, long long n_per,
                                                                              double* __restrict__ out) {
    const double* src = src_all + (size_t)blockIdx.x * n_per;
    __shared__ double sh_s;
    __shared__ long long sh_pos;
    __shared__ int sh_run;
    __shared__ unsigned long long sh_wtot[32];
    __shared__ long long sh_kmin[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { sh_s = 0.0; sh_pos = 0; sh_run = 1; }
    __syncthreads();
    long long passes = 0;
    while (true) {
        if (++passes > n_per + 8) __trap();        // every pass consumes at least one term: more passes than terms is a bug, not a wait
        const long long pos = sh_pos;
        const double s = sh_s;
        const int run = sh_run;
        if (pos >= n_per) break;
        const SeqSumFrame f = seqsum_frame(s);
        const SeqSumFpFrame g = seqsum_fp_frame(f);
        const long long end = pos + SEQSUM_WINDOW < n_per ? pos + SEQSUM_WINDOW : n_per;
        const long long k0 = pos + (long long)tid * SEQSUM_PER_THREAD;
        // all loads of the pass first (behind the branches of the classification they would be issued one L2 round trip after the
        // other: 5 us per pass instead of 1.5); a term behind the end reads as +0, which contributes nothing and never stops
        double a[SEQSUM_PER_THREAD];
#pragma unroll
        for (int j = 0; j < SEQSUM_PER_THREAD; j++) a[j] = k0 + j < end ? __ldg(src + k0 + j) : 0.0;
        unsigned long long c[SEQSUM_PER_THREAD];
        unsigned stopmask = 0;
#pragma unroll
        for (int j = 0; j < SEQSUM_PER_THREAD; j++) { bool st; c[j] = seqsum_contrib_fp(a[j], g, f.ok, st); if (st) stopmask |= 1u << j; }
        // thread total, capped at 2^53 (a thread that reaches it holds a carry, and nothing behind a carry is used): the block
        // total stays below 2^63
        unsigned long long loc = 0;
#pragma unroll
        for (int j = 0; j < SEQSUM_PER_THREAD; j++) { loc += c[j]; if (loc > SEQSUM_TOP) loc = SEQSUM_TOP; }
        unsigned long long incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) sh_wtot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const unsigned long long w = sh_wtot[lane];
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
            sh_wtot[lane] = wi - w;
        }
        __syncthreads();
        // this thread's first term that ends the pass
        unsigned long long P = sh_wtot[warp] + (incl - loc), Pstop = 0;
        long long kst = LLONG_MAX;
#pragma unroll
        for (int j = 0; j < SEQSUM_PER_THREAD; j++) {
            if (kst == LLONG_MAX && k0 + j < end) {
                const unsigned long long Pn = P + c[j];
                if ((stopmask >> j & 1u) || (f.ok && f.S + Pn >= SEQSUM_TOP)) { kst = k0 + j; Pstop = P; }
                else P = Pn;
            }
        }
        long long km = kst;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const long long v = __shfl_xor_sync(0xffffffffu, km, o); km = v < km ? v : km; }
        if (lane == 0) sh_kmin[warp] = km;
        __syncthreads();
        km = sh_kmin[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const long long v = __shfl_xor_sync(0xffffffffu, km, o); km = v < km ? v : km; }
        if (km == LLONG_MAX) {                     // the whole window folds in: the thread that holds its last term knows the total
            if (k0 < end && end <= k0 + SEQSUM_PER_THREAD) {
                sh_s = f.ok ? seqsum_value(f.ef, f.S + P) : s;
                sh_pos = end;
                sh_run = 1;
            }
        } else if (kst == km) {                    // the owner of the stopping term: everything before it at once, then real adds
            double t = f.ok ? seqsum_value(f.ef, f.S + Pstop) : s;
            long long k = km;
            const long long kend = km + run < n_per ? km + run : n_per;
            for (; k + 8 <= kend; k += 8) {
                const double a0 = src[k], a1 = src[k + 1], a2 = src[k + 2], a3 = src[k + 3], a4 = src[k + 4], a5 = src[k + 5], a6 = src[k + 6], a7 = src[k + 7];
                t += a0; t += a1; t += a2; t += a3; t += a4; t += a5; t += a6; t += a7;
            }
            for (; k < kend; k++) t += src[k];
            sh_s = t;
            sh_pos = kend;
            const long long consumed = km - pos;
            sh_run = consumed < 32 ? (run * 2 < SEQSUM_MAX_RUN ? run * 2 : SEQSUM_MAX_RUN) : 1;
        }
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = sh_s;
}

// which of the two: chr3d_set_option("seqsum", "serial") selects the serial chain (same bits, for cross-checks)
static inline bool seq_sum_serial_selected() { const char* v = dev_getenv("CHR3D_SEQSUM"); return v && v[0] == 's'; }
static inline void seq_sum_launch(const double* src, long long n_per, int nseg, double* out, cudaStream_t s) {
    if (seq_sum_serial_selected()) seq_sum_staged_kernel<<<nseg, 256, 0, s>>>(src, n_per, out);
    else seq_sum_exact_kernel<<<nseg, SEQSUM_THREADS, 0, s>>>(src, n_per, out);
}

}  // namespace chr3d
