// Host check of chromosome3d_b200/csrc/seqsum_core.h: the windowed integer-scan evaluation of a left-to-right fp64 sum must
// return the bits of the scalar loop `for (k) s += a[k]` on every input.  Built and run by tests/test_seqsum_cpu.py (g++ -O2,
// no -ffast-math).  The driver below follows seq_sum_exact_kernel (seqsum.cuh) step by step, with the block-wide scan and the
// block-wide minimum written as plain loops.
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../chromosome3d_b200/csrc/seqsum_core.h"

using namespace chr3d;

static double scalar_sum(const std::vector<double>& a) {
    volatile double s = 0.0;
    for (size_t k = 0; k < a.size(); k++) s = s + a[k];
    return s;
}

// one pass of the kernel's outer loop per iteration; window = threads * per_thread terms
static double windowed_sum(const std::vector<double>& a, long long window, long long* iterations) {
    const long long n = (long long)a.size();
    double s = 0.0;
    long long pos = 0, it = 0, run = 1;       // run: terms added one by one at a stop (grows while passes get nowhere)
    while (pos < n) {
        it++;
        const SeqSumFrame f = seqsum_frame(s);
        const SeqSumFpFrame g = seqsum_fp_frame(f);
        const long long end = pos + window < n ? pos + window : n;
        uint64_t P = 0;                 // inclusive prefix of the contributions (saturating at 2^53, as the kernel's thread-local prefix)
        long long kstop = end;
        uint64_t Pexcl = 0;
        for (long long k = pos; k < end; k++) {
            bool stop;
            const uint64_t c = seqsum_contrib_fp(a[k], g, f.ok, stop);
            {   // the integer formulation agrees wherever both fold the term in
                bool stop_i; const uint64_t ci = seqsum_contrib(a[k], f.ef, f.ok, stop_i);
                if (!stop && !stop_i && c != ci) { printf("FAIL contribution: a=%a frame ef=%d fp=%llu int=%llu\n", a[k], f.ef, (unsigned long long)c, (unsigned long long)ci); exit(1); }
                if (!stop && stop_i && !(seqsum_bits(a[k]) >> 52 == 0)) { printf("FAIL: the integer form stops at a=%a (ef=%d), the fp form does not\n", a[k], f.ef); exit(1); }
            }
            const uint64_t Pn = P + c;
            if (stop || (f.ok && f.S + Pn >= SEQSUM_TOP)) { kstop = k; Pexcl = P; break; }
            P = Pn; Pexcl = P;
        }
        if (f.ok) s = seqsum_value(f.ef, f.S + Pexcl);       // everything before kstop, at once
        if (kstop < end) {
            const long long kend = kstop + run < n ? kstop + run : n, consumed = kstop - pos;
            for (long long k = kstop; k < kend; k++) { volatile double t = s + a[k]; s = t; }
            pos = kend;
            run = consumed < 32 ? (run * 2 < 4096 ? run * 2 : 4096) : 1;
        } else { pos = end; run = 1; }
    }
    if (iterations) *iterations = it;
    return s;
}

// The kernel thread by thread (1024 threads x 8 terms, warps of 32): the same per-thread code as seq_sum_exact_kernel, with the
// warp shuffles and the block barriers replaced by loops over the threads.  Checks the scan / minimum / ownership logic that
// windowed_sum() above folds into one loop.
static double emulated_block_sum(const std::vector<double>& a, long long* passes) {
    const int T = 1024, PER = 8; const long long WINDOW = (long long)T * PER;
    const long long n_per = (long long)a.size();
    double sh_s = 0.0; long long sh_pos = 0; int sh_run = 1;
    std::vector<uint64_t> c((size_t)T * PER), loc(T), incl(T), P(T), Pstop(T);
    std::vector<unsigned> stopmask(T);
    std::vector<long long> kst(T);
    uint64_t wtot[32]; long long it = 0;
    while (true) {
        const long long pos = sh_pos; const double s = sh_s; const int run = sh_run;
        if (pos >= n_per) break;
        it++;
        const SeqSumFrame f = seqsum_frame(s);
        const SeqSumFpFrame g = seqsum_fp_frame(f);
        const long long end = pos + WINDOW < n_per ? pos + WINDOW : n_per;
        for (int tid = 0; tid < T; tid++) {
            const long long k0 = pos + (long long)tid * PER;
            stopmask[tid] = 0;
            for (int j = 0; j < PER; j++) {
                c[(size_t)tid * PER + j] = 0;
                if (k0 + j < end) { bool st; c[(size_t)tid * PER + j] = seqsum_contrib_fp(a[k0 + j], g, f.ok, st); if (st) stopmask[tid] |= 1u << j; }
            }
            uint64_t l = 0;
            for (int j = 0; j < PER; j++) { l += c[(size_t)tid * PER + j]; if (l > SEQSUM_TOP) l = SEQSUM_TOP; }
            loc[tid] = l;
        }
        for (int w = 0; w < 32; w++) {                     // warp inclusive scans
            uint64_t run_ = 0;
            for (int l = 0; l < 32; l++) { run_ += loc[w * 32 + l]; incl[w * 32 + l] = run_; }
            wtot[w] = incl[w * 32 + 31];
        }
        { uint64_t run_ = 0; for (int w = 0; w < 32; w++) { const uint64_t t = wtot[w]; wtot[w] = run_; run_ += t; } }     // exclusive over warps
        long long km = LLONG_MAX;
        for (int tid = 0; tid < T; tid++) {
            const long long k0 = pos + (long long)tid * PER;
            uint64_t p = wtot[tid >> 5] + (incl[tid] - loc[tid]); Pstop[tid] = 0; kst[tid] = LLONG_MAX;
            for (int j = 0; j < PER; j++) {
                if (kst[tid] == LLONG_MAX && k0 + j < end) {
                    const uint64_t Pn = p + c[(size_t)tid * PER + j];
                    if ((stopmask[tid] >> j & 1u) || (f.ok && f.S + Pn >= SEQSUM_TOP)) { kst[tid] = k0 + j; Pstop[tid] = p; }
                    else p = Pn;
                }
            }
            P[tid] = p;
            if (kst[tid] < km) km = kst[tid];
        }
        int writers = 0;
        for (int tid = 0; tid < T; tid++) {
            const long long k0 = pos + (long long)tid * PER;
            if (km == LLONG_MAX) {
                if (k0 < end && end <= k0 + PER) { sh_s = f.ok ? seqsum_value(f.ef, f.S + P[tid]) : s; sh_pos = end; sh_run = 1; writers++; }
            } else if (kst[tid] == km) {
                volatile double t = f.ok ? seqsum_value(f.ef, f.S + Pstop[tid]) : s;
                const long long kend = km + run < n_per ? km + run : n_per;
                for (long long k = km; k < kend; k++) t = t + a[k];
                sh_s = t; sh_pos = kend;
                const long long consumed = km - pos;
                sh_run = consumed < 32 ? (run * 2 < 4096 ? run * 2 : 4096) : 1;
                writers++;
            }
        }
        if (writers != 1) { printf("FAIL emulation: %d writers in one pass\n", writers); exit(1); }
    }
    if (passes) *passes = it;
    return sh_s;
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double urand() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }

static int check(const char* name, const std::vector<double>& a, long long window) {
    long long it = 0;
    const double ref = scalar_sum(a), got = windowed_sum(a, window, &it);
    const bool same = seqsum_bits(ref) == seqsum_bits(got) || (ref != ref && got != got);
    if (!same) { printf("FAIL %s n=%zu window=%lld ref=%a got=%a\n", name, a.size(), window, ref, got); return 1; }
    if (window == 8192) {          // the kernel's shape: also through the thread-by-thread emulation
        long long passes = 0;
        const double emu = emulated_block_sum(a, &passes);
        if (!(seqsum_bits(ref) == seqsum_bits(emu) || (ref != ref && emu != emu))) { printf("FAIL %s (emulated block) n=%zu ref=%a got=%a\n", name, a.size(), ref, emu); return 1; }
        if (passes != it) { printf("FAIL %s: emulated block took %lld passes, windowed loop %lld\n", name, passes, it); return 1; }
    }
    printf("ok   %-28s n=%-8zu window=%-6lld iterations=%lld\n", name, a.size(), window, it);
    return 0;
}

int main() {
    int bad = 0;
    const long long windows[] = {1, 7, 256, 8192};
    for (long long w : windows) {
        for (int rep = 0; rep < 6; rep++) {
            const size_t n = rep == 0 ? 1 : (rep == 1 ? 2 : (size_t)(1000 + rnd() % 300000));
            std::vector<double> a(n);
            // 1. uniform positive
            for (auto& v : a) v = urand();
            bad += check("uniform", a, w);
            // 2. Hi-C like: integer counts to a power, decaying with the distance from the diagonal, many zeros
            { const int L = (int)sqrt((double)n) + 1; a.assign((size_t)L * L, 0.0);
              const double alpha = 0.1 + 0.1 * (double)(rnd() % 20);
              for (int i = 0; i < L; i++) for (int j = 0; j < L; j++) {
                  const double cnt = floor(5000.0 * urand() / (1.0 + abs(i - j)));
                  a[(size_t)i * L + j] = pow(cnt, alpha); } }
            bad += check("hic_pow", a, w);
            a.resize(n);
            // 3. wide dynamic range
            for (auto& v : a) v = ldexp(urand() + 0.5, (int)(rnd() % 120) - 60);
            bad += check("wide_range", a, w);
            // 4. ties: few significant bits, so that half-ulp remainders are common
            for (auto& v : a) v = ldexp((double)(rnd() % 16 + 1), (int)(rnd() % 60) - 50);
            bad += check("dyadic_ties", a, w);
            // 5. deviations as assess.cu forms them: mostly zeros, differences of 3-decimal numbers
            for (auto& v : a) { if (rnd() % 4) v = 0.0; else { const double d = (double)(rnd() % 100000) / 1000.0, t = (double)(rnd() % 1000) / 10.0; v = d > t ? d - t : t - d; } }
            bad += check("deviations", a, w);
            // 6. leading zeros, then integers (exact sums), then halves
            for (size_t k = 0; k < n; k++) a[k] = k < n / 3 ? 0.0 : (k < 2 * n / 3 ? (double)(rnd() % 1000) : (double)(rnd() % 1000) + 0.5);
            bad += check("zeros_ints_halves", a, w);
            // 7. negatives, denormals, -0, spikes
            for (auto& v : a) { const unsigned r = (unsigned)(rnd() % 1000);
                v = r < 5 ? -urand() : (r < 8 ? 4.9e-324 * (double)(rnd() % 100) : (r < 10 ? -0.0 : (r < 12 ? ldexp(urand(), 80) : urand()))); }
            bad += check("mixed_signs_denormals", a, w);
            // 8. sums that run into powers of two exactly
            for (auto& v : a) v = ldexp(1.0, (int)(rnd() % 8) - 4);
            bad += check("powers_of_two", a, w);
            // 9. tiny frames: the sum stays below the range where the unit in the last place is normal
            for (auto& v : a) v = ldexp(urand() + 0.5, -1000 - (int)(rnd() % 20));
            bad += check("tiny", a, w);
            if (w == 1 && rep >= 2) break;         // window 1 is the scalar loop with extra steps: a few sizes are enough
        }
    }
    {   // non-finite terms
        std::vector<double> a(5000);
        for (auto& v : a) v = urand();
        a[1234] = INFINITY; bad += check("inf", a, 256);
        a[1234] = NAN; bad += check("nan", a, 256);
        a[1234] = 1e308; a[2000] = 1e308; bad += check("overflow", a, 256);
    }
    {   // empty
        std::vector<double> a; bad += check("empty", a, 256);
    }
    printf(bad ? "FAILED %d\n" : "all ok\n", bad);
    return bad ? 1 : 0;
}
