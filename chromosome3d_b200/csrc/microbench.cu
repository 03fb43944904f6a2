// microbench.cu — live FP32-pipe peak measurement used as the roofline denominator of the
// FP32-bound small-N kernel (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
#include <algorithm>
#include "common.h"

namespace chr3d {

__global__ void __launch_bounds__(1024, 1) ffma_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678f) out[0] = s;
}

// the same chain on the packed pipe (FFMA2, fma.rn.f32x2): two FMAs per lane per instruction
__global__ void __launch_bounds__(1024, 1) ffma2_kernel(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = make_float2(threadIdx.x * 1e-3f + q, threadIdx.x * 2e-3f + q);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
#pragma unroll
            for (int q = 0; q < 8; q++) x[q] = __ffma2_rn(x[q], a2, b2);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; q++) s += x[q].x + x[q].y;
    if (s == 12345.678f) out[0] = s;
}

}  // namespace chr3d

using namespace chr3d;

// Measured FP32 FMA throughput (TFLOP/s, FMA = 2 flop) of the whole chip; best of `reps` launches.
static int fp32_peak_impl(int reps, double* tflops, bool packed);
CHR3D_API int chr3d_fp32_peak(int reps, double* tflops) { return fp32_peak_impl(reps, tflops, false); }
// same measurement with packed FFMA2 instructions (the form ln_tile_kernel's pair term uses)
CHR3D_API int chr3d_fp32_peak_packed(int reps, double* tflops) { return fp32_peak_impl(reps, tflops, true); }
static int fp32_peak_impl(int reps, double* tflops, bool packed) {
    if (!tflops) { set_error("chr3d_fp32_peak: null output"); return CHR3D_ERR_ARG; }
    if (int rc = require_device()) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    DevBuf<float> out; CHR3D_CUDA(out.alloc(1));
    cudaEvent_t e0, e1; CHR3D_CUDA(cudaEventCreate(&e0)); CHR3D_CUDA(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 2, threads = 1024;
    double best = 0.0;
    for (int r = 0; r < reps + 1; r++) {
        CHR3D_CUDA(cudaEventRecord(e0));
        if (packed) ffma2_kernel<<<blocks, threads>>>(out.p, iters, 0.999f, 0.001f);
        else ffma_kernel<<<blocks, threads>>>(out.p, iters, 0.999f, 0.001f);
        CHR3D_CUDA(cudaEventRecord(e1));
        CHR3D_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f; CHR3D_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads * (packed ? 2.0 : 1.0);
        if (r > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return 0;
}
