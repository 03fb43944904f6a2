// seqsum.cuh — exact left-to-right fp64 summation (reproduces the rounding sequence of a scalar
// CPU loop, which the Perl's `$sum += ...` loops are): the block stages chunks through shared
// memory with coalesced loads while thread 0 adds them in order.  One block per segment.
#pragma once

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

}  // namespace chr3d
