// dff_metrics.cuh -- pairwise-distance statistics of sampled structures on the GPU.
// Replaces evaluate/evaluators.py:934-948 (get_pwd_triu_batch) + the torch.histc loops of PwdEvaluator (:241-263):
// for every bead pair (i, j) with j - i >= offset, the maximum distance over the samples and a fixed-resolution histogram.
// HBM-bound byte work: one pass over x [n, N, 3] per kernel (12 N bytes per sample), no GEMM shape anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dff {

// pair index p <-> (i, j) in torch.triu_indices(N, N, offset) order (row-major over i, then j)
__device__ __forceinline__ void pair_of(int p, int N, int offset, int& i, int& j) {
    int row = 0, left = p;
    while (true) {
        const int cnt = N - offset - row;         // pairs (row, row + offset .. N - 1)
        if (left < cnt) break;
        left -= cnt; ++row;
    }
    i = row; j = row + offset + left;
}

// |d| with the rounding of torch.norm's CPU kernel (sum of squares as a chain of fused multiply-adds, then sqrt): verified
// bit-identical on 100 000 distances of the golden set (oracle/make_golden_metrics.py)
__device__ __forceinline__ float pwd_norm(float dx, float dy, float dz) {
    return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
}

// max_out[p] = max over samples of |x_i - x_j|  (float bits compared as unsigned: distances are >= 0)
__global__ void __launch_bounds__(256)
dff_pwd_max_kernel(const float* __restrict__ x, int n, int N, int offset, int P, unsigned int* __restrict__ max_bits) {
    extern __shared__ unsigned int smax[];
    for (int p = threadIdx.x; p < P; p += blockDim.x) smax[p] = 0u;
    __syncthreads();
    const long long total = (long long)n * P;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(w / P), p = (int)(w - (long long)s * P);
        int i, j;
        pair_of(p, N, offset, i, j);
        const float* a = x + ((size_t)s * N + i) * 3;
        const float* b = x + ((size_t)s * N + j) * 3;
        const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const float d = pwd_norm(dx, dy, dz);
        atomicMax(smax + p, __float_as_uint(d));
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) if (smax[p]) atomicMax(max_bits + p, smax[p]);
}

// hist[p][b] += 1: the equal-width binning of torch.histc(pwd, bins = nbins, min = 0, max = resolution * nbins)
// (evaluators.py:244-246, 261-263).  Identical to torch's CPU result except, possibly, for a value within one float ulp
// of a bin edge, where torch's own answer depends on how its vectorised linspace rounded that edge.
__global__ void __launch_bounds__(256)
dff_pwd_hist_kernel(const float* __restrict__ x, int n, int N, int offset, int P, float resolution,
                    const int* __restrict__ nbins, int ld_hist, unsigned int* __restrict__ hist) {
    const long long total = (long long)n * P;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(w / P), p = (int)(w - (long long)s * P);
        int i, j;
        pair_of(p, N, offset, i, j);
        const float* a = x + ((size_t)s * N + i) * 3;
        const float* b = x + ((size_t)s * N + j) * 3;
        const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
        const float d = pwd_norm(dx, dy, dz);
        const int nb = nbins[p];
        const float hi = resolution * (float)nb;
        if (d < 0.f || d > hi) continue;          // histc ignores out-of-range elements
        // ATen's histc on the CPU (HistogramKernel.cpp, linear interpolation + local search): a first guess
        // (d * nbins) / (max - min), corrected against the neighbouring bin edges of linspace(min, max, nbins + 1)
        const int steps = nb + 1, half = steps / 2;
        const float step = hi / (float)nb;
        auto edge = [&](int i) { return i < half ? step * (float)i : hi - step * (float)(steps - i - 1); };
        const int guess = (int)(d * (float)nb / hi);
        const int lo_i = max(0, guess - 1), hi_i = min(guess + 2, steps);
        int bin = lo_i - 1;
        for (int i = lo_i; i < hi_i; ++i)
            if (edge(i) <= d) bin = i;
        if (bin < 0) bin = 0;
        if (bin >= nb) bin = nb - 1;               // the rightmost bin includes the right boundary
        atomicAdd(hist + (size_t)p * ld_hist + bin, 1u);
    }
}

}  // namespace dff
