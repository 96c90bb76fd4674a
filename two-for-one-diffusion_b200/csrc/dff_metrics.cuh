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

// ------------------------------------------------------------------ contact maps (evaluate/evaluators.py:735-858)
// counts[i][j] += [|x_i - x_j| < cutoff] over the samples (ContactEvaluator._get_samp_contacts + the sum of _plot_contact_normcount);
// mismatch[s] = number of pairs (i, j), j - i >= offset, whose contact state differs from the folded structure's
// (_eval_bce_dynamics: binary_cross_entropy of {0,1} inputs is 100 per mismatching pair, torch clamps log at -100).
__global__ void __launch_bounds__(256)
dff_contacts_kernel(const float* __restrict__ x, int n, int N, float cutoff, const unsigned char* __restrict__ folded, int offset,
                    unsigned int* __restrict__ counts, unsigned int* __restrict__ mismatch) {
    extern __shared__ unsigned int scnt[];
    const int NN = N * N;
    for (int p = threadIdx.x; p < NN; p += blockDim.x) scnt[p] = 0u;
    __syncthreads();
    const long long total = (long long)n * NN;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(w / NN), p = (int)(w - (long long)s * NN);
        const int i = p / N, j = p - i * N;
        const float* a = x + ((size_t)s * N + i) * 3;
        const float* b = x + ((size_t)s * N + j) * 3;
        const float d = pwd_norm(a[0] - b[0], a[1] - b[1], a[2] - b[2]);
        const bool c = d < cutoff;
        if (c) atomicAdd(scnt + p, 1u);
        if (mismatch != nullptr && folded != nullptr && j - i >= offset && c != (folded[p] != 0)) atomicAdd(mismatch + s, 1u);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < NN; p += blockDim.x) if (scnt[p]) atomicAdd(counts + p, scnt[p]);
}

// ------------------------------------------------------------------ backbone torsions (evaluate/evaluators_CGflowmatching.py:32-51)
// mdtraj.compute_dihedrals of two atom quadruples per structure (phi = beads 0-1-2-3, psi = 1-2-3-4 for alanine dipeptide):
//   b1 = x1 - x0, b2 = x2 - x1, b3 = x3 - x2, c1 = b2 x b3, c2 = b1 x b2, angle = atan2((b1 . c1) |b2|, c1 . c2)   (fp32)
// and the 2-D histogram of get_prob: np.histogram2d over edges = np.linspace(-pi, pi, n_bins) (n_bins - 1 bins per axis).
__device__ __forceinline__ float torsion_f32(const float* p0, const float* p1, const float* p2, const float* p3) {
    const float b1x = p1[0] - p0[0], b1y = p1[1] - p0[1], b1z = p1[2] - p0[2];
    const float b2x = p2[0] - p1[0], b2y = p2[1] - p1[1], b2z = p2[2] - p1[2];
    const float b3x = p3[0] - p2[0], b3y = p3[1] - p2[1], b3z = p3[2] - p2[2];
    const float c1x = __fsub_rn(__fmul_rn(b2y, b3z), __fmul_rn(b2z, b3y)), c1y = __fsub_rn(__fmul_rn(b2z, b3x), __fmul_rn(b2x, b3z)),
                c1z = __fsub_rn(__fmul_rn(b2x, b3y), __fmul_rn(b2y, b3x));
    const float c2x = __fsub_rn(__fmul_rn(b1y, b2z), __fmul_rn(b1z, b2y)), c2y = __fsub_rn(__fmul_rn(b1z, b2x), __fmul_rn(b1x, b2z)),
                c2z = __fsub_rn(__fmul_rn(b1x, b2y), __fmul_rn(b1y, b2x));
    const float n2 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(b2x, b2x), __fmul_rn(b2y, b2y)), __fmul_rn(b2z, b2z)));
    const float pp1 = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(b1x, c1x), __fmul_rn(b1y, c1y)), __fmul_rn(b1z, c1z)), n2);
    const float pp2 = __fadd_rn(__fadd_rn(__fmul_rn(c1x, c2x), __fmul_rn(c1y, c2y)), __fmul_rn(c1z, c2z));
    return atan2f(pp1, pp2);
}
// np.histogram bin of v over np.linspace(lo, hi, nb + 1): searchsorted(edges, v, 'right') - 1, the last edge inclusive; -1 = outside
__device__ __forceinline__ int linspace_bin(double v, double lo, double hi, int nb) {
    if (!(v >= lo) || !(v <= hi)) return -1;
    const double step = (hi - lo) / (double)nb;
    auto edge = [&](int i) { return i >= nb ? hi : (double)i * step + lo; };
    int b = (int)((v - lo) / step);
    b = max(0, min(b, nb - 1));
    while (b > 0 && v < edge(b)) --b;
    while (b < nb - 1 && v >= edge(b + 1)) ++b;
    return b;
}
__global__ void __launch_bounds__(256)
dff_dihedral_kernel(const float* __restrict__ x, int n, int N, const int* __restrict__ quads, int nb, float* __restrict__ tors_out,
                    unsigned int* __restrict__ hist) {
    const double kPi = 3.141592653589793;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const float* xs = x + (size_t)s * N * 3;
        const float a = torsion_f32(xs + quads[0] * 3, xs + quads[1] * 3, xs + quads[2] * 3, xs + quads[3] * 3);
        const float b = torsion_f32(xs + quads[4] * 3, xs + quads[5] * 3, xs + quads[6] * 3, xs + quads[7] * 3);
        if (tors_out != nullptr) { tors_out[(size_t)s * 2] = a; tors_out[(size_t)s * 2 + 1] = b; }
        if (hist != nullptr) {
            const int ia = linspace_bin((double)a, -kPi, kPi, nb), ib = linspace_bin((double)b, -kPi, kPi, nb);
            if (ia >= 0 && ib >= 0) atomicAdd(hist + (size_t)ia * nb + ib, 1u);
        }
    }
}

// ------------------------------------------------------------------ RMSD to a reference structure after optimal superposition
// (evaluate/evaluators.py:636-665: mdtraj.rmsd(traj, folded)).  One warp per structure: centre both, G = sum |a|^2 + sum |b|^2,
// M = sum a_i b_i^T; the optimal rotation gives  rmsd^2 = (G - 2 (s1 + s2 + sign(det M) s3)) / N  with s_k the singular values
// of M (Kabsch), obtained in fp64 from the closed-form eigenvalues of M^T M.
__global__ void __launch_bounds__(256)
dff_rmsd_kernel(const float* __restrict__ x, int n, int N, const float* __restrict__ ref, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    auto wsum = [](double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    for (int s = warp_global; s < n; s += nwarps) {
        const float* xs = x + (size_t)s * N * 3;
        double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
        for (int i = lane; i < N; i += 32)
            for (int c = 0; c < 3; ++c) { ca[c] += (double)xs[i * 3 + c]; cb[c] += (double)ref[i * 3 + c]; }
        for (int c = 0; c < 3; ++c) { ca[c] = wsum(ca[c]) / N; cb[c] = wsum(cb[c]) / N; }
        double G = 0, M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = lane; i < N; i += 32) {
            double a[3], b[3];
            for (int c = 0; c < 3; ++c) { a[c] = (double)xs[i * 3 + c] - ca[c]; b[c] = (double)ref[i * 3 + c] - cb[c]; }
            G += a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r * 3 + c] += a[r] * b[c];
        }
        G = wsum(G);
        for (int k = 0; k < 9; ++k) M[k] = wsum(M[k]);
        if (lane == 0) {
            double A[6];      // M^T M (symmetric): 00 01 02 11 12 22
            A[0] = M[0] * M[0] + M[3] * M[3] + M[6] * M[6]; A[1] = M[0] * M[1] + M[3] * M[4] + M[6] * M[7]; A[2] = M[0] * M[2] + M[3] * M[5] + M[6] * M[8];
            A[3] = M[1] * M[1] + M[4] * M[4] + M[7] * M[7]; A[4] = M[1] * M[2] + M[4] * M[5] + M[7] * M[8]; A[5] = M[2] * M[2] + M[5] * M[5] + M[8] * M[8];
            const double p1 = A[1] * A[1] + A[2] * A[2] + A[4] * A[4];
            const double q = (A[0] + A[3] + A[5]) / 3.0;
            double e1, e2, e3;
            const double p2 = (A[0] - q) * (A[0] - q) + (A[3] - q) * (A[3] - q) + (A[5] - q) * (A[5] - q) + 2.0 * p1;
            if (p2 <= 1e-300) { e1 = e2 = e3 = q; }
            else {
                const double p = sqrt(p2 / 6.0);
                const double b00 = (A[0] - q) / p, b11 = (A[3] - q) / p, b22 = (A[5] - q) / p, b01 = A[1] / p, b02 = A[2] / p, b12 = A[4] / p;
                double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
                r = fmin(1.0, fmax(-1.0, r));
                const double phi = acos(r) / 3.0;
                e1 = q + 2.0 * p * cos(phi);
                e3 = q + 2.0 * p * cos(phi + 2.0943951023931953);
                e2 = 3.0 * q - e1 - e3;
            }
            const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
            const double s1 = sqrt(fmax(e1, 0.0)), s2 = sqrt(fmax(e2, 0.0)), s3 = sqrt(fmax(e3, 0.0));
            const double msd = (G - 2.0 * (s1 + s2 + (det < 0 ? -s3 : s3))) / (double)N;
            out[s] = (float)sqrt(fmax(msd, 0.0));
        }
    }
}

}  // namespace dff
