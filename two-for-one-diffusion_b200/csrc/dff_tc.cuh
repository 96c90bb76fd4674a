// dff_tc.cuh -- 5th-generation tensor-core (tcgen05 / UMMA) building blocks for sm_100a:
// shared-memory operand descriptors, the TF32 instruction descriptor, TMEM allocation, MMA issue, commit and
// TMEM loads, plus the split-precision (3xTF32) block GEMM built from them.
//
// Operand layout (both A [M rows x K] and B [N rows x K], K-major, no swizzle): 16-byte chunks of 4 consecutive k,
// stored chunk-major:   byte_offset(r, k) = ((k / 4) * ROWS + r) * 16 + (k % 4) * 4
// i.e. UMMA "core matrices" of 8 rows x 16 B are contiguous 128-byte blocks, the next 8-row group follows at
// SBO = 128 B and the next k-chunk at LBO = ROWS * 16 B.  (Encodings follow cute/arch/mma_sm100_desc.hpp.)
#pragma once
#include <stdint.h>
#include "dff_common.cuh"

namespace dff {
namespace tc {

// ---- descriptors
__device__ __forceinline__ uint64_t smem_desc(const void* smem_ptr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) >> 4) & 0x3FFFu);          // start address, 16-byte units
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;             // leading-dimension byte offset
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;             // stride-dimension byte offset
    d |= (uint64_t)1 << 46;                                        // descriptor version (Blackwell)
    return d;                                                       // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
// kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM management (one warp allocates / frees; ncols power of two >= 32)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- MMA issue (one thread), D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
// accumulating form (enable-input-d predicate always true)
__device__ __forceinline__ void mma_tf32_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: each thread reads 8 consecutive fp32 columns of its own lane (32 lanes x 32 bit shape)
__device__ __forceinline__ void ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- operand staging: canonical chunk-major layout with a round-to-nearest TF32 hi/lo split
__device__ __forceinline__ uint32_t canon_off(int r, int k4, int rows) { return (uint32_t)(k4 * rows + r) * 4u; }   // in floats
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
    hi.x = __uint_as_float((__float_as_uint(x.x) + 0x1000u) & 0xffffe000u); lo.x = x.x - hi.x;
    hi.y = __uint_as_float((__float_as_uint(x.y) + 0x1000u) & 0xffffe000u); lo.y = x.y - hi.y;
    hi.z = __uint_as_float((__float_as_uint(x.z) + 0x1000u) & 0xffffe000u); lo.z = x.z - hi.z;
    hi.w = __uint_as_float((__float_as_uint(x.w) + 0x1000u) & 0xffffe000u); lo.w = x.w - hi.w;
}

// One thread: D[64 x N] (+)= (A_hi + A_lo)[64 x K] * (B_hi + B_lo)[N x K]^T as lo*hi + hi*lo + hi*hi TF32 MMAs.
// a_* : canonical [K/4][64][4] buffers, b_* : canonical [K/4][N][4] buffers.
__device__ __forceinline__ void issue_3xtf32(uint32_t d_tmem, const float* a_hi, const float* a_lo, const float* b_hi,
                                             const float* b_lo, int N, int K, bool accumulate_first) {
    const uint32_t idesc = idesc_tf32(64, N);
    uint32_t acc = accumulate_first ? 1u : 0u;
    for (int k = 0; k < K; k += 8) {            // one MMA consumes K = 8 (two 16-byte chunks)
        const int c = k >> 2;
        const uint64_t ah = smem_desc(a_hi + canon_off(0, c, 64), 64 * 16, 128), al = smem_desc(a_lo + canon_off(0, c, 64), 64 * 16, 128);
        const uint64_t bh = smem_desc(b_hi + canon_off(0, c, N), N * 16, 128), bl = smem_desc(b_lo + canon_off(0, c, N), N * 16, 128);
        mma_tf32_ss(d_tmem, al, bh, idesc, acc);
        mma_tf32_ss(d_tmem, ah, bl, idesc, 1u);
        mma_tf32_ss(d_tmem, ah, bh, idesc, 1u);
        acc = 1u;
    }
}

}  // namespace tc

#ifdef DFF_HOST_TU       // only the host translation unit (dff_b200.cu) carries the validation kernel
// ------------------------------------------------------------------ standalone validation / throughput kernel
// D[64 x N] = A[64 x K] * B[N x K]^T (fp32 in, fp32 out, 3xTF32 on tcgen05), repeated `reps` times for timing.
// One CTA of 128 threads; used by dff_debug_tc_gemm (tests/test_gpu_tc.py).
__global__ void __launch_bounds__(128, 1)
dff_tc_gemm_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K, int reps) {
    extern __shared__ __align__(128) float tsm[];
    float* a_hi = tsm;                       // [K/4][64][4]
    float* a_lo = a_hi + 64 * K;
    float* b_hi = a_lo + 64 * K;             // [K/4][N][4]
    float* b_lo = b_hi + N * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int idx = tid; idx < 64 * (K / 4); idx += blockDim.x) {
        const int r = idx % 64, c = idx / 64;
        const float4 x = *reinterpret_cast<const float4*>(A + (size_t)r * K + c * 4);
        float4 hi, lo;
        tc::split4(x, hi, lo);
        *reinterpret_cast<float4*>(a_hi + tc::canon_off(r, c, 64)) = hi;
        *reinterpret_cast<float4*>(a_lo + tc::canon_off(r, c, 64)) = lo;
    }
    for (int idx = tid; idx < N * (K / 4); idx += blockDim.x) {
        const int r = idx % N, c = idx / N;
        const float4 x = *reinterpret_cast<const float4*>(B + (size_t)r * K + c * 4);
        float4 hi, lo;
        tc::split4(x, hi, lo);
        *reinterpret_cast<float4*>(b_hi + tc::canon_off(r, c, N)) = hi;
        *reinterpret_cast<float4*>(b_lo + tc::canon_off(r, c, N)) = lo;
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base, 256);
    fence_proxy_async();                     // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t d_tmem = tmem_base;

    uint32_t phase = 0;
    for (int it = 0; it < reps; ++it) {
        if (tid == 0) {
            tc::issue_3xtf32(d_tmem, a_hi, a_lo, b_hi, b_lo, N, K, false);
            tc::commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        tc::fence_after_sync();
    }
    // epilogue: M = 64 accumulators live in lanes (m % 16) + 32 * (m / 16): warp w, lanes 0..15 -> rows 16 w + lane
    const int row = warp * 16 + lane;
    for (int n0 = 0; n0 < N; n0 += 8) {
        float v[8];
        tc::ld8(d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0, v);
        if (lane < 16) {
#pragma unroll
            for (int i = 0; i < 8; ++i) D[(size_t)row * N + n0 + i] = v[i];
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free(d_tmem, 256);
}

#endif  // DFF_HOST_TU

}  // namespace dff
