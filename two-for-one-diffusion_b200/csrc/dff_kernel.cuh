// dff_kernel.cuh -- the fused score-network + integrator kernel (sm_100a, 3xTF32 tensor-core GEMMs, TMA bulk
// weight streaming).  One CTA (256 threads) owns a group of S whole samples (<= R node rows) and runs, for
// each of n_steps diffusion / MD steps, the complete forward pass of the collapsed graph transformer, its
// hand-written reverse pass w.r.t. x, and the DDPM-posterior / BAOAB / Brownian update -- coordinates
// never leave shared memory between steps.  Samples never interact, so there is no grid-wide sync.
//
// Math: SURVEY.md Appendix A restated with k'_j = k_j + A x_j, v'_j = v_j + A x_j so that every head is
// plain attention over (q, k', v'):   o_i = sum_j p_ij v'_j - A x_i + c,
//   dx_r += A_h^T (dk'_r + dv'_r - do_r)   (oracle/collapsed_ref.py is the CPU statement of the same).
// Reference lines replaced: models/graph_transformer.py:87-111,143-159,178-329; models/ddpm.py:195-251;
// dynamics/langevin.py:75-92; dynamics/langevin_cgnet.py:447-542,737-771; utils.py:65-86.
//
// Template parameters: HP = hidden size padded to 64/128 output columns, R = node rows per pass (32/64),
// HC = attention heads processed per chunk (1 for R=64, 2 for R=32).
#pragma once
#include <math.h>
#include "dff_common.cuh"

#ifndef DFF_SPLIT_RN
#define DFF_SPLIT_RN 2      // 0: hi = truncation (mask); 1: cvt.rna.tf32; 2: integer add + mask (round half away)
#endif
#ifndef DFF_FLUSH
#define DFF_FLUSH 1         // 1: each k8 step accumulates into a zeroed temporary that is added to the fp32 accumulator with RN
#endif

namespace dff {

// ------------------------------------------------------------------ weight stream (TMA bulk + mbarrier ring)
struct WStream {
    float* stage_base;
    uint64_t* full;       // [kStages] TMA completion (expect_tx) barriers
    uint32_t n;           // slices consumed so far (identical in every thread)
    uint32_t nst, sh;     // ring depth (power of two) and its log2
    uint32_t stage_floats;
    // producer cursor (meaningful in thread 0 only)
    const Seg* segs;      // segment table (a shared-memory copy when it fits)
    int nseg, seg_i;
    uint32_t slice_i, issued, total;

    __device__ __forceinline__ void issue_one() {
        const uint32_t st = issued & (nst - 1);
        const Seg sg = segs[seg_i];
        mbar_expect_tx(full + st, sg.slice_bytes);
        bulk_g2s(stage_base + st * stage_floats,
                 reinterpret_cast<const char*>(sg.base) + (size_t)slice_i * sg.slice_bytes, sg.slice_bytes, full + st);
        ++issued;
        if (++slice_i == sg.n_slices) { slice_i = 0; if (++seg_i == nseg) seg_i = 0; }
    }
    __device__ __forceinline__ void prefill() {      // thread 0, once: nothing to wait for
        for (uint32_t i = 0; i < nst && issued < total; ++i) issue_one();
    }
    __device__ __forceinline__ const float* wait_slice() const {
        mbar_wait(full + (n & (nst - 1)), (n >> sh) & 1u);
        return stage_base + (n & (nst - 1)) * stage_floats;
    }
    // All warps are done with the stage after the CTA barrier; thread 0 then refills it with slice n + nst.
    // (A warp-granular full/empty mbarrier hand-off was measured and was not faster: profiles/r01/ablation.md.)
    __device__ __forceinline__ void release_slice() {
        __syncthreads();
        if (threadIdx.x == 0 && issued < total) issue_one();
        ++n;
    }
};

// ------------------------------------------------------------------ streamed-weight GEMM on the tensor cores
// C[R][NT*CW] += sA[R][K] * W[K][NT*CW] with fp32-grade accuracy: every operand is split in registers into a
// TF32-exact high part and a low part (hi = rn_tf32(x), lo = x - hi) and three m16n8k8 TF32 MMAs
// (lo*hi + hi*lo + hi*hi, fp32 accumulate) replace one fp32 product -- "3xTF32", relative error ~2^-21.
// W arrives in [KS][NT*CW + 8] slices (row pad 8 floats => conflict-free B-fragment loads); sA row strides are
// 4 mod 32 floats => conflict-free A-fragment loads.  The warps form a WRn x WCn grid (2 x 4 for 8 warps); a warp owns
// MI m16 tiles x NI n8 tiles of each of the NT column tiles.
constexpr int kWPad = 8;

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// hi = x rounded to nearest TF32 (exactly representable, so the tensor core reads it unchanged); lo = x - hi exactly.
// Round-to-nearest keeps |lo| <= 2^-12 |x| with no sign bias (a truncating split biases every product toward zero).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
#if DFF_SPLIT_RN == 1
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
#elif DFF_SPLIT_RN == 2
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;       // round half away from zero on the integer pipe
#else
    hi = __float_as_uint(x) & 0xffffe000u;
#endif
    lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int R, int CW, int NT>
struct Acc {
    static constexpr int WRn = (kWarps == 16 && R >= 64) ? 4 : 2;   // warp rows
    static constexpr int WCn = kWarps / WRn;                         // warp columns
    static constexpr int RW = R / WRn;       // rows per warp row
    static constexpr int CWW = CW / WCn;     // columns per warp column
    static constexpr int MI = RW / 16;       // m16 tiles per warp
    static constexpr int NI = CWW / 8;       // n8 tiles per warp and column tile
    static_assert(MI >= 1 && NI >= 1 && MI * 16 * WRn == R && NI * 8 * WCn == CW, "warp grid does not tile the block");
    float v[NT][MI][NI][4];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int m = 0; m < MI; ++m)
#pragma unroll
                for (int n = 0; n < NI; ++n)
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[t][m][n][e] = 0.f;
    }
};

template <int R, int CW, int NT, int KM = 1>
__device__ __forceinline__ void gemm_acc(WStream& ws, const float* __restrict__ sA, int lda, int K, Acc<R, CW, NT>& acc) {
    using A_ = Acc<R, CW, NT>;
    constexpr int NC = CW * NT;
    constexpr int NCP = NC + kWPad;
    constexpr int KS = ((NC == 384) ? 8 : (NC == 192 ? 16 : (NC == 128 ? 16 : 32))) * (NC == 64 ? 1 : KM);   // rows per streamed slice
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp / A_::WCn, wc = warp % A_::WCn;
    const int g = lane >> 2, tig = lane & 3;
    const float* a_base = sA + (wr * A_::RW + g) * lda + tig;
    const int bcol = wc * A_::CWW + g;
    for (int k0 = 0; k0 < K; k0 += KS) {
        const float* __restrict__ w = ws.wait_slice() + tig * NCP + bcol;
#pragma unroll
        for (int kk = 0; kk < KS; kk += 8) {
            uint32_t ahi[A_::MI][4], alo[A_::MI][4];
#pragma unroll
            for (int m = 0; m < A_::MI; ++m) {
                const float* p = a_base + (m * 16) * lda + k0 + kk;
                split_tf32(p[0], ahi[m][0], alo[m][0]);
                split_tf32(p[8 * lda], ahi[m][1], alo[m][1]);
                split_tf32(p[4], ahi[m][2], alo[m][2]);
                split_tf32(p[8 * lda + 4], ahi[m][3], alo[m][3]);
            }
            uint32_t bhi[NT][A_::NI][2], blo[NT][A_::NI][2];
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int n = 0; n < A_::NI; ++n) {
                    const float* q = w + kk * NCP + t * CW + n * 8;
                    split_tf32(q[0], bhi[t][n][0], blo[t][n][0]);
                    split_tf32(q[4 * NCP], bhi[t][n][1], blo[t][n][1]);
                }
#if DFF_FLUSH
            // accuracy variant: the k8 partial product is formed in a zeroed temporary and added with fp32 RN
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int n = 0; n < A_::NI; ++n)
#pragma unroll
                    for (int m = 0; m < A_::MI; ++m) {
                        float tmp[4] = {0.f, 0.f, 0.f, 0.f};
                        mma_tf32(tmp, alo[m], bhi[t][n]);
                        mma_tf32(tmp, ahi[m], blo[t][n]);
                        mma_tf32(tmp, ahi[m], bhi[t][n]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc.v[t][m][n][e] += tmp[e];
                    }
#else
#pragma unroll
            for (int t = 0; t < NT; ++t)
#pragma unroll
                for (int n = 0; n < A_::NI; ++n)
#pragma unroll
                    for (int m = 0; m < A_::MI; ++m) {
                        mma_tf32(acc.v[t][m][n], alo[m], bhi[t][n]);
                        mma_tf32(acc.v[t][m][n], ahi[m], blo[t][n]);
                        mma_tf32(acc.v[t][m][n], ahi[m], bhi[t][n]);
                    }
#endif
        }
        ws.release_slice();
    }
}

// f(t, m, n, half, row, col_in_tile, v0, v1): the thread's 1x2 output strips (half 0: row g, half 1: row g+8 of the m16 tile)
template <int R, int CW, int NT, class F>
__device__ __forceinline__ void tile_foreach(Acc<R, CW, NT>& acc, F f) {
    using A_ = Acc<R, CW, NT>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wr = warp / A_::WCn, wc = warp % A_::WCn;
    const int g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int m = 0; m < A_::MI; ++m)
#pragma unroll
            for (int n = 0; n < A_::NI; ++n) {
                const int row = wr * A_::RW + m * 16 + g, col = wc * A_::CWW + n * 8 + 2 * tig;
                f(t, m, n, 0, row, col, acc.v[t][m][n][0], acc.v[t][m][n][1]);
                f(t, m, n, 1, row + 8, col, acc.v[t][m][n][2], acc.v[t][m][n][3]);
            }
}

// ------------------------------------------------------------------ small per-(sample, head) attention products
// C[hh][s*N+i][j] = scale * sum_d A[s*N+i][hh*64 + d] * B[s*N+j][hh*64 + d].
// Interleaved row ownership keeps LDS.128 conflict free (row strides are 4 mod 32 floats).
template <int TI, int TJ>
__device__ __forceinline__ void attn_nt_t(const float* __restrict__ sA, int lda, const float* __restrict__ sB, int ldb,
                                          float* __restrict__ sC, int ldc_head, int NP, int N, int S_act, int HCn, float scale) {
    const int IG = (N + TI - 1) / TI, JG = (N + TJ - 1) / TJ;
    const int per_head = S_act * IG * JG;
    const int total = HCn * per_head;
    for (int w0 = threadIdx.x; w0 < total; w0 += kThreads) {
        const int hh = w0 / per_head;
        const int w = w0 - hh * per_head;
        const int jg = w % JG;
        const int t2 = w / JG;
        const int ig = t2 % IG, s = t2 / IG;
        const float* pa[TI];
        const float* pb[TJ];
#pragma unroll
        for (int a = 0; a < TI; ++a) pa[a] = sA + (s * N + min(ig + a * IG, N - 1)) * lda + hh * 64;
#pragma unroll
        for (int b = 0; b < TJ; ++b) pb[b] = sB + (s * N + min(jg + b * JG, N - 1)) * ldb + hh * 64;
        float acc[TI][TJ];
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) acc[a][b] = 0.f;
#pragma unroll 4
        for (int d = 0; d < kDimHead; d += 4) {
            float4 av[TI], bv[TJ];
#pragma unroll
            for (int a = 0; a < TI; ++a) av[a] = *reinterpret_cast<const float4*>(pa[a] + d);
#pragma unroll
            for (int b = 0; b < TJ; ++b) bv[b] = *reinterpret_cast<const float4*>(pb[b] + d);
#pragma unroll
            for (int a = 0; a < TI; ++a)
#pragma unroll
                for (int b = 0; b < TJ; ++b) {
                    acc[a][b] = fmaf(av[a].x, bv[b].x, acc[a][b]);
                    acc[a][b] = fmaf(av[a].y, bv[b].y, acc[a][b]);
                    acc[a][b] = fmaf(av[a].z, bv[b].z, acc[a][b]);
                    acc[a][b] = fmaf(av[a].w, bv[b].w, acc[a][b]);
                }
        }
        float* cbase = sC + hh * ldc_head;
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) {
                const int i = ig + a * IG, j = jg + b * JG;
                if (i < N && j < N) cbase[(s * N + i) * NP + j] = scale * acc[a][b];
            }
    }
}
// picks the largest register tile that still gives (most of) the CTA's threads work
__device__ __forceinline__ void attn_nt(const float* sA, int lda, const float* sB, int ldb, float* sC, int ldc_head,
                                        int NP, int N, int S_act, int HCn, float scale) {
    const int n4 = (N + 3) / 4, n2 = (N + 1) / 2;
    if (HCn * S_act * n4 * n4 >= kThreads / 2)      attn_nt_t<4, 4>(sA, lda, sB, ldb, sC, ldc_head, NP, N, S_act, HCn, scale);
    else if (HCn * S_act * n2 * n2 >= kThreads / 2) attn_nt_t<2, 2>(sA, lda, sB, ldb, sC, ldc_head, NP, N, S_act, HCn, scale);
    else                                            attn_nt_t<1, 1>(sA, lda, sB, ldb, sC, ldc_head, NP, N, S_act, HCn, scale);
}

// C[i][d4] = sum_j P[hh][s*N+i][j] * B[s*N+j][hh*64 + d4]            (TRANS=false)
// C[j][d4] = sum_i P[hh][s*N+i][j] * B[s*N+i][hh*64 + d4]            (TRANS=true)
// store(hh, s, row_in_sample, d (multiple of 4), float4 value)
template <int TI, bool TRANS, class F>
__device__ __forceinline__ void attn_pv_t(const float* __restrict__ sPm, int ldp_head, int NP, const float* __restrict__ sB,
                                          int ldb, int N, int S_act, int HCn, F store) {
    const int IG = (N + TI - 1) / TI;
    const int per_head = S_act * IG * 16;
    const int total = HCn * per_head;
    for (int w0 = threadIdx.x; w0 < total; w0 += kThreads) {
        const int hh = w0 / per_head;
        const int w = w0 - hh * per_head;
        const int dg = w & 15;
        const int t2 = w >> 4;
        const int ig = t2 % IG, s = t2 / IG;
        int rows[TI];
#pragma unroll
        for (int a = 0; a < TI; ++a) rows[a] = min(ig + a * IG, N - 1);
        float4 acc[TI];
#pragma unroll
        for (int a = 0; a < TI; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* bp = sB + (s * N) * ldb + hh * 64 + dg * 4;
        const float* pp = sPm + hh * ldp_head + (s * N) * NP;
        if (TRANS) {
#pragma unroll 2
            for (int j = 0; j < N; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(bp + j * ldb);
#pragma unroll
                for (int a = 0; a < TI; ++a) {
                    const float pv = pp[j * NP + rows[a]];
                    acc[a].x = fmaf(pv, b.x, acc[a].x);
                    acc[a].y = fmaf(pv, b.y, acc[a].y);
                    acc[a].z = fmaf(pv, b.z, acc[a].z);
                    acc[a].w = fmaf(pv, b.w, acc[a].w);
                }
            }
        } else {
            // rows of P are padded to NP (multiple of 4) with zeros: one 16-byte load covers 4 values of j
            for (int j = 0; j < N; j += 4) {
                float4 p4[TI];
#pragma unroll
                for (int a = 0; a < TI; ++a) p4[a] = *reinterpret_cast<const float4*>(pp + rows[a] * NP + j);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 b = *reinterpret_cast<const float4*>(bp + min(j + u, N - 1) * ldb);   // pad columns carry p = 0
#pragma unroll
                    for (int a = 0; a < TI; ++a) {
                        const float pv = (u == 0) ? p4[a].x : (u == 1) ? p4[a].y : (u == 2) ? p4[a].z : p4[a].w;
                        acc[a].x = fmaf(pv, b.x, acc[a].x);
                        acc[a].y = fmaf(pv, b.y, acc[a].y);
                        acc[a].z = fmaf(pv, b.z, acc[a].z);
                        acc[a].w = fmaf(pv, b.w, acc[a].w);
                    }
                }
            }
        }
#pragma unroll
        for (int a = 0; a < TI; ++a)
            if (ig + a * IG < N) store(hh, s, ig + a * IG, dg * 4, acc[a]);
    }
}
template <bool TRANS, class F>
__device__ __forceinline__ void attn_pv(const float* sPm, int ldp_head, int NP, const float* sB, int ldb, int N,
                                        int S_act, int HCn, F store) {
    if (HCn * S_act * ((N + 3) / 4) * 16 >= kThreads)      attn_pv_t<4, TRANS>(sPm, ldp_head, NP, sB, ldb, N, S_act, HCn, store);
    else if (HCn * S_act * ((N + 1) / 2) * 16 >= kThreads) attn_pv_t<2, TRANS>(sPm, ldp_head, NP, sB, ldb, N, S_act, HCn, store);
    else                                                   attn_pv_t<1, TRANS>(sPm, ldp_head, NP, sB, ldb, N, S_act, HCn, store);
}

// ------------------------------------------------------------------ configuration
template <int HP, int R, int HC, int NS = kStages, int PN = kMaxBeads, int KM = 1>
struct Cfg {
    static constexpr int kKM = KM;           // weight-slice size multiplier (stage = KM * 12.5 KB)
    static constexpr int kHP = HP, kR = R, kHC = HC;
    static constexpr int kNumStages = NS;    // weight-ring depth
    static constexpr int kPN = PN;           // largest padded bead count this configuration accepts
    static constexpr int CWQ = 64 * HC;      // q / k' / v' tile width of one head chunk
    static constexpr int NCH = kHeads / HC;  // head chunks per layer
    static constexpr int LDH = HP + 4;       // [R][HP] activation buffers
    static constexpr int LDQ = 3 * CWQ + 4;  // q | k' | v' of one head chunk
    static constexpr int LDO = CWQ + 4;      // attention output of the chunk / its gradient
    static constexpr int LDF = 128 + 4;      // FF hidden chunk (aliases the qkv buffer)
    static constexpr int EPL = HP / 32;      // columns per lane in warp-per-row phases
    static constexpr int PSZ = HC * R * PN;          // attention probabilities of the chunk [HC][R][NP]
    // shared memory carve-up (float offsets)
    static constexpr int oN = 0;
    static constexpr int oNh = oN + R * LDH;
    static constexpr int oQKV = oNh + R * LDH;
    static constexpr int oO = oQKV + R * LDQ;
    static constexpr int oP = oO + R * LDO;
    static constexpr int oDS = oP + PSZ;
    static constexpr int oW = oDS + PSZ;
    static constexpr int oX = oW + NS * KM * kStageFloats;
    static constexpr int oV = oX + R * 4;
    static constexpr int oDX = oV + R * 4;
    static constexpr int oTmp = oDX + R * 4;
    static constexpr int oSeg = oTmp + R * 4;
    static constexpr int oBar = oSeg + kSegCap * 4;
    static constexpr int kFloats = oBar + 2 * NS;
    static constexpr size_t kSmemBytes = (size_t)kFloats * sizeof(float);
    static_assert(LDQ >= LDF, "FF chunk buffer must fit in the qkv buffer");
};

struct Ctx {
    float *sN, *sNh, *sQKV, *sO, *sP, *sDS, *sX, *sV, *sDX, *sTmp;
    WStream ws;
    float* stash;
    int rows_act, S_act;
};

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * expf(-0.5f * x * x) * 0.3989422804014327f;
}

// copy an [R][W] block (W multiple of 4) from shared (stride lds) to a dense global block
template <int R>
__device__ __forceinline__ void stash_store(float* __restrict__ dst, int W, const float* __restrict__ src, int lds) {
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < R * w4; idx += kThreads) {
        const int r = idx / w4, c = idx - r * w4;
        *reinterpret_cast<float4*>(dst + (size_t)r * W + c * 4) = *reinterpret_cast<const float4*>(src + r * lds + c * 4);
    }
}
// asynchronous reload (cp.async) of a dense global [R][W] block into shared (stride lds); complete with cp_async_wait_all
template <int R>
__device__ __forceinline__ void stash_load_async(float* __restrict__ dst, int lds, const float* __restrict__ src, int W) {
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < R * w4; idx += kThreads) {
        const int r = idx / w4, c = idx - r * w4;
        cp_async16(dst + r * lds + c * 4, src + (size_t)r * W + c * 4);
    }
}

// ------------------------------------------------------------------ warp-per-row phases
// LayerNorm of sN rows -> sNh, stats (mean, rstd) -> global; optionally stashes the input rows.
template <class C>
__device__ __forceinline__ void ln_forward_rows(const float* sN, float* sNh, const float* __restrict__ gam,
                                                const float* __restrict__ bet, int H, float* st_rows, float* st_stats) {
    constexpr int R = C::kR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += kWarps) {
        float x[C::EPL];
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            x[e] = (col < H) ? sN[r * C::LDH + col] : 0.f;
            s += x[e];
        }
        const float mean = warp_sum(s) / (float)H;
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            const float d = (col < H) ? x[e] - mean : 0.f;
            q += d * d;
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + kLnEps);
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            float y = 0.f;
            if (col < H) {
                y = (x[e] - mean) * rstd * __ldg(gam + col) + __ldg(bet + col);
                if (st_rows) st_rows[(size_t)r * H + col] = x[e];
            }
            sNh[r * C::LDH + col] = y;
        }
        if (lane == 0) { st_stats[r * 2] = mean; st_stats[r * 2 + 1] = rstd; }
    }
}

// GatedResidual forward (graph_transformer.py:202-205) on rows: a = sNh, n = sN -> out -> sN,
// then (if gam) LayerNorm(out) -> sNh.  Stashes a, gate, out (and LN stats).
template <class C>
__device__ __forceinline__ void gate_ln_forward_rows(float* sN, float* sNh, const float* __restrict__ ga,
                                                     const float* __restrict__ gb, int H, float* st_a, float* st_g,
                                                     float* st_out, const float* __restrict__ gam,
                                                     const float* __restrict__ bet, float* st_stats) {
    constexpr int R = C::kR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += kWarps) {
        float a[C::EPL], n[C::EPL], o[C::EPL];
        float z = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            const bool ok = col < H;
            a[e] = ok ? sNh[r * C::LDH + col] : 0.f;
            n[e] = ok ? sN[r * C::LDH + col] : 0.f;
            if (ok) z += a[e] * __ldg(ga + col) + n[e] * __ldg(gb + col);
        }
        z = warp_sum(z);
        const float g = 1.0f / (1.0f + expf(-z));
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            o[e] = a[e] * g + n[e] * (1.0f - g);
            if (col < H) {
                st_a[(size_t)r * H + col] = a[e];
                st_out[(size_t)r * H + col] = o[e];
            } else {
                o[e] = 0.f;
            }
            sN[r * C::LDH + col] = o[e];
            s += o[e];
        }
        if (lane == 0) st_g[r] = g;
        if (gam == nullptr) continue;
        const float mean = warp_sum(s) / (float)H;
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            const float d = (col < H) ? o[e] - mean : 0.f;
            q += d * d;
        }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + kLnEps);
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            sNh[r * C::LDH + col] = (col < H) ? (o[e] - mean) * rstd * __ldg(gam + col) + __ldg(bet + col) : 0.f;
        }
        if (lane == 0) { st_stats[r * 2] = mean; st_stats[r * 2 + 1] = rstd; }
    }
}

// Reverse of [LayerNorm ->] GatedResidual on rows.
//   dout = sN (+ LN-backward of sNh through (st_ln_in, stats, gam) when gam != nullptr)
//   d(gate input a) -> sNh,  d(residual n) -> sN.      a, n, g come from the stash.
template <class C>
__device__ __forceinline__ void gate_backward_rows(float* sN, float* sNh, int H, const float* __restrict__ gam,
                                                   const float* st_ln_in, const float* st_stats, const float* st_a,
                                                   const float* st_n, const float* st_g, const float* __restrict__ ga,
                                                   const float* __restrict__ gb) {
    constexpr int R = C::kR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += kWarps) {
        float d[C::EPL], a[C::EPL], n[C::EPL];
        // issue the global (stash) loads first so their latency overlaps the shared-memory work
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            const bool ok = col < H;
            a[e] = ok ? st_a[(size_t)r * H + col] : 0.f;
            n[e] = ok ? st_n[(size_t)r * H + col] : 0.f;
            d[e] = ok ? sN[r * C::LDH + col] : 0.f;
        }
        const float g = st_g[r];
        if (gam != nullptr) {
            const float mean = st_stats[r * 2], rstd = st_stats[r * 2 + 1];
            float y[C::EPL], dy[C::EPL];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                const int col = lane * C::EPL + e;
                const bool ok = col < H;
                y[e] = ok ? (st_ln_in[(size_t)r * H + col] - mean) * rstd : 0.f;
                dy[e] = ok ? sNh[r * C::LDH + col] * __ldg(gam + col) : 0.f;
                s1 += dy[e];
                s2 += dy[e] * y[e];
            }
            s1 = warp_sum(s1) / (float)H;
            s2 = warp_sum(s2) / (float)H;
#pragma unroll
            for (int e = 0; e < C::EPL; ++e) {
                const int col = lane * C::EPL + e;
                if (col < H) d[e] += rstd * (dy[e] - s1 - y[e] * s2);
            }
        }
        float dg = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) dg += d[e] * (a[e] - n[e]);
        dg = warp_sum(dg);
        const float dz = dg * g * (1.0f - g);
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            const bool ok = col < H;
            sNh[r * C::LDH + col] = ok ? d[e] * g + dz * __ldg(ga + col) : 0.f;
            sN[r * C::LDH + col] = ok ? d[e] * (1.0f - g) + dz * __ldg(gb + col) : 0.f;
        }
    }
}

// sN += LayerNorm-backward(sNh) through (st_ln_in, stats, gam)
template <class C>
__device__ __forceinline__ void ln_backward_rows(float* sN, const float* sNh, int H, const float* __restrict__ gam,
                                                 const float* st_ln_in, const float* st_stats) {
    constexpr int R = C::kR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += kWarps) {
        const float mean = st_stats[r * 2], rstd = st_stats[r * 2 + 1];
        float y[C::EPL], dy[C::EPL];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            const bool ok = col < H;
            y[e] = ok ? (st_ln_in[(size_t)r * H + col] - mean) * rstd : 0.f;
            dy[e] = ok ? sNh[r * C::LDH + col] * __ldg(gam + col) : 0.f;
            s1 += dy[e];
            s2 += dy[e] * y[e];
        }
        s1 = warp_sum(s1) / (float)H;
        s2 = warp_sum(s2) / (float)H;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            if (col < H) sN[r * C::LDH + col] += rstd * (dy[e] - s1 - y[e] * s2);
        }
    }
}

// softmax over j of the chunk's probability rows [HC*R][NP] (in place), 8 lanes per row; pad columns are zeroed.
template <int R, int HC>
__device__ __forceinline__ void softmax_rows(float* sP, int NP, int N, int rows_act) {
    const int q = threadIdx.x & 7;
    for (int base = 0; base < HC * R; base += kThreads / 8) {
        const int hr = base + (threadIdx.x >> 3);
        const bool ok = (hr % R) < rows_act;
        float* p = sP + (ok ? hr : 0) * NP;
        float lg[kMaxBeads / 8];
        float m = -INFINITY;
#pragma unroll
        for (int t = 0; t < kMaxBeads / 8; ++t) {
            const int j = q + t * 8;
            lg[t] = (j < N) ? p[j] : -INFINITY;
            m = fmaxf(m, lg[t]);
        }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < kMaxBeads / 8; ++t) {
            const int j = q + t * 8;
            lg[t] = (j < N) ? expf(lg[t] - m) : 0.f;
            s += lg[t];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (ok) {
#pragma unroll
            for (int t = 0; t < kMaxBeads / 8; ++t) {
                const int j = q + t * 8;
                if (j < NP) p[j] = (j < N) ? lg[t] / s : 0.f;
            }
        }
    }
}

// ds = p * (dp - sum_j p dp) in place on the sDS rows, 8 lanes per row
template <int R, int HC>
__device__ __forceinline__ void softmax_backward_rows(float* sDS, const float* sP, int NP, int N, int rows_act) {
    const int q = threadIdx.x & 7;
    for (int base = 0; base < HC * R; base += kThreads / 8) {
        const int hr = base + (threadIdx.x >> 3);
        const bool ok = (hr % R) < rows_act;
        float* d = sDS + (ok ? hr : 0) * NP;
        const float* p = sP + (ok ? hr : 0) * NP;
        float pv[kMaxBeads / 8], dv[kMaxBeads / 8];
        float s = 0.f;
#pragma unroll
        for (int t = 0; t < kMaxBeads / 8; ++t) {
            const int j = q + t * 8;
            pv[t] = (j < N) ? p[j] : 0.f;
            dv[t] = (j < N) ? d[j] : 0.f;
            s += pv[t] * dv[t];
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (ok) {
#pragma unroll
            for (int t = 0; t < kMaxBeads / 8; ++t) {
                const int j = q + t * 8;
                if (j < NP) d[j] = pv[t] * (dv[t] - s);
            }
        }
    }
}

// ------------------------------------------------------------------ forward pass (energy) for one group of samples
template <class C>
__device__ void forward_pass(const ModelDev& M, Ctx& c, float t_norm, const float* __restrict__ t_rows) {
    constexpr int HP = C::kHP, R = C::kR, HC = C::kHC;
    const int tid = threadIdx.x;
    const int N = M.N, NP = M.NP, H = M.H;

    // layer-0 node stream: W_n [onehot_i, t] + b_n   (graph_transformer.py:99-103), zero padding elsewhere
    for (int idx = tid; idx < R * HP; idx += kThreads) {
        const int r = idx / HP, d = idx - r * HP;
        float v = 0.f;
        if (r < c.rows_act && d < H) v = __ldg(M.emb + (r % N) * H + d) + (t_rows != nullptr ? __ldg(t_rows + r / N) : t_norm) * __ldg(M.embt + d);
        c.sN[r * C::LDH + d] = v;
    }
    __syncthreads();
    ln_forward_rows<C>(c.sN, c.sNh, M.layer[0].ln1_g, M.layer[0].ln1_b, H, c.stash + M.off[ST_NIN],
                               c.stash + M.off[ST_STAT1]);
    __syncthreads();

    for (int l = 0; l < M.L; ++l) {
        const LayerDev& W = M.layer[l];
        float* st = c.stash + (size_t)l * M.layer_floats;

        Acc<R, HP, 1> acc_a;
        acc_a.zero();
        for (int hc = 0; hc < C::NCH; ++hc) {
            {   // q | k | v of the head chunk:  n_hat [R][H] x Wqkv_f[l][hc] [H][3*CWQ]
                Acc<R, C::CWQ, 3> acc;
                acc.zero();
                gemm_acc<R, C::CWQ, 3, C::kKM>(c.ws, c.sNh, C::LDH, H, acc);
                tile_foreach<R, C::CWQ, 3>(acc, [&](int t, int, int, int, int row, int col, float v0, float v1) {
                    const float2 b = __ldg(reinterpret_cast<const float2*>(W.bqkv + hc * 3 * C::CWQ + t * C::CWQ + col));
                    float o[2] = {v0 + b.x, v1 + b.y};
                    if (t > 0) {   // k' = k + A x_j, v' = v + A x_j
                        const float x0 = c.sX[row * 4], x1 = c.sX[row * 4 + 1], x2 = c.sX[row * 4 + 2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + (hc * C::CWQ + col + e) * 4));
                            o[e] += a4.x * x0 + a4.y * x1 + a4.z * x2;
                        }
                    }
                    *reinterpret_cast<float2*>(c.sQKV + row * C::LDQ + t * C::CWQ + col) = make_float2(o[0], o[1]);
                });
            }
            __syncthreads();
            stash_store<R>(st + M.off[ST_QKV] + (size_t)hc * R * 3 * C::CWQ, 3 * C::CWQ, c.sQKV, C::LDQ);
            // logits: s * q_i . k'_j
            attn_nt(c.sQKV, C::LDQ, c.sQKV + C::CWQ, C::LDQ, c.sP, R * NP, NP, N, c.S_act, HC, kAttnScale);
            __syncthreads();
            softmax_rows<R, HC>(c.sP, NP, N, c.rows_act);
            __syncthreads();
            {   // stash p (dense [HC][R][NP] block)
                float* dst = st + M.off[ST_P] + (size_t)hc * HC * R * NP;
                for (int idx = tid; idx < (HC * R * NP) / 4; idx += kThreads)
                    reinterpret_cast<float4*>(dst)[idx] = reinterpret_cast<const float4*>(c.sP)[idx];
            }
            // o_i = sum_j p_ij v'_j - A x_i + c
            attn_pv<false>(c.sP, R * NP, NP, c.sQKV + 2 * C::CWQ, C::LDQ, N, c.S_act, HC,
                           [&](int hh, int s, int i, int d, const float4& a) {
                const int row = s * N + i;
                const int col = hh * 64 + d;
                const float x0 = c.sX[row * 4], x1 = c.sX[row * 4 + 1], x2 = c.sX[row * 4 + 2];
                const float4 cc = __ldg(reinterpret_cast<const float4*>(W.cvec + hc * C::CWQ + col));
                float o[4] = {a.x + cc.x, a.y + cc.y, a.z + cc.z, a.w + cc.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + (hc * C::CWQ + col + e) * 4));
                    o[e] -= a4.x * x0 + a4.y * x1 + a4.z * x2;
                }
                *reinterpret_cast<float4*>(c.sO + row * C::LDO + col) = make_float4(o[0], o[1], o[2], o[3]);
            });
            __syncthreads();
            // att += o_chunk [R][CWQ] x Wo_f[l][hc] [CWQ][HP]
            gemm_acc<R, HP, 1, C::kKM>(c.ws, c.sO, C::LDO, C::CWQ, acc_a);
        }
        tile_foreach<R, HP, 1>(acc_a, [&](int, int, int, int, int row, int col, float v0, float v1) {
            const float2 b = __ldg(reinterpret_cast<const float2*>(W.bo + col));
            *reinterpret_cast<float2*>(c.sNh + row * C::LDH + col) = make_float2(v0 + b.x, v1 + b.y);
        });
        __syncthreads();
        // gated residual 1 + LayerNorm 2
        gate_ln_forward_rows<C>(c.sN, c.sNh, W.g1a, W.g1b, H, st + M.off[ST_ATT], st + M.off[ST_G1],
                                        st + M.off[ST_M], W.ln2_g, W.ln2_b, st + M.off[ST_STAT2]);
        __syncthreads();
        // feed-forward, 128 hidden columns at a time
        Acc<R, HP, 1> acc_f;
        acc_f.zero();
        float* sH1 = c.sQKV;
        for (int ch = 0; ch < M.nch; ++ch) {
            Acc<R, 128, 1> acc1;
            acc1.zero();
            gemm_acc<R, 128, 1, C::kKM>(c.ws, c.sNh, C::LDH, H, acc1);
            tile_foreach<R, 128, 1>(acc1, [&](int, int, int, int, int row, int col, float v0, float v1) {
                const float2 b = __ldg(reinterpret_cast<const float2*>(W.b1 + ch * 128 + col));
                const float2 pre = make_float2(v0 + b.x, v1 + b.y);
                *reinterpret_cast<float2*>(st + M.off[ST_H1] + (size_t)row * (4 * H) + ch * 128 + col) = pre;
                *reinterpret_cast<float2*>(sH1 + row * C::LDF + col) = make_float2(gelu_f(pre.x), gelu_f(pre.y));
            });
            __syncthreads();
            gemm_acc<R, HP, 1, C::kKM>(c.ws, sH1, C::LDF, 128, acc_f);
        }
        tile_foreach<R, HP, 1>(acc_f, [&](int, int, int, int, int row, int col, float v0, float v1) {
            const float2 b = __ldg(reinterpret_cast<const float2*>(W.b2 + col));
            *reinterpret_cast<float2*>(c.sNh + row * C::LDH + col) = make_float2(v0 + b.x, v1 + b.y);
        });
        __syncthreads();
        // gated residual 2 (+ next layer's LayerNorm 1; its input rows are the next layer's n_in stash)
        const bool last = (l + 1 == M.L);
        float* stn = st + M.layer_floats;
        gate_ln_forward_rows<C>(c.sN, c.sNh, W.g2a, W.g2b, H, st + M.off[ST_FF], st + M.off[ST_G2],
                                        stn + M.off[ST_NIN],
                                        last ? nullptr : M.layer[l + 1].ln1_g, last ? nullptr : M.layer[l + 1].ln1_b,
                                        last ? nullptr : stn + M.off[ST_STAT1]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ reverse pass: sDX[r][0..2] = d sum(E) / d x_r
template <class C>
__device__ void backward_pass(const ModelDev& M, Ctx& c) {
    constexpr int HP = C::kHP, R = C::kR, HC = C::kHC;
    const int tid = threadIdx.x;
    const int N = M.N, NP = M.NP, H = M.H;

    for (int idx = tid; idx < R * HP; idx += kThreads) {
        const int r = idx / HP, d = idx - r * HP;
        c.sN[r * C::LDH + d] = (d < H) ? __ldg(M.dec_w + d) : 0.f;      // dE_r/dn_r = w_dec  (node_decoder, :106)
    }
    for (int idx = tid; idx < R * 4; idx += kThreads) c.sDX[idx] = 0.f;
    __syncthreads();

    for (int l = M.L - 1; l >= 0; --l) {
        const LayerDev& W = M.layer[l];
        float* st = c.stash + (size_t)l * M.layer_floats;

        // gated residual 2 backward: d ff -> sNh, d m (partial) -> sN
        gate_backward_rows<C>(c.sN, c.sNh, H, nullptr, nullptr, nullptr, st + M.off[ST_FF], st + M.off[ST_M],
                                      st + M.off[ST_G2], W.g2a, W.g2b);
        __syncthreads();
        // feed-forward backward
        Acc<R, HP, 1> acc_dm;
        acc_dm.zero();
        float* sH1 = c.sQKV;
        for (int ch = 0; ch < M.nch; ++ch) {
            Acc<R, 128, 1> acc1;
            acc1.zero();
            // fetch this thread's pre-activations now; the loads complete while the GEMM runs
            float2 pre[Acc<R, 128, 1>::MI][Acc<R, 128, 1>::NI][2];
            tile_foreach<R, 128, 1>(acc1, [&](int, int m, int n, int half, int row, int col, float, float) {
                pre[m][n][half] = *reinterpret_cast<const float2*>(st + M.off[ST_H1] + (size_t)row * (4 * H) + ch * 128 + col);
            });
            gemm_acc<R, 128, 1, C::kKM>(c.ws, c.sNh, C::LDH, H, acc1);          // d act = d ff x W2_b[ch]
            tile_foreach<R, 128, 1>(acc1, [&](int, int m, int n, int half, int row, int col, float v0, float v1) {
                *reinterpret_cast<float2*>(sH1 + row * C::LDF + col) =
                    make_float2(v0 * gelu_grad_f(pre[m][n][half].x), v1 * gelu_grad_f(pre[m][n][half].y));
            });
            __syncthreads();
            gemm_acc<R, HP, 1, C::kKM>(c.ws, sH1, C::LDF, 128, acc_dm);            // d m_hat += d h1 x W1_b[ch]
        }
        tile_foreach<R, HP, 1>(acc_dm, [&](int, int, int, int, int row, int col, float v0, float v1) {
            *reinterpret_cast<float2*>(c.sNh + row * C::LDH + col) = make_float2(v0, v1);
        });
        __syncthreads();
        // LayerNorm 2 backward + gated residual 1 backward: d att -> sNh, d n_in (residual part) -> sN
        gate_backward_rows<C>(c.sN, c.sNh, H, W.ln2_g, st + M.off[ST_M], st + M.off[ST_STAT2], st + M.off[ST_ATT],
                                      st + M.off[ST_NIN], st + M.off[ST_G1], W.g1a, W.g1b);
        __syncthreads();

        Acc<R, HP, 1> acc_dn;
        acc_dn.zero();
        for (int hc = 0; hc < C::NCH; ++hc) {
            // start reloading q | k' | v' and p of this chunk; the copies land while the d o GEMM runs
            stash_load_async<R>(c.sQKV, C::LDQ, st + M.off[ST_QKV] + (size_t)hc * R * 3 * C::CWQ, 3 * C::CWQ);
            {
                const float* src = st + M.off[ST_P] + (size_t)hc * HC * R * NP;
                for (int idx = tid; idx < (HC * R * NP) / 4; idx += kThreads) cp_async16(c.sP + idx * 4, src + idx * 4);
            }
            {   // d o_chunk = d att x Wo_b[l][hc]   [R][H] x [H][CWQ]
                Acc<R, C::CWQ, 1> acc;
                acc.zero();
                gemm_acc<R, C::CWQ, 1, C::kKM>(c.ws, c.sNh, C::LDH, H, acc);
                tile_foreach<R, C::CWQ, 1>(acc, [&](int, int, int, int, int row, int col, float v0, float v1) {
                    *reinterpret_cast<float2*>(c.sO + row * C::LDO + col) = make_float2(v0, v1);
                });
            }
            cp_async_wait_all();
            __syncthreads();
            // dp_ij = do_i . v'_j
            attn_nt(c.sO, C::LDO, c.sQKV + 2 * C::CWQ, C::LDQ, c.sDS, R * NP, NP, N, c.S_act, HC, 1.0f);
            __syncthreads();
            softmax_backward_rows<R, HC>(c.sDS, c.sP, NP, N, c.rows_act);
            __syncthreads();
            if (l > 0) {   // dq_i = s sum_j ds_ij k'_j  -> v' columns (v' is dead after dp)
                attn_pv<false>(c.sDS, R * NP, NP, c.sQKV + C::CWQ, C::LDQ, N, c.S_act, HC,
                               [&](int hh, int s, int i, int d, const float4& a) {
                    *reinterpret_cast<float4*>(c.sQKV + (s * N + i) * C::LDQ + 2 * C::CWQ + hh * 64 + d) =
                        make_float4(kAttnScale * a.x, kAttnScale * a.y, kAttnScale * a.z, kAttnScale * a.w);
                });
                __syncthreads();
            }
            // dk'_j = s sum_i ds_ij q_i  -> k' columns (k' is dead after dq)
            attn_pv<true>(c.sDS, R * NP, NP, c.sQKV, C::LDQ, N, c.S_act, HC,
                          [&](int hh, int s, int j, int d, const float4& a) {
                *reinterpret_cast<float4*>(c.sQKV + (s * N + j) * C::LDQ + C::CWQ + hh * 64 + d) =
                    make_float4(kAttnScale * a.x, kAttnScale * a.y, kAttnScale * a.z, kAttnScale * a.w);
            });
            __syncthreads();
            // dv'_j = sum_i p_ij do_i  -> q columns (q is dead after dk')
            attn_pv<true>(c.sP, R * NP, NP, c.sO, C::LDO, N, c.S_act, HC,
                          [&](int hh, int s, int j, int d, const float4& a) {
                *reinterpret_cast<float4*>(c.sQKV + (s * N + j) * C::LDQ + hh * 64 + d) = a;
            });
            __syncthreads();
            // dx_r += sum over the chunk's heads of A_h^T (dk'_r + dv'_r - do_r)   (fixed summation order: deterministic)
            for (int idx = tid; idx < c.rows_act * 3; idx += kThreads) {
                const int r = idx / 3, cc = idx - r * 3;
                float s = 0.f;
#pragma unroll
                for (int hh = 0; hh < HC; ++hh) {
                    const float* dk = c.sQKV + r * C::LDQ + C::CWQ + hh * 64;
                    const float* dv = c.sQKV + r * C::LDQ + hh * 64;
                    const float* dO = c.sO + r * C::LDO + hh * 64;
                    const float* Ah = W.A + (hc * C::CWQ + hh * 64) * 4 + cc;
#pragma unroll 4
                    for (int d = 0; d < 64; d += 4) {
                        const float4 a = *reinterpret_cast<const float4*>(dk + d);
                        const float4 b = *reinterpret_cast<const float4*>(dv + d);
                        const float4 o = *reinterpret_cast<const float4*>(dO + d);
                        s = fmaf(__ldg(Ah + (d + 0) * 4), a.x + b.x - o.x, s);
                        s = fmaf(__ldg(Ah + (d + 1) * 4), a.y + b.y - o.y, s);
                        s = fmaf(__ldg(Ah + (d + 2) * 4), a.z + b.z - o.z, s);
                        s = fmaf(__ldg(Ah + (d + 3) * 4), a.w + b.w - o.w, s);
                    }
                }
                c.sDX[r * 4 + cc] += s;
            }
            if (l > 0) {   // d n_hat += [dv' | dk' | dq] [R][3*CWQ] x Wqkv_b[l][hc] [3*CWQ][HP]
                gemm_acc<R, HP, 1, C::kKM>(c.ws, c.sQKV, C::LDQ, 3 * C::CWQ, acc_dn);
            } else {
                __syncthreads();
            }
        }
        if (l > 0) {
            tile_foreach<R, HP, 1>(acc_dn, [&](int, int, int, int, int row, int col, float v0, float v1) {
                *reinterpret_cast<float2*>(c.sNh + row * C::LDH + col) = make_float2(v0, v1);
            });
            __syncthreads();
            ln_backward_rows<C>(c.sN, c.sNh, H, W.ln1_g, st + M.off[ST_NIN], st + M.off[ST_STAT1]);
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <class C, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
dff_fused_kernel(const __grid_constant__ ModelDev M, const __grid_constant__ StepArgs A) {
    constexpr int R = C::kR;
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x;
    const int N = M.N;

    Ctx c;
    c.sN = smem + C::oN;   c.sNh = smem + C::oNh; c.sQKV = smem + C::oQKV; c.sO = smem + C::oO;
    c.sP = smem + C::oP;   c.sDS = smem + C::oDS; c.sX = smem + C::oX;     c.sV = smem + C::oV;
    c.sDX = smem + C::oDX; c.sTmp = smem + C::oTmp;
    c.stash = M.scratch + (size_t)blockIdx.x * M.scratch_per_cta;

    const int n_groups = (A.B + M.S - 1) / M.S;
    int my_groups = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) ++my_groups;

    c.ws.stage_base = smem + C::oW;
    c.ws.full = reinterpret_cast<uint64_t*>(smem + C::oBar);
    c.ws.n = 0;
    c.ws.nst = C::kNumStages;
    c.ws.stage_floats = C::kKM * kStageFloats;
    c.ws.sh = (C::kNumStages == 4) ? 2u : 1u;
    c.ws.nseg = A.need_backward ? M.nseg_all : M.nseg_fwd;
    c.ws.segs = M.segs;
    if (c.ws.nseg <= kSegCap) {      // keep the producer's table off the global-memory critical path
        Seg* sseg = reinterpret_cast<Seg*>(smem + C::oSeg);
        for (int i = tid; i < c.ws.nseg; i += kThreads) sseg[i] = M.segs[i];
        c.ws.segs = sseg;
    }
    c.ws.seg_i = 0; c.ws.slice_i = 0; c.ws.issued = 0;
    c.ws.total = (uint32_t)my_groups * (uint32_t)A.n_steps * (A.need_backward ? M.nslice_all : M.nslice_fwd);
    if (tid == 0) {
        for (int i = 0; i < C::kNumStages; ++i) mbar_init(c.ws.full + i, 1);
        fence_barrier_init();
    }
    for (int idx = tid; idx < C::oW; idx += kThreads) smem[idx] = 0.f;       // activations / attention buffers
    for (int idx = C::oX + tid; idx < C::oSeg; idx += kThreads) smem[idx] = 0.f;
    __syncthreads();
    if (tid == 0) c.ws.prefill();

    uint32_t flags = 0;
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int s0 = g * M.S;
        c.S_act = min(M.S, A.B - s0);
        c.rows_act = c.S_act * N;
        for (int idx = tid; idx < R * 3; idx += kThreads) {
            const int r = idx / 3, cc = idx - r * 3;
            const bool ok = r < c.rows_act;
            c.sX[r * 4 + cc] = ok ? A.x[((size_t)s0 * N + r) * 3 + cc] : 0.f;
            c.sV[r * 4 + cc] = (ok && A.v != nullptr) ? A.v[((size_t)s0 * N + r) * 3 + cc] : 0.f;
        }
        __syncthreads();

        for (int step = 0; step < A.n_steps; ++step) {
            // center_zero (utils.py:65-70); entry check of assert_center_zero (utils.py:73-86) as a flag
            if (tid < c.S_act * 3) {
                const int s = tid / 3, cc = tid - s * 3;
                float m = 0.f;
                for (int i = 0; i < N; ++i) m += c.sX[(s * N + i) * 4 + cc];
                m = m / (float)N;
                if (A.mode == MODE_DDPM && fabsf(m) >= 1e-3f) flags |= 2u;
                for (int i = 0; i < N; ++i) c.sX[(s * N + i) * 4 + cc] -= m;
            }
            __syncthreads();
            const int it = A.t_start - step;
            const float t_norm = (A.mode == MODE_DDPM) ? (float)it / (float)A.T : A.t_norm;

            forward_pass<C>(M, c, t_norm, (A.mode == MODE_SCORE && A.t_rows != nullptr) ? A.t_rows + s0 : nullptr);
            if (A.energy_out != nullptr) {   // node_decoder (graph_transformer.py:106)
                const int lane = tid & 31, warp = tid >> 5;
                for (int r = warp; r < c.rows_act; r += kWarps) {
                    float s = 0.f;
                    for (int d = lane; d < M.H; d += 32) s += c.sN[r * C::LDH + d] * __ldg(M.dec_w + d);
                    s = warp_sum(s);
                    if (lane == 0) A.energy_out[(size_t)s0 * N + r] = s + M.dec_b;
                }
            }
            __syncthreads();
            if (A.need_backward) backward_pass<C>(M, c);
            if (!M.conservative) {
                // non-conservative head (graph_transformer.py:62-65, 112-113): the prediction is node_decoder(nodes);
                // stored negated so that the samplers below read eps = -sDX exactly as in the conservative case
                const int lane = tid & 31, warp = tid >> 5;
                for (int r = warp; r < c.rows_act; r += kWarps) {
                    float s0_ = 0.f, s1_ = 0.f, s2_ = 0.f;
                    for (int d = lane; d < M.H; d += 32) {
                        const float v = c.sN[r * C::LDH + d];
                        s0_ = fmaf(v, __ldg(M.dec_w + d), s0_); s1_ = fmaf(v, __ldg(M.dec_w + M.H + d), s1_); s2_ = fmaf(v, __ldg(M.dec_w + 2 * M.H + d), s2_);
                    }
                    s0_ = warp_sum(s0_); s1_ = warp_sum(s1_); s2_ = warp_sum(s2_);
                    if (lane == 0) { c.sDX[r * 4] = -(s0_ + M.dec_b3[0]); c.sDX[r * 4 + 1] = -(s1_ + M.dec_b3[1]); c.sDX[r * 4 + 2] = -(s2_ + M.dec_b3[2]); }
                }
            }
            __syncthreads();

            if (A.mode == MODE_SCORE) {
                if (A.eps_out != nullptr)
                    for (int idx = tid; idx < c.rows_act * 3; idx += kThreads) {
                        const int r = idx / 3, cc = idx - r * 3;
                        A.eps_out[((size_t)s0 * N + r) * 3 + cc] = -c.sDX[r * 4 + cc];
                    }
            } else if (A.mode == MODE_DDPM) {
                // p_mean_variance + p_sample + loop tail (models/ddpm.py:195-232, 248-251); eps = -dE/dx
                if (tid < c.S_act * 3) {
                    const int s = tid / 3, cc = tid - s * 3;
                    const float cr = A.sched[0][it], crm1 = A.sched[1][it], c1 = A.sched[2][it], c2 = A.sched[3][it];
                    const float sigma = (it == 0) ? 0.f : expf(0.5f * A.sched[4][it]);
                    float me = 0.f;
                    for (int i = 0; i < N; ++i) me += -c.sDX[(s * N + i) * 4 + cc];
                    me = me / (float)N;
                    float mx0 = 0.f;
                    for (int i = 0; i < N; ++i) {
                        const int o = (s * N + i) * 4 + cc;
                        const float e = -c.sDX[o] - me;
                        const float x0 = cr * c.sX[o] - crm1 * e;
                        c.sTmp[o] = x0;
                        mx0 += x0;
                    }
                    mx0 = mx0 / (float)N;
                    float mz = 0.f;
                    for (int i = 0; i < N; ++i) {
                        const int o = (s * N + i) * 4 + cc;
                        const size_t ge = ((size_t)(s0 + s) * N + i) * 3 + cc;
                        const float z = (A.noise != nullptr)
                                            ? A.noise[(size_t)step * A.B * N * 3 + ge]
                                            : philox_normal(A.seed, A.offset + (unsigned long long)step, (uint32_t)ge);
                        c.sDX[o] = z;
                        mz += z;
                    }
                    mz = mz / (float)N;
                    float mn = 0.f;
                    for (int i = 0; i < N; ++i) {
                        const int o = (s * N + i) * 4 + cc;
                        const float mean = c1 * (c.sTmp[o] - mx0) + c2 * c.sX[o];
                        float xn = mean + sigma * (c.sDX[o] - mz);
                        if (!(fabsf(xn) <= 3.0e38f)) flags |= 4u;
                        if (xn > 1000.f || xn < -1000.f) { flags |= 1u; xn = fminf(fmaxf(xn, -1000.f), 1000.f); }
                        c.sX[o] = xn;
                        mn += xn;
                    }
                    mn = mn / (float)N;
                    for (int i = 0; i < N; ++i) c.sX[(s * N + i) * 4 + cc] -= mn;
                }
            } else {
                // ForcesWrapper (dynamics/langevin.py:78-87) + _langevin_timestep / _overdamped_timestep
                for (int idx = tid; idx < c.rows_act * 3; idx += kThreads) {
                    const int r = idx / 3, cc = idx - r * 3;
                    const int o = r * 4 + cc;
                    const size_t ge = ((size_t)s0 * N + r) * 3 + cc;
                    const float z = (A.noise != nullptr)
                                        ? A.noise[(size_t)step * A.B * N * 3 + ge]
                                        : philox_normal(A.seed, A.offset + (unsigned long long)step, (uint32_t)ge);
                    const float F = (-c.sDX[o]) * A.force_scale;
                    float x = c.sX[o];
                    if (A.mode == MODE_BAOAB) {
                        const float m = __ldg(A.mass + (r % N));
                        float v = c.sV[o];
                        v = v + A.dt * F / m;                 // B
                        x = x + v * A.dt / 2.0f;              // A
                        const float eta = sqrtf(A.inv_beta / m) * z;
                        v = v * A.vscale;                      // O
                        v = v + A.noisescale * eta;
                        x = x + v * A.dt / 2.0f;              // A
                        c.sV[o] = v;
                    } else {
                        x = x + F * A.dtau + A.bd_sigma * z;       // bd_sigma = sqrt(2 dtau / beta)
                    }
                    if (!(fabsf(x) <= 3.0e38f)) flags |= 4u;
                    c.sX[o] = x;
                }
                if (A.save_interval > 0 && (step + 1) % A.save_interval == 0) {
                    __syncthreads();
                    const int f = step / A.save_interval;
                    if (A.frames != nullptr)
                        for (int idx = tid; idx < c.rows_act * 3; idx += kThreads) {
                            const int r = idx / 3, cc = idx - r * 3;
                            A.frames[((size_t)f * A.B + s0) * N * 3 + (size_t)r * 3 + cc] = c.sX[r * 4 + cc];
                        }
                    if (A.ke != nullptr && A.mode == MODE_BAOAB && tid < c.S_act) {
                        float ke = 0.f;
                        for (int i = 0; i < N; ++i) {
                            const float* v = c.sV + (tid * N + i) * 4;
                            ke += __ldg(A.mass + i) * v[0] * v[0] + __ldg(A.mass + i) * v[1] * v[1] + __ldg(A.mass + i) * v[2] * v[2];
                        }
                        A.ke[(size_t)f * A.B + s0 + tid] = 0.5f * ke;
                    }
                }
            }
            __syncthreads();
        }
        if (A.mode != MODE_SCORE) {
            for (int idx = tid; idx < c.rows_act * 3; idx += kThreads) {
                const int r = idx / 3, cc = idx - r * 3;
                A.x[((size_t)s0 * N + r) * 3 + cc] = c.sX[r * 4 + cc];
                if (A.v != nullptr) A.v[((size_t)s0 * N + r) * 3 + cc] = c.sV[r * 4 + cc];
            }
        }
        __syncthreads();
    }
    if (A.flags != nullptr && flags != 0) atomicOr(A.flags, flags);
}

}  // namespace dff
