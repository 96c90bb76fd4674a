// dff_kernel.cuh -- the fused score-network + integrator kernel (sm_100a, fp32 SIMT GEMMs, TMA bulk
// weight streaming).  One CTA owns a group of S whole samples (R = S*N node rows) and runs, for each of
// n_steps diffusion / MD steps, the complete forward pass of the collapsed graph transformer, its
// hand-written reverse pass w.r.t. x, and the DDPM-posterior / BAOAB / Brownian update -- coordinates
// never leave shared memory between steps.  Samples never interact, so there is no grid-wide sync.
//
// Math: SURVEY.md Appendix A restated with k'_j = k_j + A x_j, v'_j = v_j + A x_j so that every head is
// plain attention over (q, k', v'):   o_i = sum_j p_ij v'_j - A x_i + c,
//   dx_r += A_h^T (dk'_r + dv'_r - do_r)   (oracle/collapsed_ref.py is the CPU statement of the same).
// Reference lines replaced: models/graph_transformer.py:87-111,143-159,178-329; models/ddpm.py:195-251;
// dynamics/langevin.py:75-92; dynamics/langevin_cgnet.py:447-542,737-771; utils.py:65-86.
#pragma once
#include <math.h>
#include "dff_common.cuh"

namespace dff {

// ------------------------------------------------------------------ weight stream (TMA bulk + mbarrier ring)
struct WStream {
    float* stage_base;
    uint64_t* bars;
    uint32_t n;  // slices consumed so far (identical in every thread)
    // producer cursor (meaningful in thread 0 only)
    const Seg* segs;
    int nseg, seg_i;
    uint32_t slice_i, issued, total;

    __device__ __forceinline__ void issue_one() {
        if (issued >= total) return;
        const uint32_t st = issued & (kStages - 1);
        const Seg sg = segs[seg_i];
        mbar_expect_tx(bars + st, sg.slice_bytes);
        bulk_g2s(stage_base + st * kStageFloats,
                 reinterpret_cast<const char*>(sg.base) + (size_t)slice_i * sg.slice_bytes, sg.slice_bytes, bars + st);
        ++issued;
        if (++slice_i == sg.n_slices) { slice_i = 0; if (++seg_i == nseg) seg_i = 0; }
    }
    __device__ __forceinline__ const float* wait_slice() const {
        mbar_wait(bars + (n & (kStages - 1)), (n / kStages) & 1u);
        return stage_base + (n & (kStages - 1)) * kStageFloats;
    }
    __device__ __forceinline__ void release_slice() {
        __syncthreads();
        if (threadIdx.x == 0) issue_one();
        ++n;
    }
};

// ------------------------------------------------------------------ streamed-weight GEMM, register tiled
// acc[t][r][0..3] += sum_k sA[row(r)][k] * W[k][t*CW + cg*4 + 0..3],  W streamed in [KS][NT*CW] slices.
// Thread (cg = tid % (CW/4), rg = tid / (CW/4)) owns rows rg*TR .. rg*TR+TR-1.
template <int CW, int NT, int TR>
__device__ __forceinline__ void gemm_acc(WStream& ws, const float* __restrict__ sA, int lda, int K,
                                         float (&acc)[NT][TR][4]) {
    constexpr int NC = CW * NT;
    constexpr int KS = (NC == 384) ? 8 : (NC == 192 ? 16 : (NC == 128 ? 16 : 32));
    constexpr int CG = CW / 4;
    const int cg = threadIdx.x % CG, rg = threadIdx.x / CG;
    const float* a_base = sA + (rg * TR) * lda;
    for (int k0 = 0; k0 < K; k0 += KS) {
        const float* __restrict__ w = ws.wait_slice() + cg * 4;
#pragma unroll
        for (int kk = 0; kk < KS; kk += 4) {
            float4 a[TR];
#pragma unroll
            for (int r = 0; r < TR; ++r) a[r] = *reinterpret_cast<const float4*>(a_base + r * lda + k0 + kk);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const float4 b = *reinterpret_cast<const float4*>(w + (kk + k) * NC + t * CW);
#pragma unroll
                    for (int r = 0; r < TR; ++r) {
                        const float av = (k == 0) ? a[r].x : (k == 1) ? a[r].y : (k == 2) ? a[r].z : a[r].w;
                        acc[t][r][0] = fmaf(av, b.x, acc[t][r][0]);
                        acc[t][r][1] = fmaf(av, b.y, acc[t][r][1]);
                        acc[t][r][2] = fmaf(av, b.z, acc[t][r][2]);
                        acc[t][r][3] = fmaf(av, b.w, acc[t][r][3]);
                    }
                }
            }
        }
        ws.release_slice();
    }
}

template <int NT, int TR>
__device__ __forceinline__ void zero_acc(float (&acc)[NT][TR][4]) {
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int r = 0; r < TR; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[t][r][e] = 0.f;
}

// f(t, row, col, acc4) for every 1x4 output strip this thread owns
template <int CW, int NT, int TR, class F>
__device__ __forceinline__ void tile_foreach(float (&acc)[NT][TR][4], F f) {
    constexpr int CG = CW / 4;
    const int cg = threadIdx.x % CG, rg = threadIdx.x / CG;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int r = 0; r < TR; ++r) f(t, rg * TR + r, cg * 4, acc[t][r]);
}

// ------------------------------------------------------------------ small per-sample attention products (K or inner = 64 / N)
// C[i][j] = sum_d A[s*N+i][d] * B[s*N+j][d], d < 64.  Interleaved row ownership keeps LDS.128 conflict free
// (row strides are = 4 mod 32 floats).
template <int TI, int TJ>
__device__ __forceinline__ void attn_nt(const float* __restrict__ sA, int lda, const float* __restrict__ sB, int ldb,
                                        float* __restrict__ sC, int NP, int N, int S_act, float scale) {
    const int IG = (N + TI - 1) / TI, JG = (N + TJ - 1) / TJ;
    const int total = S_act * IG * JG;
    for (int w = threadIdx.x; w < total; w += kThreads) {
        const int jg = w % JG;
        const int t2 = w / JG;
        const int ig = t2 % IG, s = t2 / IG;
        const float* pa[TI];
        const float* pb[TJ];
#pragma unroll
        for (int a = 0; a < TI; ++a) pa[a] = sA + (s * N + min(ig + a * IG, N - 1)) * lda;
#pragma unroll
        for (int b = 0; b < TJ; ++b) pb[b] = sB + (s * N + min(jg + b * JG, N - 1)) * ldb;
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
#pragma unroll
        for (int a = 0; a < TI; ++a)
#pragma unroll
            for (int b = 0; b < TJ; ++b) {
                const int i = ig + a * IG, j = jg + b * JG;
                if (i < N && j < N) sC[(s * N + i) * NP + j] = scale * acc[a][b];
            }
    }
}

// C[i][d4] = sum_j P[s*N+i][j] * B[s*N+j][d4]            (TRANS=false)
// C[j][d4] = sum_i P[s*N+i][j] * B[s*N+i][d4]            (TRANS=true)
// store(s, row_in_sample, d (multiple of 4), float4 value)
template <int TI, bool TRANS, class F>
__device__ __forceinline__ void attn_pv(const float* __restrict__ sPm, int NP, const float* __restrict__ sB, int ldb,
                                        int N, int S_act, F store) {
    const int IG = (N + TI - 1) / TI;
    const int total = S_act * IG * 16;
    for (int w = threadIdx.x; w < total; w += kThreads) {
        const int dg = w & 15;
        const int t2 = w >> 4;
        const int ig = t2 % IG, s = t2 / IG;
        int rows[TI];
#pragma unroll
        for (int a = 0; a < TI; ++a) rows[a] = min(ig + a * IG, N - 1);
        float4 acc[TI];
#pragma unroll
        for (int a = 0; a < TI; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* bp = sB + (s * N) * ldb + dg * 4;
        const float* pp = sPm + (s * N) * NP;
        for (int j = 0; j < N; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(bp + j * ldb);
#pragma unroll
            for (int a = 0; a < TI; ++a) {
                const float pv = TRANS ? pp[j * NP + rows[a]] : pp[rows[a] * NP + j];
                acc[a].x = fmaf(pv, b.x, acc[a].x);
                acc[a].y = fmaf(pv, b.y, acc[a].y);
                acc[a].z = fmaf(pv, b.z, acc[a].z);
                acc[a].w = fmaf(pv, b.w, acc[a].w);
            }
        }
#pragma unroll
        for (int a = 0; a < TI; ++a)
            if (ig + a * IG < N) store(s, ig + a * IG, dg * 4, acc[a]);
    }
}

// ------------------------------------------------------------------ configuration
template <int HP, int R>
struct Cfg {
    static constexpr int LDH = HP + 4;       // [R][HP] activation buffers
    static constexpr int LDQ = 192 + 4;      // q | k' | v' of one head
    static constexpr int LDO = 64 + 4;       // per-head attention output / its gradient
    static constexpr int LDF = 128 + 4;      // FF hidden chunk (aliases the qkv buffer)
    static constexpr int TR64 = R / 16;      // rows per thread for 64-wide column tiles
    static constexpr int TR128 = R / 8;      // rows per thread for 128-wide column tiles
    static constexpr int TRH = (HP == 128) ? TR128 : TR64;
    static constexpr int EPL = HP / 32;      // columns per lane in warp-per-row phases
    // shared memory carve-up (float offsets)
    static constexpr int oN = 0;
    static constexpr int oNh = oN + R * LDH;
    static constexpr int oQKV = oNh + R * LDH;
    static constexpr int oO = oQKV + R * LDQ;
    static constexpr int oP = oO + R * LDO;
    static constexpr int oDS = oP + R * kMaxBeads;
    static constexpr int oW = oDS + R * kMaxBeads;
    static constexpr int oX = oW + kStages * kStageFloats;
    static constexpr int oV = oX + R * 4;
    static constexpr int oDX = oV + R * 4;
    static constexpr int oTmp = oDX + R * 4;
    static constexpr int oBar = oTmp + R * 4;
    static constexpr int kFloats = oBar + 2 * kStages;
    static constexpr size_t kSmemBytes = (size_t)kFloats * sizeof(float);
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

// copy an [R][W] block (W multiple of 4) between shared (stride lds) and a dense global block
template <int R>
__device__ __forceinline__ void stash_store(float* __restrict__ dst, int W, const float* __restrict__ src, int lds) {
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < R * w4; idx += kThreads) {
        const int r = idx / w4, c = idx - r * w4;
        *reinterpret_cast<float4*>(dst + (size_t)r * W + c * 4) = *reinterpret_cast<const float4*>(src + r * lds + c * 4);
    }
}
template <int R>
__device__ __forceinline__ void stash_load(float* __restrict__ dst, int lds, const float* __restrict__ src, int W) {
    const int w4 = W >> 2;
    for (int idx = threadIdx.x; idx < R * w4; idx += kThreads) {
        const int r = idx / w4, c = idx - r * w4;
        *reinterpret_cast<float4*>(dst + r * lds + c * 4) = *reinterpret_cast<const float4*>(src + (size_t)r * W + c * 4);
    }
}

// ------------------------------------------------------------------ warp-per-row phases
// LayerNorm of sN rows -> sNh, stats (mean, rstd) -> global; optionally stashes the input rows.
template <int HP, int R>
__device__ __forceinline__ void ln_forward_rows(const float* sN, float* sNh, const float* __restrict__ gam,
                                                const float* __restrict__ bet, int H, float* st_rows, float* st_stats) {
    using C = Cfg<HP, R>;
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
template <int HP, int R>
__device__ __forceinline__ void gate_ln_forward_rows(float* sN, float* sNh, const float* __restrict__ ga,
                                                     const float* __restrict__ gb, int H, float* st_a, float* st_g,
                                                     float* st_out, const float* __restrict__ gam,
                                                     const float* __restrict__ bet, float* st_stats) {
    using C = Cfg<HP, R>;
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
template <int HP, int R>
__device__ __forceinline__ void gate_backward_rows(float* sN, float* sNh, int H, const float* __restrict__ gam,
                                                   const float* st_ln_in, const float* st_stats, const float* st_a,
                                                   const float* st_n, const float* st_g, const float* __restrict__ ga,
                                                   const float* __restrict__ gb) {
    using C = Cfg<HP, R>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < R; r += kWarps) {
        float d[C::EPL];
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            d[e] = (col < H) ? sN[r * C::LDH + col] : 0.f;
        }
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
        float a[C::EPL], n[C::EPL];
        float dg = 0.f;
#pragma unroll
        for (int e = 0; e < C::EPL; ++e) {
            const int col = lane * C::EPL + e;
            const bool ok = col < H;
            a[e] = ok ? st_a[(size_t)r * H + col] : 0.f;
            n[e] = ok ? st_n[(size_t)r * H + col] : 0.f;
            dg += d[e] * (a[e] - n[e]);
        }
        dg = warp_sum(dg);
        const float g = st_g[r];
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
template <int HP, int R>
__device__ __forceinline__ void ln_backward_rows(float* sN, const float* sNh, int H, const float* __restrict__ gam,
                                                 const float* st_ln_in, const float* st_stats) {
    using C = Cfg<HP, R>;
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

// softmax over j of sP rows (in place), 4 lanes per row; pad columns [N,NP) are zeroed.
template <int R>
__device__ __forceinline__ void softmax_rows(float* sP, int NP, int N, int rows_act) {
    const int q = threadIdx.x & 3;
    for (int base = 0; base < R; base += kThreads / 4) {
        const int row = base + (threadIdx.x >> 2);
        const bool ok = row < rows_act;
        float* p = sP + (ok ? row : 0) * NP;
        float m = -INFINITY;
        for (int j = q; j < N; j += 4) m = fmaxf(m, p[j]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        float s = 0.f;
        for (int j = q; j < N; j += 4) s += expf(p[j] - m);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (ok) {
            for (int j = q; j < N; j += 4) p[j] = expf(p[j] - m) / s;
            for (int j = N + q; j < NP; j += 4) p[j] = 0.f;
        }
    }
}

// ds = p * (dp - sum_j p dp) in place on sDS rows, 4 lanes per row
template <int R>
__device__ __forceinline__ void softmax_backward_rows(float* sDS, const float* sP, int NP, int N, int rows_act) {
    const int q = threadIdx.x & 3;
    for (int base = 0; base < R; base += kThreads / 4) {
        const int row = base + (threadIdx.x >> 2);
        const bool ok = row < rows_act;
        float* d = sDS + (ok ? row : 0) * NP;
        const float* p = sP + (ok ? row : 0) * NP;
        float s = 0.f;
        for (int j = q; j < N; j += 4) s += p[j] * d[j];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (ok) {
            for (int j = q; j < N; j += 4) d[j] = p[j] * (d[j] - s);
            for (int j = N + q; j < NP; j += 4) d[j] = 0.f;
        }
    }
}

// ------------------------------------------------------------------ forward pass (energy) for one group of samples
template <int HP, int R>
__device__ void forward_pass(const ModelDev& M, Ctx& c, float t_norm) {
    using C = Cfg<HP, R>;
    const int tid = threadIdx.x;
    const int N = M.N, NP = M.NP, H = M.H;

    // layer-0 node stream: W_n [onehot_i, t] + b_n   (graph_transformer.py:99-103), zero padding elsewhere
    for (int idx = tid; idx < R * HP; idx += kThreads) {
        const int r = idx / HP, d = idx - r * HP;
        float v = 0.f;
        if (r < c.rows_act && d < H) v = __ldg(M.emb + (r % N) * H + d) + t_norm * __ldg(M.embt + d);
        c.sN[r * C::LDH + d] = v;
    }
    __syncthreads();
    ln_forward_rows<HP, R>(c.sN, c.sNh, M.layer[0].ln1_g, M.layer[0].ln1_b, H, c.stash + M.off[ST_NIN],
                           c.stash + M.off[ST_STAT1]);
    __syncthreads();

    for (int l = 0; l < M.L; ++l) {
        const LayerDev& W = M.layer[l];
        float* st = c.stash + (size_t)l * M.layer_floats;

        float acc_a[1][C::TRH][4];
        zero_acc<1, C::TRH>(acc_a);
        for (int h = 0; h < kHeads; ++h) {
            {   // q | k | v of head h:  n_hat [R][H] x Wqkv_f[l][h] [H][192]
                float acc[3][C::TR64][4];
                zero_acc<3, C::TR64>(acc);
                gemm_acc<64, 3, C::TR64>(c.ws, c.sNh, C::LDH, H, acc);
                tile_foreach<64, 3, C::TR64>(acc, [&](int t, int row, int col, float (&v)[4]) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(W.bqkv + h * 192 + t * 64 + col));
                    float o[4] = {v[0] + b.x, v[1] + b.y, v[2] + b.z, v[3] + b.w};
                    if (t > 0) {   // k' = k + A x_j, v' = v + A x_j
                        const float x0 = c.sX[row * 4], x1 = c.sX[row * 4 + 1], x2 = c.sX[row * 4 + 2];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + (h * 64 + col + e) * 4));
                            o[e] += a4.x * x0 + a4.y * x1 + a4.z * x2;
                        }
                    }
                    *reinterpret_cast<float4*>(c.sQKV + row * C::LDQ + t * 64 + col) = make_float4(o[0], o[1], o[2], o[3]);
                });
            }
            __syncthreads();
            stash_store<R>(st + M.off[ST_QKV] + (size_t)h * R * 192, 192, c.sQKV, C::LDQ);
            // logits: s * q_i . k'_j
            if (N <= 16) attn_nt<2, 2>(c.sQKV, C::LDQ, c.sQKV + 64, C::LDQ, c.sP, NP, N, c.S_act, kAttnScale);
            else         attn_nt<4, 4>(c.sQKV, C::LDQ, c.sQKV + 64, C::LDQ, c.sP, NP, N, c.S_act, kAttnScale);
            __syncthreads();
            softmax_rows<R>(c.sP, NP, N, c.rows_act);
            __syncthreads();
            {   // stash p (dense [R][NP] block)
                float* dst = st + M.off[ST_P] + (size_t)h * R * NP;
                for (int idx = tid; idx < (R * NP) / 4; idx += kThreads)
                    reinterpret_cast<float4*>(dst)[idx] = reinterpret_cast<const float4*>(c.sP)[idx];
            }
            // o_i = sum_j p_ij v'_j - A x_i + c
            auto store_o = [&](int s, int i, int d, const float4& a) {
                const int row = s * N + i;
                const float x0 = c.sX[row * 4], x1 = c.sX[row * 4 + 1], x2 = c.sX[row * 4 + 2];
                const float4 cc = __ldg(reinterpret_cast<const float4*>(W.cvec + h * 64 + d));
                float o[4] = {a.x + cc.x, a.y + cc.y, a.z + cc.z, a.w + cc.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + (h * 64 + d + e) * 4));
                    o[e] -= a4.x * x0 + a4.y * x1 + a4.z * x2;
                }
                *reinterpret_cast<float4*>(c.sO + row * C::LDO + d) = make_float4(o[0], o[1], o[2], o[3]);
            };
            if (N <= 16) attn_pv<2, false>(c.sP, NP, c.sQKV + 128, C::LDQ, N, c.S_act, store_o);
            else         attn_pv<4, false>(c.sP, NP, c.sQKV + 128, C::LDQ, N, c.S_act, store_o);
            __syncthreads();
            // att += o_h [R][64] x Wo_f[l][h] [64][HP]
            gemm_acc<HP, 1, C::TRH>(c.ws, c.sO, C::LDO, 64, acc_a);
        }
        tile_foreach<HP, 1, C::TRH>(acc_a, [&](int, int row, int col, float (&v)[4]) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(W.bo + col));
            *reinterpret_cast<float4*>(c.sNh + row * C::LDH + col) = make_float4(v[0] + b.x, v[1] + b.y, v[2] + b.z, v[3] + b.w);
        });
        __syncthreads();
        // gated residual 1 + LayerNorm 2
        gate_ln_forward_rows<HP, R>(c.sN, c.sNh, W.g1a, W.g1b, H, st + M.off[ST_ATT], st + M.off[ST_G1],
                                    st + M.off[ST_M], W.ln2_g, W.ln2_b, st + M.off[ST_STAT2]);
        __syncthreads();
        // feed-forward, 128 hidden columns at a time
        float acc_f[1][C::TRH][4];
        zero_acc<1, C::TRH>(acc_f);
        float* sH1 = c.sQKV;
        for (int ch = 0; ch < M.nch; ++ch) {
            float acc1[1][C::TR128][4];
            zero_acc<1, C::TR128>(acc1);
            gemm_acc<128, 1, C::TR128>(c.ws, c.sNh, C::LDH, H, acc1);
            tile_foreach<128, 1, C::TR128>(acc1, [&](int, int row, int col, float (&v)[4]) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(W.b1 + ch * 128 + col));
                const float4 pre = make_float4(v[0] + b.x, v[1] + b.y, v[2] + b.z, v[3] + b.w);
                *reinterpret_cast<float4*>(st + M.off[ST_H1] + (size_t)row * (4 * H) + ch * 128 + col) = pre;
                *reinterpret_cast<float4*>(sH1 + row * C::LDF + col) =
                    make_float4(gelu_f(pre.x), gelu_f(pre.y), gelu_f(pre.z), gelu_f(pre.w));
            });
            __syncthreads();
            gemm_acc<HP, 1, C::TRH>(c.ws, sH1, C::LDF, 128, acc_f);
        }
        tile_foreach<HP, 1, C::TRH>(acc_f, [&](int, int row, int col, float (&v)[4]) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(W.b2 + col));
            *reinterpret_cast<float4*>(c.sNh + row * C::LDH + col) = make_float4(v[0] + b.x, v[1] + b.y, v[2] + b.z, v[3] + b.w);
        });
        __syncthreads();
        // gated residual 2 (+ next layer's LayerNorm 1; its input rows are the next layer's n_in stash)
        const bool last = (l + 1 == M.L);
        float* stn = st + M.layer_floats;
        gate_ln_forward_rows<HP, R>(c.sN, c.sNh, W.g2a, W.g2b, H, st + M.off[ST_FF], st + M.off[ST_G2],
                                    stn + M.off[ST_NIN],
                                    last ? nullptr : M.layer[l + 1].ln1_g, last ? nullptr : M.layer[l + 1].ln1_b,
                                    last ? nullptr : stn + M.off[ST_STAT1]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------ reverse pass: sDX[r][0..2] = d sum(E) / d x_r
template <int HP, int R>
__device__ void backward_pass(const ModelDev& M, Ctx& c) {
    using C = Cfg<HP, R>;
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
        gate_backward_rows<HP, R>(c.sN, c.sNh, H, nullptr, nullptr, nullptr, st + M.off[ST_FF], st + M.off[ST_M],
                                  st + M.off[ST_G2], W.g2a, W.g2b);
        __syncthreads();
        // feed-forward backward
        float acc_dm[1][C::TRH][4];
        zero_acc<1, C::TRH>(acc_dm);
        float* sH1 = c.sQKV;
        for (int ch = 0; ch < M.nch; ++ch) {
            float acc1[1][C::TR128][4];
            zero_acc<1, C::TR128>(acc1);
            gemm_acc<128, 1, C::TR128>(c.ws, c.sNh, C::LDH, H, acc1);          // d act = d ff x W2_b[ch]
            tile_foreach<128, 1, C::TR128>(acc1, [&](int, int row, int col, float (&v)[4]) {
                const float4 pre = *reinterpret_cast<const float4*>(st + M.off[ST_H1] + (size_t)row * (4 * H) + ch * 128 + col);
                *reinterpret_cast<float4*>(sH1 + row * C::LDF + col) =
                    make_float4(v[0] * gelu_grad_f(pre.x), v[1] * gelu_grad_f(pre.y), v[2] * gelu_grad_f(pre.z),
                                v[3] * gelu_grad_f(pre.w));
            });
            __syncthreads();
            gemm_acc<HP, 1, C::TRH>(c.ws, sH1, C::LDF, 128, acc_dm);            // d m_hat += d h1 x W1_b[ch]
        }
        tile_foreach<HP, 1, C::TRH>(acc_dm, [&](int, int row, int col, float (&v)[4]) {
            *reinterpret_cast<float4*>(c.sNh + row * C::LDH + col) = make_float4(v[0], v[1], v[2], v[3]);
        });
        __syncthreads();
        // LayerNorm 2 backward + gated residual 1 backward: d att -> sNh, d n_in (residual part) -> sN
        gate_backward_rows<HP, R>(c.sN, c.sNh, H, W.ln2_g, st + M.off[ST_M], st + M.off[ST_STAT2], st + M.off[ST_ATT],
                                  st + M.off[ST_NIN], st + M.off[ST_G1], W.g1a, W.g1b);
        __syncthreads();

        float acc_dn[1][C::TRH][4];
        zero_acc<1, C::TRH>(acc_dn);
        for (int h = 0; h < kHeads; ++h) {
            {   // d o_h = d att x Wo_b[l][h]   [R][H] x [H][64]
                float acc[1][C::TR64][4];
                zero_acc<1, C::TR64>(acc);
                gemm_acc<64, 1, C::TR64>(c.ws, c.sNh, C::LDH, H, acc);
                tile_foreach<64, 1, C::TR64>(acc, [&](int, int row, int col, float (&v)[4]) {
                    *reinterpret_cast<float4*>(c.sO + row * C::LDO + col) = make_float4(v[0], v[1], v[2], v[3]);
                });
            }
            stash_load<R>(c.sQKV, C::LDQ, st + M.off[ST_QKV] + (size_t)h * R * 192, 192);
            {
                const float* src = st + M.off[ST_P] + (size_t)h * R * NP;
                for (int idx = tid; idx < (R * NP) / 4; idx += kThreads)
                    reinterpret_cast<float4*>(c.sP)[idx] = reinterpret_cast<const float4*>(src)[idx];
            }
            __syncthreads();
            // dp_ij = do_i . v'_j
            if (N <= 16) attn_nt<2, 2>(c.sO, C::LDO, c.sQKV + 128, C::LDQ, c.sDS, NP, N, c.S_act, 1.0f);
            else         attn_nt<4, 4>(c.sO, C::LDO, c.sQKV + 128, C::LDQ, c.sDS, NP, N, c.S_act, 1.0f);
            __syncthreads();
            softmax_backward_rows<R>(c.sDS, c.sP, NP, N, c.rows_act);
            __syncthreads();
            if (l > 0) {   // dq_i = s sum_j ds_ij k'_j  -> v' columns (v' is dead after dp)
                auto st_dq = [&](int s, int i, int d, const float4& a) {
                    *reinterpret_cast<float4*>(c.sQKV + (s * N + i) * C::LDQ + 128 + d) =
                        make_float4(kAttnScale * a.x, kAttnScale * a.y, kAttnScale * a.z, kAttnScale * a.w);
                };
                if (N <= 16) attn_pv<2, false>(c.sDS, NP, c.sQKV + 64, C::LDQ, N, c.S_act, st_dq);
                else         attn_pv<4, false>(c.sDS, NP, c.sQKV + 64, C::LDQ, N, c.S_act, st_dq);
                __syncthreads();
            }
            {   // dk'_j = s sum_i ds_ij q_i  -> k' columns (k' is dead after dq)
                auto st_dk = [&](int s, int j, int d, const float4& a) {
                    *reinterpret_cast<float4*>(c.sQKV + (s * N + j) * C::LDQ + 64 + d) =
                        make_float4(kAttnScale * a.x, kAttnScale * a.y, kAttnScale * a.z, kAttnScale * a.w);
                };
                if (N <= 16) attn_pv<2, true>(c.sDS, NP, c.sQKV, C::LDQ, N, c.S_act, st_dk);
                else         attn_pv<4, true>(c.sDS, NP, c.sQKV, C::LDQ, N, c.S_act, st_dk);
            }
            __syncthreads();
            {   // dv'_j = sum_i p_ij do_i  -> q columns (q is dead after dk')
                auto st_dv = [&](int s, int j, int d, const float4& a) {
                    *reinterpret_cast<float4*>(c.sQKV + (s * N + j) * C::LDQ + d) = a;
                };
                if (N <= 16) attn_pv<2, true>(c.sP, NP, c.sO, C::LDO, N, c.S_act, st_dv);
                else         attn_pv<4, true>(c.sP, NP, c.sO, C::LDO, N, c.S_act, st_dv);
            }
            __syncthreads();
            // dx_r += A_h^T (dk'_r + dv'_r - do_r)
            for (int idx = tid; idx < c.rows_act * 3; idx += kThreads) {
                const int r = idx / 3, cc = idx - r * 3;
                const float* dk = c.sQKV + r * C::LDQ + 64;
                const float* dv = c.sQKV + r * C::LDQ;
                const float* dO = c.sO + r * C::LDO;
                const float* Ah = W.A + h * 64 * 4 + cc;
                float s = 0.f;
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
                c.sDX[r * 4 + cc] += s;
            }
            if (l > 0) {   // d n_hat += [dv' | dk' | dq] [R][192] x Wqkv_b[l][h] [192][HP]
                gemm_acc<HP, 1, C::TRH>(c.ws, c.sQKV, C::LDQ, 192, acc_dn);
            } else {
                __syncthreads();
            }
        }
        if (l > 0) {
            tile_foreach<HP, 1, C::TRH>(acc_dn, [&](int, int row, int col, float (&v)[4]) {
                *reinterpret_cast<float4*>(c.sNh + row * C::LDH + col) = make_float4(v[0], v[1], v[2], v[3]);
            });
            __syncthreads();
            ln_backward_rows<HP, R>(c.sN, c.sNh, H, W.ln1_g, st + M.off[ST_NIN], st + M.off[ST_STAT1]);
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <int HP, int R>
__global__ void __launch_bounds__(kThreads, 1)
dff_fused_kernel(const __grid_constant__ ModelDev M, const __grid_constant__ StepArgs A) {
    using C = Cfg<HP, R>;
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
    c.ws.bars = reinterpret_cast<uint64_t*>(smem + C::oBar);
    c.ws.n = 0;
    c.ws.segs = M.segs;
    c.ws.nseg = A.need_backward ? M.nseg_all : M.nseg_fwd;
    c.ws.seg_i = 0; c.ws.slice_i = 0; c.ws.issued = 0;
    c.ws.total = (uint32_t)my_groups * (uint32_t)A.n_steps * (A.need_backward ? M.nslice_all : M.nslice_fwd);
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) mbar_init(c.ws.bars + i, 1);
        fence_barrier_init();
    }
    for (int idx = tid; idx < C::oW; idx += kThreads) smem[idx] = 0.f;       // activations / attention buffers
    for (int idx = C::oX + tid; idx < C::oBar; idx += kThreads) smem[idx] = 0.f;
    __syncthreads();
    if (tid == 0) for (int i = 0; i < kStages; ++i) c.ws.issue_one();

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

            forward_pass<HP, R>(M, c, t_norm);
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
            if (A.need_backward) backward_pass<HP, R>(M, c);
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
