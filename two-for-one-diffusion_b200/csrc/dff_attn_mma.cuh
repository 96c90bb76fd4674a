// dff_attn_mma.cuh -- the attention contractions of the fused kernel on the tensor cores, warp level (sm_100a).
//
// Reference math (models/graph_transformer.py:241-256 and its reverse mode, collapsed as in DESIGN.md section 2):
//   forward : S = s Q K'^T -> P = softmax_j(S) -> O = P V' - A x_i + c
//   reverse : dP = dO V'^T -> dS = P o (dP - rowsum(P o dP)) -> dQ = s dS K',  dK' = s dS^T Q,  dV' = P^T dO,
//             dx_j += sum_i (p_ij w_i + s ds_ij u_i) - w_j      with  u_i = A_h^T q_i,  w_i = A_h^T do_i   
//
// Attention is block diagonal: a 64-row tile holds S_act samples of N beads and a query only sees the keys of its own
// sample.  Every contraction is therefore cut into 16 x 16 output blocks per (sample, 16-row block, pair of 8-column tiles) and
// spread over the 16 compute warps; a block is two chains of `mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32` (SASS HMMA.1688.F32.TF32)
// with the 3xTF32 split done in registers (lo*hi + hi*lo + hi*hi, fp32 accumulate), operands read straight from the
// row-major fp32 shared-memory buffers the TMEM epilogues fill -- no layout change, no extra shared memory, no TMEM.
// Why warp-level MMA and not tcgen05 here (measured, profiles/r02/hmma_rate.txt): mma.sync TF32 sustains 512 MAC/cycle/SM,
// the M = 64 tcgen05 block GEMM of this kernel 455 (72 cycles per 64x64x8), and tcgen05 would multiply the full 64 x 64
// tile (6 x the work for chignolin, 3 x for trp-cage) from canonical hi/lo operands that need 8 B per element of shared
// memory: six 33 KB operand buffers do not fit next to the projection operands.
//
// Fragment conventions of m16n8k8 (g = lane >> 2, t = lane & 3):
//   A (16 x 8)  a0 (g, t)   a1 (g + 8, t)   a2 (g, t + 4)   a3 (g + 8, t + 4)
//   B ( 8 x 8)  b0 (k = t, n = g)            b1 (k = t + 4, n = g)
//   C (16 x 8)  c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
#pragma once
#include "dff_common.cuh"

namespace dff {
namespace v2 {

__device__ __forceinline__ void hmma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// round-to-nearest TF32 high part and the exact remainder
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// C[16 x 8] += A[16 x 8 ksteps] * B[8 ksteps x 8]  at fp32 grade.  la(kk, a[4]) / lb(kk, b[2]) fill the fp32 fragments of k-step kk.
// The three split products accumulate in three independent register accumulators (three dependent HMMA chains of length
// `ksteps` instead of one of length 3 * ksteps; the two small products are summed first), and the fragments of two k-steps
// are loaded before their six HMMAs are issued.
template <class LA, class LB>
__device__ __forceinline__ void tile_mma(float (&c)[4], int ksteps, LA la, LB lb) {
    float clh[4] = {0.f, 0.f, 0.f, 0.f}, chl[4] = {0.f, 0.f, 0.f, 0.f};
    int kk = 0;
    for (; kk + 1 < ksteps; kk += 2) {
        float a0[4], b0[2], a1[4], b1[2];
        la(kk, a0); lb(kk, b0); la(kk + 1, a1); lb(kk + 1, b1);
        uint32_t ah[4], al[4], bh[2], bl[2], ah1[4], al1[4], bh1[2], bl1[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { split_tf32(a0[i], ah[i], al[i]); split_tf32(a1[i], ah1[i], al1[i]); }
#pragma unroll
        for (int i = 0; i < 2; ++i) { split_tf32(b0[i], bh[i], bl[i]); split_tf32(b1[i], bh1[i], bl1[i]); }
        hmma_tf32(clh, al, bh); hmma_tf32(chl, ah, bl); hmma_tf32(c, ah, bh);
        hmma_tf32(clh, al1, bh1); hmma_tf32(chl, ah1, bl1); hmma_tf32(c, ah1, bh1);
    }
    if (kk < ksteps) {
        float a0[4], b0[2];
        la(kk, a0); lb(kk, b0);
        uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(a0[i], ah[i], al[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) split_tf32(b0[i], bh[i], bl[i]);
        hmma_tf32(clh, al, bh); hmma_tf32(chl, ah, bl); hmma_tf32(c, ah, bh);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] += clh[i] + chl[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Work decomposition.  An ITEM is (sample, 16-row block, pair of 8-column tiles) of one product; the items of a phase are
// dealt round-robin to the 16 compute warps.  Phases of one head (CTA-wide barriers between them):
//   forward : [logits items] | [softmax rows] | [P V' items -> slot]
//   reverse : [dp items + u / w items] | [ds rows] | [dx rows, dq items -> slot] | [dk' items -> slot] | [dv' items -> slot]
struct AttnGeo {
    int N, NP, S_act, MB, NK, NKP;      // beads, beads padded to 4, samples, row blocks / key tiles / key-tile pairs per sample
    __device__ __forceinline__ AttnGeo(int N_, int NP_, int S_) : N(N_), NP(NP_), S_act(S_), MB((N_ + 15) >> 4), NK((N_ + 7) >> 3) {
        NKP = (NK + 1) >> 1;
    }
};

__device__ __forceinline__ void frag_split4(const float (&a)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(a[i], hi[i], lo[i]);
}
__device__ __forceinline__ void frag_split2(const float (&b)[2], uint32_t (&hi)[2], uint32_t (&lo)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(b[i], hi[i], lo[i]);
}

// acc[nt] (+)= A[16 x 64] * B_nt[64 x 8]  for nt < NK:  A = 16 rows (clamped to the sample) of a row-major buffer (columns
// [acol, acol + 64)), B_nt(k, n) = brows[r0 + nt * 8 + n][bcol + k].  Used for the logits (q . k') and for dp (do . v').
template <int NKMAX>
__device__ __forceinline__ void rows_dot_rows(float (&acc)[NKMAX][4], const float* abuf, int lda, int acol, const float* bbuf, int ldb, int bcol,
                                              int r0, int m0, int N, int NK, int n0 = 0) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* pa0 = abuf + (r0 + min(m0 + g, N - 1)) * lda + acol + t;
    const float* pa1 = abuf + (r0 + min(m0 + g + 8, N - 1)) * lda + acol + t;
    const float* pb[NKMAX];
#pragma unroll
    for (int nt = 0; nt < NKMAX; ++nt) pb[nt] = bbuf + (r0 + min(n0 + nt * 8 + g, N - 1)) * ldb + bcol + t;
#pragma unroll
    for (int nt = 0; nt < NKMAX; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll 4
    for (int kk = 0; kk < 8; ++kk) {
        const float a[4] = {pa0[kk * 8], pa1[kk * 8], pa0[kk * 8 + 4], pa1[kk * 8 + 4]};
        uint32_t ah[4], al[4];
        frag_split4(a, ah, al);
        // two key tiles at a time: their six HMMAs alternate between two accumulators
#pragma unroll
        for (int nt = 0; nt < NKMAX; nt += 2)
            if (nt < NK) {
                const bool two = (nt + 1 < NKMAX) && (nt + 1 < NK);
                uint32_t bh0[2], bl0[2], bh1[2], bl1[2];
                const float b0[2] = {pb[nt][kk * 8], pb[nt][kk * 8 + 4]};
                frag_split2(b0, bh0, bl0);
                if (nt + 1 < NKMAX) {
                    const float b1[2] = {pb[nt + 1 < NKMAX ? nt + 1 : nt][kk * 8], pb[nt + 1 < NKMAX ? nt + 1 : nt][kk * 8 + 4]};
                    frag_split2(b1, bh1, bl1);
                }
                hmma_tf32(acc[nt], al, bh0);
                if (nt + 1 < NKMAX) { if (two) hmma_tf32(acc[nt + 1 < NKMAX ? nt + 1 : nt], al, bh1); }
                hmma_tf32(acc[nt], ah, bl0);
                if (nt + 1 < NKMAX) { if (two) hmma_tf32(acc[nt + 1 < NKMAX ? nt + 1 : nt], ah, bl1); }
                hmma_tf32(acc[nt], ah, bh0);
                if (nt + 1 < NKMAX) { if (two) hmma_tf32(acc[nt + 1 < NKMAX ? nt + 1 : nt], ah, bh1); }
            }
    }
}

// One 16 x 16 output block (two 8-column tiles dt, dt + 1) of  W[16 x keys] * B[keys x 64]:
//   TRANSPOSED = false:  A(m, k) = wbuf[r0 + m0 + m][k]        (p or ds rows, k = key)                -> P V', dS K'
//   TRANSPOSED = true :  A(m, k) = wbuf[r0 + k][m0 + m]        (m = key, k = query)                   -> dS^T Q, P^T dO
//   B(k, n) = bbuf[r0 + k][bcol + dt * 8 + n]
template <bool TRANSPOSED>
__device__ __forceinline__ void keys_dot_cols(float (&acc)[2][4], const float* wbuf, int NP, const float* bbuf, int ldb, int bcol,
                                              int r0, int m0, int N, int NK, int dt) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ma = min(m0 + g, N - 1), mb = min(m0 + g + 8, N - 1);
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const float* wb = wbuf + r0 * NP;
    const float* bb = bbuf + r0 * ldb + bcol + dt * 8 + g;
    for (int kk = 0; kk < NK; ++kk) {
        const int k0 = kk * 8 + t, k1 = k0 + 4;
        const bool v0 = k0 < N, v1 = k1 < N;
        const int c0 = min(k0, N - 1), c1 = min(k1, N - 1);
        float a[4];
        if (TRANSPOSED) {
            a[0] = v0 ? wb[c0 * NP + ma] : 0.f; a[1] = v0 ? wb[c0 * NP + mb] : 0.f;
            a[2] = v1 ? wb[c1 * NP + ma] : 0.f; a[3] = v1 ? wb[c1 * NP + mb] : 0.f;
        } else {
            a[0] = v0 ? wb[ma * NP + c0] : 0.f; a[1] = v0 ? wb[mb * NP + c0] : 0.f;
            a[2] = v1 ? wb[ma * NP + c1] : 0.f; a[3] = v1 ? wb[mb * NP + c1] : 0.f;
        }
        uint32_t ah[4], al[4];
        frag_split4(a, ah, al);
        uint32_t bh[2][2], bl[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float b[2] = {bb[c0 * ldb + j * 8], bb[c1 * ldb + j * 8]};
            frag_split2(b, bh[j], bl[j]);
        }
        hmma_tf32(acc[0], al, bh[0]); hmma_tf32(acc[1], al, bh[1]);
        hmma_tf32(acc[0], ah, bl[0]); hmma_tf32(acc[1], ah, bl[1]);
        hmma_tf32(acc[0], ah, bh[0]); hmma_tf32(acc[1], ah, bh[1]);
    }
}

// 16 x 8 tile  rows[16 x 64] * edge[64 x 4]:  (u, alpha) = (A_h^T q, a_h . q)  or  (w, beta) = (A_h^T do, a_h . do); columns 0..ncol-1
__device__ __forceinline__ void rows_dot_edge(float (&acc)[4], const float* abuf, int lda, const float* __restrict__ Aedge, int hc, int r0, int m0, int N,
                                              int ncol) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* pa0 = abuf + (r0 + min(m0 + g, N - 1)) * lda + t;
    const float* pa1 = abuf + (r0 + min(m0 + g + 8, N - 1)) * lda + t;
    const bool on = g < ncol;
    const float* pe = Aedge + (hc * 64 + t) * 4 + (on ? g : 0);
    float e[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) e[i] = on ? __ldg(pe + i * 16) : 0.f;       // all 16 loads in flight at once
    acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
    float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const float a[4] = {pa0[kk * 8], pa1[kk * 8], pa0[kk * 8 + 4], pa1[kk * 8 + 4]};
        const float b[2] = {e[2 * kk], e[2 * kk + 1]};
        uint32_t ah[4], al[4], bh[2], bl[2];
        frag_split4(a, ah, al);
        frag_split2(b, bh, bl);
        hmma_tf32(c1, al, bh); hmma_tf32(c2, ah, bl); hmma_tf32(acc, ah, bh);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += c1[i] + c2[i];
}

// ---------------------------------------------------------------------------------------------------------------
// Forward phase 1: scaled logits  sP[u][j] = s q_u . k'_j   (items: sample x row block x key-tile pair); with the distance
// channel also (u, alpha)_u = (A_h^T q_u, a_h . q_u) -> sQKV[u][192..195]  (one more item per row block)
template <class C>
__device__ __forceinline__ void attn_logits_items(float* sQKV, float* sP, const float* __restrict__ Aedge, int hc, bool dist, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ipb = G.NKP + (dist ? 1 : 0);                  // items per row block
    const int per = G.MB * ipb, items = G.S_act * per;
    for (int it = warp; it < items; it += kCW) {
        const int s = it / per, rem = it - s * per, mb = rem / ipb, np = rem - mb * ipb;
        const int r0 = s * G.N, m0 = mb * 16, n0 = np * 16;
        const int ra = m0 + g, rb = ra + 8;
        if (np == G.NKP) {
            float e4[4];
            rows_dot_edge(e4, sQKV, C::LDQ, Aedge, hc, r0, m0, G.N, 4);
            if (t < 2) {
                if (ra < G.N) *reinterpret_cast<float2*>(sQKV + (r0 + ra) * C::LDQ + 192 + 2 * t) = make_float2(e4[0], e4[1]);
                if (rb < G.N) *reinterpret_cast<float2*>(sQKV + (r0 + rb) * C::LDQ + 192 + 2 * t) = make_float2(e4[2], e4[3]);
            }
            continue;
        }
        float c[2][4];
        rows_dot_rows<2>(c, sQKV, C::LDQ, 0, sQKV, C::LDQ, 64, r0, m0, G.N, min(2, G.NK - 2 * np), n0);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = n0 + j * 8 + 2 * t;
            if (col < G.NP) {
                if (ra < G.N) *reinterpret_cast<float2*>(sP + (r0 + ra) * G.NP + col) = make_float2(kAttnScale * c[j][0], kAttnScale * c[j][1]);
                if (rb < G.N) *reinterpret_cast<float2*>(sP + (r0 + rb) * G.NP + col) = make_float2(kAttnScale * c[j][2], kAttnScale * c[j][3]);
            }
        }
    }
}
// squared distance of rows r and r0 + key (sX holds the centred coordinates, 4 floats per row)
__device__ __forceinline__ float dist2(const float* sX, int r, int rk) {
    const float dx = sX[r * 4] - sX[rk * 4], dy = sX[r * 4 + 1] - sX[rk * 4 + 1], dz = sX[r * 4 + 2] - sX[rk * 4 + 2];
    return dx * dx + dy * dy + dz * dz;
}
// Forward phase 2: softmax over the keys of the row's sample, in place; p -> stash.  One warp per row, lane = key (and key + 32).
// Distance channel: logit_uj += s alpha_u |x_u - x_j|^2  and  z_u = sum_j p_uj |x_u - x_j|^2 -> sZ[u * ldz]
template <class C>
__device__ __forceinline__ void attn_softmax_rows(float* sP, const float* sQKV, const float* sX, float* sZ, int ldz, bool dist, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = G.S_act * G.N;
    for (int r = warp; r < rows; r += kCW) {
        float* row = sP + r * G.NP;
        const int r0 = (r / G.N) * G.N;
        const bool a0 = lane < G.N, a1 = lane + 32 < G.N;
        float l0 = a0 ? row[lane] : -INFINITY, l1 = a1 ? row[lane + 32] : -INFINITY;
        float d0 = 0.f, d1 = 0.f;
        if (dist) {
            const float sa = kAttnScale * sQKV[r * C::LDQ + 195];
            if (a0) { d0 = dist2(sX, r, r0 + lane); l0 = fmaf(sa, d0, l0); }
            if (a1) { d1 = dist2(sX, r, r0 + lane + 32); l1 = fmaf(sa, d1, l1); }
        }
        float m = fmaxf(l0, l1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float e0 = a0 ? fast_exp(l0 - m) : 0.f, e1 = a1 ? fast_exp(l1 - m) : 0.f;
        const float inv = 1.0f / warp_sum(e0 + e1);
        const float p0 = e0 * inv, p1 = e1 * inv;
        if (lane < G.NP) row[lane] = p0;              // (the caller copies sP to the stash)
        if (lane + 32 < G.NP) row[lane + 32] = p1;
        if (dist) {
            const float z = warp_sum(p0 * d0 + p1 * d1);
            if (lane == 0) sZ[r * ldz] = z;
        }
    }
}
// Forward phase 3 / reverse dq phase: 16 x 16 blocks of  W[rows x keys] * B[keys x 64]  -> f_store(row, col, v0, v1)
// (items: sample x row block x column-tile pair; 4 pairs cover the 64 columns)
template <class C, class F>
__device__ __forceinline__ void attn_weighted_items(const float* sW, const float* sB, int ldb, int bcol, const AttnGeo& G, F f_store) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = G.MB * 4, items = G.S_act * per;
    for (int it = warp; it < items; it += kCW) {
        const int s = it / per, rem = it - s * per, mb = rem >> 2, dp = rem & 3;
        const int r0 = s * G.N, m0 = mb * 16;
        float c[2][4];
        keys_dot_cols<false>(c, sW, G.NP, sB, ldb, bcol, r0, m0, G.N, G.NK, 2 * dp);
        const int ra = m0 + g, rb = ra + 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = (2 * dp + j) * 8 + 2 * t;
            if (ra < G.N) f_store(r0 + ra, col, c[j][0], c[j][1]);
            if (rb < G.N) f_store(r0 + rb, col, c[j][2], c[j][3]);
        }
    }
}

// Reverse phase 1: dp_uj = do_u . v'_j -> sDS (raw);  (u, alpha)_u = (A_h^T q_u, a_h . q_u) -> sQKV[u][192..195];
// (w, beta)_u = (A_h^T do_u, a_h . do_u) -> sO[u][64..67]
template <class C>
__device__ __forceinline__ void attn_dp_uw_items(float* sQKV, float* sO, float* sDS, const float* __restrict__ Aedge, int hc, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = G.MB * (G.NKP + 2), items = G.S_act * per;
    // the u / w items go first in the deal (to the low warps), the dp items follow: with <= 16 items every warp has one
    for (int it = warp; it < items; it += kCW) {
        const int s = it / per, rem = it - s * per, mb = rem / (G.NKP + 2), np = rem - mb * (G.NKP + 2);
        const int r0 = s * G.N, m0 = mb * 16;
        const int ra = m0 + g, rb = ra + 8;
        if (np < G.NKP) {
            const int n0 = np * 16;
            float c[2][4];
            rows_dot_rows<2>(c, sO, C::LDO, 0, sQKV, C::LDQ, 128, r0, m0, G.N, min(2, G.NK - 2 * np), n0);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = n0 + j * 8 + 2 * t;
                if (col < G.NP) {
                    if (ra < G.N) *reinterpret_cast<float2*>(sDS + (r0 + ra) * G.NP + col) = make_float2(c[j][0], c[j][1]);
                    if (rb < G.N) *reinterpret_cast<float2*>(sDS + (r0 + rb) * G.NP + col) = make_float2(c[j][2], c[j][3]);
                }
            }
        } else {
            const bool is_u = np == G.NKP;
            float* dst = is_u ? sQKV : sO;
            const int ld = is_u ? C::LDQ : C::LDO, off = is_u ? 192 : 64;
            float e4[4];
            rows_dot_edge(e4, dst, ld, Aedge, hc, r0, m0, G.N, 4);
            if (t < 2) {        // columns 2t, 2t + 1 of the 8-wide tile: (u0, u1) and (u2, alpha)  /  (w0, w1) and (w2, beta)
                if (ra < G.N) *reinterpret_cast<float2*>(dst + (r0 + ra) * ld + off + 2 * t) = make_float2(e4[0], e4[1]);
                if (rb < G.N) *reinterpret_cast<float2*>(dst + (r0 + rb) * ld + off + 2 * t) = make_float2(e4[2], e4[3]);
            }
        }
    }
}
// Reverse phase 2: ds_uj = p_uj (dp_uj - sum_j p_uj dp_uj), in place in sDS.  One warp per row.
// Distance channel: dp_uj += beta_u |x_u - x_j|^2 first;  d alpha_u = s sum_j ds_uj |x_u - x_j|^2 -> sDA[u * 4]
template <class C>
__device__ __forceinline__ void attn_ds_rows(const float* sP, float* sDS, const float* sO, const float* sX, float* sDA, bool dist, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = G.S_act * G.N;
    for (int r = warp; r < rows; r += kCW) {
        const float* pr = sP + r * G.NP;
        float* dr = sDS + r * G.NP;
        const int r0 = (r / G.N) * G.N;
        const bool a0 = lane < G.N, a1 = lane + 32 < G.N;
        const float p0 = a0 ? pr[lane] : 0.f, p1 = a1 ? pr[lane + 32] : 0.f;
        float d0 = a0 ? dr[lane] : 0.f, d1 = a1 ? dr[lane + 32] : 0.f;
        float q0 = 0.f, q1 = 0.f;
        if (dist) {
            const float beta = sO[r * C::LDO + 67];
            if (a0) { q0 = dist2(sX, r, r0 + lane); d0 = fmaf(beta, q0, d0); }
            if (a1) { q1 = dist2(sX, r, r0 + lane + 32); d1 = fmaf(beta, q1, d1); }
        }
        const float tsum = warp_sum(p0 * d0 + p1 * d1);
        const float s0 = p0 * (d0 - tsum), s1 = p1 * (d1 - tsum);
        if (lane < G.NP) dr[lane] = s0;
        if (lane + 32 < G.NP) dr[lane + 32] = s1;
        if (dist) {
            const float da = kAttnScale * warp_sum(s0 * q0 + s1 * q1);
            if (lane == 0) sDA[r * 4] = da;
        }
    }
}
// Reverse phase 3: dx_j += sum_i (p_ij w_i + s ds_ij u_i) - w_j.  Four lanes per (key row, component), fixed summation order: deterministic.
// Distance channel: with E_ij = d / d|x_i - x_j|^2 = s ds_ij alpha_i + p_ij beta_i,   dx_j += 2 sum_i (E_ij + E_ji) (x_j - x_i)
template <class C>
__device__ __forceinline__ void attn_dx_rows(const float* sQKV, const float* sO, const float* sP, const float* sDS, const float* sX, float* sDX, bool dist,
                                             const AttnGeo& G) {
    const int rows = G.S_act * G.N;
    const int q = threadIdx.x & 3;
    for (int base = 0; base < rows * 3; base += kCT / 4) {         // warp-uniform trip count (the quad shuffles need whole warps)
        const int idx0 = base + (threadIdx.x >> 2);
        const bool valid = idx0 < rows * 3;
        const int idx = valid ? idx0 : 0;
        const int j = idx / 3, cc = idx - j * 3;
        const int r0 = (j / G.N) * G.N, jj = j - r0;
        float acc = 0.f, acd = 0.f, ace = 0.f;
        const float aj = dist ? kAttnScale * sQKV[j * C::LDQ + 195] : 0.f, bj = dist ? sO[j * C::LDO + 67] : 0.f, xj = sX[j * 4 + cc];
        for (int i = q; i < G.N; i += 4) {
            const float pij = sP[(r0 + i) * G.NP + jj], dij = sDS[(r0 + i) * G.NP + jj];
            acc = fmaf(pij, sO[(r0 + i) * C::LDO + 64 + cc], acc);
            acd = fmaf(dij, sQKV[(r0 + i) * C::LDQ + 192 + cc], acd);
            if (dist) {
                const float e = fmaf(kAttnScale * sQKV[(r0 + i) * C::LDQ + 195], dij, sO[(r0 + i) * C::LDO + 67] * pij)       // E_ij
                              + fmaf(aj, sDS[j * G.NP + i], bj * sP[j * G.NP + i]);                                          // E_ji
                ace = fmaf(e, xj - sX[(r0 + i) * 4 + cc], ace);
            }
        }
        float v = acc + kAttnScale * acd + 2.0f * ace;
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (valid && q == 0) sDX[j * 4 + cc] += v - sO[j * C::LDO + 64 + cc];
    }
}
// Reverse phases 4-5: 16 x 16 blocks of  dk' = s dS^T Q  (wbuf = sDS, bbuf = q columns)  or  dv' = P^T dO  (wbuf = sP, bbuf = sO)
// into registers (two items per warp and round), then into the rotating slot once the tensor core has released it -- the MMAs
// of the previous slot job overlap the arithmetic.
template <class C, class CTX>
__device__ __forceinline__ void attn_keys_to_slot_items(CTX& c, const float* wbuf, const float* bbuf, int ldb, float scale, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = G.MB * 4, items = G.S_act * per;
    for (int base = warp; base < items; base += 2 * kCW) {
        float a0[2][4], a1[2][4];
        const int it1 = base + kCW;
        {
            const int s = base / per, rem = base - s * per;
            keys_dot_cols<true>(a0, wbuf, G.NP, bbuf, ldb, 0, s * G.N, (rem >> 2) * 16, G.N, G.NK, 2 * (rem & 3));
        }
        if (it1 < items) {
            const int s = it1 / per, rem = it1 - s * per;
            keys_dot_cols<true>(a1, wbuf, G.NP, bbuf, ldb, 0, s * G.N, (rem >> 2) * 16, G.N, G.NK, 2 * (rem & 3));
        }
        c.slot_acquire();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int it = base + h * kCW;
            if (it < items) {
                const int s = it / per, rem = it - s * per, r0 = s * G.N, m0 = (rem >> 2) * 16, dp = rem & 3;
                const int ra = m0 + g, rb = ra + 8;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int col = (2 * dp + j) * 8 + 2 * t;
                    const float (&a)[2][4] = h ? a1 : a0;
                    if (ra < G.N) can_store2<C::kCS>(c.slot_hi, c.slot_lo, r0 + ra, col, scale * a[j][0], scale * a[j][1]);
                    if (rb < G.N) can_store2<C::kCS>(c.slot_hi, c.slot_lo, r0 + rb, col, scale * a[j][2], scale * a[j][3]);
                }
            }
        }
    }
}

}  // namespace v2
}  // namespace dff
