// dff_attn_mma.cuh -- the attention contractions of the fused kernel on the tensor cores, warp level (sm_100a).
//
// Reference math (models/graph_transformer.py:241-256 and its reverse mode, collapsed as in DESIGN.md section 2):
//   forward : S = s Q K'^T -> P = softmax_j(S) -> O = P V' - A x_i + c
//   reverse : dP = dO V'^T -> dS = P o (dP - rowsum(P o dP)) -> dQ = s dS K',  dK' = s dS^T Q,  dV' = P^T dO,
//             dx_j += sum_i (p_ij w_i + s ds_ij u_i) - w_j      with  u_i = A_h^T q_i,  w_i = A_h^T do_i   
//
// Attention is block diagonal: a 64-row tile holds S_act samples of N beads and a query only sees the keys of its own
// sample.  Every contraction is therefore cut into 16 x 8 output tiles per (sample, 16-row block, 8-column block) and
// spread over the 16 compute warps; a tile is a chain of `mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32` (SASS HMMA.1688.F32.TF32)
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
template <class LA, class LB>
__device__ __forceinline__ void tile_mma(float (&c)[4], int ksteps, LA la, LB lb) {
#pragma unroll 2
    for (int kk = 0; kk < ksteps; ++kk) {
        float a[4], b[2];
        la(kk, a);
        lb(kk, b);
        uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
        for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
        hmma_tf32(c, al, bh);
        hmma_tf32(c, ah, bl);
        hmma_tf32(c, ah, bh);
    }
}

// Geometry of the block-diagonal tiling of one pass (S_act samples of N beads).
struct AttnGeo {
    int N, NP, S_act, MB, NK;       // beads, beads padded to 4, samples, 16-row blocks per sample, 8-key blocks per sample
    __device__ __forceinline__ AttnGeo(int N_, int NP_, int S_) : N(N_), NP(NP_), S_act(S_), MB((N_ + 15) >> 4), NK((N_ + 7) >> 3) {}
};

// fragment loaders -----------------------------------------------------------------------------------------------
// A operand = 16 rows [rbase + m0 .. + 15] (clamped to the sample) of a row-major buffer, k = column
struct LoadA_rows {
    const float* p0; const float* p1;          // rows g and g + 8 (already clamped), at column t
    __device__ __forceinline__ LoadA_rows(const float* buf, int ld, int r0, int m0, int N, int coloff) {
        const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
        p0 = buf + (r0 + min(m0 + g, N - 1)) * ld + coloff + t;
        p1 = buf + (r0 + min(m0 + g + 8, N - 1)) * ld + coloff + t;
    }
    __device__ __forceinline__ void operator()(int kk, float (&a)[4]) const {
        a[0] = p0[kk * 8]; a[1] = p1[kk * 8]; a[2] = p0[kk * 8 + 4]; a[3] = p1[kk * 8 + 4];
    }
};
// A operand = 16 rows of a [row][key] weight buffer (p or ds), k = key index masked to the sample's N keys
struct LoadA_keys {
    const float* p0; const float* p1; int N, t;
    __device__ __forceinline__ LoadA_keys(const float* buf, int NP, int r0, int m0, int N_) : N(N_) {
        const int g = (threadIdx.x & 31) >> 2;
        t = threadIdx.x & 3;
        p0 = buf + (r0 + min(m0 + g, N - 1)) * NP;
        p1 = buf + (r0 + min(m0 + g + 8, N - 1)) * NP;
    }
    __device__ __forceinline__ void operator()(int kk, float (&a)[4]) const {
        const int k0 = kk * 8 + t, k1 = k0 + 4;
        const bool v0 = k0 < N, v1 = k1 < N;
        a[0] = v0 ? p0[k0] : 0.f; a[1] = v0 ? p1[k0] : 0.f; a[2] = v1 ? p0[k1] : 0.f; a[3] = v1 ? p1[k1] : 0.f;
    }
};
// A operand = TRANSPOSE of a [row i][key j] weight buffer: m = key j0 + .., k = query i (masked to N)
struct LoadA_keysT {
    const float* base; int NP, N, j0, j1, t;
    __device__ __forceinline__ LoadA_keysT(const float* buf, int NP_, int r0, int m0, int N_) : NP(NP_), N(N_) {
        const int g = (threadIdx.x & 31) >> 2;
        t = threadIdx.x & 3;
        base = buf + r0 * NP;
        j0 = min(m0 + g, N - 1); j1 = min(m0 + g + 8, N - 1);
    }
    __device__ __forceinline__ void operator()(int kk, float (&a)[4]) const {
        const int i0 = kk * 8 + t, i1 = i0 + 4;
        const bool v0 = i0 < N, v1 = i1 < N;
        const float* r0p = base + min(i0, N - 1) * NP;
        const float* r1p = base + min(i1, N - 1) * NP;
        a[0] = v0 ? r0p[j0] : 0.f; a[1] = v0 ? r0p[j1] : 0.f; a[2] = v1 ? r1p[j0] : 0.f; a[3] = v1 ? r1p[j1] : 0.f;
    }
};
// B operand, n = row of a row-major buffer (a key), k = column:  B(k, n) = buf[r0 + n0 + n][coloff + k]
struct LoadB_rowsN {
    const float* p;
    __device__ __forceinline__ LoadB_rowsN(const float* buf, int ld, int r0, int n0, int N, int coloff) {
        const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
        p = buf + (r0 + min(n0 + g, N - 1)) * ld + coloff + t;
    }
    __device__ __forceinline__ void operator()(int kk, float (&b)[2]) const { b[0] = p[kk * 8]; b[1] = p[kk * 8 + 4]; }
};
// B operand, k = row of a row-major buffer (a key / query; rows beyond the sample are clamped, their A entries are zero),
// n = column:  B(k, n) = buf[r0 + k][coloff + n0 + n]
struct LoadB_rowsK {
    const float* base; int ld, N, t;
    __device__ __forceinline__ LoadB_rowsK(const float* buf, int ld_, int r0, int n0, int N_, int coloff) : ld(ld_), N(N_) {
        const int g = (threadIdx.x & 31) >> 2;
        t = threadIdx.x & 3;
        base = buf + r0 * ld + coloff + n0 + g;
    }
    __device__ __forceinline__ void operator()(int kk, float (&b)[2]) const {
        b[0] = base[min(kk * 8 + t, N - 1) * ld];
        b[1] = base[min(kk * 8 + t + 4, N - 1) * ld];
    }
};
// B operand = folded edge map of the head: B(k = d, n = c) = A[hc * 64 + d][c] for c < 3 (rows of W.A are (a0, a1, a2, 0))
struct LoadB_edge {
    const float* p; bool on;
    __device__ __forceinline__ LoadB_edge(const float* __restrict__ A, int hc) {
        const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
        on = g < 3;
        p = A + (hc * 64 + t) * 4 + (on ? g : 0);
    }
    __device__ __forceinline__ void operator()(int kk, float (&b)[2]) const {
        b[0] = on ? __ldg(p + kk * 32) : 0.f;
        b[1] = on ? __ldg(p + kk * 32 + 16) : 0.f;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Forward phase 1: logits  sP[u][j] = s q_u . k'_j   (tiles: sample x row block x key block)
template <class C>
__device__ __forceinline__ void attn_logits_mma(const float* sQKV, float* sP, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = G.MB * G.NK, items = G.S_act * per;
    for (int it = warp; it < items; it += kCW) {
        const int s = it / per, rem = it - s * per, mb = rem / G.NK, nt = rem - mb * G.NK;
        const int r0 = s * G.N;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        tile_mma(c, 8, LoadA_rows(sQKV, C::LDQ, r0, mb * 16, G.N, 0), LoadB_rowsN(sQKV, C::LDQ, r0, nt * 8, G.N, 64));
        const int col = nt * 8 + 2 * t;
        if (col < G.NP) {
            const int ra = mb * 16 + g, rb = ra + 8;
            if (ra < G.N) *reinterpret_cast<float2*>(sP + (r0 + ra) * G.NP + col) = make_float2(kAttnScale * c[0], kAttnScale * c[1]);
            if (rb < G.N) *reinterpret_cast<float2*>(sP + (r0 + rb) * G.NP + col) = make_float2(kAttnScale * c[2], kAttnScale * c[3]);
        }
    }
}
// Forward phase 2: softmax over the keys of the row's sample, in place; p -> stash.  One warp per row, lane = key (and key + 32).
__device__ __forceinline__ void attn_softmax_rows(float* sP, float* st_p, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = G.S_act * G.N;
    for (int r = warp; r < rows; r += kCW) {
        float* row = sP + r * G.NP;
        const bool a0 = lane < G.N, a1 = lane + 32 < G.N;
        const float l0 = a0 ? row[lane] : -INFINITY, l1 = a1 ? row[lane + 32] : -INFINITY;
        float m = fmaxf(l0, l1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float e0 = a0 ? expf(l0 - m) : 0.f, e1 = a1 ? expf(l1 - m) : 0.f;
        const float sum = warp_sum(e0 + e1);
        const float p0 = e0 / sum, p1 = e1 / sum;
        if (lane < G.NP) { row[lane] = p0; st_p[(size_t)r * G.NP + lane] = p0; }
        if (lane + 32 < G.NP) { row[lane + 32] = p1; st_p[(size_t)r * G.NP + lane + 32] = p1; }
    }
}
// Forward phase 3: o_u = sum_j p_uj v'_j - A x_u + c  -> canonical hi/lo operand of the out-projection (the rotating slot).
// f_store(row, col, v0, v1) receives two consecutive columns (col even) of an active row.
template <class C, class F>
__device__ __forceinline__ void attn_weighted_mma(const float* sW, const float* sB, int ldb, int coloff, const AttnGeo& G, F f_store) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = G.MB * 8, items = G.S_act * per;
    for (int it = warp; it < items; it += kCW) {
        const int s = it / per, rem = it - s * per, mb = rem >> 3, nt = rem & 7;
        const int r0 = s * G.N;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        tile_mma(c, G.NK, LoadA_keys(sW, G.NP, r0, mb * 16, G.N), LoadB_rowsK(sB, ldb, r0, nt * 8, G.N, coloff));
        const int ra = mb * 16 + g, rb = ra + 8, col = nt * 8 + 2 * t;
        if (ra < G.N) f_store(r0 + ra, col, c[0], c[1]);
        if (rb < G.N) f_store(r0 + rb, col, c[2], c[3]);
    }
}

// Reverse phase 1: dp_uj = do_u . v'_j -> sDS (raw);  u_u = A_h^T q_u -> sQKV[u][192..194];  w_u = A_h^T do_u -> sO[u][64..66]
template <class C>
__device__ __forceinline__ void attn_dp_uw_mma(float* sQKV, float* sO, float* sDS, const float* __restrict__ Aedge, int hc, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = G.MB * (G.NK + 2), items = G.S_act * per;
    for (int it = warp; it < items; it += kCW) {
        const int s = it / per, rem = it - s * per, mb = rem / (G.NK + 2), nt = rem - mb * (G.NK + 2);
        const int r0 = s * G.N;
        const int ra = mb * 16 + g, rb = ra + 8;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        if (nt < G.NK) {
            tile_mma(c, 8, LoadA_rows(sO, C::LDO, r0, mb * 16, G.N, 0), LoadB_rowsN(sQKV, C::LDQ, r0, nt * 8, G.N, 128));
            const int col = nt * 8 + 2 * t;
            if (col < G.NP) {
                if (ra < G.N) *reinterpret_cast<float2*>(sDS + (r0 + ra) * G.NP + col) = make_float2(c[0], c[1]);
                if (rb < G.N) *reinterpret_cast<float2*>(sDS + (r0 + rb) * G.NP + col) = make_float2(c[2], c[3]);
            }
        } else {
            const bool is_u = nt == G.NK;
            float* dst = is_u ? sQKV : sO;
            const int ld = is_u ? C::LDQ : C::LDO, off = is_u ? 192 : 64;
            tile_mma(c, 8, LoadA_rows(dst, ld, r0, mb * 16, G.N, 0), LoadB_edge(Aedge, hc));
            if (t < 2) {        // columns 2t, 2t + 1 of the 8-wide tile: (0, 1) and (2, 3); column 3 is zero
                if (ra < G.N) *reinterpret_cast<float2*>(dst + (r0 + ra) * ld + off + 2 * t) = make_float2(c[0], c[1]);
                if (rb < G.N) *reinterpret_cast<float2*>(dst + (r0 + rb) * ld + off + 2 * t) = make_float2(c[2], c[3]);
            }
        }
    }
}
// Reverse phase 2: ds_uj = p_uj (dp_uj - sum_j p_uj dp_uj), in place in sDS.  One warp per row.
__device__ __forceinline__ void attn_ds_rows(const float* sP, float* sDS, const AttnGeo& G) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = G.S_act * G.N;
    for (int r = warp; r < rows; r += kCW) {
        const float* pr = sP + r * G.NP;
        float* dr = sDS + r * G.NP;
        const bool a0 = lane < G.N, a1 = lane + 32 < G.N;
        const float p0 = a0 ? pr[lane] : 0.f, p1 = a1 ? pr[lane + 32] : 0.f;
        const float d0 = a0 ? dr[lane] : 0.f, d1 = a1 ? dr[lane + 32] : 0.f;
        const float tsum = warp_sum(p0 * d0 + p1 * d1);
        if (lane < G.NP) dr[lane] = p0 * (d0 - tsum);
        if (lane + 32 < G.NP) dr[lane + 32] = p1 * (d1 - tsum);
    }
}
// Reverse phase 3: dx_j += sum_i (p_ij w_i + s ds_ij u_i) - w_j.  One thread per (key row, component): deterministic.
template <class C>
__device__ __forceinline__ void attn_dx_rows(const float* sQKV, const float* sO, const float* sP, const float* sDS, float* sDX, const AttnGeo& G) {
    const int rows = G.S_act * G.N;
    for (int idx = threadIdx.x; idx < rows * 3; idx += kCT) {
        const int j = idx / 3, cc = idx - j * 3;
        const int r0 = (j / G.N) * G.N, jj = j - r0;
        float acc = 0.f, acd = 0.f;
        for (int i = 0; i < G.N; ++i) {
            acc = fmaf(sP[(r0 + i) * G.NP + jj], sO[(r0 + i) * C::LDO + 64 + cc], acc);
            acd = fmaf(sDS[(r0 + i) * G.NP + jj], sQKV[(r0 + i) * C::LDQ + 192 + cc], acd);
        }
        sDX[j * 4 + cc] += (acc + kAttnScale * acd) - sO[j * C::LDO + 64 + cc];
    }
}
// Reverse phases 4-6: 16 x 8 tiles of  dq = s dS K'  (TRANSPOSED = false, weights sDS, B = k' columns) or
// dk' = s dS^T Q / dv' = P^T dO (TRANSPOSED = true) into registers (up to MAXT tiles per warp and round), then into the
// rotating slot once the tensor core has released it -- the MMAs of the previous slot job overlap the tile arithmetic.
template <class C, bool TRANSPOSED, class CTX>
__device__ __forceinline__ void attn_grad_to_slot_mma(CTX& c, const float* sW, const float* sB, int ldb, int coloff, float scale, const AttnGeo& G) {
    constexpr int MAXT = 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int per = G.MB * 8, items = G.S_act * per;
    for (int base = warp; base < items; base += kCW * MAXT) {
        float acc[MAXT][4];
#pragma unroll
        for (int q = 0; q < MAXT; ++q) {
            acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
            const int it = base + q * kCW;
            if (it < items) {
                const int s = it / per, rem = it - s * per, mb = rem >> 3, nt = rem & 7;
                const int r0 = s * G.N;
                if (TRANSPOSED) tile_mma(acc[q], G.NK, LoadA_keysT(sW, G.NP, r0, mb * 16, G.N), LoadB_rowsK(sB, ldb, r0, nt * 8, G.N, coloff));
                else tile_mma(acc[q], G.NK, LoadA_keys(sW, G.NP, r0, mb * 16, G.N), LoadB_rowsK(sB, ldb, r0, nt * 8, G.N, coloff));
            }
        }
        c.slot_acquire();
#pragma unroll
        for (int q = 0; q < MAXT; ++q) {
            const int it = base + q * kCW;
            if (it < items) {
                const int s = it / per, rem = it - s * per, mb = rem >> 3, nt = rem & 7;
                const int r0 = s * G.N, ra = mb * 16 + g, rb = ra + 8, col = nt * 8 + 2 * t;
                if (ra < G.N) can_store2<C::kCS>(c.slot_hi, c.slot_lo, r0 + ra, col, scale * acc[q][0], scale * acc[q][1]);
                if (rb < G.N) can_store2<C::kCS>(c.slot_hi, c.slot_lo, r0 + rb, col, scale * acc[q][2], scale * acc[q][3]);
            }
        }
    }
    c.slot_acquire();       // warps without a tile still take part in the hand-off bookkeeping
}

}  // namespace v2
}  // namespace dff
