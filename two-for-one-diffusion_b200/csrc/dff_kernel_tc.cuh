// dff_kernel_tc.cuh -- tcgen05 / TMEM version of the fused score-network + integrator kernel (sm_100a).
//
// Same math, same stash, same integrator epilogues as dff_kernel.cuh (the mma.sync version); what changes is
// how the dense projections run and how the CTA is organised:
//
//   * every projection is a chain of `tcgen05.mma.cta_group::1.kind::tf32` instructions (M = 64 rows, N = 64..256,
//     K = 8 per instruction) issued by ONE thread; accumulators live in TMEM (448 of 512 columns), the epilogues
//     read them back with `tcgen05.ld`;
//   * fp32-grade accuracy = 3xTF32 split precision: activations are written by their producers directly in the
//     UMMA canonical K-major layout as a TF32-exact high part and a low part, the weights are pre-split on the
//     host; lo*hi + hi*lo + hi*hi accumulate in the fp32 TMEM accumulator;
//   * weights arrive as pre-split, pre-laid-out [hi | lo] slices through a 2-4-stage ring filled by a dedicated
//     TMA producer thread (cp.async.bulk + mbarrier expect_tx) and released by `tcgen05.commit`;
//   * the CTA is warp-specialised: 16 compute warps (attention, softmax, LayerNorm, gates, GELU, integrators,
//     TMEM epilogues) + a service warpgroup (TMA producer warp, MMA issuer warp, two idle warps) that hands its registers
//     to the compute warps with setmaxnreg (96 -> 32 / 112).  The issuer follows a host-built job table, so
//     the QKV projection of head chunk h+1 and the out-projection of chunk h-1 run on the tensor core WHILE the
//     compute warps do the attention of chunk h (double-buffered TMEM accumulators);
//   * the two big stash pieces (q|k'|v' and p of a head chunk) leave and re-enter shared memory as TMA bulk copies;
//   * the folded edge term  A x_j  and the q/k/v biases ride inside the QKV GEMM as 8 extra K rows
//     (operand row = [n_hat | x0 x1 x2 1 0 0 0 0]), so that epilogue is a pure TMEM -> smem/stash copy.
//
// Shape support (TcCfg<PN, HP>): hidden 64 with N <= 64 beads, hidden 96 / 128 with N <= 56 -- every shipped checkpoint.
// Other shapes use dff_kernel.cuh.
// Reference lines replaced: the same as dff_kernel.cuh (models/graph_transformer.py:87-111,143-159,178-329;
// models/ddpm.py:195-251; dynamics/langevin.py:75-92; dynamics/langevin_cgnet.py:447-542,737-771).
#pragma once
#include "dff_kernel.cuh"
#include <stdio.h>
#include "dff_tc.cuh"

namespace dff {
namespace v2 {

constexpr int kMmaM = 64;                 // MMA M; a configuration uses kR <= kMmaM node rows per pass
#ifndef DFF_TC_SPLIT_ACC
#define DFF_TC_SPLIT_ACC 1      // separate TMEM accumulators for the split-precision correction products of the long K chains
#endif
#ifndef DFF_TC_COMPUTE_WARPS
#define DFF_TC_COMPUTE_WARPS 16
#endif
constexpr int kCW = DFF_TC_COMPUTE_WARPS;          // compute warps (multiple of 4: TMEM lane quarters)
constexpr int kCT = kCW * 32;                      // compute threads
constexpr int kComputeThreads = kCT;
// + one service warpgroup: TMA producer warp, MMA issuer warp and two idle warps.  Registers are allocated to a CTA per
// four warps (measured: a 576-thread CTA cannot launch above 96 registers), so the idle pair costs nothing, and a complete
// warpgroup is what `setmaxnreg` needs: the service warpgroup shrinks from the launch allocation of 96 registers per thread
// to 32 and the 16 compute warps grow to 112 (4 x 64 released = 16 x 16 acquired; ptxas schedules the compute phases with
// the larger budget).  Measured against the 96-register build: C2 +8.6 %, C3 +4.4 %, C4 +8.5 %, C5 +3.7 %.
// Opt-in: drop the consumed q|k'|v' / p stash lines from L2 (discard.global.L2, SASS CCTL.E.RML2) so that they are never written
// back.  Measured: DRAM bytes per C2 step 41.5 -> 25.8 MB (-38 %), C4 -9 %, but 1.3-1.7 % slower and DRAM sits at 2-8 % of peak.
#ifndef DFF_TC_DISCARD_STASH
#define DFF_TC_DISCARD_STASH 0
#endif
#ifndef DFF_TC_SETMAXNREG
#define DFF_TC_SETMAXNREG 1
#endif
constexpr int kTcThreads = kComputeThreads + (DFF_TC_SETMAXNREG ? 128 : 64);
constexpr int kRegsCompute = 112, kRegsService = 32;
constexpr int kTcStages = 3;
constexpr int kJobCap = 200;              // job-table entries (16 B each) cached in shared memory
constexpr int B_COUNT_ = 26;              // uint64 slots of the barrier block (== B_COUNT below)

// TMEM column map (fp32 columns, 64 lanes used): [0, HP) block accumulator (att / ff / d m_hat / d n_hat), then at HP the
// work area: 2 x 192 q|k'|v' double buffer, <= 256 columns of FF hidden, 2 x 64 d o double buffer.   HP + 384 <= 512.
constexpr uint32_t kColAcc = 0;
constexpr uint32_t kTmemCols = 512;

// One GEMM of the per-step schedule, in issue order (built by the host, dff_b200.cu).  16 bytes.
struct TcJob {
    uint32_t w_off;           // weight panel (floats from TcArgs::wbase): n_slices contiguous slices, each
                              // [hi image | lo image] of [ks/4][n][4] floats
    uint16_t slice_16b;       // slice size in 16-byte units
    uint8_t n_slices, ks;
    uint16_t n;               // MMA N (64, 128, 192)
    uint16_t d_col;           // TMEM column of D (double-buffered jobs: column of buffer 0)
    uint8_t flags;            // TCJ_*
    uint8_t pad_;
    uint16_t c_col;           // TMEM column of the correction accumulator (lo*hi + hi*lo); == d_col: none, everything into D
};
static_assert(sizeof(TcJob) == 16, "TcJob layout");
enum : uint32_t {
    TCJ_SLOT = 1,             // A operand = the rotating slot (else the persistent buffer: n_hat / d ff / d att)
    TCJ_ACC = 2,              // first MMA accumulates (else overwrites D)
    TCJ_WAIT_POST = 4,        // wait for the next "operand ready" post of the compute warps before issuing
    TCJ_DBUF = 8,             // D is double-buffered (QKV / d o jobs): gate on the drain counter, commit -> dq_ready[parity]
    TCJ_COMMIT_ACC = 16,      // commit -> acc_ready after the job
    TCJ_COMMIT_D1 = 32,       // commit -> d1_ready after the job
};

struct TcArgs {
    const TcJob* jobs;
    const float* wbase;
    int njobs_fwd, njobs_all;
    uint32_t nslice_fwd, nslice_all;
    long long* dbg;           // [grid][16] wait-cycle counters (DFF_TC_PROFILE builds), else NULL
};

// ------------------------------------------------------------------ configuration
// HP = 64 (hidden 64) keeps the node stream in shared memory and uses 16 KB weight stages; HP = 128 (hidden 96 / 128) keeps
// the node stream in the per-CTA global scratch (row passes only, coalesced, L2 resident) and uses 12 KB stages -- that is
// what lets the fp32 hi/lo operands of a 64-row tile (8 bytes per element) fit 227 KB.
// ATT_ selects how the attention contractions run: 0 = CUDA-core row / pair / quad routines (fastest for the small
// intrinsic-coordinate nets: chignolin, ala2, trp-cage), 1 = warp-level tensor-core tiles (dff_attn_mma.cuh: the general path).
template <int PN_, int HP_, int R_ = 64, int ATT_ = 0>
struct TcCfg {
    static constexpr int kHP = HP_, kR = R_, kHC = 1, kPN = PN_;
    static constexpr bool kAttMma = ATT_ != 0;      // kR node rows per pass (<= 64 = the MMA M; smaller frees shared memory)
    static constexpr int kCS = kR * 4 + 4;    // canonical chunk stride in floats (kR rows x 16 B + 16 B pad: bank spread)
    static constexpr int CWQ = 64, NCH = kHeads;
    static constexpr int LDH = kHP + 4;
    static constexpr int LDQ = 3 * CWQ + 4;
    static constexpr int LDO = CWQ + 4;
    static constexpr int EPL = kHP / 32;
    static constexpr int NHAT_CHUNKS = kHP / 4 + 2;          // + [x0 x1 x2 1] chunk + zero chunk (K = H + 8 for the QKV jobs)
    static constexpr int SLOT_CHUNKS = 16;
    static constexpr bool kNodeInSmem = (kHP == 64);
    // weight ring depth = what fits next to the row buffers: hidden 64 (16 KB stages): 3 (4 with 60-row buffers), 2 for N > 32; hidden 96 / 128
    // (12 KB stages): 4 when the pass has fewer than 64 rows (smaller row buffers), else 3
#ifdef DFF_EXP_STAGES
    static constexpr int kStages = DFF_EXP_STAGES;
#else
    static constexpr int kStages = (HP_ == 64) ? (PN_ > 32 ? 2 : (R_ < 64 && PN_ <= 12 ? 4 : 3)) : (R_ < 64 ? 4 : 3);
#endif
    static constexpr int kStageFloats = (kHP == 64) ? 4096 : 3072;
    static constexpr uint32_t kColD = kHP;                   // TMEM work area
    // Correction accumulators (DFF_TC_SPLIT_ACC): the two small split-precision products of a long K chain accumulate
    // apart from the hi*hi product, so only K/8 MMAs of the chain round at full magnitude; the epilogue adds the two.
    static constexpr uint32_t kCorrFF = kColD + 256;         // FF2 / d FF1 chains: right of the <= 256-column super
    static constexpr bool kCorrAttn = (kHP == 64);           // out-projection and d n_hat chains: only hidden <= 64 has the columns
    static constexpr uint32_t kCorrOut = kColD + 384;        // right of the double-buffered q|k'|v' accumulators
    static constexpr uint32_t kCorrDn = kColD + 128 + 2 * kHP;   // three of them, kHP apart, right of the three d n_hat accumulators
    // shared memory carve-up (float offsets)
    static constexpr int oN = 0;                             // [R][LDH] node stream / its gradient (HP = 64 only)
    static constexpr int oQKV = oN + (kNodeInSmem ? kR * LDH : 0);   // [R][LDQ] q|k'|v' of the head chunk; also the [R][LDH] row buffer
    static constexpr int oO = oQKV + kR * LDQ;               // [R][LDO] d o of the head chunk (reverse pass)
    static constexpr int oP = oO + kR * LDO;
    static constexpr int oDS = oP + kR * kPN;
    static constexpr int oNhatHi = oDS + kR * kPN;           // canonical persistent A operand
    static constexpr int oNhatLo = oNhatHi + NHAT_CHUNKS * kCS;
    static constexpr int oSlotHi = oNhatLo + NHAT_CHUNKS * kCS;   // canonical rotating A operand
    static constexpr int oSlotLo = oSlotHi + SLOT_CHUNKS * kCS;
    static constexpr int oW = oSlotLo + SLOT_CHUNKS * kCS;   // weight ring
    static constexpr int oX = oW + kStages * kStageFloats;
    static constexpr int oV = oX + kR * 4;
    static constexpr int oDX = oV + kR * 4;
    static constexpr int oTmp = oDX + kR * 4;
    static constexpr int oJobs = oTmp + kR * 4;
    static constexpr int oBar = oJobs + kJobCap * 4;
    static constexpr int kFloats = oBar + 2 * B_COUNT_ + 4;
    static constexpr size_t kSmemBytes = (size_t)kFloats * sizeof(float);
    static_assert(kR * LDH <= kR * LDQ, "row buffer must fit the q|k'|v' buffer");
    static_assert(oNhatHi % 4 == 0 && oSlotHi % 4 == 0 && oW % 4 == 0 && oJobs % 4 == 0 && oBar % 4 == 0, "16-byte alignment");
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");
    static_assert(kColD + 384 <= kTmemCols && kColD + 128 + 2 * kHP <= kTmemCols, "TMEM column budget");
};

// barrier block at oBar (uint64 slots).  Every hand-off of the kernel is an mbarrier (round 1 polled two progress counters with
// st.volatile / ld.acquire; the arrive / try_wait pair is the documented release / acquire pattern and costs the same).
//   B_POST  ring of 8: compute -> issuer "operand ready" (post i arrives on slot i & 7; the compute warps can never be more than
//           two posts ahead of the issuer: every slot_post first waits for the previous slot job)
//   B_DRAIN 2: compute -> issuer "double-buffered accumulator b read back"
enum { B_FULL = 0, B_EMPTY = 4, B_DQ = 8, B_ACC = 10, B_D1 = 11, B_SLOT = 12, B_RELOAD = 13, B_POST = 16, B_DRAIN = 24, B_COUNT = 26 };      // room for 4 ring stages

// ------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, %0;" ::"n"(kCT) : "memory"); }   // the compute warps
// Bounded waits: a protocol bug must fail loudly (trap -> CUDA error -> DFF_ECUDA), never hang the GPU.
// try_wait without a suspend-time hint returns after a short hardware time-out, so the loops below re-arm it; measured
// (profiles/r02/README.md): a 0.2 ms hint (NANOSLEEP.SYNCS) removes the service warps' polling instructions (a third of all
// instructions executed) but wakes later -- no gain on chignolin, -5 % on trp-cage, whose weight stream lives on wake-up latency.
constexpr uint32_t kSpinLimit = 1u << 27;
// The default build traps without a message: an out-of-line printf makes every wait a call site, and ptxas then spills
// the caller's live registers around each of the ~250 waits of a step (96-188 bytes of spill traffic per wait in the
// 96-register build, zero without the call).  -DDFF_TC_WATCHDOG_PRINTF names the wait that timed out.
#ifdef DFF_TC_WATCHDOG_PRINTF
static __device__ __noinline__ void watchdog_fail(int tag) {
    printf("dff_fused_tc_kernel watchdog: wait %d never completed (block %d, thread %d)\n", tag, (int)blockIdx.x, (int)threadIdx.x);
    __trap();
}
#else
static __device__ __forceinline__ void watchdog_fail(int) { __trap(); }
#endif
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, int tag) {
    uint32_t n = 0;
    while (!mbar_try(bar, parity))
        if (++n > kSpinLimit) watchdog_fail(tag);
}
// Optional wait-time accounting (-DDFF_TC_PROFILE): cycles spent in each kind of wait, per CTA, written to T.dbg.
#ifdef DFF_TC_PROFILE
#define TCP_BEGIN() const long long tcp_t0_ = clock64()
#define TCP_END(arr, k) (arr)[k] += clock64() - tcp_t0_
#else
#define TCP_BEGIN() do { } while (0)
#define TCP_END(arr, k) do { } while (0)
#endif
// TMEM -> registers: 16 consecutive fp32 columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// D[64 x NCOLS] at TMEM column `col`: with M = 64 row m lives in lane (m % 16) + 32 * (m / 16), so compute warp w reads
// lane quarter w & 3 (rows 16 (w & 3) .. + 15 in its lanes 0..15) and the column part w >> 2.
// f(row, col_in_tile, v[16]) is called by the lanes that own an active row (row < rows).
template <int NCOLS, class F>
__device__ __forceinline__ void tmem_foreach(uint32_t tmem_base, uint32_t col, int rows, F f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PARTS = kCW / 4;
    const int q = warp & 3, part = warp >> 2;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + col;
    const int row = q * 16 + lane;
    if (q * 16 >= rows) return;                    // warp-uniform: this lane quarter holds no active row
#pragma unroll 1
    for (int c = part * 16; c < NCOLS; c += 16 * PARTS) {      // 16-column chunks round-robin over the warp parts
        float v[16];
        tmem_ld16(taddr + (uint32_t)c, v);
        if (lane < 16 && row < rows) f(row, c, v);
    }
}
// same, D = accumulator at `col` + correction accumulator at `col2` (compile-time switch: SPLIT false -> plain read)
template <int NCOLS, bool SPLIT, class F>
__device__ __forceinline__ void tmem_foreach_sum(uint32_t tmem_base, uint32_t col, uint32_t col2, int rows, F f) {
    if constexpr (!SPLIT) {
        tmem_foreach<NCOLS>(tmem_base, col, rows, f);
    } else {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        constexpr int PARTS = kCW / 4;
        const int q = warp & 3, part = warp >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const int row = q * 16 + lane;
        if (q * 16 >= rows) return;
#pragma unroll 1
        for (int c = part * 16; c < NCOLS; c += 16 * PARTS) {
            float v[16], w[16];
            tmem_ld16(lane_base + col + (uint32_t)c, v);
            tmem_ld16(lane_base + col2 + (uint32_t)c, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w[i];
            if (lane < 16 && row < rows) f(row, c, v);
        }
    }
}

// canonical (UMMA K-major, no swizzle) operand stores with the round-to-nearest TF32 hi/lo split
template <int CS>
__device__ __forceinline__ void can_store4(float* hi, float* lo, int r, int k4, const float4& x) {
    float4 h, l;
    tc::split4(x, h, l);
    *reinterpret_cast<float4*>(hi + k4 * CS + r * 4) = h;
    *reinterpret_cast<float4*>(lo + k4 * CS + r * 4) = l;
}
template <int CS>
__device__ __forceinline__ void can_store2(float* hi, float* lo, int r, int col, float x0, float x1) {   // col even
    float2 h, l;
    h.x = __uint_as_float((__float_as_uint(x0) + 0x1000u) & 0xffffe000u); l.x = x0 - h.x;
    h.y = __uint_as_float((__float_as_uint(x1) + 0x1000u) & 0xffffe000u); l.y = x1 - h.y;
    const int o = (col >> 2) * CS + r * 4 + (col & 3);
    *reinterpret_cast<float2*>(hi + o) = h;
    *reinterpret_cast<float2*>(lo + o) = l;
}

// exp of softmax / sigmoid / Gaussian-pdf arguments: the ex2.approx path (2 ulp, plus |x| 2^-24 relative from the scaling).
// Together with the approximate division / square root this unit is compiled with: C2 +3.7 %, C3 +5.3 %, C4 +1.9 %, C5 +1.2 %,
// worst force errors unchanged (profiles/r02/exp_fast_math*.log).  The RNG (logf / sincosf) and the integrators stay precise.
__device__ __forceinline__ float fast_exp(float x) { return __expf(x); }

// ------------------------------------------------------------------ per-CTA context of the compute warps
struct Ctx2 {
    float *sN, *sNh, *sQKV, *sO, *sP, *sDS, *sX, *sV, *sDX, *sTmp;
    float *nhat_hi, *nhat_lo, *slot_hi, *slot_lo;
    uint64_t* bars;
    uint32_t tmem;
    uint32_t n_post, n_drain, n_acc, n_d1, n_slot, n_reload;   // identical in every compute thread
    bool slot_held;
    bool pairs;            // pair-local attention routines (groups with more rows than lane groups)
    bool quads;            // quad-local routines (N > 32: two keys per lane, one quad per warp)
    long long tw[8];       // DFF_TC_PROFILE: cycles waited on {dq, acc, d1, slot}
#ifdef DFF_TC_PROFILE
    long long ph[32], last;
    __device__ __forceinline__ void mark(int k) { const long long t = clock64(); ph[k] += t - last; last = t; }
#else
    __device__ __forceinline__ void mark(int) {}
#endif
    float* stash;
    int rows_act, S_act;

    // operand written (generic proxy) -> visible to the tensor core (async proxy) -> tell the issuer
    __device__ __forceinline__ void post() {
        fence_proxy_async();
        tc::fence_before_sync();
        csync();
        if (threadIdx.x == 0) mbar_arrive(bars + B_POST + (n_post & 7u));      // post number n_post -> ring slot n_post & 7
        ++n_post;
    }
    // wait for the double-buffered accumulator of the next QKV / d o job; returns its buffer index
    __device__ __forceinline__ int dq_wait() {
        const int b = n_drain & 1;
        TCP_BEGIN();
        mbar_wait_wd(bars + B_DQ + b, (n_drain >> 1) & 1u, 1);
        TCP_END(tw, 0);
        tc::fence_after_sync();
        return b;
    }
    __device__ __forceinline__ void dq_release() {      // accumulator read back (and its smem copy written)
        tc::fence_before_sync();
        csync();
        if (threadIdx.x == 0) mbar_arrive(bars + B_DRAIN + (n_drain & 1u));    // buffer n_drain & 1 is free again
        ++n_drain;
    }
    __device__ __forceinline__ void acc_wait() { TCP_BEGIN(); mbar_wait_wd(bars + B_ACC, n_acc & 1u, 2); TCP_END(tw, 1); ++n_acc; tc::fence_after_sync(); }
    __device__ __forceinline__ void d1_wait() { TCP_BEGIN(); mbar_wait_wd(bars + B_D1, n_d1 & 1u, 3); TCP_END(tw, 2); ++n_d1; tc::fence_after_sync(); }
    // the rotating slot may be overwritten once the MMAs of its previous use have completed
    __device__ __forceinline__ void slot_acquire() {
        if (!slot_held) {
            TCP_BEGIN();
            if (n_slot > 0) mbar_wait_wd(bars + B_SLOT, (n_slot - 1) & 1u, 4);
            TCP_END(tw, 3);
            slot_held = true;
        }
    }
    __device__ __forceinline__ void slot_post() { slot_acquire(); slot_held = false; ++n_slot; post(); }
};

// ------------------------------------------------------------------ warp-per-row phases writing canonical operands
// (only the active rows are processed; pad rows of the operands hold stale finite values whose products are never read)
// A lane owns E = HP / 32 consecutive columns of its row: E = 2 (hidden 64) or 4 (hidden 96 / 128; lanes beyond H idle).
template <int E> struct RowVec { float v[E]; };
template <int E>
__device__ __forceinline__ RowVec<E> row_ld(const float* p, bool ok) {
    RowVec<E> r;
    if (ok) {
        if (E == 4) { const float4 t = *reinterpret_cast<const float4*>(p); r.v[0] = t.x; r.v[1] = t.y; r.v[2 % E] = t.z; r.v[3 % E] = t.w; }
        else { const float2 t = *reinterpret_cast<const float2*>(p); r.v[0] = t.x; r.v[1] = t.y; }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) r.v[e] = 0.f;
    }
    return r;
}
template <int E>
__device__ __forceinline__ RowVec<E> row_ldg(const float* __restrict__ p, bool ok) {
    RowVec<E> r;
    if (ok) {
        if (E == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2 % E] = t.z; r.v[3 % E] = t.w; }
        else { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); r.v[0] = t.x; r.v[1] = t.y; }
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) r.v[e] = 0.f;
    }
    return r;
}
template <int E>
__device__ __forceinline__ void row_st(float* p, const RowVec<E>& r) {
    if (E == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2 % E], r.v[3 % E]);
    else *reinterpret_cast<float2*>(p) = make_float2(r.v[0], r.v[1]);
}
template <int E, int CS>
__device__ __forceinline__ void row_can_store(float* hi, float* lo, int r, int col, const RowVec<E>& x) {
    if (E == 4) can_store4<CS>(hi, lo, r, col >> 2, make_float4(x.v[0], x.v[1], x.v[2 % E], x.v[3 % E]));
    else can_store2<CS>(hi, lo, r, col, x.v[0], x.v[1]);
}
template <int E>
__device__ __forceinline__ float row_sum(const RowVec<E>& x) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s += x.v[e];
    return s;
}

// LayerNorm of sN rows -> canonical n_hat; stats -> stash; stashes the input rows.
template <class C>
__device__ __forceinline__ void ln_forward_rows_can(const float* sN, float* hi, float* lo, const float* __restrict__ gam,
                                                    const float* __restrict__ bet, int H, int rows, float* st_rows, float* st_stats) {
    constexpr int E = C::EPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = lane * E;
    const bool ok = col < H;
    for (int r = warp; r < rows; r += kCW) {
        const RowVec<E> x = row_ld<E>(sN + r * C::LDH + col, ok);
        const float mean = warp_sum(row_sum<E>(x)) / (float)H;
        RowVec<E> d;
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) { d.v[e] = ok ? x.v[e] - mean : 0.f; q += d.v[e] * d.v[e]; }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + kLnEps);
        if (ok) {
            const RowVec<E> g = row_ldg<E>(gam + col, true), b = row_ldg<E>(bet + col, true);
            RowVec<E> y;
#pragma unroll
            for (int e = 0; e < E; ++e) y.v[e] = d.v[e] * rstd * g.v[e] + b.v[e];
            row_can_store<E, C::kCS>(hi, lo, r, col, y);
            row_st<E>(st_rows + (size_t)r * H + col, x);
        }
        if (lane == 0) { st_stats[r * 2] = mean; st_stats[r * 2 + 1] = rstd; }
    }
}

// GatedResidual forward (graph_transformer.py:202-205): a = sA rows, n = sN -> out -> sN; then (if gam)
// LayerNorm(out) -> canonical operand.  Stashes a, gate, out (and LN stats).
template <class C>
__device__ __forceinline__ void gate_ln_forward_rows_can(float* sN, const float* sA, float* hi, float* lo,
                                                         const float* __restrict__ ga, const float* __restrict__ gb, int H, int rows,
                                                         float* st_a, float* st_g, float* st_out,
                                                         const float* __restrict__ gam, const float* __restrict__ bet,
                                                         float* st_stats) {
    constexpr int E = C::EPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = lane * E;
    const bool ok = col < H;
    for (int r = warp; r < rows; r += kCW) {
        const RowVec<E> a = row_ld<E>(sA + r * C::LDH + col, ok), n = row_ld<E>(sN + r * C::LDH + col, ok);
        const RowVec<E> wa = row_ldg<E>(ga + col, ok), wb = row_ldg<E>(gb + col, ok);
        float z = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) z += a.v[e] * wa.v[e] + n.v[e] * wb.v[e];
        z = warp_sum(z);
        const float g = 1.0f / (1.0f + fast_exp(-z));
        RowVec<E> o;
#pragma unroll
        for (int e = 0; e < E; ++e) o.v[e] = a.v[e] * g + n.v[e] * (1.0f - g);
        if (ok) {
            row_st<E>(st_a + (size_t)r * H + col, a);
            row_st<E>(st_out + (size_t)r * H + col, o);
            row_st<E>(sN + r * C::LDH + col, o);
        }
        if (lane == 0) st_g[r] = g;
        if (gam == nullptr) continue;
        const float mean = warp_sum(row_sum<E>(o)) / (float)H;
        RowVec<E> d;
        float q = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) { d.v[e] = ok ? o.v[e] - mean : 0.f; q += d.v[e] * d.v[e]; }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)H + kLnEps);
        if (ok) {
            const RowVec<E> gg = row_ldg<E>(gam + col, true), bb = row_ldg<E>(bet + col, true);
            RowVec<E> y;
#pragma unroll
            for (int e = 0; e < E; ++e) y.v[e] = d.v[e] * rstd * gg.v[e] + bb.v[e];
            row_can_store<E, C::kCS>(hi, lo, r, col, y);
        }
        if (lane == 0) { st_stats[r * 2] = mean; st_stats[r * 2 + 1] = rstd; }
    }
}

// Reverse of [LayerNorm ->] GatedResidual on rows.
//   dout = sN (+ LN-backward of sD rows through (st_ln_in, stats, gam) when gam != nullptr)
//   d(gate input a) -> canonical operand,  d(residual n) -> sN.      a, n, g come from the stash.
template <class C>
__device__ __forceinline__ void gate_backward_rows_can(float* sN, const float* sD, float* hi, float* lo, int H, int rows,
                                                       const float* __restrict__ gam, const float* st_ln_in,
                                                       const float* st_stats, const float* st_a, const float* st_n,
                                                       const float* st_g, const float* __restrict__ ga,
                                                       const float* __restrict__ gb) {
    constexpr int E = C::EPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = lane * E;
    const bool ok = col < H;
    for (int r = warp; r < rows; r += kCW) {
        const RowVec<E> a = row_ld<E>(st_a + (size_t)r * H + col, ok), n = row_ld<E>(st_n + (size_t)r * H + col, ok);
        RowVec<E> d = row_ld<E>(sN + r * C::LDH + col, ok);
        const float g = st_g[r];
        if (gam != nullptr) {
            const float mean = st_stats[r * 2], rstd = st_stats[r * 2 + 1];
            const RowVec<E> xin = row_ld<E>(st_ln_in + (size_t)r * H + col, ok), dd = row_ld<E>(sD + r * C::LDH + col, ok);
            const RowVec<E> gg = row_ldg<E>(gam + col, ok);
            RowVec<E> y, dy;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                y.v[e] = ok ? (xin.v[e] - mean) * rstd : 0.f;
                dy.v[e] = dd.v[e] * gg.v[e];
                s1 += dy.v[e]; s2 += dy.v[e] * y.v[e];
            }
            s1 = warp_sum(s1) / (float)H;
            s2 = warp_sum(s2) / (float)H;
#pragma unroll
            for (int e = 0; e < E; ++e) if (ok) d.v[e] += rstd * (dy.v[e] - s1 - y.v[e] * s2);
        }
        float dg = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) dg += d.v[e] * (a.v[e] - n.v[e]);
        dg = warp_sum(dg);
        const float dz = dg * g * (1.0f - g);
        if (ok) {
            const RowVec<E> wa = row_ldg<E>(ga + col, true), wb = row_ldg<E>(gb + col, true);
            RowVec<E> da, dn;
#pragma unroll
            for (int e = 0; e < E; ++e) { da.v[e] = d.v[e] * g + dz * wa.v[e]; dn.v[e] = d.v[e] * (1.0f - g) + dz * wb.v[e]; }
            row_can_store<E, C::kCS>(hi, lo, r, col, da);
            row_st<E>(sN + r * C::LDH + col, dn);
        }
    }
}

// sN += LayerNorm-backward(sD) through (st_ln_in, stats, gam)
template <class C>
__device__ __forceinline__ void ln_backward_rows_tc(float* sN, const float* sD, int H, int rows, const float* __restrict__ gam,
                                                    const float* st_ln_in, const float* st_stats) {
    constexpr int E = C::EPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = lane * E;
    const bool ok = col < H;
    for (int r = warp; r < rows; r += kCW) {
        const float mean = st_stats[r * 2], rstd = st_stats[r * 2 + 1];
        const RowVec<E> xin = row_ld<E>(st_ln_in + (size_t)r * H + col, ok), dd = row_ld<E>(sD + r * C::LDH + col, ok);
        const RowVec<E> gg = row_ldg<E>(gam + col, ok);
        RowVec<E> y, dy;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            y.v[e] = ok ? (xin.v[e] - mean) * rstd : 0.f;
            dy.v[e] = dd.v[e] * gg.v[e];
            s1 += dy.v[e]; s2 += dy.v[e] * y.v[e];
        }
        s1 = warp_sum(s1) / (float)H;
        s2 = warp_sum(s2) / (float)H;
        if (ok) {
            RowVec<E> d = row_ld<E>(sN + r * C::LDH + col, true);
#pragma unroll
            for (int e = 0; e < E; ++e) d.v[e] += rstd * (dy.v[e] - s1 - y.v[e] * s2);
            row_st<E>(sN + r * C::LDH + col, d);
        }
    }
}

// ------------------------------------------------------------------ attention contractions: warp-level tensor-core tiles
}  // namespace v2
}  // namespace dff
#include "dff_attn_mma.cuh"
namespace dff {
namespace v2 {

// ------------------------------------------------------------------ row-local attention (a group of LPR lanes owns one node row)
// LPR = 16 lanes per row for N <= 16 beads (two rows per warp), 32 above.  Inside a group the lane index is the key
// index j while logits / probabilities are formed and the output-column group (DPL = 64 / LPR columns) while values are
// accumulated, so logits -> softmax -> P V' (forward) and dp -> ds -> dq (reverse) need no CTA-wide barrier in between
// and all compute warps work on every chunk even when the CTA holds only two samples.
template <class C>
struct AttnMap {
    // Largest padded N that uses 16-lane groups.  32 (two keys per lane for N in 17..32: half the shared-memory wavefronts
    // per FMA, but half as many busy warps) was measured on trp-cage: 464 -> 399 MD steps/s -- these phases live on
    // thread-level parallelism, not on shared-memory bandwidth.
#ifndef DFF_TC_LPR16_MAX
#define DFF_TC_LPR16_MAX 16
#endif
    static constexpr int LPR = (C::kPN <= DFF_TC_LPR16_MAX) ? 16 : 32;
    static constexpr int DPL = 64 / LPR;
    static constexpr int UPW = 32 / LPR;              // rows per warp and round
    static constexpr int KPL = (C::kPN > LPR) ? 2 : 1; // keys per lane (lane j also owns key j + LPR; pair- / quad-local routines only)
};
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int LPR>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// K independent reductions at once: the K shuffles of a butterfly step are issued back to back, so their latencies
// overlap (K separate group_sum calls separated by branches run their 5-shuffle chains one after the other)
template <int LPR, int K>
__device__ __forceinline__ void group_sum_n(float (&v)[K]) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        float t[K];
#pragma unroll
        for (int k = 0; k < K; ++k) t[k] = __shfl_xor_sync(0xffffffffu, v[k], o);
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] += t[k];
    }
}
template <int LPR, int K>
__device__ __forceinline__ void group_max_n(float (&v)[K]) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        float t[K];
#pragma unroll
        for (int k = 0; k < K; ++k) t[k] = __shfl_xor_sync(0xffffffffu, v[k], o);
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = fmaxf(v[k], t[k]);
    }
}
// dots of one (or two) shared-memory rows a (a2) with the N rows b_j of the sample, for all j at once:
// every lane takes the DPL columns it owns, forms the partial products for every key j and the partials are then
// summed across the group by a transpose-reduce (LPR - 1 shuffles), after which lane j holds  a . b_j.
// Compared with "lane j walks the whole rows a and b_j" this moves 1 + N instead of 32 LDS.128 per lane through the
// shared-memory pipe, which is what bounds the attention phases.
template <int LPR>
__device__ __forceinline__ float transpose_reduce(float (&v)[LPR], int sub) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
        const bool up = (sub & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
            const float keep = up ? v[k + o] : v[k];
            const float send = up ? v[k] : v[k + o];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}
template <int LPR, int DPL, bool TWO>
__device__ __forceinline__ void group_dots(const float* __restrict__ a, const float* __restrict__ a2, const float* __restrict__ b, int ld,
                                           int N, int sub, float& r, float& r2) {
    float av[DPL], av2[DPL];
    if (DPL == 4) {
        const float4 t = *reinterpret_cast<const float4*>(a + sub * DPL);
        av[0] = t.x; av[1] = t.y; av[2 % DPL] = t.z; av[3 % DPL] = t.w;
        if (TWO) { const float4 u = *reinterpret_cast<const float4*>(a2 + sub * DPL); av2[0] = u.x; av2[1] = u.y; av2[2 % DPL] = u.z; av2[3 % DPL] = u.w; }
    } else {
        const float2 t = *reinterpret_cast<const float2*>(a + sub * DPL);
        av[0] = t.x; av[1] = t.y;
        if (TWO) { const float2 u = *reinterpret_cast<const float2*>(a2 + sub * DPL); av2[0] = u.x; av2[1] = u.y; }
    }
    float part[LPR], part2[TWO ? LPR : 1];
#pragma unroll
    for (int j = 0; j < LPR; ++j) {
        float s = 0.f, s2 = 0.f;
        if (j < N) {
            if (DPL == 4) {
                const float4 y = *reinterpret_cast<const float4*>(b + j * ld + sub * DPL);
                s = fmaf(av[0], y.x, fmaf(av[1], y.y, fmaf(av[2 % DPL], y.z, av[3 % DPL] * y.w)));
                if (TWO) s2 = fmaf(av2[0], y.x, fmaf(av2[1], y.y, fmaf(av2[2 % DPL], y.z, av2[3 % DPL] * y.w)));
            } else {
                const float2 y = *reinterpret_cast<const float2*>(b + j * ld + sub * DPL);
                s = fmaf(av[0], y.x, av[1] * y.y);
                if (TWO) s2 = fmaf(av2[0], y.x, av2[1] * y.y);
            }
        }
        part[j] = s;
        if (TWO) part2[j] = s2;
    }
    r = transpose_reduce<LPR>(part, sub);
    if (TWO) r2 = transpose_reduce<LPR>(reinterpret_cast<float (&)[LPR]>(part2), sub);
}
// lane = key variant: this lane walks the whole 64-column rows a, a2 and b_sub (32 LDS.128, no shuffles).  Cheaper than the
// column-sliced form when the group is a full warp (the transpose-reduce then costs 31 shuffles per dot set).
__device__ __forceinline__ void lane_dots2(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ b,
                                           float& r0, float& r1) {
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const float4 y = *reinterpret_cast<const float4*>(b + 4 * t);
        const float4 x = *reinterpret_cast<const float4*>(a0 + 4 * t), z = *reinterpret_cast<const float4*>(a1 + 4 * t);
        d0 = fmaf(x.x, y.x, d0); d1 = fmaf(x.y, y.y, d1); d2 = fmaf(x.z, y.z, d2); d3 = fmaf(x.w, y.w, d3);
        e0 = fmaf(z.x, y.x, e0); e1 = fmaf(z.y, y.y, e1); e2 = fmaf(z.z, y.z, e2); e3 = fmaf(z.w, y.w, e3);
    }
    r0 = (d0 + d1) + (d2 + d3);
    r1 = (e0 + e1) + (e2 + e3);
}
// lane = key variant for two keys per lane (N > 32): rows a0, a1 against this lane's keys b0, b1 in one sweep (64 LDS.128, 4 dots)
__device__ __forceinline__ void lane_dots4(const float* __restrict__ a0, const float* __restrict__ a1, const float* __restrict__ b0,
                                           const float* __restrict__ b1, float (&r)[4]) {
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f, t00 = 0.f, t01 = 0.f, t10 = 0.f, t11 = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const float4 x = *reinterpret_cast<const float4*>(a0 + 4 * t), z = *reinterpret_cast<const float4*>(a1 + 4 * t);
        const float4 y = *reinterpret_cast<const float4*>(b0 + 4 * t), w = *reinterpret_cast<const float4*>(b1 + 4 * t);
        s00 = fmaf(x.x, y.x, s00); t00 = fmaf(x.y, y.y, t00); s00 = fmaf(x.z, y.z, s00); t00 = fmaf(x.w, y.w, t00);
        s01 = fmaf(x.x, w.x, s01); t01 = fmaf(x.y, w.y, t01); s01 = fmaf(x.z, w.z, s01); t01 = fmaf(x.w, w.w, t01);
        s10 = fmaf(z.x, y.x, s10); t10 = fmaf(z.y, y.y, t10); s10 = fmaf(z.z, y.z, s10); t10 = fmaf(z.w, y.w, t10);
        s11 = fmaf(z.x, w.x, s11); t11 = fmaf(z.y, w.y, t11); s11 = fmaf(z.z, w.z, s11); t11 = fmaf(z.w, w.w, t11);
    }
    r[0] = s00 + t00; r[1] = s01 + t01; r[2] = s10 + t10; r[3] = s11 + t11;     // (row a0, key b0), (a0, b1), (a1, b0), (a1, b1)
}
// measured on protein G (C5): 169 -> 201 steps/s against two column-sliced passes with their 4 x 31-shuffle transpose-reduces
#ifndef DFF_TC_LANE_DOTS4
#define DFF_TC_LANE_DOTS4 1
#endif
// two-row dots of a group against keys [0, nk) of the rows at b (row stride ld): the lane = key form for full-warp groups
// with one key per lane (measured: trp-cage C4 367 -> 391 steps/s; with two keys per lane, protein G, it is 1.4 % slower),
// the column-sliced form otherwise
template <int LPR, int DPL, int KPL>
__device__ __forceinline__ void pair_dots(const float* a, const float* a2, const float* b, int ld, int nk, int sub, float& r, float& r2) {
    if (LPR == 32 && KPL == 1) lane_dots2(a, a2, b + min(sub, max(nk - 1, 0)) * ld, r, r2);
    else group_dots<LPR, DPL, true>(a, a2, b, ld, nk, sub, r, r2);
}
// acc[e] (+)= sum_j w_j * src_j[e]  with w_j taken from lane (group base + j) of `wreg`
template <int LPR, int DPL>
__device__ __forceinline__ void group_weighted_rows(float (&acc)[DPL], float wreg, int gbase, const float* __restrict__ src, int ld, int N) {
    for (int j = 0; j < N; ++j) {
        const float w = __shfl_sync(0xffffffffu, wreg, gbase + j);
        if (DPL == 4) {
            const float4 v = *reinterpret_cast<const float4*>(src + j * ld);
            acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]); acc[2 % DPL] = fmaf(w, v.z, acc[2 % DPL]); acc[3 % DPL] = fmaf(w, v.w, acc[3 % DPL]);
        } else {
            const float2 v = *reinterpret_cast<const float2*>(src + j * ld);
            acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
        }
    }
}
template <int DPL, int CS>
__device__ __forceinline__ void can_store_group(float* hi, float* lo, int row, int sub, const float (&v)[DPL]) {
    if (DPL == 4) can_store4<CS>(hi, lo, row, sub, make_float4(v[0], v[1], v[2 % DPL], v[3 % DPL]));
    else can_store2<CS>(hi, lo, row, 2 * sub, v[0], v[1]);
}

// Forward for head chunk hc: p_u. = softmax_j(s q_u . k'_j), o_u = sum_j p_uj v'_j - A x_u + c -> canonical slot; p -> stash.
template <class C>
__device__ __forceinline__ void attn_forward_rows(Ctx2& c, const LayerDev& W, int hc, int N, int NP, float* st_p) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gbase = lane - sub;
    const int rows = c.rows_act;
    float cc[DPL], ax[DPL][3];
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
        const int col = hc * 64 + sub * DPL + e;
        cc[e] = __ldg(W.cvec + col);
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + col * 4));
        ax[e][0] = a4.x; ax[e][1] = a4.y; ax[e][2] = a4.z;
    }
    for (int base = warp * AM::UPW; base < rows; base += kCW * AM::UPW) {
        const int u0 = base + lane / LPR;
        const bool valid = u0 < rows;
        const int u = valid ? u0 : rows - 1;
        const int r0 = (u / N) * N;
        const bool act = sub < N;
        float dot, dot_unused;
        group_dots<LPR, DPL, false>(c.sQKV + u * C::LDQ, nullptr, c.sQKV + r0 * C::LDQ + 64, C::LDQ, N, sub, dot, dot_unused);
        const float lg = act ? kAttnScale * dot : -INFINITY;
        const float m = group_max<LPR>(lg);
        const float e = fast_exp(lg - m);              // exp(-inf) = 0 for the inactive keys
        const float p = e * __frcp_rn(group_sum<LPR>(e));
        if (valid && sub < NP) st_p[(size_t)u * NP + sub] = p;
        float acc[DPL];
#pragma unroll
        for (int e2 = 0; e2 < DPL; ++e2) acc[e2] = 0.f;
        group_weighted_rows<LPR, DPL>(acc, p, gbase, c.sQKV + r0 * C::LDQ + 128 + sub * DPL, C::LDQ, N);
        const float x0 = c.sX[u * 4], x1 = c.sX[u * 4 + 1], x2 = c.sX[u * 4 + 2];
#pragma unroll
        for (int e2 = 0; e2 < DPL; ++e2) acc[e2] += cc[e2] - (ax[e2][0] * x0 + ax[e2][1] * x1 + ax[e2][2] * x2);
        c.slot_acquire();      // the previous out-projection has had the whole row computation to finish reading the slot
        if (valid) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u, sub, acc);
    }
}

// Reverse pass A for query row u: dp_uj = do_u . v'_j, ds_uj = p_uj (dp_uj - sum_j p_uj dp_uj) -> sDS;
// layers > 0 also dq_u = s sum_j ds_uj k'_j -> canonical slot.
template <class C>
__device__ __forceinline__ void attn_backward_ds_dq(Ctx2& c, int N, int NP, bool want_dq) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gbase = lane - sub;
    const int rows = c.rows_act;
    for (int base = warp * AM::UPW; base < rows; base += kCW * AM::UPW) {
        const int u0 = base + lane / LPR;
        const bool valid = u0 < rows;
        const int u = valid ? u0 : rows - 1;
        const int r0 = (u / N) * N;
        const bool act = sub < N;
        float dp, dp_unused;
        group_dots<LPR, DPL, false>(c.sO + u * C::LDO, nullptr, c.sQKV + r0 * C::LDQ + 128, C::LDQ, N, sub, dp, dp_unused);
        const float p = act ? c.sP[u * NP + sub] : 0.f;
        const float ds = p * (dp - group_sum<LPR>(p * dp));
        if (valid && sub < NP) c.sDS[u * NP + sub] = ds;
        if (want_dq) {
            float acc[DPL];
#pragma unroll
            for (int e = 0; e < DPL; ++e) acc[e] = 0.f;
            group_weighted_rows<LPR, DPL>(acc, ds, gbase, c.sQKV + r0 * C::LDQ + 64 + sub * DPL, C::LDQ, N);
#pragma unroll
            for (int e = 0; e < DPL; ++e) acc[e] *= kAttnScale;
            c.slot_acquire();
            if (valid) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u, sub, acc);
        }
    }
}
// Reverse passes B / C for key row u (= s*N + j), lanes over output columns:
//   dk'_u = s sum_i ds_iu q_i   -> slot (job d k');   dv'_u = sum_i p_iu do_i   -> slot (job d v');
//   dx_u += A_h^T (dk'_u + dv'_u - do_u)
// Both products are formed in registers first, so that the single rotating slot is only waited for when its previous
// job (d q, then d k') has had a whole computation phase to complete.
template <class C>
__device__ __forceinline__ void attn_backward_dkv(Ctx2& c, const LayerDev& W, int hc, int N, int NP, bool to_slot) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL;
    constexpr int MR = (C::kR + kCW * AM::UPW - 1) / (kCW * AM::UPW);          // rounds a warp can have
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR;
    const int rows = c.rows_act;
    float ax[DPL][3];
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + (hc * 64 + sub * DPL + e) * 4));
        ax[e][0] = a4.x; ax[e][1] = a4.y; ax[e][2] = a4.z;
    }
    float dk[MR][DPL], dv[MR][DPL];
#pragma unroll
    for (int rd = 0; rd < MR; ++rd) {
        const int base = (warp + rd * kCW) * AM::UPW;
#pragma unroll
        for (int e = 0; e < DPL; ++e) { dk[rd][e] = 0.f; dv[rd][e] = 0.f; }
        if (base < rows) {
            const int u = min(base + lane / LPR, rows - 1);
            const int r0 = (u / N) * N, jj = u - r0;
            const float* wk = c.sDS + r0 * NP + jj;
            const float* wv = c.sP + r0 * NP + jj;
            const float* qs = c.sQKV + r0 * C::LDQ + sub * DPL;
            const float* os = c.sO + r0 * C::LDO + sub * DPL;
            for (int i = 0; i < N; ++i) {
                const float a = wk[i * NP], b = wv[i * NP];
                if (DPL == 4) {
                    const float4 q = *reinterpret_cast<const float4*>(qs + i * C::LDQ), o = *reinterpret_cast<const float4*>(os + i * C::LDO);
                    dk[rd][0] = fmaf(a, q.x, dk[rd][0]); dk[rd][1] = fmaf(a, q.y, dk[rd][1]);
                    dk[rd][2 % DPL] = fmaf(a, q.z, dk[rd][2 % DPL]); dk[rd][3 % DPL] = fmaf(a, q.w, dk[rd][3 % DPL]);
                    dv[rd][0] = fmaf(b, o.x, dv[rd][0]); dv[rd][1] = fmaf(b, o.y, dv[rd][1]);
                    dv[rd][2 % DPL] = fmaf(b, o.z, dv[rd][2 % DPL]); dv[rd][3 % DPL] = fmaf(b, o.w, dv[rd][3 % DPL]);
                } else {
                    const float2 q = *reinterpret_cast<const float2*>(qs + i * C::LDQ), o = *reinterpret_cast<const float2*>(os + i * C::LDO);
                    dk[rd][0] = fmaf(a, q.x, dk[rd][0]); dk[rd][1] = fmaf(a, q.y, dk[rd][1]);
                    dv[rd][0] = fmaf(b, o.x, dv[rd][0]); dv[rd][1] = fmaf(b, o.y, dv[rd][1]);
                }
            }
#pragma unroll
            for (int e = 0; e < DPL; ++e) dk[rd][e] *= kAttnScale;
        }
    }
    if (to_slot) {          // job d k'
        c.slot_acquire();
#pragma unroll
        for (int rd = 0; rd < MR; ++rd) {
            const int u0 = (warp + rd * kCW) * AM::UPW + lane / LPR;
            if (u0 < rows) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u0, sub, dk[rd]);
        }
        c.slot_post();
    }
    // dx_u += A_h^T (dk'_u + dv'_u - do_u): runs while the tensor core consumes d k'
#pragma unroll
    for (int rd = 0; rd < MR; ++rd) {
        const int base = (warp + rd * kCW) * AM::UPW;
        if (base < rows) {
            const int u0 = base + lane / LPR;
            const int u = min(u0, rows - 1);
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll
            for (int e = 0; e < DPL; ++e) {
                const float t = dk[rd][e] + dv[rd][e] - c.sO[u * C::LDO + sub * DPL + e];
                g0 = fmaf(ax[e][0], t, g0); g1 = fmaf(ax[e][1], t, g1); g2 = fmaf(ax[e][2], t, g2);
            }
            g0 = group_sum<LPR>(g0); g1 = group_sum<LPR>(g1); g2 = group_sum<LPR>(g2);
            if (u0 < rows && sub == 0) { c.sDX[u * 4] += g0; c.sDX[u * 4 + 1] += g1; c.sDX[u * 4 + 2] += g2; }
        }
    }
    if (to_slot) {          // job d v'
        c.slot_acquire();
#pragma unroll
        for (int rd = 0; rd < MR; ++rd) {
            const int u0 = (warp + rd * kCW) * AM::UPW + lane / LPR;
            if (u0 < rows) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u0, sub, dv[rd]);
        }
        c.slot_post();
    } else {
        csync();
    }
}

// ------------------------------------------------------------------ pair-local attention (large groups)
// Same lane mapping as the row-local routines above, but a lane group owns TWO consecutive node rows of a sample: the
// k' / v' / q / d o rows that both need are read from shared memory once.  The attention phases of a full 64-row
// group are bound by the shared-memory pipe (every LDS.128 of a warp is four wavefronts), so this is ~1.5x faster
// there; for small groups (rows <= lane groups) the row-local form has the shorter dependent chain and is kept.
struct PairUnit { int r0, ia, ib; bool valid, has_b; };
template <class C>
__device__ __forceinline__ PairUnit pair_unit(int unit, int n_units, int pps, int N) {
    PairUnit u;
    u.valid = unit < n_units;
    const int uc = u.valid ? unit : n_units - 1;
    const int s = uc / pps, pi = uc - s * pps;
    u.r0 = s * N; u.ia = 2 * pi; u.has_b = u.ia + 1 < N; u.ib = u.has_b ? u.ia + 1 : u.ia;
    return u;
}
template <int LPR, int DPL>
__device__ __forceinline__ void load_cols(float (&v)[DPL], const float* __restrict__ src) {
    if (DPL == 4) {
        const float4 t = *reinterpret_cast<const float4*>(src);
        v[0] = t.x; v[1] = t.y; v[2 % DPL] = t.z; v[3 % DPL] = t.w;
    } else {
        const float2 t = *reinterpret_cast<const float2*>(src);
        v[0] = t.x; v[1] = t.y;
    }
}

template <class C>
__device__ __forceinline__ void attn_forward_pairs(Ctx2& c, const LayerDev& W, int hc, int N, int NP, float* st_p) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL, NG = kCW * AM::UPW, KPL = AM::KPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gbase = lane - sub, gid = warp * AM::UPW + lane / LPR;
    const int pps = (N + 1) >> 1, n_units = c.S_act * pps;
    float cc[DPL], ax[DPL][3];
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
        const int col = hc * 64 + sub * DPL + e;
        cc[e] = __ldg(W.cvec + col);
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + col * 4));
        ax[e][0] = a4.x; ax[e][1] = a4.y; ax[e][2] = a4.z;
    }
    for (int base = 0; base < n_units; base += NG) {
        const PairUnit u = pair_unit<C>(base + gid, n_units, pps, N);
        const int ra = u.r0 + u.ia, rb = u.r0 + u.ib;
        float la[KPL], lb[KPL];
        if (KPL == 2 && DFF_TC_LANE_DOTS4) {
            float r4[4];
            lane_dots4(c.sQKV + ra * C::LDQ, c.sQKV + rb * C::LDQ, c.sQKV + (u.r0 + min(sub, N - 1)) * C::LDQ + 64,
                       c.sQKV + (u.r0 + min(LPR + sub, N - 1)) * C::LDQ + 64, r4);
            la[0] = sub < N ? kAttnScale * r4[0] : -INFINITY; lb[0] = sub < N ? kAttnScale * r4[2] : -INFINITY;
            la[KPL - 1] = LPR + sub < N ? kAttnScale * r4[1] : -INFINITY; lb[KPL - 1] = LPR + sub < N ? kAttnScale * r4[3] : -INFINITY;
        } else {
#pragma unroll
        for (int kp = 0; kp < KPL; ++kp) {               // keys kp * LPR + sub
            const int kb = kp * LPR, nk = min(LPR, N - kb);
            float da, db;
            pair_dots<LPR, DPL, KPL>(c.sQKV + ra * C::LDQ, c.sQKV + rb * C::LDQ, c.sQKV + (u.r0 + kb) * C::LDQ + 64, C::LDQ, nk, sub, da, db);
            const bool act = sub < nk;
            la[kp] = act ? kAttnScale * da : -INFINITY;
            lb[kp] = act ? kAttnScale * db : -INFINITY;
        }
        }
        float ma = la[0], mb = lb[0];
#pragma unroll
        for (int kp = 1; kp < KPL; ++kp) { ma = fmaxf(ma, la[kp]); mb = fmaxf(mb, lb[kp]); }
        ma = group_max<LPR>(ma); mb = group_max<LPR>(mb);
        float pa[KPL], pb[KPL], sa = 0.f, sb = 0.f;
#pragma unroll
        for (int kp = 0; kp < KPL; ++kp) {
            pa[kp] = fast_exp(la[kp] - ma); pb[kp] = fast_exp(lb[kp] - mb);        // exp(-inf) = 0 for the inactive keys
            sa += pa[kp]; sb += pb[kp];
        }
        sa = __frcp_rn(group_sum<LPR>(sa)); sb = __frcp_rn(group_sum<LPR>(sb));
#pragma unroll
        for (int kp = 0; kp < KPL; ++kp) {
            pa[kp] *= sa; pb[kp] *= sb;
            const int key = kp * LPR + sub;
            if (u.valid && key < NP) {
                st_p[(size_t)ra * NP + key] = pa[kp];
                if (u.has_b) st_p[(size_t)rb * NP + key] = pb[kp];
            }
        }
        float oa[DPL], ob[DPL];
#pragma unroll
        for (int e = 0; e < DPL; ++e) { oa[e] = 0.f; ob[e] = 0.f; }
#pragma unroll
        for (int kp = 0; kp < KPL; ++kp) {
            const int kb = kp * LPR, nk = min(LPR, N - kb);
            const float* vs = c.sQKV + (u.r0 + kb) * C::LDQ + 128 + sub * DPL;
            for (int j = 0; j < nk; ++j) {
                float v[DPL];
                load_cols<LPR, DPL>(v, vs + j * C::LDQ);
                const float wa = __shfl_sync(0xffffffffu, pa[kp], gbase + j), wb = __shfl_sync(0xffffffffu, pb[kp], gbase + j);
#pragma unroll
                for (int e = 0; e < DPL; ++e) { oa[e] = fmaf(wa, v[e], oa[e]); ob[e] = fmaf(wb, v[e], ob[e]); }
            }
        }
        {
            const float x0 = c.sX[ra * 4], x1 = c.sX[ra * 4 + 1], x2 = c.sX[ra * 4 + 2];
            const float y0 = c.sX[rb * 4], y1 = c.sX[rb * 4 + 1], y2 = c.sX[rb * 4 + 2];
#pragma unroll
            for (int e = 0; e < DPL; ++e) {
                oa[e] += cc[e] - (ax[e][0] * x0 + ax[e][1] * x1 + ax[e][2] * x2);
                ob[e] += cc[e] - (ax[e][0] * y0 + ax[e][1] * y1 + ax[e][2] * y2);
            }
        }
        c.slot_acquire();
        if (u.valid) {
            can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, ra, sub, oa);
            if (u.has_b) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, rb, sub, ob);
        }
    }
}

template <class C>
__device__ __forceinline__ void attn_backward_ds_dq_pairs(Ctx2& c, int N, int NP, bool want_dq) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL, NG = kCW * AM::UPW, KPL = AM::KPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gbase = lane - sub, gid = warp * AM::UPW + lane / LPR;
    const int pps = (N + 1) >> 1, n_units = c.S_act * pps;
    for (int base = 0; base < n_units; base += NG) {
        const PairUnit u = pair_unit<C>(base + gid, n_units, pps, N);
        const int ra = u.r0 + u.ia, rb = u.r0 + u.ib;
        float pa[KPL], pb[KPL], dpa[KPL], dpb[KPL], ta = 0.f, tb = 0.f;
        if (KPL == 2 && DFF_TC_LANE_DOTS4) {
            float r4[4];
            lane_dots4(c.sO + ra * C::LDO, c.sO + rb * C::LDO, c.sQKV + (u.r0 + min(sub, N - 1)) * C::LDQ + 128,
                       c.sQKV + (u.r0 + min(LPR + sub, N - 1)) * C::LDQ + 128, r4);
            dpa[0] = r4[0]; dpa[KPL - 1] = r4[1]; dpb[0] = r4[2]; dpb[KPL - 1] = r4[3];
        }
#pragma unroll
        for (int kp = 0; kp < KPL; ++kp) {
            const int kb = kp * LPR, nk = min(LPR, N - kb);
            if (!(KPL == 2 && DFF_TC_LANE_DOTS4))
                pair_dots<LPR, DPL, KPL>(c.sO + ra * C::LDO, c.sO + rb * C::LDO, c.sQKV + (u.r0 + kb) * C::LDQ + 128, C::LDQ, nk, sub, dpa[kp], dpb[kp]);
            const bool act = sub < nk;
            pa[kp] = act ? c.sP[ra * NP + kb + sub] : 0.f;
            pb[kp] = act ? c.sP[rb * NP + kb + sub] : 0.f;
            ta += pa[kp] * dpa[kp]; tb += pb[kp] * dpb[kp];
        }
        ta = group_sum<LPR>(ta); tb = group_sum<LPR>(tb);
        float dsa[KPL], dsb[KPL];
#pragma unroll
        for (int kp = 0; kp < KPL; ++kp) {
            dsa[kp] = pa[kp] * (dpa[kp] - ta); dsb[kp] = pb[kp] * (dpb[kp] - tb);
            const int key = kp * LPR + sub;
            if (u.valid && key < NP) {
                c.sDS[ra * NP + key] = dsa[kp];
                if (u.has_b) c.sDS[rb * NP + key] = dsb[kp];
            }
        }
        if (want_dq) {
            float qa[DPL], qb[DPL];
#pragma unroll
            for (int e = 0; e < DPL; ++e) { qa[e] = 0.f; qb[e] = 0.f; }
#pragma unroll
            for (int kp = 0; kp < KPL; ++kp) {
                const int kb = kp * LPR, nk = min(LPR, N - kb);
                const float* ks = c.sQKV + (u.r0 + kb) * C::LDQ + 64 + sub * DPL;
                for (int j = 0; j < nk; ++j) {
                    float v[DPL];
                    load_cols<LPR, DPL>(v, ks + j * C::LDQ);
                    const float wa = __shfl_sync(0xffffffffu, dsa[kp], gbase + j), wb = __shfl_sync(0xffffffffu, dsb[kp], gbase + j);
#pragma unroll
                    for (int e = 0; e < DPL; ++e) { qa[e] = fmaf(wa, v[e], qa[e]); qb[e] = fmaf(wb, v[e], qb[e]); }
                }
            }
#pragma unroll
            for (int e = 0; e < DPL; ++e) { qa[e] *= kAttnScale; qb[e] *= kAttnScale; }
            c.slot_acquire();
            if (u.valid) {
                can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, ra, sub, qa);
                if (u.has_b) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, rb, sub, qb);
            }
        }
    }
}

template <class C>
__device__ __forceinline__ void attn_backward_dkv_pairs(Ctx2& c, const LayerDev& W, int hc, int N, int NP, bool to_slot) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL, NG = kCW * AM::UPW;
    constexpr int MR = 2;                                    // rounds: pairs per CTA <= 2 * lane groups (checked by the host)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gid = warp * AM::UPW + lane / LPR;
    const int pps = (N + 1) >> 1, n_units = c.S_act * pps;
    float ax[DPL][3];
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + (hc * 64 + sub * DPL + e) * 4));
        ax[e][0] = a4.x; ax[e][1] = a4.y; ax[e][2] = a4.z;
    }
    float dk[MR][2][DPL], dv[MR][2][DPL];
#pragma unroll
    for (int rd = 0; rd < MR; ++rd) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int e = 0; e < DPL; ++e) { dk[rd][h][e] = 0.f; dv[rd][h][e] = 0.f; }
        if (rd * NG < n_units) {
            const PairUnit u = pair_unit<C>(rd * NG + gid, n_units, pps, N);
            // key rows ja = u.ia, jb = u.ib of the sample; the weights of both sit next to each other (ia is even)
            const float* wk = c.sDS + u.r0 * NP + u.ia;
            const float* wv = c.sP + u.r0 * NP + u.ia;
            const float* qs = c.sQKV + u.r0 * C::LDQ + sub * DPL;
            const float* os = c.sO + u.r0 * C::LDO + sub * DPL;
            for (int i = 0; i < N; ++i) {
                const float2 a = *reinterpret_cast<const float2*>(wk + i * NP), b = *reinterpret_cast<const float2*>(wv + i * NP);
                float q[DPL], o[DPL];
                load_cols<LPR, DPL>(q, qs + i * C::LDQ);
                load_cols<LPR, DPL>(o, os + i * C::LDO);
#pragma unroll
                for (int e = 0; e < DPL; ++e) {
                    dk[rd][0][e] = fmaf(a.x, q[e], dk[rd][0][e]); dk[rd][1][e] = fmaf(a.y, q[e], dk[rd][1][e]);
                    dv[rd][0][e] = fmaf(b.x, o[e], dv[rd][0][e]); dv[rd][1][e] = fmaf(b.y, o[e], dv[rd][1][e]);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int e = 0; e < DPL; ++e) dk[rd][h][e] *= kAttnScale;
        }
    }
    if (to_slot) {          // job d k'
        c.slot_acquire();
#pragma unroll
        for (int rd = 0; rd < MR; ++rd)
            if (rd * NG + gid < n_units) {
                const PairUnit u = pair_unit<C>(rd * NG + gid, n_units, pps, N);
                can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u.r0 + u.ia, sub, dk[rd][0]);
                if (u.has_b) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u.r0 + u.ib, sub, dk[rd][1]);
            }
        c.slot_post();
    }
    // dx_j += A_h^T (dk'_j + dv'_j - do_j): runs while the tensor core consumes d k'
#pragma unroll
    for (int rd = 0; rd < MR; ++rd)
        if (rd * NG < n_units) {
            const PairUnit u = pair_unit<C>(rd * NG + gid, n_units, pps, N);
            float g[2][3];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int row = u.r0 + (h ? u.ib : u.ia);
                g[h][0] = g[h][1] = g[h][2] = 0.f;
#pragma unroll
                for (int e = 0; e < DPL; ++e) {
                    const float t = dk[rd][h][e] + dv[rd][h][e] - c.sO[row * C::LDO + sub * DPL + e];
                    g[h][0] = fmaf(ax[e][0], t, g[h][0]); g[h][1] = fmaf(ax[e][1], t, g[h][1]); g[h][2] = fmaf(ax[e][2], t, g[h][2]);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) { g[h][0] = group_sum<LPR>(g[h][0]); g[h][1] = group_sum<LPR>(g[h][1]); g[h][2] = group_sum<LPR>(g[h][2]); }
            if (u.valid && sub == 0) {
                const int ra = u.r0 + u.ia, rb = u.r0 + u.ib;
                c.sDX[ra * 4] += g[0][0]; c.sDX[ra * 4 + 1] += g[0][1]; c.sDX[ra * 4 + 2] += g[0][2];
                if (u.has_b) { c.sDX[rb * 4] += g[1][0]; c.sDX[rb * 4 + 1] += g[1][1]; c.sDX[rb * 4 + 2] += g[1][2]; }
            }
        }
    if (to_slot) {          // job d v'
        c.slot_acquire();
#pragma unroll
        for (int rd = 0; rd < MR; ++rd)
            if (rd * NG + gid < n_units) {
                const PairUnit u = pair_unit<C>(rd * NG + gid, n_units, pps, N);
                can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u.r0 + u.ia, sub, dv[rd][0]);
                if (u.has_b) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, u.r0 + u.ib, sub, dv[rd][1]);
            }
        c.slot_post();
    } else {
        csync();
    }
}

// ------------------------------------------------------------------ query-row QUADS for full-warp lane groups (N > 12)
// A lane group (a warp) owns four consecutive query rows: the two k' (or v') rows of the lane are read once per four
// rows (4 x 2 register tile: 96 LDS.128 for 8 dot products) and every v' / k' column slice once per four rows in the
// P V' / dq accumulation.  14 quads of a 56-bead sample fit the 16 warps in one round.
template <int KPL>
__device__ __forceinline__ void lane_dots8(const float* __restrict__ a, int lda, const float* __restrict__ b0, const float* __restrict__ b1,
                                           float (&r)[4][2]) {
    float s[4][2], t[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) { s[i][0] = s[i][1] = t[i][0] = t[i][1] = 0.f; }
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
        const float4 y = *reinterpret_cast<const float4*>(b0 + 4 * k);
        const float4 w = (KPL == 2) ? *reinterpret_cast<const float4*>(b1 + 4 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(a + i * lda + 4 * k);
            s[i][0] = fmaf(x.x, y.x, s[i][0]); t[i][0] = fmaf(x.y, y.y, t[i][0]); s[i][0] = fmaf(x.z, y.z, s[i][0]); t[i][0] = fmaf(x.w, y.w, t[i][0]);
            if (KPL == 2) { s[i][1] = fmaf(x.x, w.x, s[i][1]); t[i][1] = fmaf(x.y, w.y, t[i][1]); s[i][1] = fmaf(x.z, w.z, s[i][1]); t[i][1] = fmaf(x.w, w.w, t[i][1]); }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { r[i][0] = s[i][0] + t[i][0]; r[i][1] = s[i][1] + t[i][1]; }
}

struct QuadUnit { int r0, i0, cnt; bool valid; };
__device__ __forceinline__ QuadUnit quad_unit(int gid, int S_act, int N) {
    const int qps = (N + 3) >> 2, n_units = S_act * qps;
    QuadUnit u;
    u.valid = gid < n_units;
    const int uc = u.valid ? gid : n_units - 1;
    const int s = uc / qps;
    u.r0 = s * N; u.i0 = (uc - s * qps) * 4; u.cnt = u.valid ? min(4, N - u.i0) : 0;
    return u;
}

template <class C>
__device__ __forceinline__ void attn_forward_quads(Ctx2& c, const LayerDev& W, int hc, int N, int NP, float* st_p) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gbase = lane - sub;
    const QuadUnit u = quad_unit(warp * AM::UPW + lane / LPR, c.S_act, N);
    float cc[DPL], ax[DPL][3];
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
        const int col = hc * 64 + sub * DPL + e;
        cc[e] = __ldg(W.cvec + col);
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + col * 4));
        ax[e][0] = a4.x; ax[e][1] = a4.y; ax[e][2] = a4.z;
    }
    // rows of the quad (clamped for the padded tail); all four are read with one base pointer and stride LDQ
    c.mark(28);
    const int rbase = u.r0 + min(u.i0, max(N - 4, 0));               // keep the 4 rows inside the sample: shift the window back
    const int shift = u.i0 - (rbase - u.r0);                          // rows [shift, 4) of the window are this quad's rows
    float d[4][2];
    lane_dots8<AM::KPL>(c.sQKV + rbase * C::LDQ, C::LDQ, c.sQKV + (u.r0 + min(sub, N - 1)) * C::LDQ + 64,
               c.sQKV + (u.r0 + min(LPR + sub, N - 1)) * C::LDQ + 64, d);
    c.mark(29);
    const bool act0 = sub < N, act1 = LPR + sub < N;
    // softmax of the four rows: branch-free (exp(-inf) = 0 masks the inactive keys) with the four reductions interleaved
    float p[4][2], mx[4], ssum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        p[i][0] = act0 ? kAttnScale * d[i][0] : -INFINITY; p[i][1] = act1 ? kAttnScale * d[i][1] : -INFINITY;
        mx[i] = fmaxf(p[i][0], p[i][1]);
    }
    group_max_n<LPR, 4>(mx);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        p[i][0] = fast_exp(p[i][0] - mx[i]); p[i][1] = fast_exp(p[i][1] - mx[i]);
        ssum[i] = p[i][0] + p[i][1];
    }
    group_sum_n<LPR, 4>(ssum);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float inv = __frcp_rn(ssum[i]);
        p[i][0] *= inv; p[i][1] *= inv;
        const int row = rbase + i;
        if (u.valid && i >= shift && i - shift < u.cnt) {
            if (sub < NP) st_p[(size_t)row * NP + sub] = p[i][0];
            if (LPR + sub < NP) st_p[(size_t)row * NP + LPR + sub] = p[i][1];
        }
    }
    c.mark(30);
    float o[4][DPL];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < DPL; ++e) o[i][e] = 0.f;
#pragma unroll
    for (int kp = 0; kp < AM::KPL; ++kp) {
        const int kb = kp * LPR, nk = min(LPR, N - kb);
        const float* vs = c.sQKV + (u.r0 + kb) * C::LDQ + 128 + sub * DPL;
        for (int j = 0; j < nk; ++j) {
            float v[DPL];
            load_cols<LPR, DPL>(v, vs + j * C::LDQ);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w = __shfl_sync(0xffffffffu, p[i][kp], gbase + j);
#pragma unroll
                for (int e = 0; e < DPL; ++e) o[i][e] = fmaf(w, v[e], o[i][e]);
            }
        }
    }
    c.mark(31);
    c.slot_acquire();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = rbase + i;
        const float x0 = c.sX[row * 4], x1 = c.sX[row * 4 + 1], x2 = c.sX[row * 4 + 2];
#pragma unroll
        for (int e = 0; e < DPL; ++e) o[i][e] += cc[e] - (ax[e][0] * x0 + ax[e][1] * x1 + ax[e][2] * x2);
        if (u.valid && i >= shift && i - shift < u.cnt) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, row, sub, o[i]);
    }
}

template <class C>
__device__ __forceinline__ void attn_backward_ds_dq_quads(Ctx2& c, int N, int NP, bool want_dq) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gbase = lane - sub;
    const QuadUnit u = quad_unit(warp * AM::UPW + lane / LPR, c.S_act, N);
    const int rbase = u.r0 + min(u.i0, max(N - 4, 0));
    const int shift = u.i0 - (rbase - u.r0);
    float d[4][2];
    lane_dots8<AM::KPL>(c.sO + rbase * C::LDO, C::LDO, c.sQKV + (u.r0 + min(sub, N - 1)) * C::LDQ + 128,
               c.sQKV + (u.r0 + min(LPR + sub, N - 1)) * C::LDQ + 128, d);
    c.mark(23);
    const bool act0 = sub < N, act1 = LPR + sub < N;
    float ds[4][2], tsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = rbase + i;
        ds[i][0] = act0 ? c.sP[row * NP + sub] : 0.f; ds[i][1] = act1 ? c.sP[row * NP + LPR + sub] : 0.f;     // p
        tsum[i] = ds[i][0] * d[i][0] + ds[i][1] * d[i][1];
    }
    group_sum_n<LPR, 4>(tsum);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = rbase + i;
        ds[i][0] *= d[i][0] - tsum[i]; ds[i][1] *= d[i][1] - tsum[i];
        if (u.valid && i >= shift && i - shift < u.cnt) {
            if (sub < NP) c.sDS[row * NP + sub] = ds[i][0];
            if (LPR + sub < NP) c.sDS[row * NP + LPR + sub] = ds[i][1];
        }
    }
    c.mark(24);
    if (want_dq) {
        float q[4][DPL];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int e = 0; e < DPL; ++e) q[i][e] = 0.f;
#pragma unroll
        for (int kp = 0; kp < AM::KPL; ++kp) {
            const int kb = kp * LPR, nk = min(LPR, N - kb);
            const float* ks = c.sQKV + (u.r0 + kb) * C::LDQ + 64 + sub * DPL;
            for (int j = 0; j < nk; ++j) {
                float v[DPL];
                load_cols<LPR, DPL>(v, ks + j * C::LDQ);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float w = __shfl_sync(0xffffffffu, ds[i][kp], gbase + j);
#pragma unroll
                    for (int e = 0; e < DPL; ++e) q[i][e] = fmaf(w, v[e], q[i][e]);
                }
            }
        }
        c.mark(25);
        c.slot_acquire();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int e = 0; e < DPL; ++e) q[i][e] *= kAttnScale;
            if (u.valid && i >= shift && i - shift < u.cnt) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, rbase + i, sub, q[i]);
        }
    }
}

// Key-row QUADS for N > 32 (one sample per pass): a lane group owns four consecutive key rows, so the q_i / d o_i column
// slices and the ds / p weights (one aligned float4 each) are read once per four rows; 14 quads of a 56-bead sample fit the
// 16 lane groups in a single round (28 pairs need two).  Same outputs as attn_backward_dkv_pairs.
template <class C>
__device__ __forceinline__ void attn_backward_dkv_quads(Ctx2& c, const LayerDev& W, int hc, int N, int NP, bool to_slot) {
    using AM = AttnMap<C>;
    constexpr int LPR = AM::LPR, DPL = AM::DPL, NG = kCW * AM::UPW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane % LPR, gid = warp * AM::UPW + lane / LPR;
    const int qps = (N + 3) >> 2, n_units = c.S_act * qps;           // quads per sample, per pass (<= NG: checked by the caller)
    float ax[DPL][3];
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(W.A + (hc * 64 + sub * DPL + e) * 4));
        ax[e][0] = a4.x; ax[e][1] = a4.y; ax[e][2] = a4.z;
    }
    const bool valid = gid < n_units;
    const int uc = valid ? gid : n_units - 1;
    const int s = uc / qps, r0 = s * N, j0 = (uc - s * qps) * 4;
    const int cnt = valid ? min(4, N - j0) : 0;
    float dk[4][DPL], dv[4][DPL];
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
        for (int e = 0; e < DPL; ++e) { dk[h][e] = 0.f; dv[h][e] = 0.f; }
    {
        const float* wk = c.sDS + r0 * NP + j0;                       // j0 is a multiple of 4 and NP too: aligned float4 (pad columns are zero)
        const float* wv = c.sP + r0 * NP + j0;
        const float* qs = c.sQKV + r0 * C::LDQ + sub * DPL;
        const float* os = c.sO + r0 * C::LDO + sub * DPL;
        for (int i = 0; i < N; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(wk + i * NP), b = *reinterpret_cast<const float4*>(wv + i * NP);
            float q[DPL], o[DPL];
            load_cols<LPR, DPL>(q, qs + i * C::LDQ);
            load_cols<LPR, DPL>(o, os + i * C::LDO);
#pragma unroll
            for (int e = 0; e < DPL; ++e) {
                dk[0][e] = fmaf(a.x, q[e], dk[0][e]); dk[1][e] = fmaf(a.y, q[e], dk[1][e]); dk[2][e] = fmaf(a.z, q[e], dk[2][e]); dk[3][e] = fmaf(a.w, q[e], dk[3][e]);
                dv[0][e] = fmaf(b.x, o[e], dv[0][e]); dv[1][e] = fmaf(b.y, o[e], dv[1][e]); dv[2][e] = fmaf(b.z, o[e], dv[2][e]); dv[3][e] = fmaf(b.w, o[e], dv[3][e]);
            }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h)
#pragma unroll
            for (int e = 0; e < DPL; ++e) dk[h][e] *= kAttnScale;
    }
    c.mark(26);
    if (to_slot) {          // job d k'
        c.slot_acquire();
#pragma unroll
        for (int h = 0; h < 4; ++h)
            if (h < cnt) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, r0 + j0 + h, sub, dk[h]);
        c.slot_post();
    }
    c.mark(27);
    // dx_j += A_h^T (dk'_j + dv'_j - do_j): runs while the tensor core consumes d k'
    {
        float g[4][3];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const int row = r0 + min(j0 + h, N - 1);
            g[h][0] = g[h][1] = g[h][2] = 0.f;
#pragma unroll
            for (int e = 0; e < DPL; ++e) {
                const float t = dk[h][e] + dv[h][e] - c.sO[row * C::LDO + sub * DPL + e];
                g[h][0] = fmaf(ax[e][0], t, g[h][0]); g[h][1] = fmaf(ax[e][1], t, g[h][1]); g[h][2] = fmaf(ax[e][2], t, g[h][2]);
            }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) { g[h][0] = group_sum<LPR>(g[h][0]); g[h][1] = group_sum<LPR>(g[h][1]); g[h][2] = group_sum<LPR>(g[h][2]); }
        if (sub == 0) {
#pragma unroll
            for (int h = 0; h < 4; ++h)
                if (h < cnt) { const int row = r0 + j0 + h; c.sDX[row * 4] += g[h][0]; c.sDX[row * 4 + 1] += g[h][1]; c.sDX[row * 4 + 2] += g[h][2]; }
        }
    }
    if (to_slot) {          // job d v'
        c.slot_acquire();
#pragma unroll
        for (int h = 0; h < 4; ++h)
            if (h < cnt) can_store_group<DPL, C::kCS>(c.slot_hi, c.slot_lo, r0 + j0 + h, sub, dv[h]);
        c.slot_post();
    } else {
        csync();
    }
}


// GELU'(x) with the Gaussian pdf through fast_exp (the erf stays erff)
__device__ __forceinline__ float gelu_grad_tc(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * fast_exp(-0.5f * x * x) * 0.3989422804014327f;
}

// FF hidden block [rows][4H] TMEM -> shared (row stride LDF), so that the GELU phases can be spread over all threads
constexpr int kLDF = 256 + 4;

// ------------------------------------------------------------------ forward pass (compute warps)
template <class C>
__device__ void forward_pass_tc(const ModelDev& M, Ctx2& c, float t_norm, const float* __restrict__ t_rows) {
    constexpr int R = C::kR;
    const int tid = threadIdx.x;
    const int N = M.N, NP = M.NP, H = M.H;
    const int rows = c.rows_act;

    // layer-0 node stream: W_n [onehot_i, t] + b_n   (graph_transformer.py:99-103)
    for (int idx = tid; idx < rows * C::kHP; idx += kCT) {
        const int r = idx / C::kHP, d = idx - r * C::kHP;
        const float tn = t_rows != nullptr ? __ldg(t_rows + r / N) : t_norm;          // per-sample noise level (score mode) or the shared one
        float v = (d < H) ? __ldg(M.emb + (r % N) * H + d) + tn * __ldg(M.embt + d) : 0.f;
        if (M.abs_coords && d < H)        // node input [onehot_i, x_i, t] (graph_transformer.py:99-100)
            v += c.sX[r * 4] * __ldg(M.embx + d) + c.sX[r * 4 + 1] * __ldg(M.embx + H + d) + c.sX[r * 4 + 2] * __ldg(M.embx + 2 * H + d);
        c.sN[r * C::LDH + d] = v;
    }
    // augmented operand columns of this step: [x0 x1 x2 1] (chunk H/4); chunk H/4 + 1 stays zero
    if (tid < R) {
        const float4 xv = (tid < rows) ? make_float4(c.sX[tid * 4], c.sX[tid * 4 + 1], c.sX[tid * 4 + 2], 1.0f)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        can_store4<C::kCS>(c.nhat_hi, c.nhat_lo, tid, H / 4, xv);
    }
    csync();
    ln_forward_rows_can<C>(c.sN, c.nhat_hi, c.nhat_lo, M.layer[0].ln1_g, M.layer[0].ln1_b, H, rows, c.stash + M.off[ST_NIN],
                           c.stash + M.off[ST_STAT1]);
    c.post();
    c.mark(0);

    for (int l = 0; l < M.L; ++l) {
        const LayerDev& W = M.layer[l];
        float* st = c.stash + (size_t)l * M.layer_floats;

        for (int hc = 0; hc < C::NCH; ++hc) {
            {   // q | k' | v' of the head chunk: TMEM -> shared (row-major, for the attention) + stash
                const int b = c.dq_wait();
                c.mark(1);
                float* st_qkv = st + M.off[ST_QKV] + (size_t)hc * R * C::LDQ;
                tmem_foreach<192>(c.tmem, C::kColD + b * 192, rows, [&](int row, int col, const float (&v)[16]) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(c.sQKV + row * C::LDQ + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                });
                // stash copy for the reverse pass: ONE asynchronous bulk copy (shared -> global; the stash keeps the LDQ row stride),
                // issued after the hand-off.  It drains during the attention phase and keeps the load/store unit free:
                // register-staged stores of these 46 KB held up every global access of the attention phase.
                if (tid == 0) bulk_wait_read();        // p of the previous chunk has left sP
                fence_proxy_async();
                c.dq_release();
                if (tid == 0) bulk_s2g(st_qkv, c.sQKV, (uint32_t)(rows * C::LDQ) * sizeof(float));
                c.mark(2);
            }
            if constexpr (!C::kAttMma) {
                // logits, softmax, P V' - A x_i + c -> canonical operand of the out-projection (row-local, no barrier inside)
                // (p goes to sP; its stash copy is one bulk copy after the hand-off below)
                if (c.quads) attn_forward_quads<C>(c, W, hc, N, NP, c.sP);
                else if (c.pairs) attn_forward_pairs<C>(c, W, hc, N, NP, c.sP);
                else attn_forward_rows<C>(c, W, hc, N, NP, c.sP);
            } else {
                // logits (HMMA items) | softmax (rows) | P V' - A x_i + c (HMMA items) -> canonical operand of the out-projection
                const AttnGeo G(N, NP, c.S_act);
                const bool dist = M.edge_dist != 0;
                attn_logits_items<C>(c.sQKV, c.sP, W.A, hc, dist, G);
                csync();
                c.mark(23);
                attn_softmax_rows<C>(c.sP, c.sQKV, c.sX, c.sO + 64, C::LDO, dist, G);     // z -> sO pad (free in the forward pass)
                csync();
                c.mark(24);
                attn_weighted_items<C>(c.sP, c.sQKV, C::LDQ, 128, G, [&](int row, int col, float v0, float v1) {
                    const int gc = hc * 64 + col;
                    const float2 cv = __ldg(reinterpret_cast<const float2*>(W.cvec + gc));
                    const float4 e0 = __ldg(reinterpret_cast<const float4*>(W.A + gc * 4)), e1 = __ldg(reinterpret_cast<const float4*>(W.A + gc * 4 + 4));
                    const float x0 = c.sX[row * 4], x1 = c.sX[row * 4 + 1], x2 = c.sX[row * 4 + 2];
                    v0 += cv.x - (e0.x * x0 + e0.y * x1 + e0.z * x2);
                    v1 += cv.y - (e1.x * x0 + e1.y * x1 + e1.z * x2);
                    if (dist) { const float z = c.sO[row * C::LDO + 64]; v0 = fmaf(e0.w, z, v0); v1 = fmaf(e1.w, z, v1); }     // + a_h z_i
                    c.slot_acquire();      // the previous out-projection has had the whole head to finish reading the slot
                    can_store2<C::kCS>(c.slot_hi, c.slot_lo, row, col, v0, v1);
                });
            }
            c.mark(3);
            if (tid == 0) bulk_wait_read();        // q | k' | v' have left sQKV (the copy had the whole attention phase)
            c.slot_post();
            if (tid == 0) bulk_s2g(st + M.off[ST_P] + (size_t)hc * R * NP, c.sP, (uint32_t)(rows * NP) * sizeof(float));
            c.mark(4);
        }
        // attention block output -> row buffer; gated residual 1 + LayerNorm 2 -> canonical operand of FF1
        c.acc_wait();
        c.mark(5);
        tmem_foreach_sum<C::kHP, DFF_TC_SPLIT_ACC && C::kCorrAttn>(c.tmem, kColAcc, C::kCorrOut, rows, [&](int row, int col, const float (&v)[16]) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(W.bo + col + i));
                *reinterpret_cast<float4*>(c.sNh + row * C::LDH + col + i) = make_float4(v[i] + b.x, v[i + 1] + b.y, v[i + 2] + b.z, v[i + 3] + b.w);
            }
        });
        tc::fence_before_sync();
        csync();
        c.mark(6);
        gate_ln_forward_rows_can<C>(c.sN, c.sNh, c.nhat_hi, c.nhat_lo, W.g1a, W.g1b, H, rows, st + M.off[ST_ATT], st + M.off[ST_G1],
                                    st + M.off[ST_M], W.ln2_g, W.ln2_b, st + M.off[ST_STAT2]);
        c.post();
        c.mark(7);

        // feed-forward: hidden pre-activations TMEM -> shared, up to 256 columns ("super") at a time; then bias + GELU
        // 64 columns at a time, spread over all threads -> slot -> FF2 accumulates
        float* sH = c.sQKV;
        const int nch64 = (4 * H) / 64;
        for (int c0 = 0; c0 < nch64; c0 += 4) {
            const int nc = min(4, nch64 - c0);
            c.d1_wait();
            c.mark(8);
            tmem_foreach<256>(c.tmem, C::kColD, rows, [&](int row, int col, const float (&v)[16]) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(sH + row * kLDF + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            });
            tc::fence_before_sync();
            csync();
            if (c0 + 4 < nch64) c.post();            // TMEM work area drained: the issuer may start the next super's FF1
            c.mark(9);
            for (int ch = 0; ch < nc; ++ch) {
                constexpr int MG = (R * 16 + kCT - 1) / kCT;       // granules (row, 4 columns) per thread
                const int gc = (c0 + ch) * 64;           // hidden column of the chunk
                float4 gq[MG];
#pragma unroll
                for (int g = 0; g < MG; ++g) {
                    const int idx = tid + g * kCT;
                    if (idx < rows * 16) {
                        const int r = idx >> 4, k4 = idx & 15;
                        const float4 v = *reinterpret_cast<const float4*>(sH + r * kLDF + ch * 64 + k4 * 4);
                        const float4 b = __ldg(reinterpret_cast<const float4*>(W.b1 + gc + k4 * 4));
                        const float4 pre = make_float4(v.x + b.x, v.y + b.y, v.z + b.z, v.w + b.w);
                        *reinterpret_cast<float4*>(st + M.off[ST_H1] + (size_t)r * (4 * H) + gc + k4 * 4) = pre;
                        gq[g] = make_float4(gelu_f(pre.x), gelu_f(pre.y), gelu_f(pre.z), gelu_f(pre.w));
                    }
                }
                c.slot_acquire();
#pragma unroll
                for (int g = 0; g < MG; ++g) {
                    const int idx = tid + g * kCT;
                    if (idx < rows * 16) can_store4<C::kCS>(c.slot_hi, c.slot_lo, idx >> 4, idx & 15, gq[g]);
                }
                c.slot_post();
            }
        }
        c.mark(10);
        c.acc_wait();
        c.mark(5);
        tmem_foreach_sum<C::kHP, DFF_TC_SPLIT_ACC != 0>(c.tmem, kColAcc, C::kCorrFF, rows, [&](int row, int col, const float (&v)[16]) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(W.b2 + col + i));
                *reinterpret_cast<float4*>(c.sNh + row * C::LDH + col + i) = make_float4(v[i] + b.x, v[i + 1] + b.y, v[i + 2] + b.z, v[i + 3] + b.w);
            }
        });
        tc::fence_before_sync();
        csync();
        // gated residual 2 (+ next layer's LayerNorm 1; its input rows are the next layer's n_in stash)
        const bool last = (l + 1 == M.L);
        float* stn = st + M.layer_floats;
        gate_ln_forward_rows_can<C>(c.sN, c.sNh, c.nhat_hi, c.nhat_lo, W.g2a, W.g2b, H, rows, st + M.off[ST_FF], st + M.off[ST_G2],
                                    stn + M.off[ST_NIN], last ? nullptr : M.layer[l + 1].ln1_g,
                                    last ? nullptr : M.layer[l + 1].ln1_b, last ? nullptr : stn + M.off[ST_STAT1]);
        if (tid == 0) bulk_wait_read();            // the layer's last p copy has left sP
        if (!last) c.post(); else csync();
        c.mark(7);
    }
}

// ------------------------------------------------------------------ reverse pass: sDX[r][0..2] = d sum(E) / d x_r
template <class C>
__device__ void backward_pass_tc(const ModelDev& M, Ctx2& c) {
    constexpr int R = C::kR;
    const int tid = threadIdx.x, warp_id = threadIdx.x >> 5, lane_id = threadIdx.x & 31;
    const int N = M.N, NP = M.NP, H = M.H;
    const int rows = c.rows_act;

    for (int idx = tid; idx < rows * C::kHP; idx += kCT) {
        const int r = idx / C::kHP, d = idx - r * C::kHP;
        c.sN[r * C::LDH + d] = (d < H) ? __ldg(M.dec_w + d) : 0.f;      // dE_r/dn_r = w_dec  (node_decoder, :106)
    }
    for (int idx = tid; idx < R * 4; idx += kCT) c.sDX[idx] = 0.f;
    if (tid == 0) bulk_wait_all();                 // the forward pass's bulk stash copies are complete before anything is reloaded
    csync();

    for (int l = M.L - 1; l >= 0; --l) {
        const LayerDev& W = M.layer[l];
        float* st = c.stash + (size_t)l * M.layer_floats;
        const bool deep = l > 0 || M.abs_coords != 0;      // the layer's input nodes depend on x: dq / dk' / dv' feed d n_hat

        // gated residual 2 backward: d ff -> canonical operand, d m (partial) -> sN
        gate_backward_rows_can<C>(c.sN, nullptr, c.nhat_hi, c.nhat_lo, H, rows, nullptr, nullptr, nullptr, st + M.off[ST_FF],
                                  st + M.off[ST_M], st + M.off[ST_G2], W.g2a, W.g2b);
        c.post();
        c.mark(11);
        // feed-forward backward: d act TMEM -> shared, up to 256 columns at a time; times GELU'(pre) 64 columns at a time -> slot
        float* sH = c.sQKV;
        const int nch64 = (4 * H) / 64;
        for (int c0 = 0; c0 < nch64; c0 += 4) {
            const int nc = min(4, nch64 - c0);
            c.d1_wait();
            c.mark(8);
            tmem_foreach<256>(c.tmem, C::kColD, rows, [&](int row, int col, const float (&v)[16]) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(sH + row * kLDF + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            });
            tc::fence_before_sync();
            csync();
            if (c0 + 4 < nch64) c.post();
            c.mark(12);
            for (int ch = 0; ch < nc; ++ch) {
                constexpr int MG = (R * 16 + kCT - 1) / kCT;
                const int gc = (c0 + ch) * 64;
                float4 gq[MG];
#pragma unroll
                for (int g = 0; g < MG; ++g) {
                    const int idx = tid + g * kCT;
                    if (idx < rows * 16) {
                        const int r = idx >> 4, k4 = idx & 15;
                        const float4 p = *reinterpret_cast<const float4*>(st + M.off[ST_H1] + (size_t)r * (4 * H) + gc + k4 * 4);
                        const float4 v = *reinterpret_cast<const float4*>(sH + r * kLDF + ch * 64 + k4 * 4);
                        gq[g] = make_float4(v.x * gelu_grad_tc(p.x), v.y * gelu_grad_tc(p.y), v.z * gelu_grad_tc(p.z), v.w * gelu_grad_tc(p.w));
                    }
                }
                c.slot_acquire();
#pragma unroll
                for (int g = 0; g < MG; ++g) {
                    const int idx = tid + g * kCT;
                    if (idx < rows * 16) can_store4<C::kCS>(c.slot_hi, c.slot_lo, idx >> 4, idx & 15, gq[g]);
                }
                c.slot_post();
            }
        }
        c.mark(13);
        c.acc_wait();
        c.mark(5);
        tmem_foreach_sum<C::kHP, DFF_TC_SPLIT_ACC != 0>(c.tmem, kColAcc, C::kCorrFF, rows, [&](int row, int col, const float (&v)[16]) {
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(c.sNh + row * C::LDH + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        });
        tc::fence_before_sync();
        csync();
        // LayerNorm 2 backward + gated residual 1 backward: d att -> canonical operand, d n_in (residual part) -> sN
        gate_backward_rows_can<C>(c.sN, c.sNh, c.nhat_hi, c.nhat_lo, H, rows, W.ln2_g, st + M.off[ST_M], st + M.off[ST_STAT2],
                                  st + M.off[ST_ATT], st + M.off[ST_NIN], st + M.off[ST_G1], W.g1a, W.g1b);
        c.post();
        c.mark(14);

        for (int hc = 0; hc < C::NCH; ++hc) {
            // start reloading q | k' | v' and p of this chunk (two bulk copies onto one mbarrier); they land while d o is read back
            if (tid == 0) {
                const uint32_t bq = (uint32_t)(rows * C::LDQ) * sizeof(float), bp = (uint32_t)(rows * NP) * sizeof(float);
                mbar_expect_tx(c.bars + B_RELOAD, bq + bp);
                bulk_g2s(c.sQKV, st + M.off[ST_QKV] + (size_t)hc * R * C::LDQ, bq, c.bars + B_RELOAD);
                bulk_g2s(c.sP, st + M.off[ST_P] + (size_t)hc * R * NP, bp, c.bars + B_RELOAD);
            }
            {   // d o_chunk = d att x Wo_b[l][hc]: TMEM -> shared
                c.mark(15);
                const int b = c.dq_wait();
                c.mark(1);
                tmem_foreach<64>(c.tmem, C::kColD + b * 64, rows, [&](int row, int col, const float (&v)[16]) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(c.sO + row * C::LDO + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                });
                mbar_wait_wd(c.bars + B_RELOAD, c.n_reload & 1u, 9);
                ++c.n_reload;
#if DFF_TC_DISCARD_STASH
                // q | k' | v' and p of this chunk are in shared memory now and their stash copy is dead until the next step
                // rewrites it: drop the lines from L2 so that they are never written back to HBM
                discard_l2_range(st + M.off[ST_QKV] + (size_t)hc * R * C::LDQ, (size_t)(rows * C::LDQ) * sizeof(float), tid, kCT);
                discard_l2_range(st + M.off[ST_P] + (size_t)hc * R * NP, (size_t)(rows * NP) * sizeof(float), tid, kCT);
#endif
                c.dq_release();
                c.mark(16);
            }
            if constexpr (!C::kAttMma) {
                if (c.quads) attn_backward_ds_dq_quads<C>(c, N, NP, deep);
                else if (c.pairs) attn_backward_ds_dq_pairs<C>(c, N, NP, deep);
                else attn_backward_ds_dq<C>(c, N, NP, deep);
                c.mark(17);
                if (deep) c.slot_post(); else csync();          // every ds of the sample is in sDS before the key-row passes
                c.mark(18);
                if (c.quads) attn_backward_dkv_quads<C>(c, W, hc, N, NP, deep);
                else if (c.pairs) attn_backward_dkv_pairs<C>(c, W, hc, N, NP, deep);
                else attn_backward_dkv<C>(c, W, hc, N, NP, deep);
                c.mark(19);
            } else {
                const AttnGeo G(N, NP, c.S_act);
                const bool dist = M.edge_dist != 0;
                attn_dp_uw_items<C>(c.sQKV, c.sO, c.sDS, W.A, hc, G);     // dp, (u, alpha), (w, beta) (HMMA items)
                csync();
                c.mark(25);
                attn_ds_rows<C>(c.sP, c.sDS, c.sO, c.sX, c.sTmp, dist, G);
                csync();
                c.mark(17);
                attn_dx_rows<C>(c.sQKV, c.sO, c.sP, c.sDS, c.sX, c.sDX, dist, G);     // dx from p, ds, u, w (+ the distance channel)
                c.mark(26);
                if (deep) {
                    attn_weighted_items<C>(c.sDS, c.sQKV, C::LDQ, 64, G, [&](int row, int col, float v0, float v1) {     // dq -> job d q
                        v0 *= kAttnScale; v1 *= kAttnScale;
                        if (dist) {            // + d alpha_i a_h
                            const float da = c.sTmp[row * 4];
                            v0 = fmaf(da, __ldg(W.A + (hc * 64 + col) * 4 + 3), v0); v1 = fmaf(da, __ldg(W.A + (hc * 64 + col) * 4 + 7), v1);
                        }
                        c.slot_acquire();
                        can_store2<C::kCS>(c.slot_hi, c.slot_lo, row, col, v0, v1);
                    });
                    c.slot_post();
                    c.mark(18);
                    attn_keys_to_slot_items<C>(c, c.sDS, c.sQKV, C::LDQ, kAttnScale, G);     // dk' -> job d k'
                    c.slot_post();
                    c.mark(27);
                    attn_keys_to_slot_items<C>(c, c.sP, c.sO, C::LDO, 1.0f, G);              // dv' -> job d v'
                    c.slot_post();
                } else {
                    csync();          // the next head reloads the buffers
                }
                c.mark(19);
            }
        }
        if (deep) {
            c.acc_wait();
            // d n_hat = (dq Wq) + (dk' Wk) + (dv' Wv): three TMEM accumulators (columns 0, kColD + 128, kColD + 128 + HP), added here
            {
                const int q = warp_id & 3, part = warp_id >> 2;
                constexpr int PARTS = kCW / 4;
                const int row = q * 16 + lane_id;
                if (q * 16 < rows) {
                    const uint32_t lane_base = c.tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                    for (int cc = part * 16; cc < C::kHP; cc += 16 * PARTS) {      // 16-column chunks round-robin over the warp parts
                        float v0[16], v1[16], v2[16];
                        tmem_ld16(lane_base + kColAcc + (uint32_t)cc, v0);
                        tmem_ld16(lane_base + C::kColD + 128u + (uint32_t)cc, v1);
                        tmem_ld16(lane_base + C::kColD + 128u + (uint32_t)C::kHP + (uint32_t)cc, v2);
                        if constexpr (DFF_TC_SPLIT_ACC && C::kCorrAttn) {     // + the three correction accumulators (small terms first)
                            float w0[16], w1[16];
                            tmem_ld16(lane_base + C::kCorrDn + (uint32_t)cc, w0);
                            tmem_ld16(lane_base + C::kCorrDn + (uint32_t)C::kHP + (uint32_t)cc, w1);
#pragma unroll
                            for (int i = 0; i < 16; ++i) w0[i] += w1[i];
                            tmem_ld16(lane_base + C::kCorrDn + 2u * (uint32_t)C::kHP + (uint32_t)cc, w1);
#pragma unroll
                            for (int i = 0; i < 16; ++i) v0[i] += w0[i] + w1[i];
                        }
                        if (lane_id < 16 && row < rows) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                *reinterpret_cast<float4*>(c.sNh + row * C::LDH + cc + i) =
                                    make_float4((v0[i] + v1[i]) + v2[i], (v0[i + 1] + v1[i + 1]) + v2[i + 1],
                                                (v0[i + 2] + v1[i + 2]) + v2[i + 2], (v0[i + 3] + v1[i + 3]) + v2[i + 3]);
                        }
                    }
                }
            }
            tc::fence_before_sync();
            csync();
            ln_backward_rows_tc<C>(c.sN, c.sNh, H, rows, W.ln1_g, st + M.off[ST_NIN], st + M.off[ST_STAT1]);
            csync();
            c.mark(21);
        }
    }
    if (M.abs_coords) {
        // node_embedding's x columns (graph_transformer.py:99-103): dx_i += W_n[:, N..N+2]^T d n0_i   (sN = d n0 after layer 0)
        for (int r = warp_id; r < rows; r += kCW) {
            float g0 = 0.f, g1 = 0.f, g2 = 0.f;
            for (int d = lane_id; d < H; d += 32) {
                const float v = c.sN[r * C::LDH + d];
                g0 = fmaf(v, __ldg(M.embx + d), g0); g1 = fmaf(v, __ldg(M.embx + H + d), g1); g2 = fmaf(v, __ldg(M.embx + 2 * H + d), g2);
            }
            g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2);
            if (lane_id == 0) { c.sDX[r * 4] += g0; c.sDX[r * 4 + 1] += g1; c.sDX[r * 4 + 2] += g2; }
        }
        csync();
    }
}

// ------------------------------------------------------------------ the kernel
template <class C>
#ifdef DFF_TC_MAXNREG
__global__ void __maxnreg__(DFF_TC_MAXNREG)
#else
__global__ void __launch_bounds__(kTcThreads, 1)
#endif
dff_fused_tc_kernel(const __grid_constant__ ModelDev M, const __grid_constant__ StepArgs A, const __grid_constant__ TcArgs T) {
    constexpr int R = C::kR;
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int N = M.N;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::oBar);
    uint32_t* ctr = reinterpret_cast<uint32_t*>(bars + B_COUNT);     // [0] tmem base
    TcJob* jobs = reinterpret_cast<TcJob*>(smem + C::oJobs);
    const int njobs = A.need_backward ? T.njobs_all : T.njobs_fwd;
    const uint32_t nslices = A.need_backward ? T.nslice_all : T.nslice_fwd;

    // Balanced work split: this CTA owns a contiguous range of samples (B / grid, +1 for the first B % grid CTAs) and
    // walks it in the fewest passes of <= M.S samples, of near-equal size (e.g. 7 samples = 3 + 2 + 2, not 3 + 3 + 1):
    // a pass costs roughly in proportion to its rows, so equal passes and equal ranges minimise the makespan.
    const int per = A.B / (int)gridDim.x, rem = A.B % (int)gridDim.x;
    const int my_n = per + ((int)blockIdx.x < rem ? 1 : 0);
    const int my_start = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
    const int my_groups = (my_n + M.S - 1) / M.S;
    const uint32_t reps = (uint32_t)my_groups * (uint32_t)A.n_steps;

    for (int idx = tid; idx < C::oW; idx += kTcThreads) smem[idx] = 0.f;             // activations, operands
    for (int idx = C::oX + tid; idx < C::oJobs; idx += kTcThreads) smem[idx] = 0.f;
    for (int idx = tid; idx < min(njobs, kJobCap) * 4; idx += kTcThreads)          // entries past the cache are read from global memory
        reinterpret_cast<uint32_t*>(jobs)[idx] = reinterpret_cast<const uint32_t*>(T.jobs)[idx];
    if (tid == 0) {
        for (int i = 0; i < C::kStages; ++i) { mbar_init(bars + B_FULL + i, 1); mbar_init(bars + B_EMPTY + i, 1); }
        mbar_init(bars + B_DQ, 1); mbar_init(bars + B_DQ + 1, 1);
        mbar_init(bars + B_ACC, 1); mbar_init(bars + B_D1, 1); mbar_init(bars + B_SLOT, 1); mbar_init(bars + B_RELOAD, 1);
        for (int i = 0; i < 8; ++i) mbar_init(bars + B_POST + i, 1);
        mbar_init(bars + B_DRAIN, 1); mbar_init(bars + B_DRAIN + 1, 1);
        fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc(ctr, kTmemCols);
    fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = ctr[0];

#if DFF_TC_SETMAXNREG
#define DFF_REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsService))
#define DFF_REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsCompute))
#else
#define DFF_REG_DEC() do { } while (0)
#define DFF_REG_INC() do { } while (0)
#endif
    // Every warp of the service warpgroup shrinks at the top of its own role branch (same register count; the pattern of the
    // CUTLASS sm100 role-specialised kernels).  One shared setmaxnreg ahead of the role dispatch compiles to code that is no
    // faster than the 96-register build (measured on one box: 3254 vs 3559 MD steps/s on C2, 464 vs 509 on C4).
    if (warp >= kCW) {
    if (warp == kComputeThreads / 32) {
        DFF_REG_DEC();
        // ===================================================== TMA producer: streams the weight slices of every job
        if ((tid & 31) == 0) {
            uint32_t slice_i = 0;
            long long pw[1] = {0}; (void)pw;
            for (uint32_t rep = 0; rep < reps; ++rep)
                for (int j = 0; j < njobs; ++j) {
                    const TcJob jb = j < kJobCap ? jobs[j] : T.jobs[j];
                    const char* src = reinterpret_cast<const char*>(T.wbase + jb.w_off);
                    const uint32_t slice_bytes = (uint32_t)jb.slice_16b * 16u;
                    for (uint32_t s = 0; s < jb.n_slices; ++s, ++slice_i) {
                        const uint32_t stg = slice_i % C::kStages, use = slice_i / C::kStages;
                        { TCP_BEGIN(); if (use > 0) mbar_wait_wd(bars + B_EMPTY + stg, (use - 1) & 1u, 5); TCP_END(pw, 0); }
#ifdef DFF_EXP_STREAM_DIV      // timing experiment only (wrong results): stream a fraction of every weight slice
                        const uint32_t cp_bytes = max(16u, (slice_bytes / DFF_EXP_STREAM_DIV) & ~15u);
#else
                        const uint32_t cp_bytes = slice_bytes;
#endif
                        mbar_expect_tx(bars + B_FULL + stg, cp_bytes);
                        bulk_g2s(smem + C::oW + stg * C::kStageFloats, src + (size_t)s * slice_bytes, cp_bytes, bars + B_FULL + stg);
                    }
                }
            (void)nslices;
#ifdef DFF_TC_PROFILE
            if (T.dbg) T.dbg[blockIdx.x * 16 + 8] = pw[0];
#endif
        }
    } else if (warp == kComputeThreads / 32 + 1) {
        DFF_REG_DEC();
        // ===================================================== MMA issuer: walks the job table.
        // The whole warp runs the (warp-uniform) control flow and descriptor arithmetic, so that the operands of
        // tcgen05.mma stay in uniform registers; one elected lane issues the MMAs and commits.
        {
            uint32_t slice_i = 0, post_seq = 0, dq_idx = 0;
            long long iw[4] = {0, 0, 0, 0};
            const long long t_begin = clock64();
            const uint32_t lane_id = tid & 31;
            const uint32_t nhat_hi_a = smem_u32(smem + C::oNhatHi), nhat_lo_a = smem_u32(smem + C::oNhatLo);
            const uint32_t slot_hi_a = smem_u32(smem + C::oSlotHi), slot_lo_a = smem_u32(smem + C::oSlotLo);
            const uint32_t ring_a = smem_u32(smem + C::oW);
            constexpr uint64_t kDescHi = ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);      // SBO = 128 B, version 1
            constexpr uint64_t kDescA = kDescHi | ((uint64_t)((C::kCS * 4) >> 4) << 16);             // LBO = chunk stride
            constexpr uint32_t a_step = (uint32_t)(2 * C::kCS * 4) >> 4;                             // two 16-byte k-chunks per MMA
            for (uint32_t rep = 0; rep < reps; ++rep)
                for (int j = 0; j < njobs; ++j) {
                    // job fields, made warp-uniform
                    const uint32_t* jw = reinterpret_cast<const uint32_t*>(j < kJobCap ? jobs + j : T.jobs + j);      // words 1..3 of the 16-byte entry
                    const uint32_t f0 = __shfl_sync(0xffffffffu, jw[1], 0), f1 = __shfl_sync(0xffffffffu, jw[2], 0);
                    const uint32_t f2 = __shfl_sync(0xffffffffu, jw[3], 0);
                    const uint32_t n_slices = (f0 >> 16) & 0xffu, ks = f0 >> 24, n = f1 & 0xffffu;
                    uint32_t dcol = f1 >> 16;
                    const uint32_t a_slot = f2 & TCJ_SLOT, acc_first = (f2 & TCJ_ACC) ? 1u : 0u, wait_post = f2 & TCJ_WAIT_POST, dbuf = f2 & TCJ_DBUF;
                    const uint32_t commit_acc = f2 & TCJ_COMMIT_ACC, commit_d1 = f2 & TCJ_COMMIT_D1;
                    const uint32_t ccol = f2 >> 16;
                    const bool split = ccol != dcol;                 // never set on double-buffered jobs
                    if (wait_post) {
                        TCP_BEGIN(); mbar_wait_wd(bars + B_POST + (post_seq & 7u), (post_seq >> 3) & 1u, 7); TCP_END(iw, 0);
                        ++post_seq;
                    }
                    if (dbuf) {
                        // buffer dq_idx & 1 is written for the (dq_idx >> 1)-th time: its previous contents must have been read back
                        TCP_BEGIN(); if (dq_idx >= 2) mbar_wait_wd(bars + B_DRAIN + (dq_idx & 1u), ((dq_idx >> 1) - 1u) & 1u, 8); TCP_END(iw, 1);
                        dcol += (dq_idx & 1u) * n;
                    }
                    tc::fence_after_sync();
                    uint64_t dah = kDescA | (uint64_t)(((a_slot ? slot_hi_a : nhat_hi_a) >> 4) & 0x3FFFu);
                    uint64_t dal = kDescA | (uint64_t)(((a_slot ? slot_lo_a : nhat_lo_a) >> 4) & 0x3FFFu);
                    const uint32_t idesc = tc::idesc_tf32(64, (int)n);
                    const uint32_t d_tmem = tmem + dcol;
                    const uint32_t c_tmem = split ? tmem + ccol : d_tmem;    // lo*hi and hi*lo go here
                    const uint64_t descB = kDescHi | ((uint64_t)n << 16);                          // LBO = n * 16 B
                    const uint32_t b_step = 2u * n;                                               // (2 * n * 16 B) >> 4
                    const uint32_t lo_off = (ks * n * 4u) >> 4;                                   // lo image follows the hi image
                    const uint32_t ksteps = ks >> 3;
                    uint32_t acc = acc_first;
                    for (uint32_t s = 0; s < n_slices; ++s, ++slice_i) {
                        const uint32_t stg = slice_i % C::kStages, use = slice_i / C::kStages;
                        uint64_t dbh = descB | (uint64_t)(((ring_a + stg * (C::kStageFloats * 4)) >> 4) & 0x3FFFu);
                        uint64_t dbl = dbh + lo_off;
                        { TCP_BEGIN(); mbar_wait_wd(bars + B_FULL + stg, use & 1u, 6); TCP_END(iw, 2); }
                        tc::fence_after_sync();
#ifdef DFF_EXP_ONE_PASS        // timing experiment only (TF32-grade results): hi*hi alone, a third of the MMAs
                        if (tc::elect_one()) tc::mma_tf32_ss(d_tmem, dah, dbh, idesc, acc);
#else
                        if (tc::elect_one()) {
                            tc::mma_tf32_ss(c_tmem, dal, dbh, idesc, acc);          // first k-step of the slice (may overwrite D)
                            tc::mma_tf32_acc(c_tmem, dah, dbl, idesc);
                            tc::mma_tf32_ss(d_tmem, dah, dbh, idesc, split ? acc : 1u);
                        }
#endif
                        acc = 1u;
                        for (uint32_t kk = 1; kk < ksteps; ++kk) {
                            dah += a_step; dal += a_step; dbh += b_step; dbl += b_step;
#ifdef DFF_EXP_ONE_PASS
                            if (tc::elect_one()) tc::mma_tf32_acc(d_tmem, dah, dbh, idesc);
#else
                            if (tc::elect_one()) {
                                tc::mma_tf32_acc(c_tmem, dal, dbh, idesc);
                                tc::mma_tf32_acc(c_tmem, dah, dbl, idesc);
                                tc::mma_tf32_acc(d_tmem, dah, dbh, idesc);
                            }
#endif
                        }
                        dah += a_step; dal += a_step;
                        if (tc::elect_one()) tc::commit(bars + B_EMPTY + stg);
                    }
                    if (tc::elect_one()) {
                        if (a_slot) tc::commit(bars + B_SLOT);
                        if (dbuf) tc::commit(bars + B_DQ + (dq_idx & 1u));
                        if (commit_acc) tc::commit(bars + B_ACC);
                        if (commit_d1) tc::commit(bars + B_D1);
                    }
                    if (dbuf) ++dq_idx;
                    __syncwarp();
                }
#ifdef DFF_TC_PROFILE
            if (T.dbg && lane_id == 0) { T.dbg[blockIdx.x * 16 + 9] = iw[0]; T.dbg[blockIdx.x * 16 + 10] = iw[1]; T.dbg[blockIdx.x * 16 + 11] = iw[2];
                         T.dbg[blockIdx.x * 16 + 12] = clock64() - t_begin; }
#else
            (void)iw; (void)t_begin; (void)lane_id;
#endif
        }
    } else {
        DFF_REG_DEC();        // (DFF_TC_SETMAXNREG: warps kCW + 2, kCW + 3 idle)
    }
    } else {
        // ===================================================== compute warps
        DFF_REG_INC();
        Ctx2 c;
        c.sQKV = smem + C::oQKV; c.sNh = c.sQKV; c.sO = smem + C::oO;
        c.sP = smem + C::oP; c.sDS = smem + C::oDS; c.sX = smem + C::oX; c.sV = smem + C::oV;
        c.sDX = smem + C::oDX; c.sTmp = smem + C::oTmp;
        c.nhat_hi = smem + C::oNhatHi; c.nhat_lo = smem + C::oNhatLo; c.slot_hi = smem + C::oSlotHi; c.slot_lo = smem + C::oSlotLo;
        c.bars = bars; c.tmem = tmem;
        c.n_post = c.n_drain = c.n_acc = c.n_d1 = c.n_slot = c.n_reload = 0; c.slot_held = false;
        for (int i = 0; i < 8; ++i) c.tw[i] = 0;
        const long long t_begin = clock64();
#ifdef DFF_TC_PROFILE
        for (int i = 0; i < 32; ++i) c.ph[i] = 0;
        c.last = t_begin;
#endif
        c.stash = M.scratch + (size_t)blockIdx.x * M.scratch_per_cta;
        // node stream [R][LDH]: shared memory for hidden 64, the tail of this CTA's global scratch otherwise
        c.sN = C::kNodeInSmem ? smem + C::oN : c.stash + (M.scratch_per_cta - (long long)R * (128 + 4));

        uint32_t flags = 0;
        int s_next = my_start;
        for (int g = 0; g < my_groups; ++g) {
            const int s0 = s_next;
            c.S_act = my_n / my_groups + (g < my_n % my_groups ? 1 : 0);
            s_next += c.S_act;
            c.rows_act = c.S_act * N;
#ifndef DFF_TC_QUADS16
#define DFF_TC_QUADS16 0
#endif
            c.quads = (AttnMap<C>::LPR == 32 || AttnMap<C>::KPL == 2 || DFF_TC_QUADS16) && c.rows_act > kCW * AttnMap<C>::UPW && N >= 4 &&
                      c.S_act * ((N + 3) >> 2) <= kCW * AttnMap<C>::UPW;      // more rows than lane groups, one quad per group
            c.pairs = c.rows_act > kCW * AttnMap<C>::UPW || AttnMap<C>::KPL == 2;      // (the row-local routines hold one key per lane)
            for (int idx = tid; idx < R * 3; idx += kCT) {
                const int r = idx / 3, cc = idx - r * 3;
                const bool ok = r < c.rows_act;
                c.sX[r * 4 + cc] = ok ? A.x[((size_t)s0 * N + r) * 3 + cc] : 0.f;
                c.sV[r * 4 + cc] = (ok && A.v != nullptr) ? A.v[((size_t)s0 * N + r) * 3 + cc] : 0.f;
            }
            csync();

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
                csync();
                const int it = A.t_start - step;
                const float t_norm = (A.mode == MODE_DDPM) ? (float)it / (float)A.T : A.t_norm;

                c.mark(22);
                forward_pass_tc<C>(M, c, t_norm, (A.mode == MODE_SCORE && A.t_rows != nullptr) ? A.t_rows + s0 : nullptr);
                if (A.energy_out != nullptr) {   // node_decoder (graph_transformer.py:106)
                    const int lane = tid & 31;
                    for (int r = warp; r < c.rows_act; r += kCW) {
                        float s = 0.f;
                        for (int d = lane; d < M.H; d += 32) s += c.sN[r * C::LDH + d] * __ldg(M.dec_w + d);
                        s = warp_sum(s);
                        if (lane == 0) A.energy_out[(size_t)s0 * N + r] = s + M.dec_b;
                    }
                }
                csync();
                if (A.need_backward) backward_pass_tc<C>(M, c);
                if (!M.conservative) {
                    // non-conservative head (graph_transformer.py:62-65, 112-113): the prediction is node_decoder(nodes);
                    // stored negated so that the samplers below read eps = -sDX exactly as in the conservative case
                    const int lane = tid & 31;
                    for (int r = warp; r < c.rows_act; r += kCW) {
                        float s0_ = 0.f, s1_ = 0.f, s2_ = 0.f;
                        for (int d = lane; d < M.H; d += 32) {
                            const float v = c.sN[r * C::LDH + d];
                            s0_ = fmaf(v, __ldg(M.dec_w + d), s0_); s1_ = fmaf(v, __ldg(M.dec_w + M.H + d), s1_); s2_ = fmaf(v, __ldg(M.dec_w + 2 * M.H + d), s2_);
                        }
                        s0_ = warp_sum(s0_); s1_ = warp_sum(s1_); s2_ = warp_sum(s2_);
                        if (lane == 0) { c.sDX[r * 4] = -(s0_ + M.dec_b3[0]); c.sDX[r * 4 + 1] = -(s1_ + M.dec_b3[1]); c.sDX[r * 4 + 2] = -(s2_ + M.dec_b3[2]); }
                    }
                }
                csync();

                if (A.mode == MODE_SCORE) {
                    if (A.eps_out != nullptr)
                        for (int idx = tid; idx < c.rows_act * 3; idx += kCT) {
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
                    for (int idx = tid; idx < c.rows_act * 3; idx += kCT) {
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
                        csync();
                        const int f = step / A.save_interval;
                        if (A.frames != nullptr)
                            for (int idx = tid; idx < c.rows_act * 3; idx += kCT) {
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
                csync();
            }
            if (A.mode != MODE_SCORE) {
                for (int idx = tid; idx < c.rows_act * 3; idx += kCT) {
                    const int r = idx / 3, cc = idx - r * 3;
                    A.x[((size_t)s0 * N + r) * 3 + cc] = c.sX[r * 4 + cc];
                    if (A.v != nullptr) A.v[((size_t)s0 * N + r) * 3 + cc] = c.sV[r * 4 + cc];
                }
            }
            csync();
        }
        if (A.flags != nullptr && flags != 0) atomicOr(A.flags, flags);
#ifdef DFF_TC_PROFILE
        if (T.dbg && tid == 0) { for (int i = 0; i < 4; ++i) T.dbg[blockIdx.x * 16 + i] = c.tw[i]; T.dbg[blockIdx.x * 16 + 7] = clock64() - t_begin; }
        if (T.dbg && tid == 0 && blockIdx.x == 0) for (int i = 0; i < 32; ++i) T.dbg[(size_t)gridDim.x * 16 + i] = c.ph[i];
#else
        (void)t_begin;
#endif
    }

    // teardown: every MMA has completed (the compute warps waited on the last accumulator) before TMEM is freed
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_free(tmem, kTmemCols);
}

}  // namespace v2
}  // namespace dff
