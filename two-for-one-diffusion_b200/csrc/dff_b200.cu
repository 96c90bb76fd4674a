// dff_b200.cu -- host side of libdff_b200.so: weight packing (K0), launch policy and the C ABI
// declared in include/dff_b200.h.  No torch types here; PyTorch only lends device pointers.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#define DFF_HOST_TU 1
#include "../../include/dff_b200.h"
#include "dff_kernel.cuh"
#include "dff_tc.cuh"
#include "dff_kernel_tc.cuh"
#include "dff_metrics.cuh"
#include "dff_tc_configs.h"

using namespace dff;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(DFF_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Makes `device` current for the scope of an ABI call and restores the caller's device afterwards.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device) ok = cudaSetDevice(device) == cudaSuccess;
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define DEVICE_SCOPE(dev)                                                          \
    DeviceGuard dg_(dev);                                                          \
    if (!dg_.ok) return fail(DFF_ECUDA, "cudaSetDevice(%d) failed: %s", dev, cudaGetErrorString(cudaGetLastError()))

struct SegHost {
    size_t offset;      // floats into the packed buffer
    uint32_t slice_bytes, n_slices;
};

constexpr int ks_for(int nc) { return nc == 384 ? 8 : (nc == 192 ? 16 : (nc == 128 ? 16 : 32)); }

}  // namespace

struct dff_model {
    int device = 0, num_sms = 0;
    int conservative = 1;
    int N = 0, NP = 0, H = 0, HP = 0, L = 0, nch = 0;
    int max_batch = 0;
    float* d_weights = nullptr;     // every packed tensor + GEMM panel
    Seg* d_segs[2] = {nullptr, nullptr};
    float* d_scratch = nullptr;
    ModelDev md[2];                 // [0]: R=32, 2 heads per chunk; [1]: R=64, 1 head per chunk.  S is filled per launch
    long long layer_floats[2] = {0, 0};          // for R = 32, 64
    long long off[2][ST_COUNT];
    long long scratch_per_cta = 0;               // floats (sized for R = 64)
    int scratch_ctas = 0;
    int64_t launches = 0;
    // staging for the *_host entry points
    float* d_io[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t d_io_cap[6] = {0, 0, 0, 0, 0, 0};
    uint32_t* d_flags = nullptr;
    float* d_sched = nullptr;  size_t d_sched_T = 0;
    int last_R = 0, last_S = 0, last_att = 0;
    bool attn_mma = false;          // default attention flavour of the tcgen05 kernel for this model (set at create)
    bool modes_default = true;      // intrinsic coordinates only (what the mma.sync fallback kernel implements)
    const char* last_cfg = "none";
    // tcgen05 configuration (hidden = 64): job table + canonical hi/lo weight panels
    bool tc_ok = false;
    v2::TcJob* d_jobs = nullptr;
    v2::TcArgs tc{};
};

namespace {

void stash_geometry(int R, int N_pad, int H, long long* off, long long* layer_floats) {
    long long o = 0;
    auto put = [&](int region, long long n) { off[region] = o; o += (n + 3) / 4 * 4; };
    put(ST_NIN, (long long)R * H);
    put(ST_STAT1, 2LL * R);
    put(ST_QKV, 8LL * R * 196);          // rows keep the shared-memory stride (192 + 4 pad) so that a chunk moves as one bulk copy
    put(ST_P, 8LL * R * N_pad);
    put(ST_ATT, (long long)R * H);
    put(ST_G1, R);
    put(ST_M, (long long)R * H);
    put(ST_STAT2, 2LL * R);
    put(ST_H1, 4LL * R * H);
    put(ST_FF, (long long)R * H);
    put(ST_G2, R);
    *layer_floats = o;
}

struct Packer {
    std::vector<float> buf;
    size_t alloc(size_t n) {
        size_t o = buf.size();
        buf.resize(o + (n + 3) / 4 * 4, 0.f);
        return o;
    }
};

// kernels live in their own translation units (dff_tc_inst.cu, dff_legacy_inst.cu), one configuration each
#define DFF_TC_DECL(PN, HP, R, ATT) \
    extern "C" cudaError_t dff_tc_launch_##PN##_##HP##_##R##_##ATT(const ModelDev*, const StepArgs*, const v2::TcArgs*, int, cudaStream_t);
DFF_TC_CONFIGS(DFF_TC_DECL)
#undef DFF_TC_DECL
extern "C" cudaError_t dff_legacy_launch_64_0(const ModelDev*, const StepArgs*, int, cudaStream_t);
extern "C" cudaError_t dff_legacy_launch_64_1(const ModelDev*, const StepArgs*, int, cudaStream_t);
extern "C" cudaError_t dff_legacy_launch_64_2(const ModelDev*, const StepArgs*, int, cudaStream_t);
extern "C" cudaError_t dff_legacy_launch_128_0(const ModelDev*, const StepArgs*, int, cudaStream_t);
extern "C" cudaError_t dff_legacy_launch_128_1(const ModelDev*, const StepArgs*, int, cudaStream_t);
extern "C" cudaError_t dff_legacy_launch_128_2(const ModelDev*, const StepArgs*, int, cudaStream_t);

using TcLaunchFn = cudaError_t (*)(const ModelDev*, const StepArgs*, const v2::TcArgs*, int, cudaStream_t);
using LegacyLaunchFn = cudaError_t (*)(const ModelDev*, const StepArgs*, int, cudaStream_t);

TcLaunchFn tc_launcher(int PN, int HP, int R, int ATT) {
#define DFF_TC_PICK(pn, hp, r, att) if (PN == pn && HP == hp && R == r && ATT == att) return dff_tc_launch_##pn##_##hp##_##r##_##att;
    DFF_TC_CONFIGS(DFF_TC_PICK)
#undef DFF_TC_PICK
    return nullptr;
}

int launch_legacy(dff_model* m, LegacyLaunchFn fn, const ModelDev& M, const StepArgs& A, int grid, cudaStream_t stream) {
    cudaError_t e = fn(&M, &A, grid, stream);
    if (e != cudaSuccess) return fail(DFF_ECUDA, "legacy kernel launch failed: %s", cudaGetErrorString(e));
    m->launches += 1;
    return DFF_OK;
}

int launch_tc(dff_model* m, int PN, int HP, int R, int ATT, const ModelDev& M, const StepArgs& A, int grid, cudaStream_t stream) {
    TcLaunchFn fn = tc_launcher(PN, HP, R, ATT);
    if (!fn) return fail(DFF_EINVAL, "internal: no kernel instantiation TcCfg<%d,%d,%d,%d>", PN, HP, R, ATT);
    cudaError_t e = fn(&M, &A, &m->tc, grid, stream);
    if (e != cudaSuccess) return fail(DFF_ECUDA, "tcgen05 kernel launch failed: %s", cudaGetErrorString(e));
#ifdef DFF_TC_PROFILE
    if (m->tc.dbg) {     // developer build: print the per-CTA wait-cycle breakdown of this launch (synchronises)
        std::vector<long long> h((size_t)grid * 16 + 32);
        CUDA_TRY(cudaStreamSynchronize(stream));
        CUDA_TRY(cudaMemcpy(h.data(), m->tc.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double s[16] = {0};
        for (int b = 0; b < grid; ++b) for (int i = 0; i < 16; ++i) s[i] += (double)h[(size_t)b * 16 + i] / grid;
        fprintf(stderr, "[tc profile] grid %d steps %d: compute total %.0f cyc; waits dq %.0f acc %.0f d1 %.0f slot %.0f | issuer total %.0f: post %.0f drain %.0f weights %.0f | producer empty-wait %.0f\n",
                grid, A.n_steps, s[7], s[0], s[1], s[2], s[3], s[12], s[9], s[10], s[11], s[8]);
        static const char* names[32] = {"f.init+ln", "dq_wait", "f.qkv_epi", "f.attn", "f.slot_post", "acc_wait", "f.acc_epi", "f.gate+post", "d1_wait", "f.d1copy",
                                        "f.gelu", "b.gate2+post", "b.d1copy", "b.gelu'", "b.accepi+gate1+post", "b.reload_issue", "b.do_epi", "b.ds", "b.dq+post",
                                        "b.dk+post", "b.dv+post", "b.acc+lnbwd", "integrator", "f.logits|bq.dots", "f.softmax|bq.ds", "b.dp_uw|bq.dq", "b.dx|bq.dkvloop", "b.dk'|bq.dkstore", "fq.pre", "fq.dots", "fq.softmax", "fq.pv"};
        fprintf(stderr, "[tc phases, CTA 0, cycles per step]");
        for (int i = 0; i < 32; ++i) fprintf(stderr, " %s %.0f |", names[i], (double)h[(size_t)grid * 16 + i] / A.n_steps);
        fprintf(stderr, "\n");
    }
#endif
    m->launches += 1;
    return DFF_OK;
}

// Launch policy.  Four configurations:
//   tc   : tcgen05 kernel (dff_kernel_tc.cuh): 64-row passes, TMEM accumulators, asynchronous MMA issue; hidden = 64 nets.
//          The default wherever it applies (measured: C2 3179 vs 2270 steps/s tall, C3 350 vs 265 duo).
//   wide : mma.sync kernel, R = 64 rows, 1 head per chunk, 4-stage ring, 1 CTA/SM   (anything; the only one for N > 32)
//   tall : mma.sync kernel, R = 32 rows, 2 heads per chunk, 4-stage ring, 1 CTA/SM  (small batches: fewer, fatter phases)
//   duo  : mma.sync kernel, R = 32 rows, 1 head per chunk, 2-stage ring, <= 128 regs, 2 CTAs/SM
// DFF_CONFIG=tc|legacy|wide|tall|duo overrides the choice (profiling / A-B runs).
int launch(dff_model* m, StepArgs& A, cudaStream_t stream) {
    if (A.B <= 0) return DFF_OK;
    if (A.B > m->max_batch) return fail(DFF_EINVAL, "batch %d exceeds max_batch %d given to dff_model_create", A.B, m->max_batch);
    DEVICE_SCOPE(m->device);
    const int N = m->N;
    const int s64 = 64 / N, s32 = 32 / N;
    enum { WIDE, TALL, DUO, TC } cfg;
    const int need1 = (A.B + m->num_sms - 1) / m->num_sms;            // samples per CTA covering the batch with 1 CTA/SM
    const int need2 = (A.B + 2 * m->num_sms - 1) / (2 * m->num_sms);  // ... with 2 CTAs/SM
    // measured on B200 (profiles/r01/ablation.md): tall wins when the batch fits one wave at R = 32 (C2: 2244 vs 1658 duo
    // vs 1353 wide steps/s); duo wins for large batches when 32-row passes are as full as 64-row ones (C3: 269 vs 239
    // wide); wide wins when a 32-row pass would be mostly padding (trp-cage N = 20: 239 vs 197 duo).
    const bool rows32_full = s32 >= 1 && 2 * s32 * 10 >= s64 * 9;      // util(32) >= 0.9 * util(64)
    if (s32 < 1) cfg = WIDE;
    else if (need1 <= s32) cfg = TALL;
    else if (rows32_full && kThreads == 256) cfg = DUO;
    else cfg = WIDE;
    const auto legacy = cfg;           // the mma.sync kernel's choice (DFF_CONFIG=legacy)
    if (m->tc_ok) cfg = TC;
    if (const char* e = getenv("DFF_CONFIG")) {
        if (!strcmp(e, "legacy")) cfg = legacy;
        if (!strcmp(e, "wide")) cfg = WIDE;
        else if (!strcmp(e, "tall") && s32 >= 1) cfg = TALL;
        else if (!strcmp(e, "duo") && s32 >= 1 && kThreads == 256) cfg = DUO;
        else if (!strcmp(e, "tc") && m->tc_ok) cfg = TC;
    }
    if (cfg != TC && !m->modes_default)
        return fail(DFF_EINVAL, "the distance / absolute-coordinate network modes run on the tcgen05 kernel only (hidden 64 up to 64 beads, "
                                "hidden 96 / 128 up to 56 beads); this shape falls back to the mma.sync kernel, which implements intrinsic coordinates only");
    if (cfg == TC) {
        // tcgen05 kernel: 64-row passes (MMA M = 64), 1 head per chunk, 1 CTA/SM
        const int S = std::min(s64, std::max(need1, 1));
        ModelDev M = m->md[1];
        M.S = S;
        const int n_groups = (A.B + S - 1) / S;
        const int grid = std::min(n_groups, m->num_sms);
        if (grid > m->scratch_ctas) return fail(DFF_EINVAL, "internal: grid %d exceeds scratch slots %d", grid, m->scratch_ctas);
        const double slices = (double)((A.B / grid + 1 + S - 1) / S) * (double)A.n_steps * (double)m->tc.nslice_all;   // passes of the fullest CTA
        if (slices >= 4.0e9) return fail(DFF_EINVAL, "n_steps %d too large for one launch; split the call (e.g. per save interval)", A.n_steps);
        m->last_R = 64; m->last_S = S;
        // attention flavour: HMMA tiles (general path: every edge mode, the large-N nets) or the CUDA-core routines (small intrinsic nets)
        int att = m->attn_mma ? 1 : 0;
        if (const char* e = getenv("DFF_ATTN")) att = !strcmp(e, "mma") ? 1 : (!strcmp(e, "simt") ? 0 : att);
        if (M.edge_dist) att = 1;          // no CUDA-core implementation of the squared-distance channel
        m->last_cfg = att ? "tc" : "tc";
        m->last_att = att;
        int PN, R = 64;
        if (m->HP == 64) {
            if (m->NP <= 12) { PN = 12; if (S * N <= 60) R = 60; }      // 60-row buffers leave room for a 4th weight stage
            else if (m->NP <= 32) PN = 32;
            else PN = 64;
        } else {
            // hidden 96 / 128: passes of <= 60 rows (three 20-bead samples, twelve 5-bead samples) use 60-row buffers, which
            // leaves room for a 4th weight stage
            if (m->NP <= 12) { PN = 12; if (S * N <= 60) R = 60; }
            else if (m->NP <= 20 && S * N <= 60) { PN = 20; R = 60; }
            else if (m->NP <= 32) PN = 32;
            else { PN = 56; R = 56; }     // one 33..56-bead sample per pass: 56-row buffers, 4-stage ring
        }
        return launch_tc(m, PN, m->HP, R, att, M, A, grid, stream);
    }
    int R, S, ctas;
    ModelDev M;
    if (cfg == WIDE) { R = 64; S = std::min(s64, std::max(need1, 1)); M = m->md[1]; ctas = m->num_sms; }
    else if (cfg == TALL) { R = 32; S = std::min(s32, std::max(need1, 1)); M = m->md[0]; ctas = m->num_sms; }
    else {
        R = 32; S = std::min(s32, std::max(need2, 1)); M = m->md[1]; ctas = 2 * m->num_sms;
        M.layer_floats = m->layer_floats[0];
        for (int i = 0; i < ST_COUNT; ++i) M.off[i] = m->off[0][i];
    }
    M.S = S;
    const int n_groups = (A.B + S - 1) / S;
    const int grid = std::min(n_groups, ctas);
    if (grid > m->scratch_ctas) return fail(DFF_EINVAL, "internal: grid %d exceeds scratch slots %d", grid, m->scratch_ctas);
    {   // the in-kernel weight-slice counter is 32-bit: refuse launches that would overflow it
        const double slices = (double)((n_groups + grid - 1) / grid) * (double)A.n_steps * (double)(A.need_backward ? M.nslice_all : M.nslice_fwd);
        if (slices >= 4.0e9) return fail(DFF_EINVAL, "n_steps %d too large for one launch; split the call (e.g. per save interval)", A.n_steps);
    }
    m->last_R = R; m->last_S = S; m->last_cfg = cfg == WIDE ? "wide" : cfg == TALL ? "tall" : "duo";
    const int li = cfg == WIDE ? 0 : cfg == TALL ? 1 : 2;
    static const LegacyLaunchFn legacy_fn[2][3] = {{dff_legacy_launch_64_0, dff_legacy_launch_64_1, dff_legacy_launch_64_2},
                                                   {dff_legacy_launch_128_0, dff_legacy_launch_128_1, dff_legacy_launch_128_2}};
    return launch_legacy(m, legacy_fn[m->HP == 64 ? 0 : 1][li], M, A, grid, stream);
}

int ensure_io(dff_model* m, int slot, size_t floats) {
    if (m->d_io_cap[slot] >= floats) return DFF_OK;
    if (m->d_io[slot]) cudaFree(m->d_io[slot]);
    m->d_io[slot] = nullptr; m->d_io_cap[slot] = 0;
    CUDA_TRY(cudaMalloc(&m->d_io[slot], floats * sizeof(float)));
    m->d_io_cap[slot] = floats;
    return DFF_OK;
}

}  // namespace

extern "C" {

const char* dff_last_error(void) { return g_err.c_str(); }
int dff_version(void) { return 100; }
int dff_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int dff_model_create(dff_model_t** out, int device, int num_beads, int hidden, int n_layers,
                     const float* const* w, int n_weights, int max_batch) {
    return dff_model_create_ex(out, device, num_beads, hidden, n_layers, w, n_weights, max_batch, 1);
}

int dff_model_create_ex(dff_model_t** out, int device, int num_beads, int hidden, int n_layers,
                        const float* const* w, int n_weights, int max_batch, int conservative) {
    dff_model_opts_t o{conservative, 1, 0, 0};
    return dff_model_create_v2(out, device, num_beads, hidden, n_layers, w, n_weights, max_batch, &o);
}

int dff_model_create_v2(dff_model_t** out, int device, int num_beads, int hidden, int n_layers,
                        const float* const* w, int n_weights, int max_batch, const dff_model_opts_t* opts) {
    if (!out) return fail(DFF_EINVAL, "out is NULL");
    *out = nullptr;
    if (!opts) return fail(DFF_EINVAL, "opts is NULL");
    const int conservative = opts->conservative ? 1 : 0;
    const bool intr = opts->use_intrinsic_coords != 0, dist = opts->use_distances != 0, absc = opts->use_abs_coords != 0;
    if (conservative && !intr && !dist && !absc)
        return fail(DFF_EINVAL, "a conservative network without intrinsic coordinates, distances or absolute coordinates does not depend on x: "
                                "the reference raises 'Gradient after computing forces is None' (graph_transformer.py:157-158)");
    const int in_edge = 3 * intr + dist + ((!intr && !dist) ? 1 : 0);       // graph_transformer.py:54-58
    const int in_node = num_beads + 1 + (absc ? 3 : 0);                       // :53
    const int N = num_beads, H = hidden, L = n_layers;
    if (N < 2 || N > kMaxBeads) return fail(DFF_EINVAL, "num_beads %d unsupported (2..%d)", N, kMaxBeads);
    if (H < 32 || H > 128 || H % 32) return fail(DFF_EINVAL, "hidden %d unsupported (32, 64, 96, 128)", H);
    if (L < 1 || L > kMaxLayers) return fail(DFF_EINVAL, "n_layers %d unsupported (1..%d)", L, kMaxLayers);
    if (n_weights != DFF_NUM_GLOBAL_WEIGHTS + DFF_NUM_LAYER_WEIGHTS * L)
        return fail(DFF_EINVAL, "expected %d weight tensors, got %d", DFF_NUM_GLOBAL_WEIGHTS + DFF_NUM_LAYER_WEIGHTS * L, n_weights);
    for (int i = 0; i < n_weights; ++i)
        if (!w[i]) return fail(DFF_EINVAL, "weight pointer %d is NULL", i);
    if (max_batch < 1) return fail(DFF_EINVAL, "max_batch must be >= 1");
    int ndev = dff_device_count();
    if (ndev <= 0) return fail(DFF_ENODEV, "no CUDA device visible: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(DFF_EINVAL, "device %d out of range (%d visible)", device, ndev);
    DEVICE_SCOPE(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(DFF_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);

    dff_model* m = new dff_model();
    m->device = device; m->num_sms = prop.multiProcessorCount;
    m->N = N; m->NP = (N + 3) / 4 * 4; m->H = H; m->HP = (H <= 64) ? 64 : 128; m->L = L; m->nch = 4 * H / 128;
    m->max_batch = max_batch;
    m->conservative = conservative ? 1 : 0;
    m->modes_default = intr && !dist && !absc;
    m->attn_mma = dist;                      // the squared-distance channel lives in the HMMA attention path
    const int HP = m->HP, nch = m->nch;
    if ((4 * H) % 128) { delete m; return fail(DFF_EINVAL, "4*hidden must be a multiple of 128"); }

    const float* Wn = w[0]; const float* bn = w[1]; const float* We = w[2]; const float* be = w[3];
    const float* Wd = w[4]; const float* bd = w[5];

    Packer P;
    struct LayerOff { size_t ln1_g, ln1_b, bqkv[2], A, cvec, bo, g1a, g1b, ln2_g, ln2_b, b1, b2, g2a, g2b; };
    std::vector<LayerOff> lo(L);
    const int n_dec = conservative ? 1 : 3;                  // node_decoder rows: Linear(H, 1) or Linear(H, 3)
    const size_t o_emb = P.alloc((size_t)N * H), o_embt = P.alloc(H), o_dec = P.alloc((size_t)n_dec * H), o_embx = P.alloc((size_t)3 * H);
    for (int i = 0; i < N; ++i)
        for (int d = 0; d < H; ++d) P.buf[o_emb + (size_t)i * H + d] = Wn[(size_t)d * in_node + i] + bn[d];
    for (int d = 0; d < H; ++d) P.buf[o_embt + d] = Wn[(size_t)d * in_node + in_node - 1];
    if (absc)                       // node input [onehot_i, x_i, t] (graph_transformer.py:99-100): columns N .. N + 2 act on x_i
        for (int c2 = 0; c2 < 3; ++c2)
            for (int d = 0; d < H; ++d) P.buf[o_embx + (size_t)c2 * H + d] = Wn[(size_t)d * in_node + N + c2];
    for (int d = 0; d < n_dec * H; ++d) P.buf[o_dec + d] = Wd[d];

    auto LW = [&](int l, int k) { return w[DFF_NUM_GLOBAL_WEIGHTS + DFF_NUM_LAYER_WEIGHTS * l + k]; };
    const int HCs[2] = {2, 1};      // heads per chunk of configuration 0 (R=32) and 1 (R=64)
    for (int l = 0; l < L; ++l) {
        LayerOff& o = lo[l];
        const float *ln1g = LW(l, 0), *ln1b = LW(l, 1), *bq = LW(l, 3), *bkv = LW(l, 5), *Wekv = LW(l, 6), *bekv = LW(l, 7),
                    *bo = LW(l, 9), *g1 = LW(l, 10), *ln2g = LW(l, 11), *ln2b = LW(l, 12), *b1 = LW(l, 14), *b2 = LW(l, 16),
                    *g2 = LW(l, 17);
        o.ln1_g = P.alloc(H); o.ln1_b = P.alloc(H); o.bqkv[0] = P.alloc(8 * 192); o.bqkv[1] = P.alloc(8 * 192);
        o.A = P.alloc(512 * 4); o.cvec = P.alloc(512);
        o.bo = P.alloc(HP); o.g1a = P.alloc(H); o.g1b = P.alloc(H); o.ln2_g = P.alloc(H); o.ln2_b = P.alloc(H);
        o.b1 = P.alloc(4 * H); o.b2 = P.alloc(HP); o.g2a = P.alloc(H); o.g2b = P.alloc(H);
        for (int d = 0; d < H; ++d) {
            P.buf[o.ln1_g + d] = ln1g[d]; P.buf[o.ln1_b + d] = ln1b[d];
            P.buf[o.ln2_g + d] = ln2g[d]; P.buf[o.ln2_b + d] = ln2b[d];
            P.buf[o.bo + d] = bo[d]; P.buf[o.b2 + d] = b2[d];
            P.buf[o.g1a + d] = g1[d] + g1[2 * H + d]; P.buf[o.g1b + d] = g1[H + d] - g1[2 * H + d];
            P.buf[o.g2a + d] = g2[d] + g2[2 * H + d]; P.buf[o.g2b + d] = g2[H + d] - g2[2 * H + d];
        }
        for (int j = 0; j < 4 * H; ++j) P.buf[o.b1 + j] = b1[j];
        for (int ci = 0; ci < 2; ++ci) {           // bias of chunk c: [q of its heads | k | v]
            const int CWQ = 64 * HCs[ci];
            for (int c2 = 0; c2 < 512 / CWQ; ++c2)
                for (int j = 0; j < CWQ; ++j) {
                    P.buf[o.bqkv[ci] + c2 * 3 * CWQ + j] = bq[c2 * CWQ + j];
                    P.buf[o.bqkv[ci] + c2 * 3 * CWQ + CWQ + j] = bkv[c2 * CWQ + j];
                    P.buf[o.bqkv[ci] + c2 * 3 * CWQ + 2 * CWQ + j] = bkv[512 + c2 * CWQ + j];
                }
        }
        // fold edge_embedding into edges_to_kv (exact: no nonlinearity between them, graph_transformer.py:96,235,288)
        for (int j = 0; j < 512; ++j) {
            double a[4] = {0, 0, 0, 0}, cc = bekv[j];
            for (int k = 0; k < H; ++k) {
                const double wk = Wekv[(size_t)j * H + k];
                for (int e = 0; e < in_edge; ++e) a[e] += wk * We[(size_t)k * in_edge + e];
                cc += wk * be[k];
            }
            // columns of the folded map: x_j - x_i (intrinsic) and |x_j - x_i|^2 (distances); the placeholder zero feature has none
            const double ad = (intr && dist) ? a[3] : (dist ? a[0] : 0.0);
            for (int e = 0; e < 3; ++e) P.buf[o.A + j * 4 + e] = intr ? (float)a[e] : 0.f;
            P.buf[o.A + j * 4 + 3] = (float)ad;
            P.buf[o.cvec + j] = (float)cc;
        }
    }

    // GEMM panels [K][NC] in the exact order the kernel consumes them, once per configuration
    std::vector<SegHost> segs[2];
    int nseg_fwd[2] = {0, 0};
    uint32_t nslice_fwd[2] = {0, 0}, nslice_all[2] = {0, 0};
    for (int ci = 0; ci < 2; ++ci) {
        const int CWQ = 64 * HCs[ci], NCHK = 512 / CWQ;
        auto panel = [&](int K, int NC, auto&& fill /* (k, c) -> value */) {
            const int NCP = NC + kWPad;          // padded row stride: conflict-free MMA B-fragment loads
            const size_t o = P.alloc((size_t)K * NCP);
            for (int k = 0; k < K; ++k)
                for (int c2 = 0; c2 < NC; ++c2) P.buf[o + (size_t)k * NCP + c2] = fill(k, c2);
            const int KS = ks_for(NC);     // must equal gemm_acc's KS (slice multiplier KM = 1 in every shipped configuration)
            segs[ci].push_back({o, (uint32_t)(KS * NCP * sizeof(float)), (uint32_t)(K / KS)});
        };
        for (int l = 0; l < L; ++l) {
            const float *Wq = LW(l, 2), *Wkv = LW(l, 4), *Wo = LW(l, 8), *W1 = LW(l, 13), *W2 = LW(l, 15);
            for (int hc = 0; hc < NCHK; ++hc) {
                panel(H, 3 * CWQ, [&](int k, int c2) {
                    const int t = c2 / CWQ, j = c2 % CWQ;
                    return t == 0 ? Wq[(size_t)(hc * CWQ + j) * H + k]
                                  : Wkv[(size_t)((t == 1 ? 0 : 512) + hc * CWQ + j) * H + k];
                });
                panel(CWQ, HP, [&](int k, int d) { return d < H ? Wo[(size_t)d * 512 + hc * CWQ + k] : 0.f; });
            }
            for (int ch = 0; ch < nch; ++ch) {
                panel(H, 128, [&](int k, int j) { return W1[(size_t)(ch * 128 + j) * H + k]; });
                panel(128, HP, [&](int k, int d) { return d < H ? W2[(size_t)d * 4 * H + ch * 128 + k] : 0.f; });
            }
        }
        nseg_fwd[ci] = (int)segs[ci].size();
        for (int l = L - 1; l >= 0; --l) {
            const float *Wq = LW(l, 2), *Wkv = LW(l, 4), *Wo = LW(l, 8), *W1 = LW(l, 13), *W2 = LW(l, 15);
            for (int ch = 0; ch < nch; ++ch) {
                panel(H, 128, [&](int d, int j) { return W2[(size_t)d * 4 * H + ch * 128 + j]; });
                panel(128, HP, [&](int j, int d) { return d < H ? W1[(size_t)(ch * 128 + j) * H + d] : 0.f; });
            }
            for (int hc = 0; hc < NCHK; ++hc) {
                panel(H, CWQ, [&](int d, int j) { return Wo[(size_t)d * 512 + hc * CWQ + j]; });
                if (l > 0)
                    panel(3 * CWQ, HP, [&](int k, int d) {   // rows: dv' | dk' | dq of the chunk's heads
                        if (d >= H) return 0.f;
                        const int t = k / CWQ, j = k % CWQ;
                        return t == 0 ? Wkv[(size_t)(512 + hc * CWQ + j) * H + d]
                             : t == 1 ? Wkv[(size_t)(hc * CWQ + j) * H + d]
                                      : Wq[(size_t)(hc * CWQ + j) * H + d];
                    });
            }
        }
        for (int i = 0; i < (int)segs[ci].size(); ++i) {
            if (i < nseg_fwd[ci]) nslice_fwd[ci] += segs[ci][i].n_slices;
            nslice_all[ci] += segs[ci][i].n_slices;
        }
    }


    // ---- tcgen05 configuration (hidden 64 / 96 / 128, N <= 32): job table in issue order + canonical, pre-split weight
    // panels.  A panel is K/ks slices; a slice is [hi image | lo image], each the UMMA K-major no-swizzle layout of a
    // [n rows x ks] tile: float offset ((k / 4) * n + row) * 4 + k % 4.  hi = rn_tf32(w), lo = w - hi.
    struct TcJobHost { size_t offset; v2::TcJob j; };
    std::vector<TcJobHost> tcj;
    int tc_njobs_fwd = 0;
    uint32_t tc_nslice_fwd = 0, tc_nslice_all = 0;
    const bool tc_shape = ((H == 64 && N <= 64) || ((H == 96 || H == 128) && N <= 56));   // shared-memory budget of TcCfg
    if (tc_shape) {
        const int stage_bytes = (HP == 64 ? 4096 : 3072) * 4;            // must equal TcCfg::kStageFloats
        auto job = [&](int K, int NN, auto&& fill /* (k, n) -> value */, int d_col, uint32_t flags, int c_col = -1) {
            int KS = 8;
            for (int cand = 8; cand <= K && cand <= 248; cand += 8)
                if (K % cand == 0 && 2 * cand * NN * 4 <= stage_bytes) KS = cand;
            const size_t o = P.alloc((size_t)2 * K * NN);
            for (int k = 0; k < K; ++k)
                for (int n = 0; n < NN; ++n) {
                    const float v = fill(k, n);
                    uint32_t bits;
                    memcpy(&bits, &v, 4);
                    bits = (bits + 0x1000u) & 0xffffe000u;
                    float hi;
                    memcpy(&hi, &bits, 4);
                    const int s = k / KS, kk = k % KS;
                    const size_t at = o + (size_t)s * 2 * KS * NN + (size_t)((kk / 4) * NN + n) * 4 + (kk % 4);
                    P.buf[at] = hi;
                    P.buf[at + (size_t)KS * NN] = v - hi;
                }
            TcJobHost h{};
            h.offset = o;
            h.j.slice_16b = (uint16_t)((2 * KS * NN * 4) / 16);
            h.j.n_slices = (uint8_t)(K / KS); h.j.ks = (uint8_t)KS; h.j.n = (uint16_t)NN; h.j.d_col = (uint16_t)d_col;
            h.j.flags = (uint8_t)flags;
            h.j.c_col = (uint16_t)((DFF_TC_SPLIT_ACC && c_col >= 0) ? c_col : d_col);   // correction accumulator (TcCfg::kCorr*)
            tcj.push_back(h);
        };
        using namespace v2;
        const int cD = HP, cA = (int)kColAcc;          // TMEM work area starts after the HP-column block accumulator
        const int cFF = cD + 256, cOut = HP == 64 ? cD + 384 : -1, cDn = HP == 64 ? cD + 128 + 2 * HP : -1;   // == TcCfg::kCorrFF / kCorrOut / kCorrDn
        const int nch64 = 4 * H / 64;                  // FF hidden chunks of 64 columns; "supers" of <= 4 chunks share the work area
        // a 256-column super is issued as 192 + 64 column jobs when the stage holds 12 KB: both fill their weight slices completely
        // ([hi | lo] of 8 x 192 and 24 x 64 floats), 128 + 128 would leave a third of every stage (and of the bytes in flight) unused
        const int ff_split = (getenv("DFF_FF_SPLIT") ? atoi(getenv("DFF_FF_SPLIT")) : (HP == 64 ? 128 : 192));
        for (int l = 0; l < L; ++l) {
            const float *Wq = LW(l, 2), *bq = LW(l, 3), *Wkv = LW(l, 4), *bkv = LW(l, 5), *Wo = LW(l, 8), *W1 = LW(l, 13), *W2 = LW(l, 15);
            const size_t oA = lo[l].A;
            auto qkv = [&](int hc, uint32_t fl) {
                job(H + 8, 192, [&](int k, int n) -> float {
                    const int t = n / 64, j = hc * 64 + n % 64;
                    if (k < H) return t == 0 ? Wq[(size_t)j * H + k] : Wkv[(size_t)((t == 1 ? 0 : 512) + j) * H + k];
                    if (k < H + 3) return t == 0 ? 0.f : P.buf[oA + (size_t)j * 4 + (k - H)];      // + A x_j  (k', v' only)
                    if (k == H + 3) return t == 0 ? bq[j] : bkv[(t == 1 ? 0 : 512) + j];           // bias rides on the ones column
                    return 0.f;
                }, cD, TCJ_DBUF | fl);
            };
            auto outp = [&](int hc) {
                job(64, HP, [&](int k, int d) -> float { return d < H ? Wo[(size_t)d * 512 + hc * 64 + k] : 0.f; }, cA,
                    TCJ_SLOT | TCJ_WAIT_POST | (hc > 0 ? TCJ_ACC : 0) | (hc == 7 ? TCJ_COMMIT_ACC : 0), cOut);
            };
            qkv(0, TCJ_WAIT_POST); qkv(1, 0);
            for (int hc = 0; hc < 8; ++hc) {
                outp(hc);
                if (hc + 2 < 8) qkv(hc + 2, 0);
            }
            // FF1: every super (<= 256 hidden columns) as 128-column jobs into the work area; the first job of a super waits
            // for a post (LN2 output ready / work area drained), the last one commits d1_ready
            for (int c0 = 0; c0 < nch64; c0 += 4) {
                const int w = std::min(4, nch64 - c0) * 64;
                for (int q = 0; q < w; q += ff_split) {
                    const int wq = std::min(ff_split, w - q);
                    job(H, wq, [&](int k, int j) -> float { return W1[(size_t)(c0 * 64 + q + j) * H + k]; }, cD + q,
                        (q == 0 ? TCJ_WAIT_POST : 0) | (q + ff_split >= w ? TCJ_COMMIT_D1 : 0));
                }
            }
            for (int c2 = 0; c2 < nch64; ++c2)
                job(64, HP, [&](int k, int d) -> float { return d < H ? W2[(size_t)d * 4 * H + c2 * 64 + k] : 0.f; }, cA,
                    TCJ_SLOT | TCJ_WAIT_POST | (c2 > 0 ? TCJ_ACC : 0) | (c2 == nch64 - 1 ? TCJ_COMMIT_ACC : 0), cFF);
        }
        tc_njobs_fwd = (int)tcj.size();
        for (int l = L - 1; l >= 0; --l) {
            const float *Wq = LW(l, 2), *Wkv = LW(l, 4), *Wo = LW(l, 8), *W1 = LW(l, 13), *W2 = LW(l, 15);
            for (int c0 = 0; c0 < nch64; c0 += 4) {
                const int w = std::min(4, nch64 - c0) * 64;
                for (int q = 0; q < w; q += ff_split) {
                    const int wq = std::min(ff_split, w - q);
                    job(H, wq, [&](int d, int j) -> float { return W2[(size_t)d * 4 * H + c0 * 64 + q + j]; }, cD + q,
                        (q == 0 ? TCJ_WAIT_POST : 0) | (q + ff_split >= w ? TCJ_COMMIT_D1 : 0));
                }
            }
            for (int c2 = 0; c2 < nch64; ++c2)
                job(64, HP, [&](int j, int d) -> float { return d < H ? W1[(size_t)(c2 * 64 + j) * H + d] : 0.f; }, cA,
                    TCJ_SLOT | TCJ_WAIT_POST | (c2 > 0 ? TCJ_ACC : 0) | (c2 == nch64 - 1 ? TCJ_COMMIT_ACC : 0), cFF);
            auto jdo = [&](int hc, uint32_t fl) {
                job(H, 64, [&](int d, int j) -> float { return Wo[(size_t)d * 512 + hc * 64 + j]; }, cD, TCJ_DBUF | fl);
            };
            jdo(0, TCJ_WAIT_POST); jdo(1, 0);
            for (int hc = 0; hc < 8; ++hc) {
                if (l > 0 || absc) {
                    // d n_hat = dq Wq + dk' Wk + dv' Wv accumulates over the 8 head chunks in THREE TMEM accumulators (one per
                    // product) that the epilogue adds in fp32: the tensor core's accumulation chain is 192 MMAs long instead of
                    // 576 (its accumulate rounding is what separates this kernel's error from the reference's own)
                    job(64, HP, [&](int j, int d) -> float { return d < H ? Wq[(size_t)(hc * 64 + j) * H + d] : 0.f; }, cA,
                        TCJ_SLOT | TCJ_WAIT_POST | (hc > 0 ? TCJ_ACC : 0), cDn);
                    job(64, HP, [&](int j, int d) -> float { return d < H ? Wkv[(size_t)(hc * 64 + j) * H + d] : 0.f; }, cD + 128,
                        TCJ_SLOT | TCJ_WAIT_POST | (hc > 0 ? TCJ_ACC : 0), cDn < 0 ? -1 : cDn + HP);
                    job(64, HP, [&](int j, int d) -> float { return d < H ? Wkv[(size_t)(512 + hc * 64 + j) * H + d] : 0.f; }, cD + 128 + HP,
                        TCJ_SLOT | TCJ_WAIT_POST | (hc > 0 ? TCJ_ACC : 0) | (hc == 7 ? TCJ_COMMIT_ACC : 0), cDn < 0 ? -1 : cDn + 2 * HP);
                }
                if (hc + 2 < 8) jdo(hc + 2, 0);
            }
        }
        for (int i = 0; i < (int)tcj.size(); ++i) {
            if (i < tc_njobs_fwd) tc_nslice_fwd += tcj[i].j.n_slices;
            tc_nslice_all += tcj[i].j.n_slices;
        }
    }

    auto cleanup = [&]() { dff_model_destroy(m); };
    if (cudaMalloc(&m->d_weights, P.buf.size() * sizeof(float)) != cudaSuccess) { cleanup(); return fail(DFF_ENOMEM, "cudaMalloc weights failed"); }
    if (cudaMemcpy(m->d_weights, P.buf.data(), P.buf.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { cleanup(); return fail(DFF_ECUDA, "weight upload failed"); }
    for (int ci = 0; ci < 2; ++ci) {
        std::vector<Seg> hs(segs[ci].size());
        for (size_t i = 0; i < hs.size(); ++i) hs[i] = Seg{m->d_weights + segs[ci][i].offset, segs[ci][i].slice_bytes, segs[ci][i].n_slices};
        if (cudaMalloc(&m->d_segs[ci], hs.size() * sizeof(Seg)) != cudaSuccess) { cleanup(); return fail(DFF_ENOMEM, "cudaMalloc segs failed"); }
        if (cudaMemcpy(m->d_segs[ci], hs.data(), hs.size() * sizeof(Seg), cudaMemcpyHostToDevice) != cudaSuccess) { cleanup(); return fail(DFF_ECUDA, "segment upload failed"); }
    }


    if (tc_shape) {      // job tables longer than the shared-memory cache (v2::kJobCap) are read from global memory past the cap
        std::vector<v2::TcJob> hj(tcj.size());
        for (size_t i = 0; i < hj.size(); ++i) { hj[i] = tcj[i].j; hj[i].w_off = (uint32_t)tcj[i].offset; }
        if (cudaMalloc(&m->d_jobs, hj.size() * sizeof(v2::TcJob)) != cudaSuccess) { cleanup(); return fail(DFF_ENOMEM, "cudaMalloc jobs failed"); }
        if (cudaMemcpy(m->d_jobs, hj.data(), hj.size() * sizeof(v2::TcJob), cudaMemcpyHostToDevice) != cudaSuccess) { cleanup(); return fail(DFF_ECUDA, "job table upload failed"); }
        m->tc.jobs = m->d_jobs; m->tc.wbase = m->d_weights; m->tc.njobs_fwd = tc_njobs_fwd; m->tc.njobs_all = (int)hj.size();
        m->tc.nslice_fwd = tc_nslice_fwd; m->tc.nslice_all = tc_nslice_all;
        m->tc_ok = true;
#ifdef DFF_TC_PROFILE
        if (cudaMalloc(&m->tc.dbg, ((size_t)m->num_sms * 16 + 32) * sizeof(long long)) != cudaSuccess) m->tc.dbg = nullptr;
        else cudaMemset(m->tc.dbg, 0, ((size_t)m->num_sms * 16 + 32) * sizeof(long long));
#endif
    }

    stash_geometry(32, m->NP, H, m->off[0], &m->layer_floats[0]);
    stash_geometry(64, m->NP, H, m->off[1], &m->layer_floats[1]);
    m->scratch_per_cta = (long long)L * m->layer_floats[1] + 64LL * H + 64LL * (128 + 4);   // + node stream [64][132] (tc kernel, hidden > 64)
    const int n_ctas = std::min(2 * m->num_sms, std::max(1, max_batch));
    m->scratch_ctas = n_ctas;
    if (cudaMalloc(&m->d_scratch, (size_t)n_ctas * m->scratch_per_cta * sizeof(float)) != cudaSuccess) { cleanup(); return fail(DFF_ENOMEM, "cudaMalloc scratch failed"); }
    if (cudaMalloc(&m->d_flags, sizeof(uint32_t)) != cudaSuccess) { cleanup(); return fail(DFF_ENOMEM, "cudaMalloc flags failed"); }

    for (int ci = 0; ci < 2; ++ci) {
        ModelDev& M = m->md[ci];
        M = ModelDev{};
        M.N = N; M.NP = m->NP; M.H = H; M.L = L; M.S = 1; M.nch = nch;
        M.emb = m->d_weights + o_emb; M.embt = m->d_weights + o_embt; M.dec_w = m->d_weights + o_dec; M.dec_b = bd[0];
        M.conservative = m->conservative;
        M.edge_dist = dist ? 1 : 0; M.abs_coords = absc ? 1 : 0; M.embx = m->d_weights + o_embx;
        for (int i = 0; i < 3; ++i) M.dec_b3[i] = conservative ? 0.f : bd[i];
        for (int l = 0; l < L; ++l) {
            const LayerOff& o = lo[l];
            LayerDev& D = M.layer[l];
            const float* b = m->d_weights;
            D.ln1_g = b + o.ln1_g; D.ln1_b = b + o.ln1_b; D.bqkv = b + o.bqkv[ci]; D.A = b + o.A; D.cvec = b + o.cvec; D.bo = b + o.bo;
            D.g1a = b + o.g1a; D.g1b = b + o.g1b; D.ln2_g = b + o.ln2_g; D.ln2_b = b + o.ln2_b; D.b1 = b + o.b1; D.b2 = b + o.b2;
            D.g2a = b + o.g2a; D.g2b = b + o.g2b;
        }
        M.segs = m->d_segs[ci]; M.nseg_fwd = nseg_fwd[ci]; M.nseg_all = (int)segs[ci].size();
        M.nslice_fwd = nslice_fwd[ci]; M.nslice_all = nslice_all[ci];
        M.scratch = m->d_scratch; M.scratch_per_cta = m->scratch_per_cta;
        M.layer_floats = m->layer_floats[ci];
        for (int i = 0; i < ST_COUNT; ++i) M.off[i] = m->off[ci][i];
    }
    *out = m;
    return DFF_OK;
}

void dff_model_destroy(dff_model_t* m) {
    if (!m) return;
    DeviceGuard dg_(m->device);
    if (m->d_weights) cudaFree(m->d_weights);
    for (auto p : m->d_segs) if (p) cudaFree(p);
    if (m->d_scratch) cudaFree(m->d_scratch);
    if (m->d_flags) cudaFree(m->d_flags);
    if (m->d_sched) cudaFree(m->d_sched);
    if (m->d_jobs) cudaFree(m->d_jobs);
    for (auto p : m->d_io) if (p) cudaFree(p);
    delete m;
}

int dff_model_num_beads(const dff_model_t* m) { return m ? m->N : 0; }
int dff_model_hidden(const dff_model_t* m) { return m ? m->H : 0; }
int dff_model_layers(const dff_model_t* m) { return m ? m->L : 0; }
int dff_model_device(const dff_model_t* m) { return m ? m->device : -1; }
int64_t dff_model_launch_count(const dff_model_t* m) { return m ? m->launches : 0; }
const char* dff_model_last_config(const dff_model_t* m) { return m ? m->last_cfg : "none"; }

double dff_model_flops_per_sample(const dff_model_t* m) {
    if (!m) return 0;
    const double N = m->N, H = m->H, L = m->L, I = 512;
    const double per_layer = 2 * N * H * 3 * I + 2 * (2 * N * I * 3) + 2 * 2 * N * N * (I + 24) + 2 * N * I * H + 16 * N * H * H + 12 * N * H;
    const double fwd = L * per_layer + 2 * N * H;
    return fwd + (fwd - 2 * N * H * 3 * I);
}

int dff_score_dev(dff_model_t* m, const float* x_dev, float t_norm, int batch, float* eps_out_dev,
                  float* energy_out_dev, void* stream) {
    if (!m || !x_dev) return fail(DFF_EINVAL, "NULL model or x");
    if (!m->conservative && energy_out_dev) return fail(DFF_EINVAL, "a non-conservative network has no energy output (graph_transformer.py:107-113)");
    StepArgs A{};
    A.mode = MODE_SCORE; A.B = batch; A.n_steps = 1; A.need_backward = m->conservative && eps_out_dev != nullptr;
    A.x = const_cast<float*>(x_dev); A.eps_out = eps_out_dev; A.energy_out = energy_out_dev; A.t_norm = t_norm;
    return launch(m, A, (cudaStream_t)stream);
}

int dff_score_dev_t(dff_model_t* m, const float* x_dev, const float* t_norm_dev, int batch, float* eps_out_dev,
                    float* energy_out_dev, void* stream) {
    if (!m || !x_dev || !t_norm_dev) return fail(DFF_EINVAL, "NULL model, x or t");
    if (!m->conservative && energy_out_dev) return fail(DFF_EINVAL, "a non-conservative network has no energy output (graph_transformer.py:107-113)");
    StepArgs A{};
    A.mode = MODE_SCORE; A.B = batch; A.n_steps = 1; A.need_backward = m->conservative && eps_out_dev != nullptr;
    A.x = const_cast<float*>(x_dev); A.eps_out = eps_out_dev; A.energy_out = energy_out_dev; A.t_norm = 0.f; A.t_rows = t_norm_dev;
    return launch(m, A, (cudaStream_t)stream);
}

int dff_score_host(dff_model_t* m, const float* x_host, float t_norm, int batch, float* eps_out_host,
                   float* energy_out_host) {
    if (!m || !x_host) return fail(DFF_EINVAL, "NULL model or x");
    DEVICE_SCOPE(m->device);
    const size_t n = (size_t)batch * m->N;
    int rc;
    if ((rc = ensure_io(m, 0, n * 3)) || (rc = ensure_io(m, 1, n * 3)) || (rc = ensure_io(m, 2, n))) return rc;
    CUDA_TRY(cudaMemcpyAsync(m->d_io[0], x_host, n * 3 * sizeof(float), cudaMemcpyHostToDevice, 0));
    rc = dff_score_dev(m, m->d_io[0], t_norm, batch, eps_out_host ? m->d_io[1] : nullptr, energy_out_host ? m->d_io[2] : nullptr, nullptr);
    if (rc) return rc;
    if (eps_out_host) CUDA_TRY(cudaMemcpyAsync(eps_out_host, m->d_io[1], n * 3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (energy_out_host) CUDA_TRY(cudaMemcpyAsync(energy_out_host, m->d_io[2], n * sizeof(float), cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return DFF_OK;
}

int dff_ddpm_steps_dev(dff_model_t* m, float* x_dev, int batch, int t_start, int n_steps, int T,
                       const float* const* sched_dev, const float* noise_dev, uint64_t seed, uint64_t offset,
                       uint32_t* flags_dev, void* stream) {
    if (!m || !x_dev || !sched_dev) return fail(DFF_EINVAL, "NULL argument");
    if (n_steps < 0 || t_start >= T || t_start - n_steps + 1 < 0)
        return fail(DFF_EINVAL, "timestep range [%d..%d] outside schedule of length %d", t_start - n_steps + 1, t_start, T);
    if (n_steps == 0) return DFF_OK;
    StepArgs A{};
    A.mode = MODE_DDPM; A.B = batch; A.n_steps = n_steps; A.need_backward = m->conservative;
    A.x = x_dev; A.noise = noise_dev; A.t_start = t_start; A.T = T;
    for (int i = 0; i < 5; ++i) { if (!sched_dev[i]) return fail(DFF_EINVAL, "schedule pointer %d is NULL", i); A.sched[i] = sched_dev[i]; }
    A.seed = seed; A.offset = offset; A.flags = flags_dev;
    return launch(m, A, (cudaStream_t)stream);
}

int dff_langevin_steps_dev(dff_model_t* m, float* x_dev, float* v_dev, int batch, int n_steps,
                           const dff_md_params_t* p, const float* mass_dev, const float* noise_dev, uint64_t seed,
                           uint64_t offset, int save_interval, float* frames_dev, float* ke_dev, uint32_t* flags_dev,
                           void* stream) {
    if (!m || !x_dev || !p) return fail(DFF_EINVAL, "NULL argument");
    if (p->integrator == DFF_MD_BAOAB && (!v_dev || !mass_dev)) return fail(DFF_EINVAL, "BAOAB needs velocities and masses");
    if (p->integrator != DFF_MD_BAOAB && p->integrator != DFF_MD_BROWNIAN) return fail(DFF_EINVAL, "unknown integrator %d", p->integrator);
    if (save_interval > 0 && n_steps % save_interval) return fail(DFF_EINVAL, "save_interval must divide n_steps (langevin_cgnet.py:306-309)");
    if (n_steps <= 0) return DFF_OK;
    StepArgs A{};
    A.mode = p->integrator == DFF_MD_BAOAB ? MODE_BAOAB : MODE_BROWNIAN;
    A.B = batch; A.n_steps = n_steps; A.need_backward = m->conservative;
    A.x = x_dev; A.v = (A.mode == MODE_BAOAB) ? v_dev : nullptr; A.noise = noise_dev;
    A.t_norm = p->t_norm; A.force_scale = p->force_scale; A.dt = p->dt; A.vscale = p->vscale; A.noisescale = p->noisescale;
    A.inv_beta = (float)(1.0 / (double)p->beta); A.dtau = p->dtau;
    A.bd_sigma = (float)sqrt(2.0 * (double)p->dtau / (double)p->beta);
    A.mass = mass_dev; A.save_interval = save_interval; A.frames = frames_dev; A.ke = ke_dev;
    A.seed = seed; A.offset = offset; A.flags = flags_dev;
    return launch(m, A, (cudaStream_t)stream);
}

int dff_ddpm_sample_host(dff_model_t* m, float* x_host, int batch, int T, const float* const* sched_host, uint64_t seed,
                         uint32_t* flags_host) {
    if (!m || !x_host || !sched_host) return fail(DFF_EINVAL, "NULL argument");
    DEVICE_SCOPE(m->device);
    const size_t n3 = (size_t)batch * m->N * 3;
    int rc;
    if ((rc = ensure_io(m, 0, n3))) return rc;
    if (m->d_sched_T != (size_t)T) {
        if (m->d_sched) cudaFree(m->d_sched);
        m->d_sched = nullptr; m->d_sched_T = 0;
        CUDA_TRY(cudaMalloc(&m->d_sched, 5 * (size_t)T * sizeof(float)));
        m->d_sched_T = T;
    }
    const float* sp[5];
    for (int i = 0; i < 5; ++i) {
        CUDA_TRY(cudaMemcpyAsync(m->d_sched + (size_t)i * T, sched_host[i], T * sizeof(float), cudaMemcpyHostToDevice, 0));
        sp[i] = m->d_sched + (size_t)i * T;
    }
    CUDA_TRY(cudaMemcpyAsync(m->d_io[0], x_host, n3 * sizeof(float), cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemsetAsync(m->d_flags, 0, sizeof(uint32_t), 0));
    if ((rc = dff_ddpm_steps_dev(m, m->d_io[0], batch, T - 1, T, T, sp, nullptr, seed, 0, m->d_flags, nullptr))) return rc;
    CUDA_TRY(cudaMemcpyAsync(x_host, m->d_io[0], n3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (flags_host) CUDA_TRY(cudaMemcpyAsync(flags_host, m->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return DFF_OK;
}

int dff_langevin_run_host(dff_model_t* m, float* x_host, float* v_host, int batch, int n_steps, const dff_md_params_t* p,
                          const float* mass_host, uint64_t seed, int save_interval, float* frames_host, float* ke_host,
                          uint32_t* flags_host) {
    if (!m || !x_host || !p) return fail(DFF_EINVAL, "NULL argument");
    DEVICE_SCOPE(m->device);
    const size_t n3 = (size_t)batch * m->N * 3;
    const size_t nf = save_interval > 0 ? (size_t)(n_steps / save_interval) : 0;
    int rc;
    if ((rc = ensure_io(m, 0, n3)) || (rc = ensure_io(m, 1, n3)) || (rc = ensure_io(m, 2, m->N))) return rc;
    if (frames_host && nf && (rc = ensure_io(m, 3, nf * n3))) return rc;
    if (ke_host && nf && (rc = ensure_io(m, 4, nf * batch))) return rc;
    CUDA_TRY(cudaMemcpyAsync(m->d_io[0], x_host, n3 * sizeof(float), cudaMemcpyHostToDevice, 0));
    if (v_host) CUDA_TRY(cudaMemcpyAsync(m->d_io[1], v_host, n3 * sizeof(float), cudaMemcpyHostToDevice, 0));
    else CUDA_TRY(cudaMemsetAsync(m->d_io[1], 0, n3 * sizeof(float), 0));
    if (mass_host) CUDA_TRY(cudaMemcpyAsync(m->d_io[2], mass_host, m->N * sizeof(float), cudaMemcpyHostToDevice, 0));
    CUDA_TRY(cudaMemsetAsync(m->d_flags, 0, sizeof(uint32_t), 0));
    rc = dff_langevin_steps_dev(m, m->d_io[0], m->d_io[1], batch, n_steps, p, m->d_io[2], nullptr, seed, 0, save_interval,
                                (frames_host && nf) ? m->d_io[3] : nullptr, (ke_host && nf) ? m->d_io[4] : nullptr, m->d_flags, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(x_host, m->d_io[0], n3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (v_host) CUDA_TRY(cudaMemcpyAsync(v_host, m->d_io[1], n3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (frames_host && nf) CUDA_TRY(cudaMemcpyAsync(frames_host, m->d_io[3], nf * n3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (ke_host && nf) CUDA_TRY(cudaMemcpyAsync(ke_host, m->d_io[4], nf * batch * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (flags_host) CUDA_TRY(cudaMemcpyAsync(flags_host, m->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, 0));
    CUDA_TRY(cudaStreamSynchronize(0));
    return DFF_OK;
}

int dff_pwd_num_pairs(int num_beads, int offset) {
    int p = 0;
    for (int i = 0; i < num_beads; ++i) p += std::max(0, num_beads - offset - i);
    return p;
}

int dff_pwd_max_dev(const float* x_dev, int n, int num_beads, int offset, float* max_out_dev, void* stream) {
    if (!x_dev || !max_out_dev) return fail(DFF_EINVAL, "NULL argument");
    const int P = dff_pwd_num_pairs(num_beads, offset);
    if (n <= 0 || P <= 0) return DFF_OK;
    if (P > 12000) return fail(DFF_EINVAL, "too many bead pairs (%d) for the shared-memory reduction", P);
    const long long total = (long long)n * P;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
    dff_pwd_max_kernel<<<grid, 256, (size_t)P * sizeof(unsigned int), (cudaStream_t)stream>>>(
        x_dev, n, num_beads, offset, P, reinterpret_cast<unsigned int*>(max_out_dev));
    CUDA_TRY(cudaGetLastError());
    return DFF_OK;
}

int dff_pwd_hist_dev(const float* x_dev, int n, int num_beads, int offset, float resolution, const int* nbins_dev,
                     int ld_hist, uint32_t* hist_dev, void* stream) {
    if (!x_dev || !nbins_dev || !hist_dev) return fail(DFF_EINVAL, "NULL argument");
    if (!(resolution > 0.f) || ld_hist <= 0) return fail(DFF_EINVAL, "resolution and ld_hist must be positive");
    const int P = dff_pwd_num_pairs(num_beads, offset);
    if (n <= 0 || P <= 0) return DFF_OK;
    const long long total = (long long)n * P;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
    dff_pwd_hist_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, num_beads, offset, P, resolution, nbins_dev, ld_hist, hist_dev);
    CUDA_TRY(cudaGetLastError());
    return DFF_OK;
}

int dff_contacts_dev(const float* x_dev, int n, int num_beads, float cutoff, const unsigned char* folded_dev, int offset,
                     uint32_t* counts_dev, uint32_t* mismatch_dev, void* stream) {
    if (!x_dev || !counts_dev) return fail(DFF_EINVAL, "NULL argument");
    if (num_beads < 2 || num_beads > 104) return fail(DFF_EINVAL, "num_beads %d unsupported for the shared-memory contact counters (2..104)", num_beads);
    if (mismatch_dev && !folded_dev) return fail(DFF_EINVAL, "mismatch counts need the folded contact map");
    if (n <= 0) return DFF_OK;
    const long long total = (long long)n * num_beads * num_beads;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 8);
    dff_contacts_kernel<<<grid, 256, (size_t)num_beads * num_beads * sizeof(unsigned int), (cudaStream_t)stream>>>(
        x_dev, n, num_beads, cutoff, folded_dev, offset, counts_dev, mismatch_dev);
    CUDA_TRY(cudaGetLastError());
    return DFF_OK;
}

int dff_dihedrals_dev(const float* x_dev, int n, int num_beads, const int* quads_dev, int nbins_axis, float* torsions_dev,
                      uint32_t* hist_dev, void* stream) {
    if (!x_dev || !quads_dev) return fail(DFF_EINVAL, "NULL argument");
    if (num_beads < 4 || (hist_dev && nbins_axis < 1)) return fail(DFF_EINVAL, "need >= 4 beads and >= 1 bin per axis");
    if (n <= 0) return DFF_OK;
    const int grid = std::min((n + 255) / 256, 148 * 8);
    dff_dihedral_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, num_beads, quads_dev, nbins_axis, torsions_dev, hist_dev);
    CUDA_TRY(cudaGetLastError());
    return DFF_OK;
}

int dff_rmsd_dev(const float* x_dev, int n, int num_beads, const float* ref_dev, float* rmsd_dev, void* stream) {
    if (!x_dev || !ref_dev || !rmsd_dev) return fail(DFF_EINVAL, "NULL argument");
    if (num_beads < 1) return fail(DFF_EINVAL, "num_beads must be positive");
    if (n <= 0) return DFF_OK;
    const int grid = (int)std::min<long long>(((long long)n * 32 + 255) / 256, 148LL * 8);
    dff_rmsd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, num_beads, ref_dev, rmsd_dev);
    CUDA_TRY(cudaGetLastError());
    return DFF_OK;
}

int64_t dff_debug_read_stash(dff_model_t* m, float* out_host, int64_t cap) {
    if (!m || !out_host) return fail(DFF_EINVAL, "NULL argument");
    DEVICE_SCOPE(m->device);
    const int64_t n = std::min<int64_t>(cap, m->scratch_per_cta);
    if (cudaDeviceSynchronize() != cudaSuccess) return fail(DFF_ECUDA, "sync failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (cudaMemcpy(out_host, m->d_scratch, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
        return fail(DFF_ECUDA, "stash copy failed");
    return n;
}

int dff_debug_tc_gemm(const float* a_host, const float* b_host, float* d_host, int n, int k, int reps, float* ms_out) {
    if (!a_host || !b_host || !d_host) return fail(DFF_EINVAL, "NULL argument");
    if (n < 8 || n > 256 || n % 8 || k < 8 || k % 8) return fail(DFF_EINVAL, "need 8 <= N <= 256 (multiple of 8) and K a multiple of 8");
    const size_t smem = (size_t)(64 + n) * k * 2 * sizeof(float);
    if (smem > 220 * 1024) return fail(DFF_EINVAL, "operands need %zu bytes of shared memory (> 220 KB)", smem);
    float *dA = nullptr, *dB = nullptr, *dD = nullptr;
    CUDA_TRY(cudaMalloc(&dA, 64 * (size_t)k * 4)); CUDA_TRY(cudaMalloc(&dB, (size_t)n * k * 4)); CUDA_TRY(cudaMalloc(&dD, 64 * (size_t)n * 4));
    CUDA_TRY(cudaMemcpy(dA, a_host, 64 * (size_t)k * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dB, b_host, (size_t)n * k * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaFuncSetAttribute(dff_tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    dff_tc_gemm_test_kernel<<<1, 128, smem>>>(dA, dB, dD, n, k, 1);            // warm-up
    CUDA_TRY(cudaEventRecord(e0));
    dff_tc_gemm_test_kernel<<<1, 128, smem>>>(dA, dB, dD, n, k, reps < 1 ? 1 : reps);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaDeviceSynchronize());
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (ms_out) *ms_out = ms;
    CUDA_TRY(cudaMemcpy(d_host, dD, 64 * (size_t)n * 4, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return DFF_OK;
}

int dff_debug_stash_layout(const dff_model_t* m, int* rows, int* samples, int* npad, int64_t* layer_floats, int64_t offsets[11]) {
    if (!m) return fail(DFF_EINVAL, "NULL model");
    const int ri = (m->last_R == 64) ? 1 : 0;
    if (rows) *rows = m->last_R;
    if (samples) *samples = m->last_S;
    if (npad) *npad = m->NP;
    if (layer_floats) *layer_floats = m->layer_floats[ri];
    if (offsets) for (int i = 0; i < ST_COUNT; ++i) offsets[i] = m->off[ri][i];
    return DFF_OK;
}

}  // extern "C"
