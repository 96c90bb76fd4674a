/*
 * dff_b200.h -- C ABI of the B200-native "denoising force field" hot path.
 *
 * The reference (microsoft/two-for-one-diffusion) is pure Python/PyTorch and has NO FFI of its own
 * (SURVEY.md 2a, 8b); its boundary for this path is the Python object API.  This library is what a
 * binding for that API calls: each entry point below replaces one reference call, cited as
 * file:line into the reference tree.  The Python mirror of the reference classes
 * (two-for-one-diffusion_b200/models, dynamics) binds these symbols with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success, a negative DFF_E* code on failure
 *     (message via dff_last_error()); no exceptions cross the boundary.
 *   - "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors), fp32,
 *     contiguous, 16-byte aligned; "host" pointers are ordinary host memory.
 *   - work is enqueued on the caller's stream (cudaStream_t passed as void*); no host sync inside
 *     the *_dev calls.  The *_host calls copy in, run, copy out and synchronise the stream.
 *   - one handle per (model, device); a handle is not thread-safe.
 *   - there is no CPU fallback: without a CUDA device dff_model_create fails with DFF_ENODEV.
 */
#ifndef DFF_B200_H
#define DFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFF_OK          0
#define DFF_EINVAL     -1   /* bad argument / unsupported shape */
#define DFF_ENODEV     -2   /* no CUDA device / wrong architecture */
#define DFF_ECUDA      -3   /* CUDA runtime error (see dff_last_error) */
#define DFF_ENOMEM     -4

/* Number of weight tensors expected by dff_model_create: 6 + 18 * n_layers, in this order
 * (names are the reference's state-dict keys under "ema_model.model."):
 *   0 node_embedding.weight [H, N+1]      1 node_embedding.bias [H]
 *   2 edge_embedding.weight [H, 3]        3 edge_embedding.bias [H]
 *   4 node_decoder.weight   [1, H]        5 node_decoder.bias   [1]
 * then for each layer l, with P = "graphtransformer.layers.{l}.":
 *   +0  P0.0.norm.weight [H]              +1  P0.0.norm.bias [H]
 *   +2  P0.0.fn.to_q.weight [512, H]      +3  P0.0.fn.to_q.bias [512]
 *   +4  P0.0.fn.to_kv.weight [1024, H]    +5  P0.0.fn.to_kv.bias [1024]
 *   +6  P0.0.fn.edges_to_kv.weight [512,H]+7  P0.0.fn.edges_to_kv.bias [512]
 *   +8  P0.0.fn.to_out.weight [H, 512]    +9  P0.0.fn.to_out.bias [H]
 *   +10 P0.1.proj.0.weight [1, 3H]
 *   +11 P1.0.norm.weight [H]              +12 P1.0.norm.bias [H]
 *   +13 P1.0.fn.0.weight [4H, H]          +14 P1.0.fn.0.bias [4H]
 *   +15 P1.0.fn.2.weight [H, 4H]          +16 P1.0.fn.2.bias [H]
 *   +17 P1.1.proj.0.weight [1, 3H]
 */
#define DFF_NUM_GLOBAL_WEIGHTS 6
#define DFF_NUM_LAYER_WEIGHTS  18

typedef struct dff_model dff_model_t;

/* Integrator selected by dff_langevin_steps_* (dynamics/langevin_cgnet.py:427-445). */
#define DFF_MD_BAOAB     0   /* friction given:  _langevin_timestep   (langevin_cgnet.py:447-479) */
#define DFF_MD_BROWNIAN  1   /* friction None:   _overdamped_timestep (langevin_cgnet.py:481-500) */

/* Flags OR-ed into *flags_dev by the samplers (replace the reference's per-step host syncs). */
#define DFF_FLAG_CLAMPED      1u  /* |x| > 1000 was clamped            (models/ddpm.py:248-250) */
#define DFF_FLAG_CENTER       2u  /* |mean over beads| >= 1e-3 on entry (utils.py:73-86)        */
#define DFF_FLAG_NONFINITE    4u  /* NaN/Inf coordinate seen */

const char* dff_last_error(void);
int dff_version(void);
/* Number of CUDA devices visible (0 on a CPU-only box); never fails. */
int dff_device_count(void);

/* Builds the device-resident model: folds edge_embedding into edges_to_kv (A = W_ekv W_e,
 * c = W_ekv b_e + b_ekv), lays every projection out as [K][N] panels in consumption order and
 * allocates per-CTA scratch for `max_batch` simultaneous samples.
 * Replaces: models/__init__.py:4-18 get_model + GraphTransformer.__init__ (graph_transformer.py:23-75)
 *           + load_state_dict of the "ema" weights (sample.py:157-167).
 * Supported: conservative, intrinsic-coordinate nets (every shipped checkpoint): heads=8,
 * dim_head=64, H in {32..128, multiple of 32}, N <= 64, n_layers <= 8.
 * weights_host: array of n_weights host pointers in the order above. */
int dff_model_create(dff_model_t** out, int device, int num_beads, int hidden, int n_layers,
                     const float* const* weights_host, int n_weights, int max_batch);
/* Same, for either output head of the reference network (graph_transformer.py:62-65):
 *   conservative = 1: node_decoder is Linear(H, 1), the prediction is -d sum(E)/dx (hand-written reverse pass);
 *   conservative = 0: node_decoder is Linear(H, 3) (weights_host[4] is [3, H], weights_host[5] is [3]); the prediction is
 *                     the decoder output itself, forward only; there is no energy output. */
int dff_model_create_ex(dff_model_t** out, int device, int num_beads, int hidden, int n_layers,
                        const float* const* weights_host, int n_weights, int max_batch, int conservative);
/* Every network mode of the reference constructor (graph_transformer.py:23-75, 116-140; main_train.py:149-166):
 *   use_intrinsic_coords : edge features carry x_j - x_i                (edge_embedding input 3, or 4 with distances)
 *   use_distances        : edge features carry |x_j - x_i|^2            (edge_embedding input 1, or 4 with intrinsic)
 *   use_abs_coords       : the node input is [onehot_i, x_i, t]         (node_embedding input N + 4): layer 0 depends on x
 *   neither edge flag    : the edge feature is one constant zero (edge_embedding input 1)
 * weights_host[0] is then [H, N + 1 + 3 * use_abs_coords] and weights_host[2] is [H, 3 * intrinsic + distances (or 1)].
 * The squared-distance channel is collapsed like the intrinsic one and runs on the HMMA attention path. */
typedef struct dff_model_opts {
    int conservative;
    int use_intrinsic_coords;
    int use_distances;
    int use_abs_coords;
} dff_model_opts_t;
int dff_model_create_v2(dff_model_t** out, int device, int num_beads, int hidden, int n_layers,
                        const float* const* weights_host, int n_weights, int max_batch, const dff_model_opts_t* opts);
void dff_model_destroy(dff_model_t* m);

/* Shape queries (mirror of the attributes the reference modules expose). */
int dff_model_num_beads(const dff_model_t* m);
int dff_model_hidden(const dff_model_t* m);
int dff_model_layers(const dff_model_t* m);
int dff_model_device(const dff_model_t* m);
/* Kernel-launch counter of this handle (bench.py reports it as gpu_launches). */
int64_t dff_model_launch_count(const dff_model_t* m);
/* Launch configuration the last call of this handle used: "tc" (tcgen05/TMEM kernel, csrc/dff_kernel_tc.cuh) or
 * "wide" / "tall" / "duo" (mma.sync kernel, csrc/dff_kernel.cuh); "none" before the first launch. */
const char* dff_model_last_config(const dff_model_t* m);
/* Algorithmic FLOPs of one force evaluation for one sample in the collapsed formulation the
 * kernel executes (fwd + backward w.r.t. x). */
double dff_model_flops_per_sample(const dff_model_t* m);

/* == GraphTransformer.forward(x, h=eye(N), t, return_energy)   (graph_transformer.py:77-114)
 * x_dev [B,N,3]; t_norm = diffusion index / T, shared by the batch (always true on this path:
 * ddpm.py:245, langevin.py:76).  Outputs (either may be NULL):
 *   eps_out_dev    [B,N,3]  = -d sum(E)/dx  (the "forces"/epsilon prediction, compute_forces :143-159)
 *   energy_out_dev [B,N]    per-bead energies (return_energy=True, :109-110) */
int dff_score_dev(dff_model_t* m, const float* x_dev, float t_norm, int batch,
                  float* eps_out_dev, float* energy_out_dev, void* stream);
/* Same with one noise level PER SAMPLE: t_norm_dev [B] (device).  GraphTransformer.forward embeds t per sample
 * (graph_transformer.py:91 `t.reshape(-1,1,1).repeat(1,N,1)`); the samplers always pass a uniform t, p_losses-style evaluation
 * (models/ddpm.py:296-312) does not. */
int dff_score_dev_t(dff_model_t* m, const float* x_dev, const float* t_norm_dev, int batch,
                    float* eps_out_dev, float* energy_out_dev, void* stream);
int dff_score_host(dff_model_t* m, const float* x_host, float t_norm, int batch,
                   float* eps_out_host, float* energy_out_host);

/* == n_steps iterations of GaussianDiffusion.p_sample_loop's body (models/ddpm.py:244-251):
 *    p_mean_variance (:195-219) -> p_sample (:221-232) -> clamp +-1000 (:248-250) -> center_zero (:251),
 * for timesteps t_start, t_start-1, ..., t_start-n_steps+1, in place on x_dev [B,N,3].
 * sched_dev: 5 device arrays of length T, in this order: sqrt_recip_alphas_cumprod,
 *   sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2,
 *   posterior_log_variance_clipped (the checkpoint's buffers, ddpm.py:76-99).
 * noise_dev: [n_steps,B,N,3] raw N(0,1) draws (what torch.randn_like returns at :228), or NULL to
 *   draw on the device (Philox4x32-10, keyed by seed; element counter = offset + step).
 * flags_dev: optional uint32, OR of DFF_FLAG_*. */
int dff_ddpm_steps_dev(dff_model_t* m, float* x_dev, int batch, int t_start, int n_steps, int T,
                       const float* const* sched_dev, const float* noise_dev,
                       uint64_t seed, uint64_t offset, uint32_t* flags_dev, void* stream);

/* == n_steps iterations of Langevin.simulate's loop body (dynamics/langevin_cgnet.py:737-771):
 *    center_zero (:739) -> ForcesWrapper.forward (dynamics/langevin.py:75-92) -> _timestep (:427-500)
 *    -> _save_timepoint every save_interval (:502-542).
 * x_dev [B,N,3] in place (normalised units); v_dev [B,N,3] in place (BAOAB) or NULL (Brownian).
 * force_scale = -1 / (kbt_inv * sqrt(1 - alphas_cumprod[t]))     (langevin.py:78-87)
 * mass_dev [N] bead masses.  BAOAB: dt, vscale=exp(-dt*friction), noisescale=sqrt(1-vscale^2), beta
 * (langevin_cgnet.py:327-330, :463-477).  Brownian: dtau = diffusion*dt, beta (:481-500).
 * noise_dev [n_steps,B,N,3] raw N(0,1) draws (what torch.randn returns at :470) or NULL (device Philox).
 * frames_dev [n_steps/save_interval, B,N,3] receives x_new (not re-centred, as the reference saves it),
 * ke_dev [n_steps/save_interval, B] the kinetic energies (:539-542); both may be NULL;
 * save_interval <= 0 disables saving. */
typedef struct dff_md_params {
    int   integrator;      /* DFF_MD_BAOAB / DFF_MD_BROWNIAN */
    float t_norm;          /* t / diffusion_steps */
    float force_scale;
    float dt;
    float vscale;
    float noisescale;
    float beta;
    float dtau;
} dff_md_params_t;

int dff_langevin_steps_dev(dff_model_t* m, float* x_dev, float* v_dev, int batch, int n_steps,
                           const dff_md_params_t* prm, const float* mass_dev,
                           const float* noise_dev, uint64_t seed, uint64_t offset,
                           int save_interval, float* frames_dev, float* ke_dev,
                           uint32_t* flags_dev, void* stream);

/* Host-buffer variants used by the end-to-end path (copies inside, synchronous). */
int dff_ddpm_sample_host(dff_model_t* m, float* x_host /* in: x_T, out: x_0 */, int batch, int T,
                         const float* const* sched_host, uint64_t seed, uint32_t* flags_host);
int dff_langevin_run_host(dff_model_t* m, float* x_host, float* v_host, int batch, int n_steps,
                          const dff_md_params_t* prm, const float* mass_host, uint64_t seed,
                          int save_interval, float* frames_host, float* ke_host, uint32_t* flags_host);

/* == Pairwise-distance statistics of sampled structures, the arithmetic of
 *    get_pwd_triu_batch (evaluate/evaluators.py:934-948) + the torch.histc loops of PwdEvaluator (:241-263).
 * Pairs are ordered like torch.triu_indices(N, N, offset): (i, j) with j - i >= offset, row-major.
 * x_dev [n, N, 3] fp32 (any units).  No model handle is needed.
 *   dff_pwd_num_pairs : number of pairs P
 *   dff_pwd_max_dev   : max_out_dev [P] (zero-initialised by the caller) = per-pair maximum distance over the n structures
 *   dff_pwd_hist_dev  : hist_dev [P, ld_hist] uint32 (zero-initialised by the caller) += histogram of pair p with
 *                       nbins_dev[p] <= ld_hist equal-width bins over [0, resolution * nbins[p]] (torch.histc semantics:
 *                       out-of-range values are ignored, the upper edge falls in the last bin). */
int dff_pwd_num_pairs(int num_beads, int offset);
int dff_pwd_max_dev(const float* x_dev, int n, int num_beads, int offset, float* max_out_dev, void* stream);
int dff_pwd_hist_dev(const float* x_dev, int n, int num_beads, int offset, float resolution, const int* nbins_dev,
                     int ld_hist, uint32_t* hist_dev, void* stream);

/* ---- the other structure metrics of the reference's analysis suite, on device-resident samples x_dev [n, num_beads, 3] (Angstrom)
 *
 * Contact maps (evaluate/evaluators.py:735-858, ContactEvaluator): counts_dev [N*N] += number of samples with |x_i - x_j| < cutoff
 * (caller zeroes it; normalised count = counts / n, :800-802); mismatch_dev [n] (nullable, caller zeroes) += number of pairs
 * j - i >= offset whose contact state differs from folded_dev [N*N] (0/1 bytes) -- the per-frame binary cross entropy of
 * _eval_bce_dynamics (:829-858) is 100 * mismatch / n_pairs. */
int dff_contacts_dev(const float* x_dev, int n, int num_beads, float cutoff, const unsigned char* folded_dev, int offset,
                     uint32_t* counts_dev, uint32_t* mismatch_dev, void* stream);
/* Two backbone torsions per structure (mdtraj.compute_dihedrals formula in fp32; evaluators_CGflowmatching.py:32-38) of the atom
 * quadruples quads_dev [2][4], written to torsions_dev [n][2] (nullable) and binned into hist_dev [nbins_axis][nbins_axis]
 * (nullable, caller zeroes) over np.linspace(-pi, pi, nbins_axis + 1) like get_prob's np.histogram2d (:41-51). */
int dff_dihedrals_dev(const float* x_dev, int n, int num_beads, const int* quads_dev, int nbins_axis, float* torsions_dev,
                      uint32_t* hist_dev, void* stream);
/* Minimal RMSD to ref_dev [num_beads, 3] after optimal superposition (mdtraj.rmsd of evaluators.py:655-660; Kabsch in fp64). */
int dff_rmsd_dev(const float* x_dev, int n, int num_beads, const float* ref_dev, float* rmsd_dev, void* stream);

/* Test hook: after a dff_score_* call with batch <= samples-per-CTA-group, copies the first CTA's
 * activation stash (the per-layer intermediates kept for the backward pass) to host.
 * Returns the number of floats written (<= cap) or a negative error. */
int64_t dff_debug_read_stash(dff_model_t* m, float* out_host, int64_t cap);
/* Stash geometry: rows per pass, padded N, floats per layer, and the per-layer offsets
 * {n_in, stats1, qkv, p, att, g1, m, stats2, h1, ff, g2}. */
int dff_debug_stash_layout(const dff_model_t* m, int* rows, int* samples, int* npad,
                           int64_t* layer_floats, int64_t offsets[11]);

/* Test hook for the tcgen05 building block (csrc/dff_tc.cuh): D[64,N] = A[64,K] * B[N,K]^T with 3xTF32 split
 * precision on the 5th-generation tensor cores (operands staged in shared memory, accumulator in TMEM), executed
 * `reps` times by one CTA; *ms_out receives the CUDA-event time of the timed launch. Host pointers. */
int dff_debug_tc_gemm(const float* a_host, const float* b_host, float* d_host, int n, int k, int reps, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* DFF_B200_H */
