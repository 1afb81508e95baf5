/* dne.h -- C ABI of libdne.so: the sm_100a ES/GA rollout-and-update engine.
 *
 * The reference (uber-research/deep-neuroevolution) is Python and has no C ABI for this path; its
 * "plugin boundary" is (a) the Python API es_distributed.{es,ga,nses}.run_master/run_worker +
 * policies.Policy + SharedNoiseTable and (b) on its GPU path the TF custom-op registry
 * (gpu_implementation/gym_tensorflow/ops/indexedmatmul.cpp:303-344).  Each entry point below names the
 * reference code it replaces (paths relative to the reference root).  The Python host side
 * (deep-neuroevolution_b200/dne/_ffi.py) binds exactly these symbols through ctypes; INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions
 *   - every call returns int: 0 = ok, <0 = error (text via dne_last_error(), thread-local); never throws,
 *     never aborts.
 *   - device pointers are BORROWED (the caller -- torch -- owns and frees them).  Kernels are enqueued
 *     on the caller's cudaStream_t (passed as void*) and the call returns without synchronising.
 *   - no hidden allocation on the hot path: workspaces are passed in; sizes come from the *_ws_bytes
 *     queries.  A dne_ctx owns only a small fixed scratch buffer allocated at creation.
 *   - one host thread per context; handles are not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     DNE_ERR_CUDA.
 */
#ifndef DNE_H_
#define DNE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNE_OK            0
#define DNE_ERR_ARG      -1
#define DNE_ERR_CUDA     -2
#define DNE_ERR_WS       -3   /* workspace too small */
#define DNE_ERR_UNSUP    -4   /* layer shape not supported by the compiled kernels */

#define DNE_MAX_LAYERS    8

/* layer kinds / activations / batch-norm flavours */
#define DNE_CONV   0
#define DNE_DENSE  1
#define DNE_ACT_NONE 0
#define DNE_ACT_RELU 1
#define DNE_ACT_TANH 2
#define DNE_BN_NONE 0
#define DNE_BN_TF   1          /* contrib.layers.batch_norm(scale=True, decay=0, eps=1e-3): policies.py:322 */
#define DNE_BN_GPU  2          /* ModelVirtualBN (gpu_implementation/neuroevolution/models/batchnorm.py:64-93): layer without
                                  bias, (x - mean) / sqrt(var + 1e-3) + b, no gamma; off_b is that post-normalisation bias */

/* observation kinds */
#define DNE_OB_ATARI_U8 0      /* uint8 [slots,84,84,4], scaled by 1/255 (atari_wrappers.py:186) */
#define DNE_OB_VECTOR   1      /* float32 [slots,ob_dim], clip((o-mean)/std,-5,5) (policies.py:151) */

typedef struct dne_layer_desc {
    int32_t kind;              /* DNE_CONV | DNE_DENSE */
    int32_t cin, cout;         /* conv: channels; dense: fan-in / fan-out */
    int32_t ksize, stride;     /* conv only (square kernel) */
    int32_t hin, hout, pad;    /* conv only: square input/output size, TF-SAME pad_before */
    int32_t act;               /* DNE_ACT_* */
    int32_t bn;                /* DNE_BN_* */
    int32_t bn_off;            /* offset of this layer's (mean[cout], var[cout]) in a slot's vbn stats vector */
    int32_t _pad;
    int64_t off_w, off_b;      /* element offsets into the flat parameter vector; off_b < 0: no bias */
    int64_t off_beta, off_gamma; /* bn == DNE_BN_TF only */
} dne_layer_desc;

/* Flat layout = variable creation order of the reference policy (tf_util.py:224-246;
 * gpu_implementation/neuroevolution/models/base.py:165-192). */
typedef struct dne_net_desc {
    int32_t n_layers;
    int32_t ob_kind;           /* DNE_OB_* */
    int32_t ob_dim;            /* DNE_OB_VECTOR: observation length; ATARI: 84*84*4 */
    int32_t n_out;             /* logits / action dimension */
    int32_t vbn_len;           /* floats of virtual-batch-norm statistics per slot (0 if none) */
    int32_t _pad;
    int64_t num_params;
    dne_layer_desc layers[DNE_MAX_LAYERS];
} dne_net_desc;

typedef struct dne_ctx dne_ctx;

/* ---- context ------------------------------------------------------------------------------------ */
int         dne_ctx_create(int device, dne_ctx** out);
int         dne_ctx_destroy(dne_ctx* ctx);
const char* dne_last_error(void);
int         dne_version(void);
/* ABI self-check for FFI bindings: sizeof(dne_layer_desc), sizeof(dne_net_desc). */
int         dne_abi_sizes(int* layer_desc_bytes, int* net_desc_bytes);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------------
 * dne_launch_count: kernels launched by this library in this process so far (reset != 0 zeroes it).
 * dne_profile_enable/read: CUDA-event timing of every launch of the dominant HBM-bound kernel
 * (dense_noise_gemv) on the stream it is launched on; read() synchronises the device. */
long long   dne_launch_count(int reset);
/* Runtime switches (process-wide; A/B measurement and referee paths only):
 *   "conv_tc" = 2 (default): shifted-window tcgen05 convolutions, images / weights by TMA (conv_s2d.cu), also used by the
 *               virtual-batch-norm reference pass; 1: im2col-staged tcgen05 kind::tf32 convolutions (tc_conv.cu); 0: fp32 SIMT
 *               kernels everywhere (the parity referee).
 *   "theta_tma" = 1 (default): TMA-fed shared-theta GEMM when a prepared region is current (dne_theta_prepare).
 *   "theta_mc" = 0 (default): cluster-multicast variant of it (measured slower).
 *   "fuse_head" = 1 (default): combine + output head + argmax in one kernel.
 *   "fold_theta" = 1 (default): the theta GEMM's split-K partials are folded into the noise GEMV's output.
 *   "pdl" = 1 (default): the tick's kernels are chained by programmatic dependent launch (griddepcontrol).
 *   "chain_ticks" = 0 (default): 1 = the tick's first convolution is a dependent launch too; only valid when the stream's
 *               previous kernel is the previous tick's last kernel (nothing that writes theta / the noise table / the slot table).
 *   "gemv_balance" = 1 (default): the GEMV grid size is chosen to balance the round-robin deal of work items.
 *   "gemv_bulk" = 1 (default): noise GEMV through the cp.async.bulk shared-memory ring; 0 = plain-LDG kernel.
 *   "gemv_ctas_per_sm" = 1|2 (default 2), "gemv_stages" = 2..8 (default 6), "gemv_grid" (default 0 = no cap),
 *   "gemv_chunk_kb" (work-item size), "gemv_prefetch" = 0..256 (default 0; L2 prefetch distance, measured slower). */
int         dne_set_option(const char* name, int value);
int         dne_profile_enable(dne_ctx* ctx, int on, int capacity);
int         dne_profile_read(dne_ctx* ctx, int* n_launches, double* total_ms);

/* Replaces SharedNoiseTable (es_distributed/es.py:51-67): the table lives in HBM; `count` floats, the
 * allocation must extend at least 8 floats past `count` (aligned vector loads of unaligned slices). */
int dne_noise_bind(dne_ctx* ctx, const float* d_noise, int64_t count);

/* ---- rollout side ------------------------------------------------------------------------------- */
/* Bytes of workspace dne_perturb_forward_* needs for `n_slots` slots of `net`. */
int dne_forward_ws_bytes(const dne_net_desc* net, int n_slots, size_t* out_bytes);

/* Replaces, per env tick and for all slots at once:
 *   v = noise_stdev*noise.get(idx,P); policy.set_trainable_flat(theta +/- v)   (es.py:412-419)
 *   policy.act(ob)  -> conv/dense forward + argmax                              (policies.py:319-330,403,449-459;
 *                                                                                models/dqn.py:25-47; indexedmatmul.cpp:148-213)
 * Slot s evaluates weights theta + d_scale[s]*noise[d_noise_idx[s] : +P] (never materialised in HBM).
 * paired == 1 asserts slots (2p, 2p+1) share d_noise_idx (antithetic pair): the slice is then read once.
 * paired == 2 asserts slots (2p, 2p+1) share d_theta_idx (GA siblings of one parent): the parent row is read once.
 * d_theta_idx (nullable): per-slot row into d_theta [n_theta, P] (GA parents); NULL = row 0 for every slot.
 * d_active (nullable): uint8 per slot; inactive slots are skipped and their outputs left untouched.
 * d_vbn: per-slot virtual-batch-norm statistics from dne_vbn_reference_pass (NULL if the net has none).
 * Outputs: d_actions int32[n_slots] (argmax, first max on ties), d_logits float[n_slots, n_out] (nullable). */
int dne_perturb_forward_conv(dne_ctx* ctx, const dne_net_desc* net, const float* d_theta,
                             const int64_t* d_noise_idx, const float* d_scale, const int32_t* d_theta_idx,
                             const uint8_t* d_active, int n_slots, int paired,
                             const uint8_t* d_obs, const float* d_vbn,
                             int32_t* d_actions, float* d_logits,
                             void* d_ws, size_t ws_bytes, void* stream);

/* Optional, once per generation: relays out the shared weight matrices of the net's large dense layers
 * (theta_w[K, N] of the fc layer: the x . theta_w half of x . (theta_w + s*noise), policies.py:327 / dqn.py:46) into the
 * workspace in the tensor-core operand layout (TF32 hi / lo planes), so that the following dne_perturb_forward_conv calls
 * on the SAME (d_ws, d_theta, n_slots) feed that GEMM by TMA instead of staging it through threads.  The entry stays
 * current until dne_adam_step / dne_sgd_step on this context rewrite d_theta (they drop it themselves); after any OTHER
 * write to d_theta call it again.  Without a current entry the forward is still correct (thread-staged GEMM). */
int dne_theta_prepare(dne_ctx* ctx, const dne_net_desc* net, const float* d_theta, int n_slots, void* d_ws, size_t ws_bytes,
                      void* stream);
/* Entries are keyed by the workspace ADDRESS: forget them when a workspace is allocated or freed (an allocator may hand a
 * freed workspace's address to a new one).  d_ws == NULL forgets every entry of the context. */
int dne_theta_forget(dne_ctx* ctx, const void* d_ws);

/* Phase-shifted double buffering of two slot tables on two CUDA streams: the NEXT dne_perturb_forward_* call on ctx
 * makes its stream wait for wait_event (cudaEvent_t, nullable) before its first kernel and records record_event
 * (nullable) right before its first HBM-bound noise GEMV.  Table A: (wait eB, record eA); table B: (wait eA, record eB):
 * the tensor-core conv phase of one table then runs under the HBM-bound phase of the other (mode 0).
 * mode 1 instead waits right before the GEMV and records right after it: the tables take turns on the memory system
 * while their conv chains free-run (use with >= 3 tables). */
int dne_set_phase_events(dne_ctx* ctx, void* wait_event, void* record_event, int mode);

/* MujocoPolicy variant (policies.py:150-162,195-196,202-206): float observations, ob normalisation, tanh MLP,
 * continuous head.  d_actions_out float[n_slots, n_out] (action noise is added by the caller's stream). */
int dne_perturb_forward_mlp(dne_ctx* ctx, const dne_net_desc* net, const float* d_theta,
                            const int64_t* d_noise_idx, const float* d_scale, const int32_t* d_theta_idx,
                            const uint8_t* d_active, int n_slots, int paired,
                            const float* d_obs, const float* d_ob_mean, const float* d_ob_std,
                            float* d_actions_out, void* d_ws, size_t ws_bytes, void* stream);

/* Observation statistics of the running normaliser (es.py:356-363 rollout_and_update_ob_stat; RunningStat es.py:26-48):
 * adds the observations d_obs[slot, :] (float32 [*, ob_dim], the unnormalised vectors fed to this tick's forward) of the m
 * listed slots -- the slots whose episode was sampled with probability calc_obstat_prob -- into float64 running sums
 * d_sum[ob_dim], d_sumsq[ob_dim].  The episode count stays with the caller (m per tick). */
int dne_ob_stat_accumulate(const float* d_obs, int ob_dim, const int32_t* d_slots, int m, double* d_sum, double* d_sumsq,
                           void* stream);

/* Virtual batch norm reference pass, per member before each episode (policies.py:322-328,399;
 * es.py:105-113): forwards the shared reference batch d_ref [n_ref,84,84,4] through every listed slot's
 * perturbed weights with batch statistics and stores (mean, biased var) per BN layer in d_vbn[slot]. */
int dne_vbn_ws_bytes(const dne_net_desc* net, int n_slots, int n_ref, size_t* out_bytes);
int dne_vbn_reference_pass(dne_ctx* ctx, const dne_net_desc* net, const float* d_theta,
                           const int64_t* d_noise_idx, const float* d_scale, const int32_t* d_theta_idx,
                           const uint8_t* d_active, int n_slots,
                           const uint8_t* d_ref, int n_ref, float* d_vbn,
                           void* d_ws, size_t ws_bytes, void* stream);

/* Observation preprocess (atari_wrappers.py:105,167-180 mode 0; tf_atari.py:90 + stack_frames.py:33-43 mode 1):
 * max over two 84x84 uint8 frames, then frame-stack k=4 in place in d_stack [n_slots,84,84,4]. */
int dne_preprocess_atari(const uint8_t* d_prev, const uint8_t* d_cur, uint8_t* d_stack,
                         const uint8_t* d_reset_mask, int n_slots, int mode, void* stream);

/* The 210x160 -> 84x84 warp in front of the frame stack, both reference flavours (one 84x84 frame per raw frame pair):
 * dne_warp_atari_rgb      atari_wrappers.py:105,138-142: d_raw uint8 [n, 2, 210, 160, 3] (the last two RGB frames) ->
 *                         per-channel max -> gray float32 -> PIL BILINEAR (area-scaled triangle filter, two passes, double
 *                         accumulation) -> uint8 [n, 84, 84] (truncating cast).  Bit-exact with Pillow on the same gray.
 * dne_warp_atari_palette  tf_atari.py:88-92,149: d_raw uint8 [n, 2, 210, 160] NTSC palette indices, d_gray_lut float[256]
 *                         -> max of the two gray frames -> resize_bilinear(align_corners=True) -> float32 [n, 84, 84]
 *                         (d_out_f32, nullable) and / or its uint8 quantisation round(255*x) (d_out_u8, nullable) for
 *                         the uint8 frame-stack pipeline.
 * Feed the result to dne_preprocess_atari with d_prev = NULL (the max has already been taken on the raw frames). */
int dne_warp_atari_rgb(const uint8_t* d_raw, uint8_t* d_out, int n_frames, void* stream);
int dne_warp_atari_palette(const uint8_t* d_raw, const float* d_gray_lut, float* d_out_f32, uint8_t* d_out_u8,
                           int n_frames, void* stream);

/* ---- update side -------------------------------------------------------------------------------- */
/* compute_ranks / compute_centered_ranks (es.py:70-85) over the flattened returns; stable tie rule. */
int dne_centered_rank(const float* d_returns, int count, float* d_centered, int32_t* d_ranks, void* stream);

/* batched_weighted_sum + normalise (es.py:115-122,291-296):
 *   g[j] = (1/denom) * sum_i (d_proc[2i]-d_proc[2i+1]) * noise[d_noise_idx[i] + j],  j in [0,P)
 * d_proc: centred ranks (or any processed returns) [n,2]; denom = returns_n2.size of the WHOLE generation
 * (es.py:296); float64 accumulation, one float32 rounding.
 * accumulate != 0 adds into d_g instead of overwriting it. */
int dne_es_grad(dne_ctx* ctx, const float* d_proc_n2, const int64_t* d_noise_idx, int n, int64_t P,
                double denom, float* d_g, int accumulate, void* stream);

/* optimizer.update(-g + l2coeff*theta) (es.py:298; optimizers.py:10-17,35-50).  t is the 1-based step
 * count AFTER the increment at optimizers.py:11.  d_update_ratio: float32 scalar ||step||/||theta_old||. */
int dne_adam_step(dne_ctx* ctx, float* d_theta, float* d_m, float* d_v, const float* d_g, int64_t P,
                  double l2coeff, double stepsize, double beta1, double beta2, double epsilon, int t,
                  float* d_update_ratio, void* stream);
/* SGD with EMA momentum (optimizers.py:23-32). */
int dne_sgd_step(dne_ctx* ctx, float* d_theta, float* d_v, const float* d_g, int64_t P,
                 double l2coeff, double stepsize, double momentum, float* d_update_ratio, void* stream);

/* ---- GA / novelty -------------------------------------------------------------------------------- */
/* Seed chain -> weights.
 * mode 0 (gpu path, models/base.py:140-146,155-156; dqn.py:26-28): theta = noise[seed0]*scale_by; theta += power_k*noise[seed_k]
 * mode 1 (cpu path, ga.py:256-264; policies.py:42-44; tf_util.py:122-130): theta = column-normalise(noise[seed0]),
 *        biases 0; theta += power_k*noise[seed_k].
 * d_seeds int64[len], d_powers float[len] (powers[0] unused) on the device; h_std double[n_layers] on the HOST:
 * init std per layer (NULL = 1.0). */
int dne_ga_materialize(dne_ctx* ctx, const dne_net_desc* net, const int64_t* d_seeds, const float* d_powers,
                       int len, const double* h_std, int mode, float* d_theta_out, void* stream);
/* theta_out = theta_parent + power*noise[seed] (models/base.py:155-156) -- one mutation on a cached parent. */
int dne_ga_mutate(dne_ctx* ctx, const float* d_parent, int64_t seed, float power, int64_t P,
                  float* d_theta_out, void* stream);
/* Truncation selection (ga.py:145-149; gpu_implementation/ga.py:180): indices of the top-T fitness values,
 * descending, ties by arrival order. */
int dne_ga_truncate(const float* d_fitness, int pop, int T, int32_t* d_selected, void* stream);

/* k-NN novelty (nses.py:12-32): BC sequences are [*, t_max, D] uint8 padded with their LAST row, with
 * true lengths; distance over rows t < max(len_q, len_a); novelty = mean of the k smallest distances. */
int dne_knn_ws_bytes(int q, int A, size_t* out_bytes);
int dne_knn_novelty(const uint8_t* d_bc, const int32_t* d_bc_len, int q,
                    const uint8_t* d_archive, const int32_t* d_archive_len, int A,
                    int t_max, int D, int k, float* d_novelty, void* d_ws, size_t ws_bytes, void* stream);

/* Same for vector BCs (MujocoPolicy: final (x, y) position / trajectory, policies.py:292-299): float64 [*, D] of equal
 * length, plain L2 distance in float64 (nses.py:12-20 with equal lengths), mean of the k smallest. */
int dne_knn_novelty_vec(const double* d_bc, int q, const double* d_archive, int A, int D, int k, float* d_novelty,
                        void* d_ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DNE_H_ */
