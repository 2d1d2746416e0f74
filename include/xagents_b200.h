/*
 * xagents_b200 -- C ABI of the B200-native on-policy rollout-to-update hot path.
 *
 * The reference (abstractguy/xagents, pure Python on TF2) has no FFI of its own: the seam is the
 * method-override surface of its agent classes.  Each entry point below replaces the arithmetic of
 * one reference method (cited as file:line relative to the reference tree); the Python classes in
 * xagents_b200/agents/ keep the reference signatures and call these through ctypes, passing device
 * pointers taken from DLPack capsules (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked "host";
 *   - the caller owns every buffer; the library allocates nothing the caller sees;
 *   - calls are stream-ordered and asynchronous: `stream` is a cudaStream_t (NULL = legacy default);
 *   - return 0 on success, a negative XA_E* code for argument errors, a positive cudaError_t for
 *     launch errors; the message is in xa_last_error() (thread-local); nothing aborts or throws;
 *   - layouts: "time-major" = [T, E, ...] as the rollout is produced; "env-major flat" index
 *     b = e*T + t is what BaseAgent.concat_step_batches (xagents/base.py:549-564) produces.  Kernels
 *     taking (n_steps, n_envs) read time-major buffers through env-major indices
 *     (row = (b % T)*E + b / T), so that flatten copy never happens; n_steps = 0 disables the remap.
 *   - there is no CPU fallback: without a CUDA device every compute call returns an error.
 */
#ifndef XAGENTS_B200_H
#define XAGENTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XA_VERSION 10000 /* 1.0.0 */

#define XA_OK 0
#define XA_EINVAL (-1)    /* null pointer, non-positive size, unknown mode */
#define XA_EALIGN (-2)    /* pointer not aligned for its element type */
#define XA_EOVERFLOW (-3) /* size does not fit the index type of the kernel */
#define XA_ENOSPACE (-4)  /* workspace too small */

typedef void* xa_stream_t;

int xa_version(void);
const char* xa_last_error(void);
/* sm_count / cc of `device` (host out-params; any may be NULL). Fails when no CUDA device exists. */
int xa_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);
/* Chained launches (programmatic dependent launch, csrc/xa_common.cuh): kernels of the rollout step, the network pass and the
 * loss -> optimiser chain may be scheduled while their predecessor on the stream drains; each waits for the predecessor's
 * completion before its first global-memory access, so results do not depend on the level.  0 = never, 1 = small launches
 * (default; XA_PDL in the environment sets the initial value), 2 = every launch, 3 = small tensor-core / rollout launches
 * only.  Returns the previous level; a level outside 0..3 only queries.  Process-wide; the reference has no counterpart
 * (TensorFlow orders its kernels on one stream, xagents/a2c/agent.py:96-139 / ppo/agent.py:215-225 being the loops served). */
int xa_set_chained_launches(int level);

/* ---- returns -------------------------------------------------------------------------------- */
#define XA_SCAN_AUTO 0       /* pick by shape */
#define XA_SCAN_SEQUENTIAL 1 /* one thread walks T per env: bit-identical to the reference's op order */
#define XA_SCAN_CHUNKED 2    /* T split across warps, carries combined as affine maps (re-associates) */

/* PPO.calculate_returns, xagents/ppo/agent.py:80-94 (the bootstrap model call at :72-79 is the
 * caller's: pass its output as last_values).  rewards [T,E]; values [T,E]; last_values [E];
 * dones [T+1,E] fp32, row t+1 gates step t; returns [T,E]; advantages [T,E] or NULL.
 * gamma / lam are doubles because the reference folds gamma*lam as Python floats before the product
 * meets the fp32 arrays (ppo/agent.py:92): gamma_lam = (float)(gamma*lam), gamma = (float)gamma. */
int xa_gae_f32(const float* rewards, const float* values, const float* last_values, const float* dones,
               float* returns, float* advantages, int n_steps, int n_envs, double gamma, double lam,
               int mode, xa_stream_t stream);

/* A2C.calculate_returns, xagents/a2c/agent.py:165-171.  returns [T,E]. */
int xa_nstep_returns_f32(const float* rewards, const float* dones, const float* last_values,
                         float* returns, int n_steps, int n_envs, double gamma, int mode,
                         xa_stream_t stream);

/* ACER.calculate_returns (Retrace), xagents/acer/agent.py:198-208, on time-major [T,E] fields (row "next":
 * the same reverse-scan machinery reused by ACER).  dones [T+1,E] as above; importance is the raw ratio
 * (min(1, .) is applied inside); q_selected = critic output at the taken action. */
int xa_retrace_f32(const float* rewards, const float* dones, const float* values, const float* last_values,
                   const float* q_selected, const float* importance, float* returns, int n_steps, int n_envs,
                   double gamma, xa_stream_t stream);

/* ---- permute-gather ------------------------------------------------------------------------- */
#define XA_GATHER_AUTO 0
#define XA_GATHER_BULK 1   /* TMA bulk copies global->shared->global; needs 16-B aligned rows */
#define XA_GATHER_VECTOR 2 /* 128-bit (or narrower, by alignment) LDG/STG */

/* tf.gather(item, batch_indices) of PPO.get_mini_batches, xagents/ppo/agent.py:149-154, fused with
 * the env-major flatten of xagents/base.py:559-564.  dst[i, :] = src[row(idx[i]), :], whole rows of
 * row_bytes bytes, bit-exact.  n_src_rows bounds idx (precondition: 0 <= idx[i] < n_src_rows). */
int xa_gather_rows(const void* src, const int32_t* idx, void* dst, int64_t n_idx, int64_t row_bytes,
                   int64_t n_src_rows, int n_steps, int n_envs, int mode, xa_stream_t stream);

/* Same gather for up to XA_MAX_FIELDS fp32 scalar-per-sample fields at once (actions, returns,
 * old values, old log-probs: the four scalar tf.gather calls of xagents/ppo/agent.py:149-154).  src/dst: HOST arrays of n_fields device pointers. */
#define XA_MAX_FIELDS 8
int xa_gather_fields_f32(const float* const* src, float* const* dst, int n_fields, const int32_t* idx,
                         int64_t n_idx, int n_steps, int n_envs, xa_stream_t stream);

/* One launch for a whole minibatch (or epoch): observation rows + scalar fields -- everything one iteration of the
 * loop in PPO.get_mini_batches produces (xagents/ppo/agent.py:139-155). */
int xa_gather_minibatch(const void* obs_src, void* obs_dst, int64_t row_bytes, int64_t n_src_rows,
                        const float* const* field_src, float* const* field_dst, int n_fields,
                        const int32_t* idx, int64_t n_idx, int n_steps, int n_envs, int mode,
                        xa_stream_t stream);

/* Gather with fine-grained completion, for pipelines that consume minibatches while one long launch is still moving the
 * later ones (the reference materialises every minibatch before the first update, xagents/ppo/agent.py:149-155).  TMA bulk
 * path only (16-byte aligned src/dst/row_bytes).  Destination row i belongs to minibatch
 *   m = e * ceil(rows_per_epoch / mb_rows) + (g - e * rows_per_epoch) / mb_rows,  g = row_offset + i, e = g / rows_per_epoch
 * and progress[m] (device uint32, never reset: cyclic) grows by xa_gather_progress_units(row_bytes) per finished row, after
 * the row's bytes are written.  A consumer orders its stream after minibatch m of the s-th launch over these counters with
 * xa_stream_wait_geq_u32(stream, progress + m, s * rows_in_m * units)  (cuStreamWaitValue32, cyclic >=).  `progress`
 * may be NULL (no counters).  `work` (NULL, or two zero-filled device uint32 words that the kernel leaves zeroed again, not
 * shared by launches that can run concurrently) switches from static dealing of the copy items to a work counter: the launch
 * then no longer runs at the pace of its slowest CTA when other kernels share the SMs. */
int xa_gather_progress_units(int64_t row_bytes);
int xa_gather_rows_progress(const void* src, const int32_t* idx, void* dst, int64_t n_idx, int64_t row_bytes,
                            int64_t n_src_rows, int n_steps, int n_envs, uint32_t* progress, int64_t row_offset,
                            int64_t rows_per_epoch, int64_t mb_rows, uint32_t* work, xa_stream_t stream);
int xa_stream_wait_geq_u32(xa_stream_t stream, const uint32_t* addr, uint32_t value);
/* The same wait as a one-warp spin kernel on `stream` (bounded, ~2 s: on expiry *status -- device int, may be NULL -- is set
 * to 1 and the stream goes on).  Wakes within a microsecond of the publication; the front-end wait above re-polls long waits
 * only every few milliseconds on B200. */
int xa_wait_progress_u32(const uint32_t* addr, uint32_t target, int* status, xa_stream_t stream);

/* BaseAgent.get_model_outputs image scaling, xagents/base.py:505-506, fused behind the gather:
 * dst fp32 [n_idx, row_bytes] = float(src u8) / 255.0f (true division). */
int xa_gather_rows_u8_scaled_f32(const uint8_t* src, const int32_t* idx, float* dst, int64_t n_idx,
                                 int64_t row_bytes, int64_t n_src_rows, int n_steps, int n_envs,
                                 xa_stream_t stream);

/* ---- advantage moments ---------------------------------------------------------------------- */
/* Per-minibatch count / mean / M2 of adv = returns - old_values (PPO.run_ppo_epochs,
 * xagents/ppo/agent.py:180-183; population std = sqrt(M2/count)).  Minibatch m covers
 * idx[mb_offsets[m] .. mb_offsets[m+1]) (mb_offsets: HOST array of n_minibatches+1); idx NULL =
 * identity.  moments: [n_minibatches][XA_MOMENT_STRIDE] doubles {count, mean, M2, 0}.  Exposed
 * separately from the loss so that a cross-rank all-gather of the partial moments can sit between. */
#define XA_MOMENT_STRIDE 4
int xa_adv_moments_f32(const float* returns, const float* old_values, const int32_t* idx,
                       const int64_t* mb_offsets, int n_minibatches, int n_steps, int n_envs,
                       double* moments, xa_stream_t stream);

/* advantages[i] = ((returns - old_values)[row(i)] - mean) / (std + eps) with mean / population std
 * combined from n_moment_parts parts (part p at moments + p*moment_part_stride doubles), i.e. the
 * `advantages_mb` PPO.run_ppo_epochs hands to update_gradients (xagents/ppo/agent.py:180-183). */
int xa_normalize_adv_f32(const float* returns, const float* old_values, const int32_t* idx, int64_t n,
                         int n_steps, int n_envs, const double* moments, int n_moment_parts,
                         int64_t moment_part_stride, double adv_eps, float* advantages,
                         xa_stream_t stream);

/* ---- losses --------------------------------------------------------------------------------- */
#define XA_ACTOR_LOGITS 0 /* Categorical(logits=)            a2c/agent.py:63 */
#define XA_ACTOR_PROBS 1  /* Categorical(probs=)             a2c/agent.py:61-62 */
#define XA_ACTOR_NORMAL 2 /* MultivariateNormalDiag(loc)     a2c/agent.py:59-60 */

typedef struct xa_loss_args {
  /* model outputs for the minibatch (forward pass at ppo/agent.py:113-115) */
  const float* actor_out; /* [n, n_actions] logits | probs | loc */
  const float* values;    /* [n] critic output */
  /* per-sample rollout fields: already gathered [n] when idx == NULL, else full buffers read at
   * row(idx[i]) (time-major when n_steps > 0) so the scalar gathers never materialise */
  const float* actions;       /* [.] fp32-encoded ints, or [., n_actions] for XA_ACTOR_NORMAL */
  const float* old_log_probs; /* PPO only */
  const float* old_values;
  const float* returns;
  const int32_t* idx; /* [n] or NULL */
  int32_t n_steps, n_envs;
  /* advantage normalisation (PPO): either supplied advantages [n] (already normalised, gathered),
   * or moment parts to combine: part p at moments + p*moment_part_stride (doubles). */
  const float* advantages;
  const double* moments;
  int32_t n_moment_parts;
  int64_t moment_part_stride;
  int64_t n;
  int32_t n_actions;
  int32_t actor_kind;
  float clip;       /* PPO clip_norm, used for ratio AND value clip (ppo/agent.py:117-127) */
  float ent_coef;   /* entropy_coef */
  float vf_coef;    /* value_loss_coef */
  float adv_eps;    /* advantage_epsilon */
  /* outputs */
  float* out_scalars;    /* [4] loss, pg_loss, value_loss, entropy */
  float* d_actor;        /* [n, n_actions] d loss / d actor_out, or NULL */
  float* d_values;       /* [n] d loss / d values, or NULL */
  float* advantages_out; /* [n] normalised advantages as used, or NULL */
  void* workspace;       /* xa_loss_workspace_bytes(n) bytes, zero-filled once; kernels leave it zeroed */
  int64_t workspace_bytes;
} xa_loss_args;

int64_t xa_loss_workspace_bytes(int64_t n);
/* PPO.update_gradients forward + the gradients the tape hands back, xagents/ppo/agent.py:112-134. */
int xa_ppo_loss_f32(const xa_loss_args* args, xa_stream_t stream);
/* A2C.train_step loss, xagents/a2c/agent.py:202-215 (advantages = returns - old_values, raw). */
int xa_a2c_loss_f32(const xa_loss_args* args, xa_stream_t stream);

/* ---- rollout-time policy step (row "next": the step before the path) --------------------------- */
/* distribution.sample() / log_prob() / entropy() of A2C.get_model_outputs, xagents/a2c/agent.py:80-94,
 * for one environment step of n envs.  Categorical: Gumbel-max over the log-softmax; `noise` [n, A] holds
 * uniforms in (0,1) (standard normal draws for XA_ACTOR_NORMAL), or NULL to generate them in-kernel with
 * Philox4x32-10 keyed by (seed, offset) -- advance `offset` by at least (A+3)/4 (2A... for NORMAL) per call.
 * actions: [n] fp32-encoded ids, or [n, A] for XA_ACTOR_NORMAL; entropies may be NULL. */
int xa_policy_step_f32(const float* actor_out, int actor_kind, const float* noise, uint64_t seed,
                       uint64_t offset, float* actions, float* log_probs, float* entropies, int64_t n,
                       int n_actions, xa_stream_t stream);
/* The same step with the Philox offset in DEVICE memory: the kernel draws at *offset_dev + offset; when `advance` is not zero a
 * one-thread kernel behind it adds `advance` to *offset_dev: a CUDA graph that captured a whole rollout draws fresh noise on
 * every replay (its steps pass offset = t * step, advance = 0, and one xa_bump_u64 closes the rollout). */
int xa_policy_step_counter_f32(const float* actor_out, int actor_kind, uint64_t seed, uint64_t* offset_dev,
                               uint64_t offset, uint64_t advance, float* actions, float* log_probs, float* entropies, int64_t n,
                               int n_actions, xa_stream_t stream);

/* *counter += delta, in stream order (counter: device memory, 8-byte aligned). */
int xa_bump_u64(uint64_t* counter, uint64_t delta, xa_stream_t stream);

/* One step of `n_envs` device-resident synthetic Atari environments (SURVEY.md 8d's synthetic distribution behind the env
 * interface) fused with the bookkeeping BaseAgent.step_envs does around env.step / env.reset, xagents/base.py:408-426:
 * new_states[e] (may be NULL) = the frame the step returns -- the terminal frame of a finished episode --, states[e] = the frame
 * the environment holds afterwards (after the reset when done), rewards / dones [n_envs] fp32, episode_sums (may be NULL)
 * += reward, copied to sums_log (may be NULL), then cut at dones.  Frames are rows of `row_bytes` drawn from `pool`
 * [pool_size rows]; draws are Philox4x32-10 keyed by seed at counter (e, *offset_dev + offset) (offset_dev may be NULL). */
int xa_synth_env_step_u8(const uint8_t* pool, int pool_size, int64_t row_bytes, uint8_t* states, uint8_t* new_states,
                         float* rewards, float* dones, float* episode_sums, float* sums_log, int n_envs, float p_reward,
                         float p_done, uint64_t seed, const uint64_t* offset_dev, uint64_t offset, xa_stream_t stream);

/* ---- dense contractions on the tensor cores (row "next": the policy/value network) ------------- */
/* C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (ReLU), bf16 operands, fp32 accumulation in TMEM (tcgen05),
 * operands staged by TMA: the Dense layers ModelReader builds (xagents/utils/common.py:239-258; FC 3136->512
 * and the actor/critic heads of ppo/models/cnn-actor-critic.cfg:30-42) in forward (A = activations,
 * B = weights) and, on transposed copies, both backward products.  A, B row-major with K contiguous,
 * 16-B aligned, K % 8 == 0; C row-major with pitch ldc, fp32 (out_bf16 = 0) or bf16 (1). */
int xa_gemm_bf16_tn(const void* a, const void* b, void* c, const float* bias, int64_t m, int64_t n,
                    int64_t k, int64_t ldc, int out_bf16, int relu, const void* relu_mask, void* workspace,
                    int64_t workspace_bytes, xa_stream_t stream);

/* Backward of a Keras Dense layer (the tape's part of xagents/ppo/agent.py:134): xa_gemm_bf16_tn with an output column
 * map and a separate mask pitch: with col_group > 0 (a multiple of 32) column j
 * of the product is stored at c[row*ldc + (j / col_group) * col_group_pitch + j % col_group], which writes a
 * [B, h*w*ch] gradient straight onto a zero-bordered [B, H, W, ch] grid (col_group = w*ch, col_group_pitch = W*ch,
 * ldc = H*W*ch) -- the layout xa_conv_wgrad_nhwc_bf16 reads.  relu_mask is [m, mask_ld] in PRODUCT columns. */
int xa_gemm_bf16_tn_ex(const void* a, const void* b, void* c, const float* bias, int64_t m, int64_t n, int64_t k,
                       int64_t ldc, int out_bf16, int relu, const void* relu_mask, int64_t mask_ld, int64_t col_group,
                       int64_t col_group_pitch, void* workspace, int64_t workspace_bytes, xa_stream_t stream);
/* Scratch for split-K (few output tiles, long K: the weight-gradient products); 0 when the shape does not split.
 * Passing workspace = NULL simply disables splitting.  relu_mask: optional bf16 [m, ldc], c *= (mask > 0). */
int64_t xa_gemm_workspace_bytes(int64_t m, int64_t n, int64_t k);
/* c [m, ldc] bf16 = (a b^T) * (mask > 0) with the ReLU-derivative mask as BITS: bit j of word i of mask_bits <=> element
 * 32 i + j of the [m, mask_ld] activation is > 0 (what xa_conv2d_nhwc_bf16_ex writes through relu_bits_out).  The data
 * gradient of a Dense layer behind a ReLU (tape.gradient through Dense + relu, xagents/ppo/agent.py:134): 1/16 of the mask
 * bytes, and the epilogue stores 32 x 32 boxes by TMA.  n and mask_ld multiples of 32, ldc a multiple of 8, c 16-byte
 * aligned; col_group / col_group_pitch as in xa_gemm_bf16_tn_ex. */
int xa_gemm_bf16_tn_maskbits(const void* a, const void* b, void* c, int64_t m, int64_t n, int64_t k, int64_t ldc,
                             const uint32_t* mask_bits, int64_t mask_ld, int64_t col_group, int64_t col_group_pitch,
                             xa_stream_t stream);
/* The split-K product without its reduction pass: fp32 partial tiles [*splits_out, m, n] in `workspace` for a consumer that
 * adds them itself (xa_heads_forward_partial_bf16).  *splits_out = 1 and nothing is launched when the shape is not split
 * or the workspace is too small: take xa_gemm_bf16_tn(_ex) then.  (Keras Dense of the trunk, xagents/utils/common.py:239-258.) */
int xa_gemm_bf16_tn_partial(const void* a, const void* b, int64_t m, int64_t n, int64_t k, void* workspace, int64_t workspace_bytes,
                            int* splits_out, xa_stream_t stream);

/* Stride-1 NHWC convolution as an implicit GEMM on tcgen05, no im2col buffer.  Two kernels behind one entry point:
 * the flat kernel (csrc/conv_flat_tc.cu: pixels flattened, one window per 128-pixel tile, each tap a row-shifted view
 * of it, weights resident in shared memory) for unpadded layers and zero-bordered full-padding gradients with
 * n_out in {32, 64, 128}; otherwise the per-tap kernel (csrc/conv_tc.cu: per kernel tap TMA fetches the shifted box
 * of the activation, zero-filled outside the image).  This is the network's convolutional trunk
 * (ppo/models/cnn-actor-critic.cfg:1-21) after space-to-depth turns its strided layers into stride-1 ones, and,
 * with pad = k-1 and flipped weights, its data-gradient (relu_mask, output layout, applies the ReLU derivative of
 * the layer below).  x [B,H,W,C] bf16, w [n_out, kh*kw*C] bf16 with K ordered (kh, kw, c), y [B,OH,OW,n_out] bf16
 * -- or, with out_s2d, [B,OH/2,OW/2,4*n_out] with channel block (oy%2, ox%2), the next layer's space-to-depth
 * input.  Needs C % 64 == 0, n_out % 32 == 0, OW <= 128. */
int xa_conv2d_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, int batch, int height,
                        int width, int channels, int kh, int kw, int n_out, int pad_y, int pad_x, int relu,
                        int out_s2d, const void* relu_mask, xa_stream_t stream);

/* xa_conv2d_nhwc_bf16 with an explicit output extent and output pixel grid (what the backward pass needs to keep every
 * gradient on the zero-bordered grid xa_conv_wgrad_nhwc_bf16 reads):
 *   out_h, out_w            > 0: compute only the top-left out_h x out_w corner of the padded output
 *   out_grid_h, out_grid_w  > 0: pixel (y, x) is stored at ((b*out_grid_h + y)*out_grid_w + x); pixels outside the
 *                           written extent are left untouched (the caller zero-fills the buffer once)
 *   out_mode                0 natural; 1 pack 2x2 pixels into channels (the old out_s2d); 2 unpack: the n_out channels
 *                           are (dy, dx, n_out/4) and land on pixels (2y+dy, 2x+dx) of a [B, grid_h, grid_w, n_out/4] grid
 *   flags                   XA_CONV_INPUT_ZERO_BORDER: the caller guarantees that the last pad_x columns and pad_y rows
 *                           of every input image are zero (a gradient on a zero-bordered grid); padded convolutions
 *                           may then use the flat kernel that fetches every input pixel once (csrc/conv_flat_tc.cu)
 *                           XA_CONV_MASK_BITS: relu_mask points to BIT masks (uint32 words: bit j of word i <=> element 32 i + j
 *                           of the compact activation is > 0) instead of the bf16 activation itself
 *   relu_bits_out           != NULL (ReLU layers with a compact or 2x2-packed output, flat kernel): such a bit mask of THIS
 *                           layer's output, written from the epilogue (1/16 of the output's bytes) for the data gradient of
 *                           the layer above -- the backward pass then never re-reads an activation only for its sign
 * relu_mask always has the natural compact [B, out_h, out_w, n_out] layout. */
#define XA_CONV_INPUT_ZERO_BORDER 1
#define XA_CONV_MASK_BITS 2
int xa_conv2d_nhwc_bf16_ex(const void* x, const void* w, const float* bias, void* y, int batch, int height, int width,
                           int channels, int kh, int kw, int n_out, int pad_y, int pad_x, int relu, int out_mode,
                           const void* relu_mask, int out_h, int out_w, int out_grid_h, int out_grid_w,
                           uint32_t* relu_bits_out, int flags, xa_stream_t stream);

/* The first (4x4-strided) layer straight from the uint8 frames: cast + /255 (xagents/base.py:505-506), space-to-depth and the
 * convolution in one kernel -- raw uint8 windows arrive by TMA, converter warps write the bf16 SWIZZLE_128B operand tiles
 * the tensor core reads, so the bf16 space-to-depth tensor (2x the frames' bytes) never exists in HBM.  frames
 * [B, height, width, 4] uint8; a kh x kw stride-1 kernel over the 4x4 space-to-depth grid (the 8x8/4 layer: kh = kw = 2);
 * w [n_out, kh*kw*64] bf16, K ordered (kh, kw, dy, dx, c); y as xa_conv2d_nhwc_bf16.  x_s2d_out (may be NULL): the converted
 * tiles are also stored as the bf16 [B, height/4, width/4, 64] space-to-depth tensor (what xa_conv_wgrad_nhwc_bf16 reads
 * in the backward pass), from shared memory, without a second pass over the frames.  Results are bit-identical to
 * xa_space_to_depth_u8_bf16 followed by xa_conv2d_nhwc_bf16. */
int xa_conv2d_u8_s2d_bf16(const uint8_t* frames, const void* w, const float* bias, void* y, void* x_s2d_out, int batch,
                          int height, int width, int kh, int kw, int n_out, int relu, int out_s2d, xa_stream_t stream);
/* The same layer with every option: frame_idx (may be NULL) as in xa_conv2d_u8_s2d_bf16_indexed below; relu_bits_out (may be NULL)
 * as in xa_conv2d_nhwc_bf16_ex. */
int xa_conv2d_u8_s2d_bf16_ex(const uint8_t* frames, int64_t n_frames, const int32_t* frame_idx, int n_steps, int n_envs,
                             const void* w, const float* bias, void* y, void* x_s2d_out, uint32_t* relu_bits_out, int batch,
                             int height, int width, int kh, int kw, int n_out, int relu, int out_s2d, xa_stream_t stream);
/* The same layer on the `batch` frames frames[frame_idx[0 .. batch)] of a store of n_frames frames, read through the
 * permutation by per-frame TMA bulk copies: get_mini_batches' tf.gather of the states (xagents/ppo/agent.py:139-155) folded
 * into the layer -- the gathered minibatch never exists in HBM.  n_steps > 0: ids are env-major sample ids of a time-major
 * [n_steps, n_envs] rollout (row = (id % n_steps) * n_envs + id / n_steps, xagents/base.py:559-564); ids must be in range.
 * Bit-identical to xa_gather_rows followed by xa_conv2d_u8_s2d_bf16. */
int xa_conv2d_u8_s2d_bf16_indexed(const uint8_t* frames, int64_t n_frames, const int32_t* frame_idx, int n_steps, int n_envs,
                                  const void* w, const float* bias, void* y, void* x_s2d_out, int batch, int height,
                                  int width, int kh, int kw, int n_out, int relu, int out_s2d, xa_stream_t stream);

/* uint8 NHWC frames -> bf16 (optionally /255, xagents/base.py:505-506) rearranged block x block -> channels:
 * dst[b, y/s, x/s, (y%s, x%s, c)]. */
int xa_space_to_depth_u8_bf16(const uint8_t* src, void* dst, int batch, int height, int width, int channels,
                              int block, int scale_255, xa_stream_t stream);

/* After optimizer.apply_gradients (xagents/ppo/agent.py:137) the kernels' operand copies must follow the fp32 weights:
 * dst[i] = src[map[i]] as bf16 (out_bf16) or fp32, 0 where map[i] < 0: re-derives every weight layout of the
 * tensor-core network (all permutations of the fp32 parameters) in one launch after an optimiser step. */
int xa_gather_cast_f32(const float* src, const int32_t* map, void* dst, int64_t n, int out_bf16, xa_stream_t stream);

/* c [m, n] fp32 (pitch ldc) = a^T b for ROW-MAJOR a [k, m] and b [k, n] bf16: the Dense weight gradient dW = dY^T X
 * (Dense layers of the .cfg model, xagents/utils/common.py:239-258; gradient taken at xagents/ppo/agent.py:134)
 * with dY [batch, out] and X [batch, in] as the other kernels leave them -- no transposed copies (MN-major UMMA
 * operands, csrc/gemm_atb_tc.cu).  m and n multiples of 8.  workspace (xa_gemm_atb_workspace_bytes, may be NULL)
 * enables the deterministic split over k for shapes with few output tiles. */
int64_t xa_gemm_atb_workspace_bytes(int64_t m, int64_t n, int64_t k);
int xa_gemm_bf16_atb(const void* a, const void* b, float* c, int64_t m, int64_t n, int64_t k, int64_t ldc, void* workspace,
                     int64_t workspace_bytes, xa_stream_t stream);

/* Weight and bias gradient of a stride-1 NHWC convolution (conv sections of the .cfg model, xagents/utils/common.py:
 * 218-237, read as Conv2D: README.md:243-259) from the NATURAL tensors (no transposes, no im2col):
 * x [q_total, channels] bf16 with q = (b*H + y)*grid_w + x the pixels of the INPUT grid, dy_grid [q_total, n_out] bf16
 * = dY placed on that same grid (zero where there is no output pixel; xa_conv2d_nhwc_bf16_ex / xa_gemm_bf16_tn_ex
 * write it there).  dw [n_out, kh*kw*channels] fp32 (K ordered kh, kw, c), db [n_out] fp32 (may be NULL).
 * n_out 32 or 64, channels a multiple of 64; both operands are read once (MN-major UMMA operands, csrc/wgrad_mn_tc.cu). */
int64_t xa_conv_wgrad_nhwc_workspace_bytes(int n_out, int channels, int kh, int kw);
int xa_conv_wgrad_nhwc_bf16(const void* x, const void* dy_grid, float* dw, float* db, int n_out, int channels, int kh,
                            int kw, int grid_w, int64_t q_total, void* workspace, int64_t workspace_bytes,
                            xa_stream_t stream);

/* The minibatch gather for the tensor-core network: tf.gather (ppo/agent.py:154) + flatten (base.py:559-564) +
 * cast/255 (base.py:505-506) + space-to-depth in ONE pass: dst [n_idx, H/s, W/s, s*s*C] bf16 from the time-major uint8
 * rollout (28 KB read + 56 KB written per frame instead of 56 + 84 for gather then space-to-depth). */
int xa_gather_s2d_u8_bf16(const uint8_t* src, const int32_t* idx, void* dst, int64_t n_idx, int64_t n_src_rows,
                          int n_steps, int n_envs, int height, int width, int channels, int block, int scale_255,
                          xa_stream_t stream);

/* Split partials only (no reduction pass): xa_conv_wgrad_nhwc_bf16 / xa_gemm_bf16_atb leave fp32 partial sums
 * [splits, n_out, ld_partial] (column kh*kw*channels of a row = the bias gradient) / [splits, m, n] for the caller's own
 * reduction -- xa_grad_finalize_f32 adds the partials of every layer of the network in one launch.  The *_plan calls
 * return the split count (host out-params) for given sizes on the current device. */
int xa_conv_wgrad_nhwc_plan(int n_out, int channels, int kh, int kw, int64_t q_total, int* splits, int* ld_partial);
int xa_conv_wgrad_nhwc_bf16_partial(const void* x, const void* dy_grid, int n_out, int channels, int kh, int kw,
                                    int grid_w, int64_t q_total, float* partial, int64_t partial_bytes,
                                    xa_stream_t stream);
int xa_gemm_atb_plan(int64_t m, int64_t n, int64_t k, int* splits);
int xa_gemm_bf16_atb_partial(const void* a, const void* b, int64_t m, int64_t n, int64_t k, float* partial,
                             int64_t partial_bytes, xa_stream_t stream);

/* The actor / critic heads (two Dense layers on the 512-wide trunk output, ppo/models/cnn-actor-critic.cfg:30-42) in one
 * pass over h each way.  wh [8, hidden] bf16 = both heads stacked (rows n_actions+1.. zero), bh [8] fp32.
 * forward:  actor [batch, n_actions], critic [batch] fp32 (tf.squeeze'd, a2c/agent.py:84).
 * backward: d_actor / d_critic fp32 from the loss kernel -> dh [batch, hidden] bf16 = (d_out wh) * (h > 0), and per-CTA
 * partial blocks [xa_heads_backward_blocks(batch)][10][hidden] fp32: rows 0-7 dW_heads, row 8 db of the layer that
 * produced h (column sums of dh), row 9 db_heads in its first 8 entries.  hidden must be 512, n_actions <= 7. */
int xa_heads_backward_blocks(int batch);
int xa_heads_forward_bf16(const void* h, const void* wh, const float* bh, float* actor, float* critic, int batch,
                          int hidden, int n_actions, xa_stream_t stream);
/* The same forward when h does not exist yet: `partial` [splits, batch, hidden] fp32 are the split-K partial products of the
 * layer that produces h (xa_gemm_bf16_tn_partial); they are added in split order, `bias` (may be NULL) added, ReLU applied and
 * the result rounded to bf16 -- bit for bit what the GEMM's own reduction pass writes -- stored to h_out (may be NULL) and fed
 * to the heads.  One launch less per rollout step (a2c/agent.py:96-139, the model call inside get_batch's loop). */
int xa_heads_forward_partial_bf16(const float* partial, int splits, const float* bias, void* h_out, const void* wh, const float* bh,
                                  float* actor, float* critic, int batch, int hidden, int n_actions, xa_stream_t stream);
int xa_heads_backward_bf16(const float* d_actor, const float* d_critic, const void* h, const void* wh, void* dh,
                           float* partial, int64_t partial_floats, int batch, int hidden, int n_actions,
                           xa_stream_t stream);

/* tape.gradient's results into the flat gradient buffer of the optimiser step (xagents/ppo/agent.py:134-137):
 * grad[dest[j]] = sum_{s < splits} src[map[j] + s*split_stride] (0 where map[j] < 0; dest = NULL: grad[j]) with splits /
 * split_stride constant over segments of consecutive j (segment i covers [dest_begin_i, dest_begin_{i+1}); the first
 * starts at 0, the last ends at n).  Enumerate each segment in the order of its sources (map ascending) so that the
 * partial sums are read coalesced.  `wide` = lanes per output: 0 one thread, 4 / 8 / 16 / 32 lanes each adding a contiguous
 * share of the splits, combined by a fixed tree (1 = 4); use 32 for few outputs with hundreds of splits.
 * segments: HOST array. */
#define XA_MAX_GRAD_SEGMENTS 16
typedef struct xa_grad_segment_t {
  int64_t dest_begin;
  int64_t split_stride;
  int32_t splits;
  int32_t wide;
} xa_grad_segment_t;
int xa_grad_finalize_f32(const float* src, const int32_t* map, const int32_t* dest, const xa_grad_segment_t* segments,
                         int n_segments, float* grad, int64_t n, xa_stream_t stream);

/* The documented PPO/A2C network (README.md:243-259; ppo/models/cnn-actor-critic.cfg) as two calls: every launch of the
 * forward / backward pass issued from native code over buffers the caller owns (HOST struct of device pointers).
 *   operands     bf16 weight layouts and fp32 biases (agents/tc_operands.py)
 *   activations  x1 [B,21,21,64] (space-to-depth input; unused when the frames arrive in that form), x2 [B,10,10,128],
 *                x3 [B,9,9,64], y3 [B,7,7,64], h [B,512] bf16; actor [B,A], critic [B] fp32
 *   gradients    dh [B,512], g3 [B,9,9,64], g2 [B,10,10,64], g1 [B,21,21,32] bf16 (the g* zero-filled once by the caller:
 *                only their valid corners are ever written); scratch: fp32 partial sums at the given offsets
 *   grad_map / segments   for xa_grad_finalize_f32 over `scratch`
 * forward: frames uint8 [B,84,84,4] (frames_s2d = 0) or bf16 [B,21,21,64] already scaled and space-to-depth'd (1).
 * backward: d_actor [B,A], d_critic [B] fp32 -> flat_grad [n_grad] fp32, every element written. */
typedef struct xa_nature_cnn_t {
  int32_t batch, n_actions;
  const void *w1, *w2, *w3, *w2_flip, *w3_flip, *wf, *wf_t, *wh;
  const float *b1, *b2, *b3, *bf, *bh;
  void *x1, *x2, *x3, *y3, *h;
  float *actor, *critic;
  void *dh, *g3, *g2, *g1;
  void* gemm_ws; /* optional split-K scratch of the FC forward product (small batches) */
  int64_t gemm_ws_bytes;
  float* scratch;
  int64_t scratch_floats;
  int64_t off_c1, off_c2, off_c3, off_fc, off_heads; /* float offsets into scratch */
  const int32_t* grad_map;  /* source offset of enumeration index j ... */
  const int32_t* grad_dest; /* ... and its place in flat_grad */
  xa_grad_segment_t segments[XA_MAX_GRAD_SEGMENTS];
  int32_t n_segments, reserved;
  int64_t n_grad;
  uint32_t *relu_bits2, *relu_bits3; /* optional (both or none): ReLU-derivative bit masks of x2 / x3 (numel / 8 bytes each), written
                                        by the forward pass and read by the data gradients instead of the activations */
  uint32_t* relu_bitsf;              /* optional: the same for y3 (batch * 3136 / 8 bytes), read by the FC layer's data gradient */
} xa_nature_cnn_t;
int xa_nature_cnn_forward(const xa_nature_cnn_t* net, const void* frames, int frames_s2d, xa_stream_t stream);
/* forward on the minibatch frames[frame_idx[0 .. batch)] of a frame store [n_frames, 84, 84, 4] uint8 without gathering it: the
 * tf.gather of the states in get_mini_batches (xagents/ppo/agent.py:139-155) folded into the first layer (ids as in
 * xa_gather_rows: env-major sample ids of a time-major [n_steps, n_envs] rollout when n_steps > 0, plain rows otherwise). */
int xa_nature_cnn_forward_indexed(const xa_nature_cnn_t* net, const void* frames, int64_t n_frames,
                                  const int32_t* frame_idx, int n_steps, int n_envs, xa_stream_t stream);
int xa_nature_cnn_backward(const xa_nature_cnn_t* net, const void* frames_s2d_or_null, const float* d_actor,
                           const float* d_critic, float* flat_grad, xa_stream_t stream);

/* Operand preparation for the Dense / convolution products above (no reference counterpart: the reference computes in
 * fp32 throughout).  fp32 | bf16 [rows, cols] -> bf16, same orientation (dst pitch ld_dst >= cols) or transposed into
 * [cols, ld_dst >= rows]: operand preparation for the backward products (dW = dY^T X needs both transposed). */
int xa_to_bf16(const void* src, int src_is_f32, void* dst, int64_t rows, int64_t cols, int64_t ld_dst,
               int transpose, xa_stream_t stream);

/* ---- optimiser step (row "next": the step right after the path) ------------------------------ */
/* tf.clip_by_global_norm + Keras Adam.apply_gradients, xagents/ppo/agent.py:135-137,
 * xagents/a2c/agent.py:216-218, over ONE flat fp32 buffer holding every trainable tensor.
 * xa_grad_sumsq_f32 leaves sum(g^2) in the workspace (zero-filled once; kernels leave it reusable);
 * xa_clip_adam_f32 then applies g' = g * grad_scale * clip*min(1/|g*grad_scale|, 1/clip) (clip_norm
 * <= 0: no clipping, workspace unused) and the bias-corrected-lr Adam update for 1-based `step`. */
int64_t xa_clip_adam_workspace_bytes(int64_t n);
int xa_grad_sumsq_f32(const float* grads, int64_t n, void* workspace, int64_t workspace_bytes,
                      xa_stream_t stream);
int xa_clip_adam_f32(float* param, const float* grad, float* m, float* v, int64_t n,
                     const void* workspace, double lr, double beta1, double beta2, double eps,
                     double clip_norm, int64_t step, double grad_scale, xa_stream_t stream);

/* ---- collective C1 fused with the optimiser over NVLink peer memory (SURVEY.md 8e + 8f-2) --------- */
/* Replaces ncclAllReduce(gradients) + xa_grad_sumsq_f32 + xa_clip_adam_f32 by one cooperative kernel per update:
 * reduce-scatter by P2P loads (rank r sums shard r over every rank's gradient buffer, in rank order), global norm
 * from G exchanged partial sums, clip + Keras Adam on the owned shard (m, v exist only for it), all-gather by P2P
 * stores of the updated weights into every rank's parameter buffer.  The reference's step is
 * tf.clip_by_global_norm + Adam.apply_gradients (xagents/ppo/agent.py:135-137) in one process; averaging over ranks
 * is the data-parallel extension.  grad[r], param[r], sumsq[r], flags[r] are THIS process's mappings of rank r's
 * buffers (entries >= world unused): grad/param [world * shard_len] floats, sumsq [XA_MAX_PEERS] doubles, flags
 * xa_peer_adam_flag_bytes() bytes, zero-filled before the first call.  m, v: local [shard_len]; workspace: local,
 * xa_peer_adam_workspace_bytes(), zero-filled once.  `epoch` must be the same on every rank and increase by one per
 * call (it is what the peers' flags are compared with).  Every rank must call with the same step / hyper-parameters;
 * waits are bounded (~2 s) and a timeout is reported in workspace word 2 (status: 0 = ok, 1..3 = phase that timed out). */
#define XA_MAX_PEERS 8
typedef struct xa_peer_adam_args {
  void* grad[XA_MAX_PEERS];
  void* param[XA_MAX_PEERS];
  void* sumsq[XA_MAX_PEERS];
  void* flags[XA_MAX_PEERS];
  float* m;
  float* v;
  void* workspace;
  int64_t shard_len;
  int32_t rank, world;
  uint32_t epoch;
  float lr_t, beta1, beta2, eps, clip, inv_world; /* filled by the call */
} xa_peer_adam_args;
/* Peer-mappable device memory through CUDA IPC (zero-filled cudaMalloc + a 64-byte handle; the peers open the handle). */
int xa_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64);
int xa_ipc_open(const unsigned char* handle64, void** ptr);
int xa_ipc_close(void* ptr);
int xa_ipc_free(void* ptr);
int64_t xa_peer_adam_workspace_bytes(void);
int64_t xa_peer_adam_flag_bytes(void);
int xa_peer_allreduce_adam_f32(const xa_peer_adam_args* args, double lr, double beta1, double beta2, double eps,
                               double clip_norm, int64_t step, xa_stream_t stream);

/* Collective C2 over peer memory: all-gather of a small fp64 block per rank (the per-minibatch advantage moments of
 * xagents/ppo/agent.py:180-183, so that normalisation uses the statistics of the GLOBAL minibatch).  recv[r] / flags[r]: this
 * process's mappings of rank r's receive buffer (2 generations x world x n_per_rank doubles) and flag words (XA_MAX_PEERS
 * uint32), zero-filled before the first call; `epoch` as in xa_peer_allreduce_adam_f32 (same on every rank, +1 per call);
 * status: optional local device int set to 1 if a wait gave up (~2 s).  out: local [world * n_per_rank] doubles. */
typedef struct xa_peer_gather_args {
  void* recv[XA_MAX_PEERS];
  void* flags[XA_MAX_PEERS];
  int* status;
  int64_t n_per_rank;
  int32_t rank, world;
  uint32_t epoch;
} xa_peer_gather_args;
int xa_peer_allgather_f64(const xa_peer_gather_args* args, const double* src, double* out, xa_stream_t stream);

/* Benchmark stand-in for the network backward (bench.py): grad[j] = d_actor[j mod n*A] + 0.5 * d_values[j mod n].  The reference
 * obtains the parameter gradients from its tape (xagents/ppo/agent.py:134); the headline metric excludes the network, so
 * this kernel supplies what the gradient all-reduce and the optimiser need from it: a flat gradient that depends on this
 * minibatch's loss outputs. */
int xa_grad_from_outputs_f32(const float* d_actor, const float* d_values, int64_t n, int n_actions,
                             float* grad, int64_t n_params, xa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* XAGENTS_B200_H */
