// The documented PPO / A2C network (README.md:243-259; ppo/models/cnn-actor-critic.cfg read as Conv2D: 32x8/4 -> 64x4/2 ->
// 64x3/1 -> FC512 -> actor / critic heads) as two native calls: the host side of a forward or backward pass is a
// sequence of launches over buffers with fixed addresses, so it is issued from here -- 6 / 9 launches per call with the
// tensor maps encoded in C -- instead of ~40 Python wrapper calls per minibatch (each re-deriving its arguments through
// DLPack), which at 0.8 ms of GPU work per minibatch would leave the GPU waiting for the interpreter.
// Which kernel does what: agents/tc_cnn.py (the autograd form of the same pipeline, kept as the checker).
#include "xa_common.cuh"

extern "C" {

#define XA_TRY(call)         \
  do {                       \
    if (int rc_ = (call)) return rc_; \
  } while (0)

static int check_net(const xa_nature_cnn_t* n, const char* what) {
  XA_REQUIRE(n != nullptr, XA_EINVAL, "%s: null network", what);
  XA_REQUIRE(n->batch > 0 && n->n_actions > 0 && n->n_actions <= 7, XA_EINVAL, "%s: batch=%d n_actions=%d", what, n->batch, n->n_actions);
  XA_REQUIRE(n->w1 && n->w2 && n->w3 && n->wf && n->wh && n->b1 && n->b2 && n->b3 && n->bf && n->bh, XA_EINVAL, "%s: null operand", what);
  XA_REQUIRE(n->x2 && n->x3 && n->y3 && n->h && n->actor && n->critic, XA_EINVAL, "%s: null activation buffer", what);
  return XA_OK;
}

static int forward_impl(const char* what, const xa_nature_cnn_t* n, const void* frames, int frames_s2d, int64_t n_frames, const int32_t* frame_idx,
                        int n_steps, int n_envs, xa_stream_t stream) {
  XA_TRY(check_net(n, what));
  XA_REQUIRE(frames != nullptr, XA_EINVAL, "%s: null frames", what);
  const int B = n->batch;
  XA_REQUIRE((n->relu_bits2 == nullptr) == (n->relu_bits3 == nullptr), XA_EINVAL, "%s: relu_bits2 and relu_bits3 go together", what);
  if (frames_s2d)
    XA_TRY(xa_conv2d_nhwc_bf16_ex(frames, n->w1, n->b1, n->x2, B, 21, 21, 64, 2, 2, 32, 0, 0, 1, 1, nullptr, 0, 0, 0, 0, n->relu_bits2, 0, stream));
  else  // /255 + space-to-depth + conv1 in one kernel; x1 (only the backward pass reads it) is written from shared memory.  With
        // frame_idx: the minibatch gather folded in as well -- frames = the whole rollout, read through the permutation
    XA_TRY(xa_conv2d_u8_s2d_bf16_ex(static_cast<const uint8_t*>(frames), n_frames, frame_idx, n_steps, n_envs, n->w1, n->b1, n->x2, n->x1,
                                    n->relu_bits2, B, 84, 84, 2, 2, 32, 1, 1, stream));
  XA_TRY(xa_conv2d_nhwc_bf16_ex(n->x2, n->w2, n->b2, n->x3, B, 10, 10, 128, 2, 2, 64, 0, 0, 1, 0, nullptr, 0, 0, 0, 0, n->relu_bits3, 0, stream));
  XA_TRY(xa_conv2d_nhwc_bf16_ex(n->x3, n->w3, n->b3, n->y3, B, 9, 9, 64, 3, 3, 64, 0, 0, 1, 0, nullptr, 0, 0, 0, 0, n->relu_bitsf, 0, stream));
  // small batches (rollout inference): the FC product is split along K; its partial sums are added by the heads kernel
  int splits = 1;
  XA_TRY(xa_gemm_bf16_tn_partial(n->y3, n->wf, B, 512, 3136, n->gemm_ws, n->gemm_ws_bytes, &splits, stream));
  if (splits > 1)
    return xa_heads_forward_partial_bf16(static_cast<const float*>(n->gemm_ws), splits, n->bf, n->h, n->wh, n->bh, n->actor, n->critic, B, 512,
                                         n->n_actions, stream);
  XA_TRY(xa_gemm_bf16_tn_ex(n->y3, n->wf, n->h, n->bf, B, 512, 3136, 512, 1, 1, nullptr, 512, 0, 0, n->gemm_ws, n->gemm_ws_bytes, stream));
  return xa_heads_forward_bf16(n->h, n->wh, n->bh, n->actor, n->critic, B, 512, n->n_actions, stream);
}

int xa_nature_cnn_forward(const xa_nature_cnn_t* n, const void* frames, int frames_s2d, xa_stream_t stream) {
  return forward_impl("xa_nature_cnn_forward", n, frames, frames_s2d, 0, nullptr, 0, 0, stream);
}

int xa_nature_cnn_forward_indexed(const xa_nature_cnn_t* n, const void* frames, int64_t n_frames, const int32_t* frame_idx, int n_steps,
                                  int n_envs, xa_stream_t stream) {
  XA_REQUIRE(frame_idx != nullptr, XA_EINVAL, "xa_nature_cnn_forward_indexed: null frame_idx");
  return forward_impl("xa_nature_cnn_forward_indexed", n, frames, 0, n_frames, frame_idx, n_steps, n_envs, stream);
}

int xa_nature_cnn_backward(const xa_nature_cnn_t* n, const void* frames_s2d_or_null, const float* d_actor, const float* d_critic,
                           float* flat_grad, xa_stream_t stream) {
  const char* what = "xa_nature_cnn_backward";
  XA_TRY(check_net(n, what));
  XA_REQUIRE(d_actor && d_critic && flat_grad, XA_EINVAL, "%s: null gradient pointer", what);
  XA_REQUIRE(n->w2_flip && n->w3_flip && n->wf_t && n->dh && n->g3 && n->g2 && n->g1 && n->scratch && n->grad_map && n->grad_dest, XA_EINVAL,
             "%s: null backward buffer", what);
  const void* x1 = frames_s2d_or_null ? frames_s2d_or_null : n->x1;
  XA_REQUIRE(x1 != nullptr, XA_EINVAL, "%s: no first-layer input", what);
  const int B = n->batch;
  float* sc = n->scratch;
  const int64_t room = n->scratch_floats;
  auto bytes_from = [&](int64_t off) { return (room - off) * static_cast<int64_t>(sizeof(float)); };
  XA_REQUIRE(n->off_c1 >= 0 && n->off_c2 >= 0 && n->off_c3 >= 0 && n->off_fc >= 0 && n->off_heads >= 0 && n->off_c1 < room && n->off_c2 < room &&
                 n->off_c3 < room && n->off_fc < room && n->off_heads < room,
             XA_EINVAL, "%s: scratch offsets out of range", what);
  // heads: dh (ReLU derivative of the FC layer applied), dW_heads, db_heads, db_fc
  XA_TRY(xa_heads_backward_bf16(d_actor, d_critic, n->h, n->wh, n->dh, sc + n->off_heads, room - n->off_heads, B, 512, n->n_actions, stream));
  // FC512: dW = dh^T y3 (split partials), dX = dh Wf masked by y3 > 0, written as 7x7 onto conv3's zero-bordered 9x9 grid
  XA_TRY(xa_gemm_bf16_atb_partial(n->dh, n->y3, 512, 3136, B, sc + n->off_fc, bytes_from(n->off_fc), stream));
  if (n->relu_bitsf != nullptr && xa::gemm_tma_store_enabled())   // y3's signs as bits (written by conv3's forward epilogue), 32 x 32 boxes stored by TMA
    XA_TRY(xa_gemm_bf16_tn_maskbits(n->dh, n->wf_t, n->g3, B, 3136, 512, 9 * 9 * 64, n->relu_bitsf, 3136, 7 * 64, 9 * 64, stream));
  else
    XA_TRY(xa_gemm_bf16_tn_ex(n->dh, n->wf_t, n->g3, nullptr, B, 3136, 512, 9 * 9 * 64, 1, 0, n->y3, 3136, 7 * 64, 9 * 64, nullptr, 0, stream));
  // conv3, conv2: weight gradient from the natural NHWC tensors, data gradient = flat convolution with flipped weights
  XA_TRY(xa_conv_wgrad_nhwc_bf16_partial(n->x3, n->g3, 64, 64, 3, 3, 9, static_cast<int64_t>(B) * 81, sc + n->off_c3, bytes_from(n->off_c3), stream));
  const bool bits = n->relu_bits2 != nullptr && n->relu_bits3 != nullptr;  // ReLU derivatives from the forward pass's bit masks
  XA_TRY(xa_conv2d_nhwc_bf16_ex(n->g3, n->w3_flip, nullptr, n->g2, B, 9, 9, 64, 3, 3, 64, 2, 2, 0, 0, bits ? static_cast<const void*>(n->relu_bits3) : n->x3,
                                9, 9, 10, 10, nullptr, XA_CONV_INPUT_ZERO_BORDER | (bits ? XA_CONV_MASK_BITS : 0), stream));
  XA_TRY(xa_conv_wgrad_nhwc_bf16_partial(n->x2, n->g2, 64, 128, 2, 2, 10, static_cast<int64_t>(B) * 100, sc + n->off_c2, bytes_from(n->off_c2), stream));
  XA_TRY(xa_conv2d_nhwc_bf16_ex(n->g2, n->w2_flip, nullptr, n->g1, B, 10, 10, 64, 2, 2, 128, 1, 1, 0, 2, bits ? static_cast<const void*>(n->relu_bits2) : n->x2,
                                10, 10, 21, 21, nullptr, XA_CONV_INPUT_ZERO_BORDER | (bits ? XA_CONV_MASK_BITS : 0), stream));
  XA_TRY(xa_conv_wgrad_nhwc_bf16_partial(x1, n->g1, 32, 64, 2, 2, 21, static_cast<int64_t>(B) * 441, sc + n->off_c1, bytes_from(n->off_c1), stream));
  return xa_grad_finalize_f32(sc, n->grad_map, n->grad_dest, n->segments, n->n_segments, flat_grad, n->n_grad, stream);
}

}  // extern "C"
