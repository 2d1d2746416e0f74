// Fused global-norm clip + Adam over one flat fp32 parameter buffer (the step right after the loss).
//
// Replaces tf.clip_by_global_norm + Keras Adam.apply_gradients at xagents/ppo/agent.py:135-137 and
// xagents/a2c/agent.py:216-218 (Adam built at xagents/utils/common.py:476,589-594).  All trainable
// tensors live back to back in one buffer (1.69 M floats for the Nature CNN), so "multi-tensor" is one
// launch: pass 1 reduces sum(g^2) deterministically (fp64 partials, last block finishes), pass 2 reads
// g, m, v, theta once and writes m, v, theta once (28 B per parameter) with the clip scale and an
// optional 1/world_size factor (gradient averaging after the all-reduce) folded in.
#include <math.h>
#include <stdlib.h>

#include "xa_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kVec = 4;
constexpr int kMaxBlocks = 148 * 8;

struct OptimWorkspace {
  unsigned int ticket;
  unsigned int pad;
  double sumsq;
  double partials[1];  // [kMaxBlocks]
};

__global__ void __launch_bounds__(kThreads) grad_sumsq_kernel(const float* __restrict__ g, int64_t n, OptimWorkspace* ws) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  __shared__ double scratch[kThreads / 32];
  __shared__ bool s_last;
  double acc = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t n4 = n / kVec;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const uint64_t keep = xa::policy_evict_last();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    const float4 x = xa::ld_keep(g4 + i, keep);
    acc += static_cast<double>(x.x) * x.x + static_cast<double>(x.y) * x.y + static_cast<double>(x.z) * x.z +
           static_cast<double>(x.w) * x.w;
  }
  for (int64_t i = n4 * kVec + static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride)
    acc += static_cast<double>(g[i]) * g[i];
  const double total = xa::block_sum<kThreads / 32>(acc, scratch);
  if (threadIdx.x == 0) {
    ws->partials[blockIdx.x] = total;
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double t = 0.0;
  for (unsigned k = threadIdx.x; k < gridDim.x; k += kThreads) t += const_cast<const volatile double*>(ws->partials)[k];
  t = xa::block_sum<kThreads / 32>(t, scratch);
  if (threadIdx.x == 0) {
    ws->sumsq = t;
    ws->ticket = 0;
  }
}

struct AdamParams {
  float* param;
  const float* grad;
  float* m;
  float* v;
  int64_t n;
  const OptimWorkspace* ws;
  float lr_t, beta1, beta2, eps, clip, grad_scale;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamParams& a, float scale) {
  g *= scale;
  m = a.beta1 * m + (1.0f - a.beta1) * g;
  v = a.beta2 * v + (1.0f - a.beta2) * g * g;
  p -= a.lr_t * m / (sqrtf(v) + a.eps);
}

__global__ void __launch_bounds__(kThreads) clip_adam_kernel(const AdamParams a) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  float scale = a.grad_scale;
  if (a.clip > 0.0f) {
    // tf.clip_by_global_norm: g * clip * min(1/norm, 1/clip)
    const float norm = a.grad_scale * static_cast<float>(sqrt(a.ws->sumsq));
    scale *= a.clip * fminf(1.0f / norm, 1.0f / a.clip);
  }
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t n4 = a.n / kVec;
  const uint64_t keep = xa::policy_evict_last();   // parameters, moments and gradient stay in L2 from one minibatch to the next
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < n4; i += stride) {
    float4 p = xa::ld_keep(reinterpret_cast<const float4*>(a.param) + i, keep);
    const float4 g = xa::ld_keep(reinterpret_cast<const float4*>(a.grad) + i, keep);
    float4 m = xa::ld_keep(reinterpret_cast<const float4*>(a.m) + i, keep);
    float4 v = xa::ld_keep(reinterpret_cast<const float4*>(a.v) + i, keep);
    adam_one(p.x, g.x, m.x, v.x, a, scale);
    adam_one(p.y, g.y, m.y, v.y, a, scale);
    adam_one(p.z, g.z, m.z, v.z, a, scale);
    adam_one(p.w, g.w, m.w, v.w, a, scale);
    xa::st_keep(reinterpret_cast<float4*>(a.param) + i, p, keep);
    xa::st_keep(reinterpret_cast<float4*>(a.m) + i, m, keep);
    xa::st_keep(reinterpret_cast<float4*>(a.v) + i, v, keep);
  }
  for (int64_t i = n4 * kVec + static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i < a.n; i += stride)
    adam_one(a.param[i], a.grad[i], a.m[i], a.v[i], a, scale);
}

// Benchmark stand-in for the network's backward pass (bench.py; the metric excludes the network contractions): every
// "parameter gradient" is a fixed combination of the loss kernel's outputs, so the flat gradient buffer -- the input of
// collective C1 and of the optimiser -- is PRODUCED by a kernel that consumes d loss / d(actor_out, critic_out) of this
// minibatch, as in training.  6.75 MB written per call for the Nature CNN's 1.69 M parameters.
__global__ void __launch_bounds__(kThreads) grad_from_outputs_kernel(const float* __restrict__ d_actor, const float* __restrict__ d_values,
                                                                     uint32_t n, uint32_t na, float* __restrict__ grad, int64_t n_params) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  // one float4 of "parameter gradients" per thread and pass; 32-bit index arithmetic (n_params, n * A < 2^31 are checked)
  const uint32_t n4 = static_cast<uint32_t>(n_params / 4);
  const uint32_t stride = gridDim.x * kThreads;
  const uint64_t keep = xa::policy_evict_last();
  for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q < n4; q += stride) {
    const uint32_t j = q * 4u;
    uint32_t ia = j % na, iv = j % n;
    float out[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      out[c] = __ldg(d_actor + ia) + 0.5f * __ldg(d_values + iv);
      ia = ia + 1 == na ? 0 : ia + 1;
      iv = iv + 1 == n ? 0 : iv + 1;
    }
    xa::st_keep(reinterpret_cast<float4*>(grad) + q, make_float4(out[0], out[1], out[2], out[3]), keep);
  }
  if (blockIdx.x == 0 && threadIdx.x < static_cast<uint32_t>(n_params % 4)) {
    const uint32_t j = n4 * 4u + threadIdx.x;
    grad[j] = __ldg(d_actor + j % na) + 0.5f * __ldg(d_values + j % n);
  }
}

// grid of the optimiser-chain kernels: XA_OPT_BLOCKS in the environment caps it (tuning: these kernels run beside the gather's
// one CTA per SM; read once)
int opt_block_cap() {
  static const int cap = [] {
    const char* e = getenv("XA_OPT_BLOCKS");
    const int v = e != nullptr ? atoi(e) : 0;
    return v > 0 && v < kMaxBlocks ? v : kMaxBlocks;
  }();
  return cap;
}

unsigned blocks_for(int64_t n) {
  const int64_t want = (n + kThreads * kVec - 1) / (kThreads * kVec);
  const int cap = opt_block_cap();
  return static_cast<unsigned>(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" {

int64_t xa_clip_adam_workspace_bytes(int64_t n) {
  (void)n;
  return 16 + static_cast<int64_t>(sizeof(double)) * kMaxBlocks;
}

int xa_grad_sumsq_f32(const float* grads, int64_t n, void* workspace, int64_t workspace_bytes, xa_stream_t stream) {
  XA_REQUIRE(grads && workspace, XA_EINVAL, "xa_grad_sumsq_f32: null pointer");
  XA_REQUIRE(n > 0, XA_EINVAL, "xa_grad_sumsq_f32: n=%lld", static_cast<long long>(n));
  XA_REQUIRE(workspace_bytes >= xa_clip_adam_workspace_bytes(n), XA_ENOSPACE, "xa_grad_sumsq_f32: workspace too small");
  XA_REQUIRE(xa::aligned(grads, 16) && xa::aligned(workspace, 16), XA_EALIGN, "xa_grad_sumsq_f32: 16-byte alignment required");
  xa::launch_chained(xa::kChainElementwise, grad_sumsq_kernel, dim3(blocks_for(n)), dim3(kThreads), 0, static_cast<cudaStream_t>(stream), grads, n, static_cast<OptimWorkspace*>(workspace));
  return xa::check_launch("xa_grad_sumsq_f32");
}

int xa_grad_from_outputs_f32(const float* d_actor, const float* d_values, int64_t n, int n_actions, float* grad, int64_t n_params,
                             xa_stream_t stream) {
  XA_REQUIRE(d_actor && d_values && grad, XA_EINVAL, "xa_grad_from_outputs_f32: null pointer");
  XA_REQUIRE(n > 0 && n_actions > 0 && n_params > 0, XA_EINVAL, "xa_grad_from_outputs_f32: n=%lld n_actions=%d n_params=%lld",
             static_cast<long long>(n), n_actions, static_cast<long long>(n_params));
  XA_REQUIRE(n_params < (int64_t(1) << 31) && n * n_actions < (int64_t(1) << 31), XA_EOVERFLOW, "xa_grad_from_outputs_f32: sizes exceed 32-bit indices");
  XA_REQUIRE(xa::aligned(grad, 16), XA_EALIGN, "xa_grad_from_outputs_f32: grad must be 16-byte aligned");
  const int64_t want = (n_params / 4 + kThreads - 1) / kThreads;
  const unsigned grid = static_cast<unsigned>(want < 1 ? 1 : (want > opt_block_cap() ? opt_block_cap() : want));
  xa::launch_chained(xa::kChainElementwise, grad_from_outputs_kernel, dim3(grid), dim3(kThreads), 0, static_cast<cudaStream_t>(stream), d_actor, d_values, static_cast<uint32_t>(n),
                     static_cast<uint32_t>(n * n_actions), grad, n_params);
  return xa::check_launch("xa_grad_from_outputs_f32");
}

int xa_clip_adam_f32(float* param, const float* grad, float* m, float* v, int64_t n, const void* workspace, double lr,
                     double beta1, double beta2, double eps, double clip_norm, int64_t step, double grad_scale,
                     xa_stream_t stream) {
  XA_REQUIRE(param && grad && m && v, XA_EINVAL, "xa_clip_adam_f32: null pointer");
  XA_REQUIRE(n > 0 && step > 0, XA_EINVAL, "xa_clip_adam_f32: n=%lld step=%lld must be positive", static_cast<long long>(n),
             static_cast<long long>(step));
  XA_REQUIRE(clip_norm <= 0.0 || workspace != nullptr, XA_EINVAL, "xa_clip_adam_f32: clipping needs the sumsq workspace");
  XA_REQUIRE(xa::aligned(param, 16) && xa::aligned(grad, 16) && xa::aligned(m, 16) && xa::aligned(v, 16), XA_EALIGN,
             "xa_clip_adam_f32: 16-byte alignment required");
  AdamParams a;
  a.param = param;
  a.grad = grad;
  a.m = m;
  a.v = v;
  a.n = n;
  a.ws = static_cast<const OptimWorkspace*>(workspace);
  // Keras Adam: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
  a.lr_t = static_cast<float>(lr * sqrt(1.0 - pow(beta2, static_cast<double>(step))) / (1.0 - pow(beta1, static_cast<double>(step))));
  a.beta1 = static_cast<float>(beta1);
  a.beta2 = static_cast<float>(beta2);
  a.eps = static_cast<float>(eps);
  a.clip = static_cast<float>(clip_norm);
  a.grad_scale = static_cast<float>(grad_scale);
  xa::launch_chained(xa::kChainElementwise, clip_adam_kernel, dim3(blocks_for(n)), dim3(kThreads), 0, static_cast<cudaStream_t>(stream), a);
  return xa::check_launch("xa_clip_adam_f32");
}

}  // extern "C"
