// Rollout-time policy step: sample an action, its log-prob and the entropy for a batch of environments.
//
// Replaces, per environment step, `distribution.sample()`, `distribution.log_prob(actions)` and
// `distribution.entropy()` of A2C.get_model_outputs (xagents/a2c/agent.py:80-94; tfp Categorical /
// MultivariateNormalDiag built at :50-63).  One thread per environment; results go straight into row t of
// the time-major rollout buffers (the caller passes the row pointers), so nothing is staged on the host.
// Categorical sampling is Gumbel-max over the log-softmax (what tf.random.categorical does); the noise is
// either supplied (tests: identical draws on both sides) or generated in-kernel with Philox4x32-10.
#include <math.h>

#include "xa_common.cuh"

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ float u01(uint32_t bits) { return (static_cast<float>(bits >> 8) + 0.5f) * (1.0f / 16777216.0f); }

struct PolicyParams {
  const float* actor_out;
  const float* noise;
  uint64_t seed, offset;
  const uint64_t* offset_ptr;  // when set, the Philox offset is read from device memory (a replayed CUDA graph must not repeat its draws)
  float* actions;
  float* log_probs;
  float* entropies;
  int64_t n;
  int n_actions;
};

template <int kActorKind>
__global__ void __launch_bounds__(128) policy_step_kernel(const PolicyParams p) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int A = p.n_actions;
  const float* row = p.actor_out + i * A;
  const uint2 key = make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));
  const uint64_t offset = (p.offset_ptr ? *p.offset_ptr : 0) + p.offset;
  uint4 rnd = make_uint4(0, 0, 0, 0);
  auto draw = [&](int j) -> uint32_t {  // j-th 32-bit word of this sample's Philox stream
    if ((j & 3) == 0)
      rnd = philox4x32_10(make_uint4(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32),
                                     static_cast<uint32_t>(offset + (j >> 2)), static_cast<uint32_t>((offset + (j >> 2)) >> 32)),
                          key);
    const int k = j & 3;
    return k == 0 ? rnd.x : (k == 1 ? rnd.y : (k == 2 ? rnd.z : rnd.w));
  };

  if (kActorKind == XA_ACTOR_NORMAL) {
    // MultivariateNormalDiag(loc), identity scale: a = loc + z
    const float half_k_log2pi = 0.5f * static_cast<float>(A) * 1.8378770664093453f;
    float ss = 0.0f;
    for (int j = 0; j < A; ++j) {
      float z;
      if (p.noise) {
        z = p.noise[i * A + j];
      } else {  // Box-Muller
        const float u1 = u01(draw(2 * j)), u2 = u01(draw(2 * j + 1));
        z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
      }
      p.actions[i * A + j] = row[j] + z;
      ss += z * z;
    }
    p.log_probs[i] = -0.5f * ss - half_k_log2pi;
    if (p.entropies) p.entropies[i] = 0.5f * static_cast<float>(A) + half_k_log2pi;
    return;
  }

  float zmax = -INFINITY;
  for (int j = 0; j < A; ++j) zmax = fmaxf(zmax, kActorKind == XA_ACTOR_PROBS ? logf(row[j]) : row[j]);
  float se = 0.0f;
  for (int j = 0; j < A; ++j) se += expf((kActorKind == XA_ACTOR_PROBS ? logf(row[j]) : row[j]) - zmax);
  const float lse = logf(se);
  float best = -INFINITY, best_lsm = 0.0f, ent = 0.0f;
  int arg = 0;
  for (int j = 0; j < A; ++j) {
    const float l = ((kActorKind == XA_ACTOR_PROBS ? logf(row[j]) : row[j]) - zmax) - lse;
    const float pl = expf(l);
    ent -= pl > 0.0f ? pl * l : 0.0f;  // tfp's multiply_no_nan: a probability of exactly 0 contributes 0, not NaN
    const float u = p.noise ? p.noise[i * A + j] : u01(draw(j));
    const float score = l - logf(-logf(u));  // Gumbel-max
    if (score > best) {
      best = score;
      best_lsm = l;
      arg = j;
    }
  }
  p.actions[i] = static_cast<float>(arg);  // fp32-encoded ids, as the rollout stores them (ppo/agent.py:202-210)
  p.log_probs[i] = best_lsm;
  if (p.entropies) p.entropies[i] = ent;
}

__global__ void bump_u64_kernel(uint64_t* p, uint64_t delta) {
  xa::pdl_trigger();
  xa::pdl_wait();
  *p += delta;
}

int launch_policy(const PolicyParams& p, int actor_kind, cudaStream_t s, const char* what) {
  const unsigned grid = static_cast<unsigned>((p.n + 127) / 128);
  switch (actor_kind) {
    case XA_ACTOR_LOGITS: xa::launch_chained(xa::kChainSmall, policy_step_kernel<XA_ACTOR_LOGITS>, dim3(grid), dim3(128), 0, s, p); break;
    case XA_ACTOR_PROBS: xa::launch_chained(xa::kChainSmall, policy_step_kernel<XA_ACTOR_PROBS>, dim3(grid), dim3(128), 0, s, p); break;
    default: xa::launch_chained(xa::kChainSmall, policy_step_kernel<XA_ACTOR_NORMAL>, dim3(grid), dim3(128), 0, s, p); break;
  }
  return xa::check_launch(what);
}

}  // namespace

// The same step with the Philox offset held in DEVICE memory: the kernel draws at *offset_dev + offset and, when `advance` is
// not zero, a one-thread kernel behind it adds `advance` to *offset_dev, so a CUDA graph that captured the call draws fresh
// noise on every replay.  A captured rollout passes offset = t * step and advance = 0 for its steps and advances the counter
// once at its end (xa_bump_u64): one launch per step instead of two.
extern "C" int xa_policy_step_counter_f32(const float* actor_out, int actor_kind, uint64_t seed, uint64_t* offset_dev, uint64_t offset,
                                          uint64_t advance, float* actions, float* log_probs, float* entropies, int64_t n, int n_actions, xa_stream_t stream) {
  XA_REQUIRE(n > 0 && n_actions > 0, XA_EINVAL, "xa_policy_step_counter_f32: n=%lld n_actions=%d", static_cast<long long>(n), n_actions);
  XA_REQUIRE(actor_out && actions && log_probs && offset_dev, XA_EINVAL, "xa_policy_step_counter_f32: null pointer");
  XA_REQUIRE(xa::aligned(offset_dev, 8), XA_EALIGN, "xa_policy_step_counter_f32: offset_dev must be 8-byte aligned");
  XA_REQUIRE(actor_kind >= XA_ACTOR_LOGITS && actor_kind <= XA_ACTOR_NORMAL, XA_EINVAL, "xa_policy_step_counter_f32: unknown actor_kind %d", actor_kind);
  PolicyParams p{actor_out, nullptr, seed, offset, offset_dev, actions, log_probs, entropies, n, n_actions};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = launch_policy(p, actor_kind, s, "xa_policy_step_counter_f32")) return rc;
  if (advance == 0) return XA_OK;
  xa::launch_chained(xa::kChainSmall, bump_u64_kernel, dim3(1), dim3(1), 0, s, offset_dev, advance);
  return xa::check_launch("xa_policy_step_counter_f32");
}

extern "C" int xa_policy_step_f32(const float* actor_out, int actor_kind, const float* noise, uint64_t seed, uint64_t offset,
                                  float* actions, float* log_probs, float* entropies, int64_t n, int n_actions,
                                  xa_stream_t stream) {
  XA_REQUIRE(n >= 0 && n_actions > 0, XA_EINVAL, "xa_policy_step_f32: n=%lld n_actions=%d", static_cast<long long>(n), n_actions);
  if (n == 0) return XA_OK;
  XA_REQUIRE(actor_out && actions && log_probs, XA_EINVAL, "xa_policy_step_f32: null pointer");
  XA_REQUIRE(actor_kind >= XA_ACTOR_LOGITS && actor_kind <= XA_ACTOR_NORMAL, XA_EINVAL, "xa_policy_step_f32: unknown actor_kind %d",
             actor_kind);
  PolicyParams p{actor_out, noise, seed, offset, nullptr, actions, log_probs, entropies, n, n_actions};
  return launch_policy(p, actor_kind, static_cast<cudaStream_t>(stream), "xa_policy_step_f32");
}
