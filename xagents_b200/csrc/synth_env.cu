// One environment step of the device-resident synthetic Atari environments (xagents_b200/envs.py, BatchedSyntheticAtari)
// fused with the bookkeeping BaseAgent.step_envs does around it (xagents/base.py:408-426): ONE launch per rollout step
// instead of ~19 elementwise / index_select / where launches of a few microseconds each (the rollout at 256 environments
// is launch-bound: ncu launch list in profiles/r2_rollout_launches.md).
//
// Per environment e (SURVEY.md 8d's synthetic distribution: i.i.d. frames from a pool, sparse +-1 rewards, geometric episodes):
//   u0, u1, u2, i_new, i_reset  <- Philox4x32-10 keyed by (seed), counter (e, *offset_dev + offset)
//   reward  = u0 < p_reward ? (u1 < 0.5 ? +1 : -1) : 0
//   done    = u2 < p_done
//   new_states[e] = pool[i_new]                      the frame the step returns (the TERMINAL frame of a finished episode)
//   states[e]     = done ? pool[i_reset] : pool[i_new]   what the environment holds afterwards (already reset, base.py:419-424)
//   sums[e] += reward; sums_log[e] = sums[e]; sums[e] *= 1 - done     (episode reward totals, base.py:414-423)
// Frame rows move as 16-byte vectors; two CTAs per environment keep 148 SMs busy at 256 environments.
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

struct SynthParams {
  const uint8_t* pool;
  uint8_t* states;
  uint8_t* new_states;
  float* rewards;
  float* dones;
  float* sums;
  float* sums_log;
  const uint64_t* offset_dev;
  uint64_t seed, offset;
  int64_t row_bytes;
  int n_envs, pool_size, ctas_per_env;
  float p_reward, p_done;
};

__global__ void __launch_bounds__(256) synth_env_step_kernel(const SynthParams p) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  const int e = blockIdx.x / p.ctas_per_env, part = blockIdx.x - e * p.ctas_per_env;
  const uint2 key = make_uint2(static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));
  const uint64_t off = (p.offset_dev ? *p.offset_dev : 0) + p.offset;
  const uint4 a = philox4x32_10(make_uint4(static_cast<uint32_t>(e), 0u, static_cast<uint32_t>(off), static_cast<uint32_t>(off >> 32)), key);
  const uint4 b = philox4x32_10(make_uint4(static_cast<uint32_t>(e), 1u, static_cast<uint32_t>(off), static_cast<uint32_t>(off >> 32)), key);
  const float reward = u01(a.x) < p.p_reward ? (u01(a.y) < 0.5f ? 1.0f : -1.0f) : 0.0f;
  const bool done = u01(a.z) < p.p_done;
  const uint32_t i_new = __umulhi(b.x, static_cast<uint32_t>(p.pool_size)), i_reset = __umulhi(b.y, static_cast<uint32_t>(p.pool_size));
  if (part == 0 && threadIdx.x == 0) {
    p.rewards[e] = reward;
    p.dones[e] = done ? 1.0f : 0.0f;
    if (p.sums != nullptr) {
      const float s = p.sums[e] + reward;
      if (p.sums_log != nullptr) p.sums_log[e] = s;
      p.sums[e] = s * (1.0f - (done ? 1.0f : 0.0f));
    }
  }
  const int64_t vecs = p.row_bytes >> 4;
  const int4* src_new = reinterpret_cast<const int4*>(p.pool + static_cast<int64_t>(i_new) * p.row_bytes);
  const int4* src_held = reinterpret_cast<const int4*>(p.pool + static_cast<int64_t>(done ? i_reset : i_new) * p.row_bytes);
  int4* dst_new = p.new_states ? reinterpret_cast<int4*>(p.new_states + static_cast<int64_t>(e) * p.row_bytes) : nullptr;
  int4* dst_held = reinterpret_cast<int4*>(p.states + static_cast<int64_t>(e) * p.row_bytes);
  for (int64_t v = static_cast<int64_t>(part) * blockDim.x + threadIdx.x; v < vecs; v += static_cast<int64_t>(p.ctas_per_env) * blockDim.x) {
    const int4 x = __ldg(src_new + v);
    if (dst_new != nullptr) xa::st_stream(dst_new + v, x);
    xa::st_stream(dst_held + v, done ? __ldg(src_held + v) : x);
  }
}

__global__ void bump_u64_kernel(uint64_t* p, uint64_t delta) {
  xa::pdl_trigger();
  xa::pdl_wait();
  *p += delta;
}

}  // namespace

extern "C" int xa_synth_env_step_u8(const uint8_t* pool, int pool_size, int64_t row_bytes, uint8_t* states, uint8_t* new_states, float* rewards,
                                    float* dones, float* episode_sums, float* sums_log, int n_envs, float p_reward, float p_done, uint64_t seed,
                                    const uint64_t* offset_dev, uint64_t offset, xa_stream_t stream) {
  const char* what = "xa_synth_env_step_u8";
  XA_REQUIRE(n_envs > 0 && pool_size > 0 && row_bytes > 0, XA_EINVAL, "%s: n_envs=%d pool_size=%d row_bytes=%lld", what, n_envs, pool_size,
             static_cast<long long>(row_bytes));
  XA_REQUIRE(pool && states && rewards && dones, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(row_bytes % 16 == 0 && xa::aligned(pool, 16) && xa::aligned(states, 16) && xa::aligned(new_states, 16), XA_EALIGN,
             "%s: frame rows must be 16-byte aligned multiples of 16 bytes", what);
  XA_REQUIRE(xa::aligned(offset_dev, 8), XA_EALIGN, "%s: offset_dev must be 8-byte aligned", what);
  XA_REQUIRE(states != new_states, XA_EINVAL, "%s: states and new_states must be different buffers", what);
  SynthParams p{pool, states, new_states, rewards, dones, episode_sums, sums_log, offset_dev, seed, offset, row_bytes, n_envs, pool_size, 1,
                p_reward, p_done};
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  int per = (4 * sms + n_envs - 1) / n_envs;  // about four CTAs per SM when there are few environments
  const int64_t max_per = (row_bytes / 16 + 255) / 256;
  if (per > max_per) per = static_cast<int>(max_per);
  if (per < 1) per = 1;
  p.ctas_per_env = per;
  xa::launch_chained(xa::kChainSmall, synth_env_step_kernel, dim3(static_cast<unsigned>(n_envs) * per), dim3(256), 0, static_cast<cudaStream_t>(stream), p);
  return xa::check_launch(what);
}

// *counter += delta on the stream: Philox offsets kept in device memory are advanced once per captured rollout, not per step.
extern "C" int xa_bump_u64(uint64_t* counter, uint64_t delta, xa_stream_t stream) {
  XA_REQUIRE(counter != nullptr && xa::aligned(counter, 8), XA_EINVAL, "xa_bump_u64: counter must be a non-null 8-byte aligned device pointer");
  xa::launch_chained(xa::kChainSmall, bump_u64_kernel, dim3(1), dim3(1), 0, static_cast<cudaStream_t>(stream), counter, delta);
  return xa::check_launch("xa_bump_u64");
}
