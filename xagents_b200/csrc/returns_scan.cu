// Reverse segmented scans over the rollout: GAE returns (PPO) and discounted n-step returns (A2C).
//
// Replaces PPO.calculate_returns (xagents/ppo/agent.py:80-94) and A2C.calculate_returns
// (xagents/a2c/agent.py:165-171), which walk T in a Python loop of NumPy calls on [E] vectors.
//
// Both recurrences are x_t = b_t + c_t * x_{t+1} with c_t = 0 wherever the next step is terminal, so a
// done flag cuts the carry exactly like a segment head.  Layout is time-major
// [T, E]: lanes run along E (every load/store is a coalesced 128-B line), warps run along T.
// Warp w of a block owns a chunk of kChunk consecutive steps for 32 envs; it loads the chunk once into
// registers (3*kChunk+1 independent loads per thread, issued one whole span AHEAD of their use so the
// latency hides behind the current span's arithmetic and barriers), reduces it to the affine map (C, B),
// exchanges maps through shared memory, derives its carry-in from the later chunks and then replays its
// chunk from registers.  Inputs are read once and outputs written once: 16 B per env-step for GAE,
// 12 B for n-step returns.  With one warp per block (XA_SCAN_SEQUENTIAL) no map is ever composed and the
// arithmetic is the reference's, operation for operation (explicit _rn intrinsics: no FMA contraction),
// so results are bit-identical to the NumPy loop.
#include "xa_common.cuh"

namespace {

constexpr int kChunk = 8;
constexpr int kMaxWarps = 16;

struct ScanParams {
  const float* rewards;
  const float* values;       // GAE only
  const float* last_values;  // [E] bootstrap
  const float* dones;        // [T+1, E]
  float* returns;
  float* advantages;  // GAE only, may be null
  int n_steps;
  int n_envs;
  float gamma;
  float gamma_lam;  // fp32(double(gamma) * double(lam)), folded as the reference folds it
};

// One chunk of raw rollout values, exactly as loaded (kept apart from the derived maps so that the NEXT
// chunk's loads can be in flight while the current chunk is reduced, exchanged and replayed).
template <bool kNstep>
struct RawChunk {
  float rew[kChunk];
  float done[kChunk];
  float val[kNstep ? 1 : kChunk];
  float v_next;
};

// Loads are unconditional: rows past the chunk end are clamped to its last valid row (the value is loaded
// and ignored), so there is no divergence and all 3*kChunk+1 requests leave back to back.
template <bool kNstep>
__device__ __forceinline__ void load_chunk(RawChunk<kNstep>& q, const ScanParams& p, int t0, int len, int env) {
  const int T = p.n_steps;
  const size_t E = static_cast<size_t>(p.n_envs);
#pragma unroll
  for (int j = 0; j < kChunk; ++j) {
    const int t = t0 + (j < len ? j : len - 1);
    const size_t o = static_cast<size_t>(t) * E + env;
    q.rew[j] = p.rewards[o];
    q.done[j] = p.dones[o + E];  // row t+1 gates step t (a2c/agent.py:116,129,138)
    if (!kNstep) q.val[j] = p.values[o];
  }
  if (!kNstep) {
    const int t1 = t0 + len;
    q.v_next = (t1 == T) ? p.last_values[env] : p.values[static_cast<size_t>(t1) * E + env];
  }
}

template <bool kNstep, bool kMulti>
__global__ void __launch_bounds__(32 * kMaxWarps) returns_scan_kernel(const ScanParams p) {
  __shared__ float s_c[kMulti ? kMaxWarps : 1][32];
  __shared__ float s_b[kMulti ? kMaxWarps : 1][32];
  __shared__ float s_carry[32];
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory

  const int lane = threadIdx.x;
  const int w = threadIdx.y;
  const int n_warps = blockDim.y;
  const int env_raw = blockIdx.x * 32 + lane;
  const bool active = env_raw < p.n_envs;
  const int env = active ? env_raw : p.n_envs - 1;  // inactive lanes shadow the last env and never store
  const int T = p.n_steps;
  const size_t E = static_cast<size_t>(p.n_envs);
  const float gamma = p.gamma;
  const float gl = p.gamma_lam;
  const int span = n_warps * kChunk;

  // value flowing in from the future: R_T for n-step returns (a2c/agent.py:165-166), 0 for GAE (ppo/agent.py:81)
  float carry = kNstep ? p.last_values[env] : 0.0f;

  // chunk of this warp inside the span ending at `hi` (exclusive): [t0, t0+len)
  auto chunk_of = [&](int hi, int& t0, int& len) {
    const int t1 = hi - (n_warps - 1 - w) * kChunk;
    t0 = t1 - kChunk > 0 ? t1 - kChunk : 0;
    len = t1 > t0 ? t1 - t0 : 0;
  };

  RawChunk<kNstep> nxt;
  int t0, len;
  chunk_of(T, t0, len);
  if (len > 0) load_chunk<kNstep>(nxt, p, t0, len, env);

  for (int hi = T; hi > 0; hi -= span) {
    const RawChunk<kNstep> cur = nxt;
    const int c_t0 = t0, c_len = len;
    if (hi - span > 0) {  // put the next span's loads in flight before touching this one
      chunk_of(hi - span, t0, len);
      if (len > 0) load_chunk<kNstep>(nxt, p, t0, len, env);
    }

    float coef[kChunk];  // GAE: gamma*lam*alive ; n-step: alive
    float bias[kChunk];  // GAE: delta          ; n-step: reward
#pragma unroll
    for (int j = 0; j < kChunk; ++j) {
      const float alive = __fsub_rn(1.0f, cur.done[j]);  // ppo/agent.py:85
      if (kNstep) {
        coef[j] = alive;
        bias[j] = cur.rew[j];
      } else {
        const float vn = (j + 1 < kChunk && j + 1 < c_len) ? cur.val[j + 1 < kChunk ? j + 1 : j] : cur.v_next;
        // delta = r + gamma*V_{t+1}*alive - V_t, evaluated left to right (ppo/agent.py:87-91)
        bias[j] = __fsub_rn(__fadd_rn(cur.rew[j], __fmul_rn(__fmul_rn(gamma, vn), alive)), cur.val[j]);
        coef[j] = __fmul_rn(gl, alive);  // ppo/agent.py:92
      }
    }

    float x = carry;
    if (kMulti) {
      // reduce the chunk to x -> B + C*x and publish it
      float cc = 1.0f, bb = 0.0f;
#pragma unroll
      for (int j = kChunk - 1; j >= 0; --j) {
        if (j < c_len) {
          const float c = kNstep ? __fmul_rn(gamma, coef[j]) : coef[j];
          bb = __fadd_rn(bias[j], __fmul_rn(c, bb));
          cc = __fmul_rn(cc, c);
        }
      }
      s_c[w][lane] = cc;
      s_b[w][lane] = bb;
      __syncthreads();
      for (int w2 = n_warps - 1; w2 > w; --w2) x = __fadd_rn(s_b[w2][lane], __fmul_rn(s_c[w2][lane], x));
    }

#pragma unroll
    for (int j = kChunk - 1; j >= 0; --j) {
      if (j < c_len) {
        const size_t o = static_cast<size_t>(c_t0 + j) * E + env;
        if (kNstep) {
          // R_t = r_t + gamma*R_{t+1}*(1-d_{t+1})   (a2c/agent.py:168-170)
          x = __fadd_rn(bias[j], __fmul_rn(__fmul_rn(gamma, x), coef[j]));
          if (active) p.returns[o] = x;
        } else {
          // last_lam = delta + gamma*lam*alive*last_lam ; returns = last_lam + V   (ppo/agent.py:92-94)
          x = __fadd_rn(bias[j], __fmul_rn(coef[j], x));
          if (active) {
            p.returns[o] = __fadd_rn(x, cur.val[j]);
            if (p.advantages != nullptr) p.advantages[o] = x;
          }
        }
      }
    }

    if (kMulti) {
      if (w == 0) s_carry[lane] = x;  // earliest chunk of this span: its value is what flows further back
      __syncthreads();
      carry = s_carry[lane];
    } else {
      carry = x;
    }
  }
}

// ACER Retrace targets (xagents/acer/agent.py:198-208): a reverse recurrence with an importance-weighted
// correction, R_t = r_t + gamma*c*(1-d_{t+1}); out_t = R_t; c = min(1,rho_t)*(R_t - Q_t) + V_t.  One thread per
// env walks T with the next group of steps already loaded (same prefetch idea as above); the arithmetic is the
// reference's, operation for operation.  ACER rollouts are short (n_steps ~ 16-20), so T is not split.
struct RetraceParams {
  const float* rewards;
  const float* dones;  // [T+1, E]
  const float* values;
  const float* last_values;
  const float* q_selected;
  const float* importance;
  float* returns;
  int n_steps, n_envs;
  float gamma;
};

constexpr int kRetraceGroup = 4;

struct RetraceRaw {
  float rew[kRetraceGroup], done[kRetraceGroup], val[kRetraceGroup], q[kRetraceGroup], imp[kRetraceGroup];
};

__device__ __forceinline__ void load_retrace(RetraceRaw& r, const RetraceParams& p, int t1, int env) {
  const size_t E = static_cast<size_t>(p.n_envs);
#pragma unroll
  for (int j = 0; j < kRetraceGroup; ++j) {
    const int t = t1 - 1 - j > 0 ? t1 - 1 - j : 0;  // clamped: loaded and ignored past the front
    const size_t o = static_cast<size_t>(t) * E + env;
    r.rew[j] = p.rewards[o];
    r.done[j] = p.dones[o + E];
    r.val[j] = p.values[o];
    r.q[j] = p.q_selected[o];
    r.imp[j] = p.importance[o];
  }
}

__global__ void __launch_bounds__(128) retrace_kernel(const RetraceParams p) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= p.n_envs) return;
  const size_t E = static_cast<size_t>(p.n_envs);
  float current = p.last_values[env];
  RetraceRaw nxt;
  load_retrace(nxt, p, p.n_steps, env);
  for (int t1 = p.n_steps; t1 > 0; t1 -= kRetraceGroup) {
    const RetraceRaw cur = nxt;
    if (t1 - kRetraceGroup > 0) load_retrace(nxt, p, t1 - kRetraceGroup, env);
#pragma unroll
    for (int j = 0; j < kRetraceGroup; ++j) {
      const int t = t1 - 1 - j;
      if (t >= 0) {
        current = __fadd_rn(cur.rew[j], __fmul_rn(__fmul_rn(p.gamma, current), __fsub_rn(1.0f, cur.done[j])));
        p.returns[static_cast<size_t>(t) * E + env] = current;
        current = __fadd_rn(__fmul_rn(fminf(1.0f, cur.imp[j]), __fsub_rn(current, cur.q[j])), cur.val[j]);
      }
    }
  }
}

int pick_warps(int mode, int n_steps, int n_envs) {
  const int chunks = (n_steps + kChunk - 1) / kChunk;
  if (mode == XA_SCAN_SEQUENTIAL || chunks <= 1) return 1;
  const int max_w = chunks < kMaxWarps ? chunks : kMaxWarps;
  if (mode == XA_SCAN_CHUNKED) return max_w > 1 ? max_w : 2;
  // AUTO: with the next span prefetched, ~12 resident warps per SM already saturate HBM (measured: one
  // warp per block is the fastest once there are >= 12 blocks per SM); split T across warps only below that
  const long blocks = (n_envs + 31) / 32;
  const long want = 12L * (xa::sm_count() > 0 ? xa::sm_count() : 148);
  long wps = (want + blocks - 1) / blocks;
  if (wps < 1) wps = 1;
  if (wps > max_w) wps = max_w;
  return static_cast<int>(wps);
}

template <bool kNstep>
int launch(const ScanParams& p, int mode, cudaStream_t stream, const char* what) {
  const int n_warps = pick_warps(mode, p.n_steps, p.n_envs);
  const dim3 block(32, n_warps);
  const dim3 grid((p.n_envs + 31) / 32);
  if (n_warps == 1)
    xa::launch_chained(xa::kChainElementwise, returns_scan_kernel<kNstep, false>, dim3(grid), dim3(block), 0, stream, p);
  else
    xa::launch_chained(xa::kChainElementwise, returns_scan_kernel<kNstep, true>, dim3(grid), dim3(block), 0, stream, p);
  return xa::check_launch(what);
}

int check_shape(const char* what, int n_steps, int n_envs, int mode) {
  XA_REQUIRE(n_steps > 0 && n_envs > 0, XA_EINVAL, "%s: n_steps=%d n_envs=%d must be positive", what, n_steps, n_envs);
  XA_REQUIRE(mode >= XA_SCAN_AUTO && mode <= XA_SCAN_CHUNKED, XA_EINVAL, "%s: unknown mode %d", what, mode);
  XA_REQUIRE((static_cast<int64_t>(n_steps) + 1) * n_envs < (int64_t(1) << 40), XA_EOVERFLOW, "%s: rollout too large", what);
  return XA_OK;
}

}  // namespace

extern "C" {

int xa_gae_f32(const float* rewards, const float* values, const float* last_values, const float* dones, float* returns,
               float* advantages, int n_steps, int n_envs, double gamma, double lam, int mode, xa_stream_t stream) {
  if (int rc = check_shape("xa_gae_f32", n_steps, n_envs, mode)) return rc;
  XA_REQUIRE(rewards && values && last_values && dones && returns, XA_EINVAL, "xa_gae_f32: null pointer");
  XA_REQUIRE(xa::aligned(rewards, 4) && xa::aligned(values, 4) && xa::aligned(last_values, 4) && xa::aligned(dones, 4) &&
                 xa::aligned(returns, 4) && xa::aligned(advantages, 4),
             XA_EALIGN, "xa_gae_f32: pointers must be 4-byte aligned");
  ScanParams p{rewards, values, last_values, dones, returns, advantages, n_steps, n_envs, static_cast<float>(gamma),
               static_cast<float>(gamma * lam)};
  return launch<false>(p, mode, static_cast<cudaStream_t>(stream), "xa_gae_f32");
}

int xa_nstep_returns_f32(const float* rewards, const float* dones, const float* last_values, float* returns, int n_steps,
                         int n_envs, double gamma, int mode, xa_stream_t stream) {
  if (int rc = check_shape("xa_nstep_returns_f32", n_steps, n_envs, mode)) return rc;
  XA_REQUIRE(rewards && dones && last_values && returns, XA_EINVAL, "xa_nstep_returns_f32: null pointer");
  XA_REQUIRE(xa::aligned(rewards, 4) && xa::aligned(dones, 4) && xa::aligned(last_values, 4) && xa::aligned(returns, 4),
             XA_EALIGN, "xa_nstep_returns_f32: pointers must be 4-byte aligned");
  ScanParams p{rewards, nullptr, last_values, dones, returns, nullptr, n_steps, n_envs, static_cast<float>(gamma), 0.0f};
  return launch<true>(p, mode, static_cast<cudaStream_t>(stream), "xa_nstep_returns_f32");
}

int xa_retrace_f32(const float* rewards, const float* dones, const float* values, const float* last_values,
                   const float* q_selected, const float* importance, float* returns, int n_steps, int n_envs, double gamma,
                   xa_stream_t stream) {
  if (int rc = check_shape("xa_retrace_f32", n_steps, n_envs, XA_SCAN_AUTO)) return rc;
  XA_REQUIRE(rewards && dones && values && last_values && q_selected && importance && returns, XA_EINVAL,
             "xa_retrace_f32: null pointer");
  RetraceParams p{rewards, dones, values, last_values, q_selected, importance, returns, n_steps, n_envs, static_cast<float>(gamma)};
  retrace_kernel<<<(n_envs + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return xa::check_launch("xa_retrace_f32");
}

}  // extern "C"
