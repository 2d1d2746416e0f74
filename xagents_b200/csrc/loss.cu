// Advantage moments and the fused PPO / A2C loss (forward + the gradients autodiff would return).
//
// Replaces, per minibatch, the ~30 small TF ops of PPO.run_ppo_epochs / PPO.update_gradients
// (xagents/ppo/agent.py:180-183 and :112-134) and of A2C.train_step (xagents/a2c/agent.py:202-215),
// including the tfp Categorical / MultivariateNormalDiag log_prob + entropy of a2c/agent.py:50-94.
//
// One thread owns one sample: it reads the actor row, the critic value and the four rollout scalars
// (optionally straight from the time-major rollout through the minibatch indices, so the scalar gathers
// never materialise), evaluates log-softmax / entropy / ratio / both clips, writes dlogits and dV, and
// contributes to three block sums.  Every mean divides by the known n, so forward and backward are a
// single pass: (8A + 24) B per sample.  Scalar sums are carried in fp64, one partial per block, and the
// last block to finish (ticket counter in the workspace) adds the partials in block order -- the result
// does not depend on scheduling.
#include <math.h>

#include "xa_common.cuh"

namespace {

constexpr int kLossThreads = 256;
constexpr int kLossWarps = kLossThreads / 32;
constexpr int kMomentThreads = 1024;
constexpr int kMaxRegActions = 8;  // actor rows up to this width live in registers

constexpr int kMaxMomentBatches = 64;  // minibatch offsets ride in kernel-parameter space

struct MomentOffsets {
  int64_t v[kMaxMomentBatches + 1];
};

struct MomentParams {
  const float* returns;
  const float* old_values;
  const int32_t* idx;
  int n_steps, n_envs;
  double* moments;
};

// One block per minibatch, two passes (mean, then squared deviations) exactly like reduce_std's
// definition; the 2nd pass re-reads a few KB that are still in L1/L2.
__global__ void __launch_bounds__(kMomentThreads) adv_moments_kernel(const MomentParams p, const MomentOffsets offsets) {
  __shared__ double scratch[kMomentThreads / 32];
  __shared__ double s_mean;
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  const int m = blockIdx.x;
  const int64_t lo = offsets.v[m], hi = offsets.v[m + 1];
  const int64_t n = hi - lo;
  double acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kMomentThreads) {
    const int64_t row = p.idx ? xa::sample_row(p.idx[i], p.n_steps, p.n_envs) : i;
    acc += static_cast<double>(__fsub_rn(p.returns[row], p.old_values[row]));  // adv = R - V_old (ppo/agent.py:180)
  }
  double total = xa::block_sum<kMomentThreads / 32>(acc, scratch);
  if (threadIdx.x == 0) s_mean = n > 0 ? total / static_cast<double>(n) : 0.0;
  __syncthreads();
  const double mean = s_mean;
  acc = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kMomentThreads) {
    const int64_t row = p.idx ? xa::sample_row(p.idx[i], p.n_steps, p.n_envs) : i;
    const double d = static_cast<double>(__fsub_rn(p.returns[row], p.old_values[row])) - mean;
    acc += d * d;
  }
  total = xa::block_sum<kMomentThreads / 32>(acc, scratch);
  if (threadIdx.x == 0) {
    double* out = p.moments + static_cast<int64_t>(m) * XA_MOMENT_STRIDE;
    out[0] = static_cast<double>(n);
    out[1] = mean;
    out[2] = total;
    out[3] = 0.0;
  }
}

// (adv - mean) / (std + eps) for one minibatch from its moment parts (ppo/agent.py:181-183)
struct NormParams {
  const float* returns;
  const float* old_values;
  const int32_t* idx;
  int n_steps, n_envs;
  const double* moments;
  int n_parts;
  int64_t part_stride;
  float eps;
  float* out;
  int64_t n;
};

__device__ __forceinline__ void combine_moments(const double* moments, int n_parts, int64_t stride, float eps, float* mean,
                                                float* denom) {
  // Chan's pairwise combination of (count, mean, M2) parts, in part order, in fp64
  double cn = 0.0, cm = 0.0, c2 = 0.0;
  for (int q = 0; q < n_parts; ++q) {
    const double* part = moments + static_cast<int64_t>(q) * stride;
    const double qn = part[0], qm = part[1], q2 = part[2];
    if (qn <= 0.0) continue;
    const double tot = cn + qn, delta = qm - cm;
    c2 += q2 + delta * delta * cn * qn / tot;
    cm += delta * qn / tot;
    cn = tot;
  }
  *mean = static_cast<float>(cm);
  *denom = __fadd_rn(static_cast<float>(sqrt(cn > 0.0 ? c2 / cn : 0.0)), eps);  // population std + eps
}

__global__ void __launch_bounds__(256) normalize_adv_kernel(const NormParams p) {
  __shared__ float s_mean, s_denom;
  if (threadIdx.x == 0) combine_moments(p.moments, p.n_parts, p.part_stride, p.eps, &s_mean, &s_denom);
  __syncthreads();
  const float mean = s_mean, denom = s_denom;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = p.idx ? xa::sample_row(p.idx[i], p.n_steps, p.n_envs) : i;
    p.out[i] = __fdiv_rn(__fsub_rn(__fsub_rn(p.returns[row], p.old_values[row]), mean), denom);
  }
}

// p * x with tfp's multiply_no_nan rule (Categorical.entropy): a probability of exactly 0 -- a saturated softmax-output
// model, Categorical(probs=) -- contributes 0, not 0 * (-inf) = NaN
__device__ __forceinline__ float plogp(float p, float x) { return p > 0.0f ? p * x : 0.0f; }

struct LossWorkspace {  // layout of xa_loss_args.workspace
  unsigned int ticket;
  unsigned int pad[3];
  double partials[1];  // [n_blocks][3]
};

template <bool kPpo, int kActorKind, int kRegActions /*0 = loop over global memory*/>
__global__ void __launch_bounds__(kLossThreads) loss_kernel(const xa_loss_args a) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  __shared__ double scratch[kLossWarps];
  __shared__ float s_mean, s_denom;
  __shared__ bool s_last;

  const int64_t n = a.n;
  const int A = a.n_actions;
  const float inv_n = 1.0f / static_cast<float>(n);

  float adv_mean = 0.0f, adv_denom = 1.0f;
  if (kPpo && a.advantages == nullptr) {
    // (adv - mean) / (std + eps), population std (ppo/agent.py:181-183)
    if (threadIdx.x == 0) combine_moments(a.moments, a.n_moment_parts, a.moment_part_stride, a.adv_eps, &s_mean, &s_denom);
    __syncthreads();
    adv_mean = s_mean;
    adv_denom = s_denom;
  }

  double sum_pg = 0.0, sum_vl = 0.0, sum_ent = 0.0;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kLossThreads + threadIdx.x;
  if (i < n) {
    const int64_t row = a.idx ? xa::sample_row(a.idx[i], a.n_steps, a.n_envs) : i;
    const float ret = a.returns[row];
    const float v_old = a.old_values[row];
    const float v = a.values[i];
    const float* actor = a.actor_out + i * A;
    float* d_actor = a.d_actor ? a.d_actor + i * A : nullptr;

    float adv;
    if (kPpo) {
      adv = a.advantages ? a.advantages[i] : __fdiv_rn(__fsub_rn(__fsub_rn(ret, v_old), adv_mean), adv_denom);
      if (a.advantages_out) a.advantages_out[i] = adv;
    } else {
      adv = __fsub_rn(ret, v_old);  // a2c/agent.py:202
    }

    float logp, ent;
    // ---- distribution: log_prob(action), entropy, and what is needed for d/d actor_out ----------
    float lsm[kRegActions > 0 ? kRegActions : 1];
    float lse = 0.0f, zmax = 0.0f;
    int act = 0;
    if (kActorKind == XA_ACTOR_NORMAL) {
      // MultivariateNormalDiag(loc), identity scale (a2c/agent.py:59-60)
      const float* action = a.actions + row * A;
      float ss = 0.0f;
      for (int j = 0; j < A; ++j) {
        const float d = action[j] - actor[j];
        ss += d * d;
      }
      const float half_k_log2pi = 0.5f * static_cast<float>(A) * 1.8378770664093453f;
      logp = -0.5f * ss - half_k_log2pi;
      ent = 0.5f * static_cast<float>(A) + half_k_log2pi;
    } else {
      act = static_cast<int>(a.actions[row]);  // fp32-encoded action id (ppo/agent.py:202-210)
      if (kRegActions > 0) {
#pragma unroll
        for (int j = 0; j < kRegActions; ++j)
          if (j < A) lsm[j] = kActorKind == XA_ACTOR_PROBS ? logf(actor[j]) : actor[j];
        zmax = lsm[0];
#pragma unroll
        for (int j = 1; j < kRegActions; ++j)
          if (j < A) zmax = fmaxf(zmax, lsm[j]);
        float se = 0.0f;
#pragma unroll
        for (int j = 0; j < kRegActions; ++j)
          if (j < A) se += expf(lsm[j] - zmax);
        lse = logf(se);
        ent = 0.0f;
        logp = 0.0f;
#pragma unroll
        for (int j = 0; j < kRegActions; ++j)
          if (j < A) {
            lsm[j] = (lsm[j] - zmax) - lse;
            ent -= plogp(expf(lsm[j]), lsm[j]);
            if (j == act) logp = lsm[j];
          }
      } else {
        zmax = -INFINITY;
        for (int j = 0; j < A; ++j) zmax = fmaxf(zmax, kActorKind == XA_ACTOR_PROBS ? logf(actor[j]) : actor[j]);
        float se = 0.0f;
        for (int j = 0; j < A; ++j) se += expf((kActorKind == XA_ACTOR_PROBS ? logf(actor[j]) : actor[j]) - zmax);
        lse = logf(se);
        ent = 0.0f;
        logp = 0.0f;
        for (int j = 0; j < A; ++j) {
          const float l = ((kActorKind == XA_ACTOR_PROBS ? logf(actor[j]) : actor[j]) - zmax) - lse;
          ent -= plogp(expf(l), l);
          if (j == act) logp = l;
        }
      }
    }

    // ---- policy term ---------------------------------------------------------------------------
    float pg, g_logp;  // g_logp = d(sum of pg_i)/d logp_i
    if (kPpo) {
      const float ratio = expf(__fsub_rn(logp, a.old_log_probs[row]));                      // ppo/agent.py:123
      const float s1 = -adv * ratio;                                                      // :124
      const float s2 = -adv * fminf(fmaxf(ratio, 1.0f - a.clip), 1.0f + a.clip);          // :125-127
      pg = fmaxf(s1, s2);                                                                 // :128
      g_logp = (s1 >= s2) ? s1 : 0.0f;  // ties -> first argument; s2 > s1 only outside the clip band (zero grad)
    } else {
      pg = -adv * logp;  // a2c/agent.py:208
      g_logp = -adv;
    }

    // ---- value term ----------------------------------------------------------------------------
    float vl, g_v;
    const float e1d = __fsub_rn(v, ret);
    if (kPpo) {
      const float dv = __fsub_rn(v, v_old);
      const float vc = __fadd_rn(v_old, fminf(fmaxf(dv, -a.clip), a.clip));               // :117-119
      const float e2d = __fsub_rn(vc, ret);
      const float e1 = e1d * e1d, e2 = e2d * e2d;                                         // :120-121
      vl = fmaxf(e1, e2);                                                                 // :122
      const bool inside = fabsf(dv) <= a.clip;
      g_v = (e1 >= e2) ? 2.0f * e1d : (inside ? 2.0f * e2d : 0.0f);
      g_v *= 0.5f * a.vf_coef * inv_n;
    } else {
      vl = e1d * e1d;  // a2c/agent.py:209
      g_v = 2.0f * a.vf_coef * e1d * inv_n;
    }
    if (a.d_values) a.d_values[i] = g_v;

    // ---- d loss / d actor_out ------------------------------------------------------------------
    if (d_actor) {
      if (kActorKind == XA_ACTOR_NORMAL) {
        const float* action = a.actions + row * A;
        for (int j = 0; j < A; ++j) d_actor[j] = g_logp * (action[j] - actor[j]) * inv_n;  // entropy is constant
      } else if (kRegActions > 0) {
#pragma unroll
        for (int j = 0; j < kRegActions; ++j)
          if (j < A) {
            const float pj = expf(lsm[j]);
            // g*(onehot - p) + c_e * p * (lsm + H)        
            float g = (g_logp * ((j == act ? 1.0f : 0.0f) - pj) + a.ent_coef * plogp(pj, lsm[j] + ent)) * inv_n;
            if (kActorKind == XA_ACTOR_PROBS) g = actor[j] > 0.0f ? __fdiv_rn(g, actor[j]) : 0.0f;  // chain through logits = log(probs)
            d_actor[j] = g;
          }
      } else {
        for (int j = 0; j < A; ++j) {
          const float l = ((kActorKind == XA_ACTOR_PROBS ? logf(actor[j]) : actor[j]) - zmax) - lse;
          const float pj = expf(l);
          float g = (g_logp * ((j == act ? 1.0f : 0.0f) - pj) + a.ent_coef * plogp(pj, l + ent)) * inv_n;
          if (kActorKind == XA_ACTOR_PROBS) g = actor[j] > 0.0f ? __fdiv_rn(g, actor[j]) : 0.0f;
          d_actor[j] = g;
        }
      }
    }
    sum_pg = pg;
    sum_vl = vl;
    sum_ent = ent;
  }

  // ---- block sums -> partials -> last block finishes -----------------------------------------------
  LossWorkspace* ws = static_cast<LossWorkspace*>(a.workspace);
  const double b_pg = xa::block_sum<kLossWarps>(sum_pg, scratch);
  const double b_vl = xa::block_sum<kLossWarps>(sum_vl, scratch);
  const double b_ent = xa::block_sum<kLossWarps>(sum_ent, scratch);
  if (threadIdx.x == 0) {
    double* mine = ws->partials + static_cast<int64_t>(blockIdx.x) * 3;
    mine[0] = b_pg;
    mine[1] = b_vl;
    mine[2] = b_ent;
    __threadfence();
    s_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double t_pg = 0.0, t_vl = 0.0, t_ent = 0.0;
  for (unsigned k = threadIdx.x; k < gridDim.x; k += kLossThreads) {
    const volatile double* q = ws->partials + static_cast<int64_t>(k) * 3;
    t_pg += q[0];
    t_vl += q[1];
    t_ent += q[2];
  }
  t_pg = xa::block_sum<kLossWarps>(t_pg, scratch);
  t_vl = xa::block_sum<kLossWarps>(t_vl, scratch);
  t_ent = xa::block_sum<kLossWarps>(t_ent, scratch);
  if (threadIdx.x == 0) {
    const double dn = static_cast<double>(n);
    const float pg = static_cast<float>(t_pg / dn);
    const float ent = static_cast<float>(t_ent / dn);
    const float vl = kPpo ? 0.5f * static_cast<float>(t_vl / dn) : static_cast<float>(t_vl / dn);
    // loss = pg - entropy*c_e + value_loss*c_v   (ppo/agent.py:129-133, a2c/agent.py:210-214)
    a.out_scalars[0] = __fadd_rn(__fsub_rn(pg, __fmul_rn(ent, a.ent_coef)), __fmul_rn(vl, a.vf_coef));
    a.out_scalars[1] = pg;
    a.out_scalars[2] = vl;
    a.out_scalars[3] = ent;
    ws->ticket = 0;  // leave the workspace ready for the next launch
  }
}

int64_t loss_blocks(int64_t n) { return (n + kLossThreads - 1) / kLossThreads; }

template <bool kPpo, int kActorKind>
void dispatch_width(const xa_loss_args& a, unsigned grid, cudaStream_t stream) {
  if (kActorKind != XA_ACTOR_NORMAL && a.n_actions <= kMaxRegActions)
    xa::launch_chained(xa::kChainElementwise, loss_kernel<kPpo, kActorKind, kMaxRegActions>, dim3(grid), dim3(kLossThreads), 0, stream, a);
  else
    xa::launch_chained(xa::kChainElementwise, loss_kernel<kPpo, kActorKind, 0>, dim3(grid), dim3(kLossThreads), 0, stream, a);
}

template <bool kPpo>
int run_loss(const xa_loss_args* args, cudaStream_t stream, const char* what) {
  XA_REQUIRE(args != nullptr, XA_EINVAL, "%s: null args", what);
  const xa_loss_args& a = *args;
  XA_REQUIRE(a.n > 0 && a.n_actions > 0, XA_EINVAL, "%s: n=%lld n_actions=%d must be positive", what, static_cast<long long>(a.n),
             a.n_actions);
  XA_REQUIRE(a.n <= (int64_t(1) << 31), XA_EOVERFLOW, "%s: n too large", what);
  XA_REQUIRE(a.actor_out && a.values && a.actions && a.old_values && a.returns && a.out_scalars, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(a.actor_kind >= XA_ACTOR_LOGITS && a.actor_kind <= XA_ACTOR_NORMAL, XA_EINVAL, "%s: unknown actor_kind %d", what, a.actor_kind);
  XA_REQUIRE(a.n_steps >= 0 && a.n_envs >= 0, XA_EINVAL, "%s: negative n_steps/n_envs", what);
  if (kPpo) {
    XA_REQUIRE(a.old_log_probs != nullptr, XA_EINVAL, "%s: null old_log_probs", what);
    XA_REQUIRE(a.advantages != nullptr || (a.moments != nullptr && a.n_moment_parts > 0), XA_EINVAL,
               "%s: need advantages or moments", what);
    XA_REQUIRE(a.clip >= 0.0f, XA_EINVAL, "%s: negative clip", what);
  }
  XA_REQUIRE(a.workspace != nullptr && a.workspace_bytes >= xa_loss_workspace_bytes(a.n), XA_ENOSPACE,
             "%s: workspace of %lld bytes needed, got %lld", what, static_cast<long long>(xa_loss_workspace_bytes(a.n)),
             static_cast<long long>(a.workspace_bytes));
  XA_REQUIRE(xa::aligned(a.workspace, 16), XA_EALIGN, "%s: workspace must be 16-byte aligned", what);
  const unsigned grid = static_cast<unsigned>(loss_blocks(a.n));
  switch (a.actor_kind) {
    case XA_ACTOR_LOGITS: dispatch_width<kPpo, XA_ACTOR_LOGITS>(a, grid, stream); break;
    case XA_ACTOR_PROBS: dispatch_width<kPpo, XA_ACTOR_PROBS>(a, grid, stream); break;
    default: dispatch_width<kPpo, XA_ACTOR_NORMAL>(a, grid, stream); break;
  }
  return xa::check_launch(what);
}

}  // namespace

extern "C" {

int xa_adv_moments_f32(const float* returns, const float* old_values, const int32_t* idx, const int64_t* mb_offsets,
                       int n_minibatches, int n_steps, int n_envs, double* moments, xa_stream_t stream) {
  XA_REQUIRE(n_minibatches > 0 && n_minibatches <= 65535, XA_EINVAL, "xa_adv_moments_f32: n_minibatches=%d", n_minibatches);
  XA_REQUIRE(returns && old_values && mb_offsets && moments, XA_EINVAL, "xa_adv_moments_f32: null pointer");
  XA_REQUIRE(xa::aligned(moments, 8), XA_EALIGN, "xa_adv_moments_f32: moments must be 8-byte aligned");
  XA_REQUIRE(n_steps >= 0 && n_envs >= 0, XA_EINVAL, "xa_adv_moments_f32: negative n_steps/n_envs");
  for (int m = 0; m < n_minibatches; ++m)
    XA_REQUIRE(mb_offsets[m] <= mb_offsets[m + 1], XA_EINVAL, "xa_adv_moments_f32: mb_offsets not ascending at %d", m);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int first = 0; first < n_minibatches; first += kMaxMomentBatches) {
    const int count = n_minibatches - first < kMaxMomentBatches ? n_minibatches - first : kMaxMomentBatches;
    MomentOffsets off{};
    for (int m = 0; m <= count; ++m) off.v[m] = mb_offsets[first + m];
    MomentParams p{returns, old_values, idx, n_steps, n_envs, moments + static_cast<int64_t>(first) * XA_MOMENT_STRIDE};
    xa::launch_chained(xa::kChainElementwise, adv_moments_kernel, dim3(count), dim3(kMomentThreads), 0, s, p, off);
    if (int rc = xa::check_launch("xa_adv_moments_f32")) return rc;
  }
  return XA_OK;
}

int xa_normalize_adv_f32(const float* returns, const float* old_values, const int32_t* idx, int64_t n, int n_steps, int n_envs,
                         const double* moments, int n_moment_parts, int64_t moment_part_stride, double adv_eps,
                         float* advantages, xa_stream_t stream) {
  XA_REQUIRE(n >= 0 && n <= (int64_t(1) << 31), XA_EINVAL, "xa_normalize_adv_f32: n=%lld", static_cast<long long>(n));
  if (n == 0) return XA_OK;
  XA_REQUIRE(returns && old_values && moments && advantages, XA_EINVAL, "xa_normalize_adv_f32: null pointer");
  XA_REQUIRE(n_moment_parts > 0, XA_EINVAL, "xa_normalize_adv_f32: n_moment_parts=%d", n_moment_parts);
  XA_REQUIRE(n_steps >= 0 && n_envs >= 0, XA_EINVAL, "xa_normalize_adv_f32: negative n_steps/n_envs");
  NormParams p{returns, old_values, idx, n_steps, n_envs, moments, n_moment_parts, moment_part_stride,
               static_cast<float>(adv_eps), advantages, n};
  const int64_t want = (n + 255) / 256;
  normalize_adv_kernel<<<static_cast<unsigned>(want < 1184 ? want : 1184), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return xa::check_launch("xa_normalize_adv_f32");
}

int64_t xa_loss_workspace_bytes(int64_t n) {
  if (n < 0) n = 0;
  return 16 + loss_blocks(n) * 3 * static_cast<int64_t>(sizeof(double));
}

int xa_ppo_loss_f32(const xa_loss_args* args, xa_stream_t stream) {
  return run_loss<true>(args, static_cast<cudaStream_t>(stream), "xa_ppo_loss_f32");
}

int xa_a2c_loss_f32(const xa_loss_args* args, xa_stream_t stream) {
  return run_loss<false>(args, static_cast<cudaStream_t>(stream), "xa_a2c_loss_f32");
}

}  // extern "C"
