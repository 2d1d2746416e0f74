// The actor / critic heads of the policy-value network (two Dense layers on the shared 512-wide trunk output,
// ppo/models/cnn-actor-critic.cfg:30-42 as built by ModelReader, xagents/utils/common.py:239-258), forward and backward,
// as two bandwidth-bound CUDA-core kernels.
//
// With n_actions + 1 <= 8 output columns the heads are 0.07 GFLOP per 8192-frame minibatch: on the tensor-core GEMM they
// cost one launch each for the forward product, the weight gradient, the data gradient, a cast and two bias
// reductions, every one of them latency-, not throughput-bound (14-24 us each, 27 launches under 10 us around them).
// Here the forward is ONE pass over h (1 KB per frame) and the backward ONE pass over h that produces, together,
//   dh   [B, 512] bf16   = (d_out W_heads) * (h > 0)       the FC layer's output gradient (ReLU derivative applied)
//   dW_heads [R, 512], db_heads [R], db_fc [512] = column sums of dh      as per-CTA fp32 partials, added in CTA order
// by xa_grad_finalize_f32 (deterministic).  d_out = [d_actor | d_critic] comes straight from the loss kernel in fp32.
#include <cuda_bf16.h>

#include "xa_common.cuh"

namespace {

constexpr int kHidden = 512;
constexpr int kRows = 8;           // stacked head rows: n_actions + 1 <= 8 (the rest are zero)
constexpr int kFwdThreads = 256;
constexpr int kBwdThreads = 256;
constexpr int kBwdGroups = 2;                  // row groups of a backward CTA
constexpr int kBwdMaxRowsPerCta = 64;
constexpr int kChunkThreads = kHidden / 4;     // 128 threads cover one row of h with 8-byte loads

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack4(const uint2& v, float (&f)[4]) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y;
}

// One warp per PAIR of frames: lane l holds columns [128 j + 4 l, 128 j + 4 l + 4), j = 0..3, of h (8-byte loads, a warp
// reads 256 contiguous bytes); the stacked head matrix sits in shared memory as fp32 and is read as 16-byte vectors at
// stride 16 bytes across the lanes (conflict-free: the first version's 32-byte lane stride made every read a 2-way
// conflict, and with 32 vector reads per frame the kernel was bound by shared-memory wavefronts -- ncu: 2.7 M of them, 10 us),
// each weight vector feeding both frames; R dot products are reduced with a butterfly.
// kPartial: h does not exist yet -- the trunk's last Dense layer left `splits` fp32 partial products [splits, batch, 512]
// (xa_gemm_bf16_tn_partial); this kernel adds them in split order, adds that layer's bias, applies its ReLU and rounds to bf16
// -- the arithmetic of the GEMM's own reduction pass, bit for bit -- writes h (the backward pass reads it) and goes on as above:
// one launch less per rollout step.
struct HeadsPartial {
  const float* partial;   // [splits, batch, 512]
  const float* bias;      // [512] bias of the layer that produces h (may be NULL)
  __nv_bfloat16* h_out;   // [batch, 512] (may be NULL)
  int splits;
};

template <bool kPartial>
__device__ __forceinline__ void load_h_row(const __nv_bfloat16* __restrict__ h, const HeadsPartial& hp, int64_t row, int64_t batch, int lane,
                                           float (&x)[4][4]) {
  if (!kPartial) {
    const uint2* hr = reinterpret_cast<const uint2*>(h + row * kHidden);
    uint2 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldg(hr + 32 * j + lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) unpack4(v[j], x[j]);
  } else {
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const float4* p0 = reinterpret_cast<const float4*>(hp.partial + row * kHidden);
    const int64_t stride4 = batch * (kHidden / 4);
    for (int s = 0; s < hp.splits; ++s) {   // split order, as splitk_reduce_kernel adds them
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = __ldg(p0 + s * stride4 + 32 * j + lane);
        acc[j].x += v.x, acc[j].y += v.y, acc[j].z += v.z, acc[j].w += v.w;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 b = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      if (hp.bias != nullptr) b = __ldg(reinterpret_cast<const float4*>(hp.bias) + 32 * j + lane);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(fmaxf(acc[j].x + b.x, 0.0f), fmaxf(acc[j].y + b.y, 0.0f));
      const __nv_bfloat162 hi = __floats2bfloat162_rn(fmaxf(acc[j].z + b.z, 0.0f), fmaxf(acc[j].w + b.w, 0.0f));
      uint2 packed;
      packed.x = *reinterpret_cast<const uint32_t*>(&lo), packed.y = *reinterpret_cast<const uint32_t*>(&hi);
      if (hp.h_out != nullptr) reinterpret_cast<uint2*>(hp.h_out + row * kHidden)[32 * j + lane] = packed;
      unpack4(packed, x[j]);
    }
  }
}

template <bool kPartial>
__global__ void __launch_bounds__(kFwdThreads) heads_forward_kernel(const __nv_bfloat16* __restrict__ h, const HeadsPartial hp,
                                                                  const __nv_bfloat16* __restrict__ wh, const float* __restrict__ bh,
                                                                  float* __restrict__ actor, float* __restrict__ critic, int batch, int n_actions) {
  __shared__ __align__(16) float s_w[kRows][kHidden];
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  for (int i = threadIdx.x; i < kRows * kHidden / 8; i += kFwdThreads) {   // 512 16-byte pieces, two per thread, both in flight
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(wh) + i), f);
    float4* dst = reinterpret_cast<float4*>(&s_w[0][0] + 8 * i);
    dst[0] = make_float4(f[0], f[1], f[2], f[3]);
    dst[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int warps = gridDim.x * (kFwdThreads / 32);
  const float my_bias = lane < kRows && lane <= n_actions ? bh[lane] : 0.0f;
  for (int row = 2 * (blockIdx.x * (kFwdThreads / 32) + warp); row < batch; row += 2 * warps) {
    const bool two = row + 1 < batch;
    float x0[4][4], x1[4][4];
    load_h_row<kPartial>(h, hp, row, batch, lane, x0);
    load_h_row<kPartial>(h, hp, row + (two ? 1 : 0), batch, lane, x1);
    float acc0[kRows], acc1[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // columns in ascending order within the lane, as before
        const float4 w = *reinterpret_cast<const float4*>(&s_w[r][128 * j + 4 * lane]);
        s0 = fmaf(x0[j][0], w.x, s0), s0 = fmaf(x0[j][1], w.y, s0), s0 = fmaf(x0[j][2], w.z, s0), s0 = fmaf(x0[j][3], w.w, s0);
        s1 = fmaf(x1[j][0], w.x, s1), s1 = fmaf(x1[j][1], w.y, s1), s1 = fmaf(x1[j][2], w.z, s1), s1 = fmaf(x1[j][3], w.w, s1);
      }
      acc0[r] = s0, acc1[r] = s1;
    }
    float mine0 = 0.0f, mine1 = 0.0f;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float s0 = acc0[r], s1 = acc1[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, o), s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      if (lane == r) mine0 = s0, mine1 = s1;
    }
    if (lane < n_actions) {
      actor[static_cast<int64_t>(row) * n_actions + lane] = mine0 + my_bias;
      if (two) actor[static_cast<int64_t>(row + 1) * n_actions + lane] = mine1 + my_bias;
    } else if (lane == n_actions) {
      critic[row] = mine0 + my_bias;
      if (two) critic[row + 1] = mine1 + my_bias;
    }
  }
}

// 256 threads = 2 row groups x 128 column chunks (4 columns each); a CTA owns `rows_per_cta` consecutive frames (chosen so that
// all CTAs are resident at once: two per SM), row group g the frames g, g+2, ...  Per frame and thread: 8 x 4 FMAs for dh,
// 8 x 4 for dW_heads.  (The first version gave a thread 8 columns of 16 frames: 1024 warps in all, 7 per SM, 170 registers --
// bound by instruction issue at that occupancy, 14 us for 16 MB of traffic.)  The two row groups are then added in order
// through shared memory and the CTA writes one partial block [kRows + 2, 512]:
//   rows 0..7  dW_heads        row 8  db_fc (column sums of the bf16-rounded dh, what the FC weight gradient also sees)
//   row 9      db_heads in its first 8 entries
__global__ void __launch_bounds__(kBwdThreads, 2) heads_backward_kernel(const float* __restrict__ d_actor, const float* __restrict__ d_critic,
                                                                      const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ wh,
                                                                      __nv_bfloat16* __restrict__ dh, float* __restrict__ partial, int batch,
                                                                      int n_actions, int rows_per_cta) {
  __shared__ __align__(16) float s_d[kBwdMaxRowsPerCta][kRows];
  __shared__ float s_red[kRows + 1][4][kChunkThreads];   // row group 1's accumulators (18 KB)
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  const int chunk = threadIdx.x % kChunkThreads, grp = threadIdx.x / kChunkThreads;
  const int row0 = blockIdx.x * rows_per_cta;
  const int rows = min(rows_per_cta, batch - row0);
  for (int i = threadIdx.x; i < rows_per_cta * kRows; i += kBwdThreads) {
    const int r = i / kRows, c = i % kRows;
    float v = 0.0f;
    if (r < rows) {
      if (c < n_actions)
        v = d_actor[static_cast<int64_t>(row0 + r) * n_actions + c];
      else if (c == n_actions)
        v = d_critic[row0 + r];
    }
    s_d[r][c] = v;
  }
  float w[kRows][4];
#pragma unroll
  for (int r = 0; r < kRows; ++r) unpack4(__ldg(reinterpret_cast<const uint2*>(wh + r * kHidden) + chunk), w[r]);
  float dw[kRows][4], dbf[4];
#pragma unroll
  for (int r = 0; r < kRows; ++r)
#pragma unroll
    for (int e = 0; e < 4; ++e) dw[r][e] = 0.0f;
#pragma unroll
  for (int e = 0; e < 4; ++e) dbf[e] = 0.0f;
  __syncthreads();

  const int per_group = (rows_per_cta + kBwdGroups - 1) / kBwdGroups;   // loads issued four at a time
  for (int i0 = 0; i0 < per_group; i0 += 4) {
    uint2 hv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = (i0 + u) * kBwdGroups + grp;
      hv[u] = r < rows ? __ldg(reinterpret_cast<const uint2*>(h + static_cast<int64_t>(row0 + r) * kHidden) + chunk) : make_uint2(0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = (i0 + u) * kBwdGroups + grp;
      if (r >= rows) continue;
      float x[4], d[kRows];
      unpack4(hv[u], x);
      const float4 d_lo = *reinterpret_cast<const float4*>(&s_d[r][0]), d_hi = *reinterpret_cast<const float4*>(&s_d[r][4]);
      d[0] = d_lo.x, d[1] = d_lo.y, d[2] = d_lo.z, d[3] = d_lo.w, d[4] = d_hi.x, d[5] = d_hi.y, d[6] = d_hi.z, d[7] = d_hi.w;
      __nv_bfloat162 out[2];
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int q = 0; q < kRows; ++q) s0 = fmaf(d[q], w[q][e], s0), s1 = fmaf(d[q], w[q][e + 1], s1);
        s0 = x[e] > 0.0f ? s0 : 0.0f;          // ReLU derivative of the FC layer
        s1 = x[e + 1] > 0.0f ? s1 : 0.0f;
        out[e / 2] = __floats2bfloat162_rn(s0, s1);
        const float2 back = __bfloat1622float2(out[e / 2]);
        dbf[e] += back.x, dbf[e + 1] += back.y;
      }
      reinterpret_cast<uint2*>(dh + static_cast<int64_t>(row0 + r) * kHidden)[chunk] = *reinterpret_cast<uint2*>(out);
#pragma unroll
      for (int q = 0; q < kRows; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) dw[q][e] = fmaf(d[q], x[e], dw[q][e]);
    }
  }
  // row group 1 is added onto group 0 (fixed association: the result does not depend on timing)
  if (grp == 1) {
#pragma unroll
    for (int q = 0; q < kRows; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) s_red[q][e][chunk] = dw[q][e];
#pragma unroll
    for (int e = 0; e < 4; ++e) s_red[kRows][e][chunk] = dbf[e];
  }
  __syncthreads();
  float* mine = partial + static_cast<int64_t>(blockIdx.x) * (kRows + 2) * kHidden;
  if (grp == 0) {
#pragma unroll
    for (int q = 0; q < kRows; ++q)
      *reinterpret_cast<float4*>(mine + q * kHidden + 4 * chunk) =
          make_float4(dw[q][0] + s_red[q][0][chunk], dw[q][1] + s_red[q][1][chunk], dw[q][2] + s_red[q][2][chunk], dw[q][3] + s_red[q][3][chunk]);
    *reinterpret_cast<float4*>(mine + kRows * kHidden + 4 * chunk) = make_float4(
        dbf[0] + s_red[kRows][0][chunk], dbf[1] + s_red[kRows][1][chunk], dbf[2] + s_red[kRows][2][chunk], dbf[3] + s_red[kRows][3][chunk]);
  } else if (chunk < kRows) {   // db_heads[c] = sum of d_out[:, c] over this CTA's frames, in frame order
    float s = 0.0f;
    for (int r = 0; r < rows; ++r) s += s_d[r][chunk];
    mine[(kRows + 1) * kHidden + chunk] = s;
  }
}

// frames per backward CTA: all CTAs resident at once (two per SM), an even count of at most kBwdMaxRowsPerCta
int bwd_rows_per_cta(int batch) {
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  int rows = (batch + 2 * sms - 1) / (2 * sms);
  rows = (rows + 1) & ~1;
  if (rows < 2) rows = 2;
  if (rows > kBwdMaxRowsPerCta) rows = kBwdMaxRowsPerCta;
  return rows;
}

}  // namespace

extern "C" {

int xa_heads_backward_blocks(int batch) {
  if (batch <= 0) return 0;
  const int rows = bwd_rows_per_cta(batch);
  return (batch + rows - 1) / rows;
}

int xa_heads_forward_bf16(const void* h, const void* wh, const float* bh, float* actor, float* critic, int batch, int hidden, int n_actions,
                          xa_stream_t stream) {
  const char* what = "xa_heads_forward_bf16";
  XA_REQUIRE(h && wh && bh && actor && critic, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(batch > 0 && hidden == kHidden && n_actions > 0 && n_actions + 1 <= kRows, XA_EINVAL,
             "%s: batch=%d hidden=%d (must be %d) n_actions=%d (at most %d)", what, batch, hidden, kHidden, n_actions, kRows - 1);
  XA_REQUIRE(xa::aligned(h, 16) && xa::aligned(wh, 16), XA_EALIGN, "%s: h and wh must be 16-byte aligned", what);
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int want = (batch + 2 * (kFwdThreads / 32) - 1) / (2 * (kFwdThreads / 32));
  const int grid = want < 4 * sms ? want : 4 * sms;  // a pair of frames per warp, up to four CTAs per SM
  xa::launch_chained(batch <= 1024 ? xa::kChainSmall : xa::kChainElementwise, heads_forward_kernel<false>, dim3(grid), dim3(kFwdThreads), 0,
                     static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(h), HeadsPartial{}, static_cast<const __nv_bfloat16*>(wh), bh, actor,
                     critic, batch, n_actions);
  return xa::check_launch(what);
}

int xa_heads_forward_partial_bf16(const float* partial, int splits, const float* bias, void* h_out, const void* wh, const float* bh, float* actor,
                                  float* critic, int batch, int hidden, int n_actions, xa_stream_t stream) {
  const char* what = "xa_heads_forward_partial_bf16";
  XA_REQUIRE(partial && wh && bh && actor && critic, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(batch > 0 && splits >= 1 && hidden == kHidden && n_actions > 0 && n_actions + 1 <= kRows, XA_EINVAL,
             "%s: batch=%d splits=%d hidden=%d (must be %d) n_actions=%d (at most %d)", what, batch, splits, hidden, kHidden, n_actions, kRows - 1);
  XA_REQUIRE(xa::aligned(partial, 16) && xa::aligned(wh, 16) && xa::aligned(bias, 16) && xa::aligned(h_out, 8), XA_EALIGN,
             "%s: partial, bias and wh must be 16-byte aligned", what);
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int want = (batch + 2 * (kFwdThreads / 32) - 1) / (2 * (kFwdThreads / 32));
  const int grid = want < 4 * sms ? want : 4 * sms;
  HeadsPartial hp{partial, bias, static_cast<__nv_bfloat16*>(h_out), splits};
  xa::launch_chained(batch <= 1024 ? xa::kChainSmall : xa::kChainElementwise, heads_forward_kernel<true>, dim3(grid), dim3(kFwdThreads), 0,
                     static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(nullptr), hp, static_cast<const __nv_bfloat16*>(wh), bh, actor,
                     critic, batch, n_actions);
  return xa::check_launch(what);
}

int xa_heads_backward_bf16(const float* d_actor, const float* d_critic, const void* h, const void* wh, void* dh, float* partial,
                           int64_t partial_floats, int batch, int hidden, int n_actions, xa_stream_t stream) {
  const char* what = "xa_heads_backward_bf16";
  XA_REQUIRE(d_actor && d_critic && h && wh && dh && partial, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(batch > 0 && hidden == kHidden && n_actions > 0 && n_actions + 1 <= kRows, XA_EINVAL,
             "%s: batch=%d hidden=%d (must be %d) n_actions=%d (at most %d)", what, batch, hidden, kHidden, n_actions, kRows - 1);
  XA_REQUIRE(xa::aligned(h, 16) && xa::aligned(wh, 16) && xa::aligned(dh, 16) && xa::aligned(partial, 16), XA_EALIGN,
             "%s: 16-byte alignment required", what);
  const int grid = xa_heads_backward_blocks(batch);
  XA_REQUIRE(partial_floats >= static_cast<int64_t>(grid) * (kRows + 2) * kHidden, XA_ENOSPACE, "%s: partial buffer too small", what);
  xa::launch_chained(xa::kChainElementwise, heads_backward_kernel, dim3(grid), dim3(kBwdThreads), 0, static_cast<cudaStream_t>(stream), d_actor, d_critic,
                     static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(wh), static_cast<__nv_bfloat16*>(dh), partial, batch, n_actions,
                     bwd_rows_per_cta(batch));
  return xa::check_launch(what);
}

}  // extern "C"
