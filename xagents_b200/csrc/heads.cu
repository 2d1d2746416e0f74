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
constexpr int kBwdRowsPerCta = 64;
constexpr int kChunkThreads = kHidden / 8;   // 64 threads cover one row of h with 16-byte loads

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}

// One warp per frame: lane l holds columns [8l, 8l+8) and [256 + 8l, ...) of h; the stacked head matrix sits in shared
// memory as fp32 (16 KB, conflict-free 32-byte reads); R dot products are reduced with a butterfly.
__global__ void __launch_bounds__(kFwdThreads) heads_forward_kernel(const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ wh,
                                                                  const float* __restrict__ bh, float* __restrict__ actor,
                                                                  float* __restrict__ critic, int batch, int n_actions) {
  __shared__ float s_w[kRows][kHidden];
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
  for (int row = blockIdx.x * (kFwdThreads / 32) + warp; row < batch; row += warps) {
    const uint4* hr = reinterpret_cast<const uint4*>(h + static_cast<int64_t>(row) * kHidden);
    const uint4 v0 = __ldg(hr + lane), v1 = __ldg(hr + 32 + lane);
    float x[16];
    unpack8(v0, *reinterpret_cast<float(*)[8]>(x));
    unpack8(v1, *reinterpret_cast<float(*)[8]>(x + 8));
    float acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const float4* w0 = reinterpret_cast<const float4*>(&s_w[r][8 * lane]);
      const float4* w1 = reinterpret_cast<const float4*>(&s_w[r][256 + 8 * lane]);
      const float4 a = w0[0], b = w0[1], c = w1[0], d = w1[1];
      float s = x[0] * a.x;
      s = fmaf(x[1], a.y, s), s = fmaf(x[2], a.z, s), s = fmaf(x[3], a.w, s);
      s = fmaf(x[4], b.x, s), s = fmaf(x[5], b.y, s), s = fmaf(x[6], b.z, s), s = fmaf(x[7], b.w, s);
      s = fmaf(x[8], c.x, s), s = fmaf(x[9], c.y, s), s = fmaf(x[10], c.z, s), s = fmaf(x[11], c.w, s);
      s = fmaf(x[12], d.x, s), s = fmaf(x[13], d.y, s), s = fmaf(x[14], d.z, s), s = fmaf(x[15], d.w, s);
      acc[r] = s;
    }
    float mine = 0.0f;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      float s = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == r) mine = s;
    }
    if (lane < n_actions)
      actor[static_cast<int64_t>(row) * n_actions + lane] = mine + my_bias;
    else if (lane == n_actions)
      critic[row] = mine + my_bias;
  }
}

// 256 threads = 4 row groups x 64 column chunks (8 columns each); a CTA owns 64 consecutive frames, row group g the frames
// g, g+4, ...  Per frame and thread: 8 x 8 FMAs for dh, 8 x 8 for dW_heads; the four row groups are then added in order
// through shared memory and the CTA writes one partial block [kRows + 2, 512]:
//   rows 0..7  dW_heads        row 8  db_fc (column sums of the bf16-rounded dh, what the FC weight gradient also sees)
//   row 9      db_heads in its first 8 entries
__global__ void __launch_bounds__(kBwdThreads) heads_backward_kernel(const float* __restrict__ d_actor, const float* __restrict__ d_critic,
                                                                   const __nv_bfloat16* __restrict__ h, const __nv_bfloat16* __restrict__ wh,
                                                                   __nv_bfloat16* __restrict__ dh, float* __restrict__ partial, int batch,
                                                                   int n_actions) {
  __shared__ float s_d[kBwdRowsPerCta][kRows];
  __shared__ float s_red[kRows + 1][8][kChunkThreads];   // one row group's accumulators at a time (18 KB)
  const int chunk = threadIdx.x % kChunkThreads, grp = threadIdx.x / kChunkThreads;
  const int row0 = blockIdx.x * kBwdRowsPerCta;
  const int rows = min(kBwdRowsPerCta, batch - row0);
  for (int i = threadIdx.x; i < kBwdRowsPerCta * kRows; i += kBwdThreads) {
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
  float w[kRows][8];
#pragma unroll
  for (int r = 0; r < kRows; ++r) unpack8(__ldg(reinterpret_cast<const uint4*>(wh + r * kHidden) + chunk), w[r]);
  float dw[kRows][8], dbf[8];
#pragma unroll
  for (int r = 0; r < kRows; ++r)
#pragma unroll
    for (int e = 0; e < 8; ++e) dw[r][e] = 0.0f;
#pragma unroll
  for (int e = 0; e < 8; ++e) dbf[e] = 0.0f;
  __syncthreads();

  constexpr int kPerGroup = kBwdRowsPerCta / 4;   // 16 frames per row group, loads issued four at a time
  for (int i0 = 0; i0 < kPerGroup; i0 += 4) {
    uint4 hv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = (i0 + u) * 4 + grp;
      hv[u] = r < rows ? __ldg(reinterpret_cast<const uint4*>(h + static_cast<int64_t>(row0 + r) * kHidden) + chunk) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = (i0 + u) * 4 + grp;
      if (r >= rows) continue;
      float x[8], d[kRows];
      unpack8(hv[u], x);
#pragma unroll
      for (int q = 0; q < kRows; ++q) d[q] = s_d[r][q];
      __nv_bfloat162 out[4];
#pragma unroll
      for (int e = 0; e < 8; e += 2) {
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int q = 0; q < kRows; ++q) s0 = fmaf(d[q], w[q][e], s0), s1 = fmaf(d[q], w[q][e + 1], s1);
        s0 = x[e] > 0.0f ? s0 : 0.0f;          // ReLU derivative of the FC layer
        s1 = x[e + 1] > 0.0f ? s1 : 0.0f;
        out[e / 2] = __floats2bfloat162_rn(s0, s1);
        const float2 back = __bfloat1622float2(out[e / 2]);
        dbf[e] += back.x, dbf[e + 1] += back.y;
      }
      reinterpret_cast<uint4*>(dh + static_cast<int64_t>(row0 + r) * kHidden)[chunk] = *reinterpret_cast<uint4*>(out);
#pragma unroll
      for (int q = 0; q < kRows; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) dw[q][e] = fmaf(d[q], x[e], dw[q][e]);
    }
  }
  // row groups 1..3 are added onto group 0 in order (fixed association: the result does not depend on timing)
  for (int g = 1; g < 4; ++g) {
    __syncthreads();
    if (grp == g) {
#pragma unroll
      for (int q = 0; q < kRows; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) s_red[q][e][chunk] = dw[q][e];
#pragma unroll
      for (int e = 0; e < 8; ++e) s_red[kRows][e][chunk] = dbf[e];
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
      for (int q = 0; q < kRows; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) dw[q][e] += s_red[q][e][chunk];
#pragma unroll
      for (int e = 0; e < 8; ++e) dbf[e] += s_red[kRows][e][chunk];
    }
  }
  float* mine = partial + static_cast<int64_t>(blockIdx.x) * (kRows + 2) * kHidden;
  if (grp == 0) {
#pragma unroll
    for (int q = 0; q < kRows; ++q) {
      float4* dst = reinterpret_cast<float4*>(mine + q * kHidden + 8 * chunk);
      dst[0] = make_float4(dw[q][0], dw[q][1], dw[q][2], dw[q][3]);
      dst[1] = make_float4(dw[q][4], dw[q][5], dw[q][6], dw[q][7]);
    }
    float4* dst = reinterpret_cast<float4*>(mine + kRows * kHidden + 8 * chunk);
    dst[0] = make_float4(dbf[0], dbf[1], dbf[2], dbf[3]);
    dst[1] = make_float4(dbf[4], dbf[5], dbf[6], dbf[7]);
  } else if (grp == 1 && chunk < kRows) {   // db_heads[c] = sum of d_out[:, c] over this CTA's frames, in frame order
    float s = 0.0f;
    for (int r = 0; r < rows; ++r) s += s_d[r][chunk];
    mine[(kRows + 1) * kHidden + chunk] = s;
  }
}

}  // namespace

extern "C" {

int xa_heads_backward_blocks(int batch) { return batch > 0 ? (batch + kBwdRowsPerCta - 1) / kBwdRowsPerCta : 0; }

int xa_heads_forward_bf16(const void* h, const void* wh, const float* bh, float* actor, float* critic, int batch, int hidden, int n_actions,
                          xa_stream_t stream) {
  const char* what = "xa_heads_forward_bf16";
  XA_REQUIRE(h && wh && bh && actor && critic, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(batch > 0 && hidden == kHidden && n_actions > 0 && n_actions + 1 <= kRows, XA_EINVAL,
             "%s: batch=%d hidden=%d (must be %d) n_actions=%d (at most %d)", what, batch, hidden, kHidden, n_actions, kRows - 1);
  XA_REQUIRE(xa::aligned(h, 16) && xa::aligned(wh, 16), XA_EALIGN, "%s: h and wh must be 16-byte aligned", what);
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int want = (batch + kFwdThreads / 32 - 1) / (kFwdThreads / 32);
  const int grid = want < 8 * sms ? want : 8 * sms;  // one frame per warp up to a full wave of 8 CTAs per SM: the kernel is latency-bound
  heads_forward_kernel<<<grid, kFwdThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(h),
                                                                                   static_cast<const __nv_bfloat16*>(wh), bh, actor, critic, batch,
                                                                                   n_actions);
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
  heads_backward_kernel<<<grid, kBwdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      d_actor, d_critic, static_cast<const __nv_bfloat16*>(h), static_cast<const __nv_bfloat16*>(wh), static_cast<__nv_bfloat16*>(dh), partial,
      batch, n_actions);
  return xa::check_launch(what);
}

}  // extern "C"
