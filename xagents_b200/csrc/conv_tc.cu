// Stride-1 NHWC convolution as an implicit GEMM on the tensor cores (forward; the policy/value network's
// convolutional trunk, ModelReader's conv sections: ppo/models/cnn-actor-critic.cfg:1-21, README.md:243-259).
//
// No im2col buffer exists.  Operand A of the GEMM  Y[(b,y,x), n] = sum_k X_col[(b,y,x), k] W[n, k],  k = (kh, kw, c),
// is fetched by TMA straight from the NHWC activation: one K block = 64 channels of ONE kernel tap (kh, kw), i.e.
// the box (64 channels, bx pixels along x, by rows, bb images) of the rank-4 tensor shifted by the tap -- up to
// 128 GEMM rows per tile.  Coordinates that fall outside the image are zero-filled by TMA, which is how padding
// works; with full padding and flipped weights the same kernel is the data-gradient convolution (then `mask`
// applies the previous layer's ReLU derivative in the epilogue).  Strided convolutions are turned into stride-1 ones by space-to-depth (8x8/4 on 84x84x4
// becomes 2x2/1 on 21x21x64; 4x4/2 on 20x20x32 becomes 2x2/1 on 10x10x128), which xa_space_to_depth_u8_bf16 and
// this kernel's epilogue (out_s2d) produce directly, together with bias + ReLU.
// Same machinery as gemm_tc.cu: persistent CTAs, TMA producer warp, single-thread tcgen05.mma issuer, fp32
// accumulators double-buffered in TMEM, four epilogue warps.
#include "tc_common.cuh"

#include <cstdlib>

int xa_conv_flat_try(const void* x, const void* w, const float* bias, void* y, int batch, int height, int width, int channels, int kh, int kw,
                     int n_out, int pad_y, int pad_x, int relu, int out_mode, const void* relu_mask, int mask_is_bits, uint32_t* relu_bits_out,
                     int OH, int OW, int PH, int PW, xa_stream_t stream);

namespace {

using namespace xa_tc;

struct ConvParams {
  __nv_bfloat16* y;
  const float* bias;
  int B, H, W, C, KH, KW, N, OH, OW;
  int bx, by, bb;  // tile box over (x, y, image)
  int relu, out_s2d;  // out_s2d: 0 natural, 1 pack 2x2 pixels into channels, 2 unpack channels into 2x2 pixels
  int PH, PW;         // pixel grid the output is written on (>= the written extent; the rest is left untouched)
  int pad_y, pad_x;
  const __nv_bfloat16* mask;  // optional, output layout: result *= (mask > 0)
};

template <int BN>
struct ConvSmem {
  static constexpr int kStageA = kBlockM * kBlockK * 2;
  static constexpr int kStageB = BN * kBlockK * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (160 * 1024) / kStage > 8 ? 8 : (160 * 1024) / kStage;
  static constexpr int kBytes = kStages * kStage + 1024 + 256 + 2048 /*bias*/;
};

// 2 + 8 warps: TMA producer, MMA issuer, and TWO epilogue groups of four warps (one per TMEM lane quadrant).  Group g
// drains accumulator g, i.e. every other tile, so two tiles' epilogues are in flight per SM: with one warp per
// scheduler the epilogue of these short-K convolutions was a chain of exposed latencies (tcgen05.ld -> mask/bias loads
// -> stores) and bounded the kernel at ~10 % tensor-pipe utilisation.
constexpr int kConvThreads = 64 + 8 * 32;

template <int BN>
__global__ void __launch_bounds__(kConvThreads) conv_fwd_kernel(const __grid_constant__ CUtensorMap map_x,
                                                             const __grid_constant__ CUtensorMap map_w, const ConvParams p) {
  using S = ConvSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStage);
  uint64_t* empty = full + S::kStages;
  uint64_t* acc_full = empty + S::kStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + S::kStages * S::kStage + 256);  // [N <= 512]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.N; i += kConvThreads) s_bias[i] = p.bias != nullptr ? p.bias[i] : 0.0f;
  const int kc_blocks = p.C / kBlockK;        // 64-channel K blocks per kernel tap
  const int k_blocks = p.KH * p.KW * kc_blocks;
  const int tiles_y = (p.OH + p.by - 1) / p.by;
  const int tiles_b = (p.B + p.bb - 1) / p.bb;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int n_tiles = tiles_y * tiles_b * tiles_n;
  const int rows = p.bx * p.by * p.bb;        // GEMM rows actually filled per tile (<= 128)
  constexpr uint32_t kAccStride = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemCols = 2 * kAccStride;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < S::kStages; ++s) {
      xa::mbar_init(full + s, 1);
      xa::mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      xa::mbar_init(acc_full + a, 1);
      xa::mbar_init(acc_empty + a, 4);
    }
    xa::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xa::smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // tile -> (n tile, image block, row block); images vary fastest so neighbouring CTAs share the weights in L2
  auto decode = [&](int tile, int& tn, int& tb, int& ty) {
    tb = tile % tiles_b;
    ty = (tile / tiles_b) % tiles_y;
    tn = tile / (tiles_b * tiles_y);
  };

  if (warp == 0) {
    if (elect_one()) {  // ---- TMA producer
      // No division or modulo in the per-K-block path: this single thread feeds the whole pipeline, and three runtime
      // divisions per K block (tap / channel block / stage) cost more cycles than the four MMAs the block feeds.
      uint32_t s = 0, round = 0;
      const uint32_t stage_bytes = static_cast<uint32_t>(rows) * 128u + S::kStageB;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int tn, tb, ty;
        decode(tile, tn, tb, ty);
        const int y0 = ty * p.by - p.pad_y, b0 = tb * p.bb, n0 = tn * BN;
        int kh = 0, kw = 0, kc = 0;
        for (int kb = 0; kb < k_blocks; ++kb) {
          if (round > 0) mbar_wait_wd(empty + s, (round - 1) & 1);
          uint8_t* a_dst = smem + s * S::kStage;
          xa::mbar_expect_tx(full + s, stage_bytes);
          tma_load_4d(a_dst, &map_x, kc * kBlockK, kw - p.pad_x, y0 + kh, b0, full + s);
          tma_load_2d(a_dst + S::kStageA, &map_w, kb * kBlockK, n0, full + s);
          if (++kc == kc_blocks) {
            kc = 0;
            if (++kw == p.KW) kw = 0, ++kh;
          }
          if (++s == S::kStages) s = 0, ++round;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ---- MMA issuer
      constexpr uint32_t idesc = make_idesc(kBlockM, BN);
      uint32_t s = 0, phase = 0, lt = 0;
      const uint64_t desc0 = make_smem_desc(smem);  // descriptor of stage 0; stage s adds s * kStage to the address field
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t acc = lt & 1, use = lt >> 1;
        if (use > 0) {
          mbar_wait_wd(acc_empty + acc, (use - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait_wd(full + s, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = desc0 + static_cast<uint64_t>(s * (S::kStage >> 4));
          const uint64_t db = da + (S::kStageA >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(empty + s);
          if (++s == S::kStages) s = 0, phase ^= 1;
        }
        umma_commit(acc_full + acc);
      }
    }
  } else {
    // ---- epilogue: group `grp` (warps 2-5 / 6-9) owns accumulator `grp` = the CTA's tiles of that parity
    const int quad = warp & 3;         // TMEM lane quadrant this warp may read
    const uint32_t grp = (warp - 2) >> 2;
    // GEMM row r of a tile = pixel (x, y, image) in box order (x fastest): the same for every tile of this thread
    const int r = quad * 32 + lane;
    const int ox = r % p.bx;
    const int ry = (r / p.bx) % p.by;
    const int rb = r / (p.bx * p.by);
    for (uint32_t lt = grp;; lt += 2) {
      const int tile = blockIdx.x + static_cast<int>(lt) * static_cast<int>(gridDim.x);
      if (tile >= n_tiles) break;
      int tn, tb, ty;
      decode(tile, tn, tb, ty);
      const uint32_t acc = grp;
      const int oy = ty * p.by + ry;
      const int ob = tb * p.bb + rb;
      const bool valid = r < rows && oy < p.OH && ob < p.B;
      int64_t out_off = 0, mask_off = 0;
      if (valid) {
        mask_off = ((static_cast<int64_t>(ob) * p.OH + oy) * p.OW + ox) * p.N;  // the mask has the natural compact layout
        if (p.out_s2d == 1)  // [B, OH/2, OW/2, 4N], channel block (oy%2, ox%2): the next layer's space-to-depth input
          out_off = ((static_cast<int64_t>(ob) * (p.OH / 2) + oy / 2) * (p.OW / 2) + ox / 2) * (4 * p.N) + ((oy & 1) * 2 + (ox & 1)) * p.N;
        else if (p.out_s2d == 2)  // [B, PH, PW, N/4]: channel block (dy, dx) of pixel (oy, ox) is pixel (2oy+dy, 2ox+dx)
          out_off = ((static_cast<int64_t>(ob) * p.PH + 2 * oy) * p.PW + 2 * ox) * (p.N / 4);
        else
          out_off = ((static_cast<int64_t>(ob) * p.PH + oy) * p.PW + ox) * p.N;
      }
      // the ReLU-derivative mask does not depend on the accumulator: fetch this row's BN values before waiting for it
      uint4 mraw[BN / 8];
      const bool use_mask = p.mask != nullptr && valid;
      if (use_mask) {
#pragma unroll
        for (int j = 0; j < BN / 8; ++j)
          if (tn * BN + j * 8 < p.N) mraw[j] = __ldg(reinterpret_cast<const uint4*>(p.mask + mask_off + tn * BN) + j);
      }
      mbar_wait_wd(acc_full + acc, (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride + c0, v);
        const int col0 = tn * BN + c0;
        if (valid && col0 < p.N) {
          int64_t off = out_off + col0;
          if (p.out_s2d == 2) {
            const int n4 = p.N / 4, sub = col0 / n4;
            off = out_off + (static_cast<int64_t>(sub >> 1) * p.PW + (sub & 1)) * n4 + (col0 - sub * n4);
          }
          uint4* dst = reinterpret_cast<uint4*>(p.y + off);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h[4];
            const __nv_bfloat162* mk = reinterpret_cast<const __nv_bfloat162*>(&mraw[c0 / 8 + j]);
            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + col0 + 8 * j);
            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + col0 + 8 * j + 4);
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float a = __uint_as_float(v[8 * j + 2 * q]) + bv[2 * q], b = __uint_as_float(v[8 * j + 2 * q + 1]) + bv[2 * q + 1];
              if (p.relu) {
                a = fmaxf(a, 0.0f);
                b = fmaxf(b, 0.0f);
              }
              if (use_mask) {  // ReLU derivative of the layer below
                if (!(__low2float(mk[q]) > 0.0f)) a = 0.0f;
                if (!(__high2float(mk[q]) > 0.0f)) b = 0.0f;
              }
              h[q] = __floats2bfloat162_rn(a, b);
            }
            dst[j] = *reinterpret_cast<uint4*>(h);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xa::smem_u32(acc_empty + acc)) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// uint8 NHWC frames -> bf16, scaled by 1/255 (true division in fp32, base.py:505-506), rearranged block x block
// -> channels: out[b, y/s, x/s, (y%s, x%s, c)].  One thread per output pixel-channel group of 8.
__global__ void __launch_bounds__(256) space_to_depth_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B,
                                                              int H, int W, int C, int s, int scale) {
  const int OH = H / s, OW = W / s, OC = s * s * C;
  const int64_t total = static_cast<int64_t>(B) * OH * OW * OC;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int oc = static_cast<int>(i % OC);
    const int64_t pix = i / OC;
    const int ox = static_cast<int>(pix % OW), oy = static_cast<int>((pix / OW) % OH);
    const int64_t b = pix / (static_cast<int64_t>(OW) * OH);
    const int c = oc % C, dx = (oc / C) % s, dy = oc / (C * s);
    const float v = static_cast<float>(src[((b * H + oy * s + dy) * W + ox * s + dx) * C + c]);
    dst[i] = __float2bfloat16_rn(scale ? __fdiv_rn(v, 255.0f) : v);
  }
}

// Fast path for block*C == 16 (84x84x4 frames, block 4): one thread moves the 16 contiguous input bytes of one
// (output pixel, dy) -- (dx, c) for dx = 0..3 -- into 16 bf16 = 32 contiguous output bytes.  Reads are 16-B
// vectors along input rows, writes are fully coalesced (a pixel's 64 channels = 4 threads x 32 B).
// With `idx` the kernel is also the minibatch gather (PPO.get_mini_batches, ppo/agent.py:149-154, fused with the flatten
// of base.py:559-564 and the cast/255 of base.py:505-506): output image b is frame row(idx[b]) of the time-major rollout.
// Blocks walk over frames; the 16-byte pieces of one frame are spread over the block's threads, all loads of a thread
// are issued before the first conversion (seven in flight), and the index arithmetic is 32-bit and per-frame.  x/255
// is computed as x * (1/255): for all 256 byte values the bf16 result is bit-identical to the rounded true division
// (tests/test_gpu_conv.py checks every value), without sixteen IEEE divisions per thread.
__global__ void __launch_bounds__(256) space_to_depth16_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B,
                                                                int H, int W, int s, int scale, const int32_t* __restrict__ idx,
                                                                int n_steps, int n_envs) {
  const int OH = H / s, OW = W / s;
  const int per_frame = OH * OW * s;  // (pixel, dy) pairs = 16-byte input pieces of one frame
  const int in_pitch = W * (16 / s);  // bytes per input row
  const float mul = scale ? 1.0f / 255.0f : 1.0f;
  constexpr int kUnroll = 7;          // 84x84x4: 1764 pieces = 6.9 x 256
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const int64_t row = idx != nullptr ? xa::sample_row(__ldg(idx + b), n_steps, n_envs) : b;
    const uint8_t* frame = src + row * (static_cast<int64_t>(H) * in_pitch);
    __nv_bfloat16* out_frame = dst + static_cast<int64_t>(b) * per_frame * 16;
    for (int t0 = threadIdx.x; t0 < per_frame; t0 += kUnroll * 256) {
      uint4 in[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int t = t0 + u * 256;
        if (t < per_frame) {
          const int dy = t % s, pix = t / s;
          const int ox = pix % OW, oy = pix / OW;
          in[u] = __ldg(reinterpret_cast<const uint4*>(frame + (oy * s + dy) * in_pitch + ox * 16));
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int t = t0 + u * 256;
        if (t < per_frame) {
          const uint32_t words[4] = {in[u].x, in[u].y, in[u].z, in[u].w};
          __nv_bfloat162 out[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float a = static_cast<float>((words[k] >> (16 * h)) & 0xFF) * mul;
              const float c = static_cast<float>((words[k] >> (16 * h + 8)) & 0xFF) * mul;
              out[2 * k + h] = __floats2bfloat162_rn(a, c);
            }
          }
          uint4* d = reinterpret_cast<uint4*>(out_frame + static_cast<int64_t>(t) * 16);
          d[0] = *reinterpret_cast<uint4*>(out);
          d[1] = *reinterpret_cast<uint4*>(out + 4);
        }
      }
    }
  }
}

template <int BN>
int launch_conv(const CUtensorMap& mx, const CUtensorMap& mw, const ConvParams& p, cudaStream_t stream, const char* what) {
  auto kernel = conv_fwd_kernel<BN>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<BN>::kBytes);
  if (e != cudaSuccess) {
    xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  const int64_t tiles = static_cast<int64_t>((p.OH + p.by - 1) / p.by) * ((p.B + p.bb - 1) / p.bb) * ((p.N + BN - 1) / BN);
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  kernel<<<static_cast<unsigned>(tiles < sms ? tiles : sms), kConvThreads, ConvSmem<BN>::kBytes, stream>>>(mx, mw, p);
  return xa::check_launch(what);
}

}  // namespace

extern "C" {

int xa_conv2d_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, int batch, int height, int width, int channels,
                        int kh, int kw, int n_out, int pad_y, int pad_x, int relu, int out_s2d, const void* relu_mask,
                        xa_stream_t stream) {
  return xa_conv2d_nhwc_bf16_ex(x, w, bias, y, batch, height, width, channels, kh, kw, n_out, pad_y, pad_x, relu, out_s2d, relu_mask, 0, 0,
                                0, 0, nullptr, 0, stream);
}

int xa_conv2d_nhwc_bf16_ex(const void* x, const void* w, const float* bias, void* y, int batch, int height, int width, int channels,
                           int kh, int kw, int n_out, int pad_y, int pad_x, int relu, int out_mode, const void* relu_mask, int out_h,
                           int out_w, int out_grid_h, int out_grid_w, uint32_t* relu_bits_out, int flags, xa_stream_t stream) {
  const char* what = "xa_conv2d_nhwc_bf16";
  const int out_s2d = out_mode;
  XA_REQUIRE(x && w && y, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(batch > 0 && channels > 0 && kh > 0 && kw > 0 && n_out > 0 && pad_y >= 0 && pad_x >= 0 && pad_y < kh && pad_x < kw,
             XA_EINVAL, "%s: bad shape", what);
  XA_REQUIRE(height + 2 * pad_y >= kh && width + 2 * pad_x >= kw, XA_EINVAL, "%s: kernel larger than the padded image", what);
  XA_REQUIRE(channels % kBlockK == 0, XA_EINVAL, "%s: channels=%d must be a multiple of 64 (one K block per tap)", what, channels);
  XA_REQUIRE(n_out % 32 == 0 && n_out <= 512, XA_EINVAL, "%s: n_out=%d must be a multiple of 32, at most 512", what, n_out);
  XA_REQUIRE(xa::aligned(x, 16) && xa::aligned(w, 16) && xa::aligned(y, 16) && xa::aligned(relu_mask, 16), XA_EALIGN,
             "%s: 16-byte alignment required", what);
  ConvParams p{};
  p.y = static_cast<__nv_bfloat16*>(y);
  p.bias = bias;
  p.B = batch, p.H = height, p.W = width, p.C = channels, p.KH = kh, p.KW = kw, p.N = n_out;
  p.pad_y = pad_y, p.pad_x = pad_x;
  p.OH = height + 2 * pad_y - kh + 1, p.OW = width + 2 * pad_x - kw + 1;
  XA_REQUIRE(out_h >= 0 && out_w >= 0 && out_h <= p.OH && out_w <= p.OW, XA_EINVAL, "%s: out_h/out_w exceed the padded output %dx%d", what,
             p.OH, p.OW);
  if (out_h > 0) p.OH = out_h;  // only the top-left corner of the padded output (a data gradient read from a zero-bordered grid)
  if (out_w > 0) p.OW = out_w;
  XA_REQUIRE(out_mode >= 0 && out_mode <= 2, XA_EINVAL, "%s: out_mode=%d", what, out_mode);
  const int mul = out_mode == 2 ? 2 : 1;
  p.PH = out_grid_h > 0 ? out_grid_h : mul * p.OH, p.PW = out_grid_w > 0 ? out_grid_w : mul * p.OW;
  XA_REQUIRE(p.PH >= mul * p.OH && p.PW >= mul * p.OW && (out_mode != 1 || (out_grid_h == 0 && out_grid_w == 0)), XA_EINVAL,
             "%s: output grid %dx%d smaller than the output", what, p.PH, p.PW);
  XA_REQUIRE(out_mode != 2 || (n_out % 128 == 0), XA_EINVAL, "%s: unpacking needs n_out %% 128 == 0", what);
  p.relu = relu, p.out_s2d = out_s2d;
  p.mask = static_cast<const __nv_bfloat16*>(relu_mask);
  XA_REQUIRE(p.OW <= kBlockM, XA_EINVAL, "%s: output width %d exceeds one tile (128)", what, p.OW);
  XA_REQUIRE(xa::aligned(relu_bits_out, 16), XA_EALIGN, "%s: relu_bits_out must be 16-byte aligned", what);
  XA_REQUIRE(out_s2d != 1 || (p.OH % 2 == 0 && p.OW % 2 == 0 && relu_mask == nullptr), XA_EINVAL,
             "%s: out_s2d needs even output height/width and no mask", what);
  // Preferred: the flat kernel (conv_flat_tc.cu), every input pixel fetched once per tile.  It needs the taps' row shifts
  // to read zeros where they wrap: always true without padding, the caller's promise (zero border) with padding.
  static const bool flat_enabled = !(std::getenv("XA_CONV_FLAT") && std::atoi(std::getenv("XA_CONV_FLAT")) == 0);
  if (flat_enabled && ((pad_y == 0 && pad_x == 0) || (flags & XA_CONV_INPUT_ZERO_BORDER))) {
    const int rc = xa_conv_flat_try(x, w, bias, y, batch, height, width, channels, kh, kw, n_out, pad_y, pad_x, relu, out_mode, relu_mask,
                                    (flags & XA_CONV_MASK_BITS) != 0 && relu_mask != nullptr, relu_bits_out, p.OH, p.OW, p.PH, p.PW, stream);
    if (rc != 1) return rc;  // launched (or failed with an error); 1 = shape not covered, use the per-tap kernel below
  }
  XA_REQUIRE(relu_bits_out == nullptr && !((flags & XA_CONV_MASK_BITS) && relu_mask), XA_EINVAL,
             "%s: ReLU mask bits are implemented by the flat kernel only (unpadded or zero-bordered input, n_out 32 / 64 / 128)", what);
  // tile box: whole output rows (bx = OW), by rows (a divisor of OH), bb images; maximise filled GEMM rows <= 128
  p.bx = p.OW;
  int best = 0;
  for (int by = 1; by <= p.OH; ++by) {
    if (p.OH % by) continue;
    for (int bb = 1; bb <= 16; ++bb) {
      const int rows = p.bx * by * bb;
      if (rows <= kBlockM && rows > best) best = rows, p.by = by, p.bb = bb;
    }
  }
  EncodeTiledFn fn = encode_fn();
  XA_REQUIRE(fn != nullptr, XA_EINVAL, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
  CUtensorMap mx, mw;
  {
    // the NHWC activation as a rank-4 tensor (c, x, y, b); boxes shifted by the kernel tap, zero-filled outside
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(channels), static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(height),
                                static_cast<cuuint64_t>(batch)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(channels) * 2, static_cast<cuuint64_t>(width) * channels * 2,
                                   static_cast<cuuint64_t>(height) * width * channels * 2};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(p.bx), static_cast<cuuint32_t>(p.by),
                               static_cast<cuuint32_t>(p.bb)};
    const cuuint32_t elem[4] = {1, 1, 1, 1};
    const CUresult r = fn(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, elem,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    XA_REQUIRE(r == CUDA_SUCCESS, XA_EINVAL, "%s: cuTensorMapEncodeTiled(activation) failed with %d", what, static_cast<int>(r));
  }
  const int bn = n_out >= 64 ? 64 : 32;
  if (int rc = make_map_2d(&mw, w, n_out, static_cast<int64_t>(kh) * kw * channels, bn, what)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bn == 64 ? launch_conv<64>(mx, mw, p, s, what) : launch_conv<32>(mx, mw, p, s, what);
}

int xa_space_to_depth_u8_bf16(const uint8_t* src, void* dst, int batch, int height, int width, int channels, int block, int scale_255,
                              xa_stream_t stream) {
  XA_REQUIRE(src && dst, XA_EINVAL, "xa_space_to_depth_u8_bf16: null pointer");
  XA_REQUIRE(batch > 0 && block > 0 && height % block == 0 && width % block == 0 && channels > 0, XA_EINVAL,
             "xa_space_to_depth_u8_bf16: %dx%dx%d is not divisible into %dx%d blocks", height, width, channels, block, block);
  const int64_t total = static_cast<int64_t>(batch) * height * width * channels;
  const int64_t want = (total + 255) / 256;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int64_t cap = static_cast<int64_t>(sms) * 32;
  if (block * channels == 16 && xa::aligned(src, 16) && xa::aligned(dst, 16) && (width * channels) % 16 == 0) {
    const int64_t want16 = batch;  // one frame per block iteration
    space_to_depth16_kernel<<<static_cast<unsigned>(want16 < cap ? want16 : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<__nv_bfloat16*>(dst), batch, height, width, block, scale_255, nullptr, 0, 0);
    return xa::check_launch("xa_space_to_depth_u8_bf16");
  }
  space_to_depth_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), batch, height, width, channels, block, scale_255);
  return xa::check_launch("xa_space_to_depth_u8_bf16");
}

int xa_gather_s2d_u8_bf16(const uint8_t* src, const int32_t* idx, void* dst, int64_t n_idx, int64_t n_src_rows, int n_steps, int n_envs,
                          int height, int width, int channels, int block, int scale_255, xa_stream_t stream) {
  const char* what = "xa_gather_s2d_u8_bf16";
  XA_REQUIRE(n_idx >= 0 && n_src_rows > 0 && n_src_rows <= INT32_MAX, XA_EINVAL, "%s: bad sizes", what);
  if (n_idx == 0) return XA_OK;
  XA_REQUIRE(src && idx && dst, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(block * channels == 16 && height % block == 0 && width % block == 0 && (width * channels) % 16 == 0, XA_EINVAL,
             "%s: built for block*channels == 16 (84x84x4 frames, block 4)", what);
  XA_REQUIRE(n_steps >= 0 && n_envs >= 0 && (n_steps == 0 || static_cast<int64_t>(n_steps) * n_envs == n_src_rows), XA_EINVAL,
             "%s: n_steps*n_envs must equal n_src_rows", what);
  XA_REQUIRE(xa::aligned(src, 16) && xa::aligned(dst, 16) && xa::aligned(idx, 4), XA_EALIGN, "%s: alignment", what);
  const int64_t total = n_idx * height * width * channels / 16;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  (void)total;
  const int64_t want = n_idx, cap = static_cast<int64_t>(sms) * 32;  // one frame per block iteration
  space_to_depth16_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), static_cast<int>(n_idx), height, width, block, scale_255, idx, n_steps, n_envs);
  return xa::check_launch(what);
}

}  // extern "C"
