// Stride-1 NHWC convolution on the tensor cores, "flat" form: every input pixel is fetched ONCE per tile.
//
// conv_tc.cu fetches one TMA box per kernel tap, so a KHxKW convolution pulls every input pixel KH*KW times through L2
// -- for the short-K layers of this network (K = 256..576) that L2 -> shared-memory traffic, not the tensor pipe, set
// the pace.  Here the pixels of the INPUT grid are flattened, q = (b*H + y)*W + x, and a tile is 128 consecutive q.  The
// operand rows of tap (kh, kw) are then rows q + (kh - pad_y)*W + (kw - pad_x) of the same [Q, C] matrix, so one
// window of 128 + (KH-1)*W + (KW-1) rows per 64-channel block is loaded per tile and each tap is the SAME shared
// memory tile read through a descriptor whose start address is advanced by whole 128-byte rows (the SWIZZLE_128B
// pattern is a function of the absolute shared-memory address, so a row-shifted window of a swizzled tile is still a
// valid swizzled operand; wgrad_mn_tc.cu relies on the same fact).  The weights of these layers (<= 72 KB) stay
// resident in shared memory for the whole persistent kernel.
// Output row q is pixel (b, y, x) of the input grid; rows with y >= OH or x >= OW are computed and dropped (forward:
// 9-40 % of the rows), a data gradient over a zero-bordered dY grid uses every row.  Where a tap's shift wraps into
// the previous/next grid row, the operand must read zeros: true without padding for all rows that are kept, and with
// padding when the caller guarantees a zero border of pad_x columns / pad_y rows (XA_CONV_INPUT_ZERO_BORDER).
// Pipeline and epilogue as in conv_tc.cu (two epilogue groups, one per TMEM accumulator).
#include "tc_common.cuh"

namespace {

using namespace xa_tc;

constexpr int kFlatThreads = 64 + 8 * 32;

__device__ __forceinline__ uint4 lds_u4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int64_t lds_i64(uint32_t a) {
  int64_t v;
  asm volatile("ld.shared.s64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_i64(uint32_t a, int64_t v) { asm volatile("st.shared.s64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
constexpr int kMaxEntries = 40;

struct FlatParams {
  __nv_bfloat16* y;
  const float* bias;
  const __nv_bfloat16* mask;
  int B, H, W, N, OH, OW, PH, PW;
  int relu, out_mode;
  int n_entries;   // taps x 64-channel blocks
  int kc_blocks;   // 64-channel blocks of the input
  int win_rows;    // rows of the window box (multiple of 8)
  int min_shift;   // row shift of the first tap: -(pad_y*W + pad_x)
  int stages;
  int64_t Q;       // B*H*W
  uint32_t w_bytes, stage_bytes;
  uint32_t a_units[kMaxEntries];  // (byte offset of the entry's A operand inside a stage) >> 4
  uint32_t b_units[kMaxEntries];  // (byte offset of the entry's weight tile) >> 4
};

template <int BN>
__global__ void __launch_bounds__(kFlatThreads) conv_flat_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                 const __grid_constant__ CUtensorMap map_w,
                                                                 const __grid_constant__ FlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + p.w_bytes;  // weights first, then the ring of window stages
  uint8_t* tail = ring + static_cast<size_t>(p.stages) * p.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + 8;
  uint64_t* acc_full = empty + 8;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  uint32_t* s_a = reinterpret_cast<uint32_t*>(tail + 256);  // [kMaxEntries]
  uint32_t* s_b = s_a + kMaxEntries;
  float* s_bias = reinterpret_cast<float*>(tail + 256 + 2 * kMaxEntries * 4);  // [N <= 128]
  uint8_t* epi_stage = tail + 2048;  // 8 warps x (32 rows x (2 BN + 16) bytes + 512)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < p.N; i += kFlatThreads) s_bias[i] = p.bias != nullptr ? p.bias[i] : 0.0f;
  for (int i = threadIdx.x; i < p.n_entries; i += kFlatThreads) s_a[i] = p.a_units[i], s_b[i] = p.b_units[i];
  const int n_tiles = static_cast<int>((p.Q + kBlockM - 1) / kBlockM);
  constexpr uint32_t kAccStride = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemCols = 2 * kAccStride;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      xa::mbar_init(full + s, 1);
      xa::mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      xa::mbar_init(acc_full + a, 1);
      xa::mbar_init(acc_empty + a, 4);
    }
    xa::mbar_init(w_full, 1);
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

  if (warp == 0) {
    if (elect_one()) {  // ---- TMA producer: the weights once, then one window per tile
      xa::mbar_expect_tx(w_full, p.w_bytes);
      for (int e = 0; e < p.n_entries; ++e) tma_load_2d(smem + e * (BN * 128), &map_w, e * kBlockK, 0, w_full);
      int s = 0;
      uint32_t round = 0;
      const uint32_t blk_bytes = static_cast<uint32_t>(p.win_rows) * 128u;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (round > 0) mbar_wait_wd(empty + s, (round - 1) & 1);
        uint8_t* dst = ring + static_cast<size_t>(s) * p.stage_bytes;
        xa::mbar_expect_tx(full + s, p.stage_bytes);
        const int q0 = tile * kBlockM + p.min_shift;  // may be negative: rows before the tensor read as zeros
        for (int c = 0; c < p.kc_blocks; ++c) tma_load_2d(dst + c * blk_bytes, &map_x, c * kBlockK, q0, full + s);
        if (++s == p.stages) s = 0, ++round;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ---- MMA issuer: per tile, taps x channel blocks x 4 K steps into one accumulator
      constexpr uint32_t idesc = make_idesc(kBlockM, BN);
      const uint64_t dw0 = make_smem_desc(smem);
      const uint64_t da0 = make_smem_desc(ring);
      const uint32_t stage_units = p.stage_bytes >> 4;
      int s = 0;
      uint32_t phase = 0, lt = 0;
      mbar_wait_wd(w_full, 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
        const uint32_t acc = lt & 1, use = lt >> 1;
        if (use > 0) {
          mbar_wait_wd(acc_empty + acc, (use - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        mbar_wait_wd(full + s, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t das = da0 + static_cast<uint64_t>(s * stage_units);
#pragma unroll 1
        for (int e = 0; e < p.n_entries; ++e) {
          const uint64_t da = das + s_a[e], db = dw0 + s_b[e];
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (e | k) != 0);
        }
        umma_commit(empty + s);
        umma_commit(acc_full + acc);
        if (++s == p.stages) s = 0, phase ^= 1;
      }
    }
  } else {
    // ---- epilogue: group `grp` (warps 2-5 / 6-9) owns accumulator `grp` = the CTA's tiles of that parity.
    // tcgen05.ld hands every thread one ROW (32 consecutive columns): stored directly, a warp-wide 16-byte access touches
    // 32 different lines, and those load/store wavefronts -- not the tensor pipe, not HBM -- bounded the data-gradient
    // layers (ncu: LSU 80 % busy).  So each warp transposes through its own padded shared-memory tile: phase 1 writes
    // rows as they come out of TMEM (thread = row), phase 2 moves 16-byte pieces with consecutive lanes on consecutive
    // pieces of a row, so that mask loads and output stores are coalesced (4 wavefronts per access instead of 32).
    const int quad = warp & 3;
    const uint32_t grp = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t hw = static_cast<uint32_t>(p.H) * p.W;
    constexpr int kPitch = BN * 2 + 16;   // bytes; +16 keeps 16-byte accesses of 32 rows conflict-free
    constexpr int kPieces = BN / 8;       // 16-byte pieces per row
    uint8_t* my_stage = epi_stage + static_cast<size_t>(warp - 2) * (32 * kPitch + 512);
    // explicit shared-state-space accesses: through the aligned-up base pointer the compiler only sees generic addresses
    const uint32_t my_stage_u32 = xa::smem_u32(my_stage), s_bias_u32 = xa::smem_u32(s_bias);
    const uint32_t row_out_u32 = my_stage_u32 + 32 * kPitch;  // [32] int64 output offset of the row, -1 = dropped
    const uint32_t row_msk_u32 = row_out_u32 + 256;            // [32] int64 mask offset of the row
    // phase-2 coordinates: iteration `it` moves 32 consecutive 16-byte pieces = kRowsPerIt rows, so a thread keeps its
    // piece (column group) for the whole kernel and only its row advances -- all per-iteration address arithmetic is
    // an add of a compile-time constant.
    constexpr int kRowsPerIt = 32 / kPieces;
    const int piece = lane % kPieces, row0 = lane / kPieces;
    int col_delta = piece * 8;  // element offset of my piece inside the output row
    if (p.out_mode == 2) {      // (dy, dx, c) channel blocks land on pixels (2y+dy, 2x+dx): see xa_conv2d_nhwc_bf16_ex
      constexpr int kN4 = BN / 4;
      const int sub = col_delta / kN4;
      col_delta = ((sub >> 1) * p.PW + (sub & 1)) * kN4 + (col_delta - sub * kN4);
    }
    const uint32_t my_piece_u32 = my_stage_u32 + row0 * kPitch + piece * 16;
    for (uint32_t lt = grp;; lt += 2) {
      const int tile = blockIdx.x + static_cast<int>(lt) * static_cast<int>(gridDim.x);
      if (tile >= n_tiles) break;
      const uint32_t acc = grp;
      {  // phase 0: where does my row go?
        const int64_t q = static_cast<int64_t>(tile) * kBlockM + r;
        const uint32_t qq = static_cast<uint32_t>(q < p.Q ? q : 0);
        const int ob = static_cast<int>(qq / hw);
        const uint32_t rem = qq - static_cast<uint32_t>(ob) * hw;
        const int oy = static_cast<int>(rem / p.W), ox = static_cast<int>(rem - (rem / p.W) * p.W);
        const bool valid = q < p.Q && oy < p.OH && ox < p.OW;
        int64_t out_off = -1, mask_off = 0;
        if (valid) {
          mask_off = ((static_cast<int64_t>(ob) * p.OH + oy) * p.OW + ox) * p.N;
          if (p.out_mode == 1)
            out_off = ((static_cast<int64_t>(ob) * (p.OH / 2) + oy / 2) * (p.OW / 2) + ox / 2) * (4 * p.N) + ((oy & 1) * 2 + (ox & 1)) * p.N;
          else if (p.out_mode == 2)
            out_off = ((static_cast<int64_t>(ob) * p.PH + 2 * oy) * p.PW + 2 * ox) * (p.N / 4);
          else
            out_off = ((static_cast<int64_t>(ob) * p.PH + oy) * p.PW + ox) * p.N;
        }
        sts_i64(row_out_u32 + lane * 8, out_off);
        sts_i64(row_msk_u32 + lane * 8, mask_off);
      }
      __syncwarp();
      // the ReLU-derivative mask does not depend on the accumulator: fetch it (coalesced) before waiting
      uint4 mraw[kPieces];
      if (p.mask != nullptr) {
#pragma unroll
        for (int it = 0; it < kPieces; ++it) {
          const int row = it * kRowsPerIt + row0;
          if (lds_i64(row_out_u32 + row * 8) >= 0)
            mraw[it] = __ldg(reinterpret_cast<const uint4*>(p.mask + lds_i64(row_msk_u32 + row * 8)) + piece);
        }
      }
      mbar_wait_wd(acc_full + acc, (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float lo = p.relu ? 0.0f : -INFINITY;  // ReLU as one max per element, no per-element branch on the flag
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {  // phase 1: TMEM -> (+bias, ReLU) -> bf16 rows in shared memory
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride + c0, v);
        const uint32_t dst = my_stage_u32 + lane * kPitch + c0 * 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b0 = lds_f4(s_bias_u32 + (c0 + 8 * j) * 4), b1 = lds_f4(s_bias_u32 + (c0 + 8 * j + 4) * 4);
          __nv_bfloat162 h[4];
          h[0] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, lo), fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, lo));
          h[1] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, lo), fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, lo));
          h[2] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, lo), fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, lo));
          h[3] = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, lo), fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, lo));
          sts_u4(dst + j * 16, *reinterpret_cast<uint4*>(h));
        }
      }
      // the accumulator is drained: hand it back before the global stores
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xa::smem_u32(acc_empty + acc)) : "memory");
#pragma unroll
      for (int it = 0; it < kPieces; ++it) {  // phase 2: coalesced mask application and stores
        const int64_t off0 = lds_i64(row_out_u32 + (it * kRowsPerIt + row0) * 8);
        if (off0 >= 0) {
          uint4 val = lds_u4(my_piece_u32 + it * (kRowsPerIt * kPitch));
          if (p.mask != nullptr) {  // ReLU derivative of the layer below: zero where its activation was <= 0
            const __nv_bfloat162* mk = reinterpret_cast<const __nv_bfloat162*>(&mraw[it]);
            const __nv_bfloat162 zero = __floats2bfloat162_rn(0.0f, 0.0f);
            val.x &= __hgt2_mask(mk[0], zero);
            val.y &= __hgt2_mask(mk[1], zero);
            val.z &= __hgt2_mask(mk[2], zero);
            val.w &= __hgt2_mask(mk[3], zero);
          }
          *reinterpret_cast<uint4*>(p.y + off0 + col_delta) = val;
        }
      }
      __syncwarp();  // the tile's staging rows are free again
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <int BN>
int launch_flat(const CUtensorMap& mx, const CUtensorMap& mw, const FlatParams& p, size_t smem, cudaStream_t stream, const char* what) {
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_flat_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_dev = dev;
  }
  const int64_t tiles = (p.Q + kBlockM - 1) / kBlockM;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  conv_flat_kernel<BN><<<static_cast<unsigned>(tiles < sms ? tiles : sms), kFlatThreads, smem, stream>>>(mx, mw, p);
  return xa::check_launch(what);
}

}  // namespace

// Returns XA_OK after launching, or 1 when the shape does not fit this kernel (the caller falls back to conv_tc.cu's).
int xa_conv_flat_try(const void* x, const void* w, const float* bias, void* y, int batch, int height, int width, int channels, int kh, int kw,
                     int n_out, int pad_y, int pad_x, int relu, int out_mode, const void* relu_mask, int OH, int OW, int PH, int PW,
                     xa_stream_t stream) {
  const char* what = "xa_conv2d_nhwc_bf16";
  if (!(n_out == 32 || n_out == 64 || n_out == 128)) return 1;
  // Output pixels are indexed on the input grid, and a kept output (y < OH, x < OW) must never read below / right of its
  // image -- there the flattened index would run into the next row / image instead of padding.  True for unpadded
  // convolutions and for full padding (pad = k - 1, the data gradients); 'same'-style padding goes to the per-tap kernel.
  if (OH > height || OW > width) return 1;
  if (OH > height + pad_y - kh + 1 || OW > width + pad_x - kw + 1) return 1;
  const int kc_blocks = channels / kBlockK;
  const int n_entries = kh * kw * kc_blocks;
  if (n_entries > kMaxEntries) return 1;
  const int win_rows = ((kBlockM + (kh - 1) * width + (kw - 1) + 7) / 8) * 8;
  if (win_rows > 256) return 1;
  const int64_t Q = static_cast<int64_t>(batch) * height * width;
  if (Q >= (int64_t(1) << 31) - 4096) return 1;
  FlatParams p{};
  p.w_bytes = static_cast<uint32_t>(n_entries) * n_out * 128u;
  p.stage_bytes = static_cast<uint32_t>(kc_blocks) * win_rows * 128u;
  const int64_t epi_bytes = 8 * (32 * (2 * n_out + 16) + 512);  // the epilogue warps' transposing tiles
  const int64_t budget = 227 * 1024 - 1024 /*alignment*/ - 2048 /*barriers, tables, bias*/ - epi_bytes - p.w_bytes;
  int stages = static_cast<int>(budget / p.stage_bytes);
  if (stages < 2) return 1;
  if (stages > 6) stages = 6;
  p.y = static_cast<__nv_bfloat16*>(y), p.bias = bias, p.mask = static_cast<const __nv_bfloat16*>(relu_mask);
  p.B = batch, p.H = height, p.W = width, p.N = n_out, p.OH = OH, p.OW = OW, p.PH = PH, p.PW = PW;
  p.relu = relu, p.out_mode = out_mode;
  p.n_entries = n_entries, p.kc_blocks = kc_blocks, p.win_rows = win_rows, p.stages = stages, p.Q = Q;
  p.min_shift = -(pad_y * width + pad_x);
  int e = 0;
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j)
      for (int c = 0; c < kc_blocks; ++c, ++e) {
        // weights are [N, (kh, kw, c)]: entry e = K block e of the weight matrix; A = window block c, shifted by the tap
        p.a_units[e] = (static_cast<uint32_t>(c) * win_rows * 128u + static_cast<uint32_t>(i * width + j) * 128u) >> 4;
        p.b_units[e] = (static_cast<uint32_t>(e) * n_out * 128u) >> 4;
      }
  CUtensorMap mx, mw;
  if (int rc = make_map_2d_box(&mx, x, Q, channels, win_rows, 64, what)) return rc;
  if (int rc = make_map_2d(&mw, w, n_out, static_cast<int64_t>(kh) * kw * channels, n_out, what)) return rc;
  const size_t smem = 1024 + p.w_bytes + static_cast<size_t>(stages) * p.stage_bytes + 2048 + epi_bytes;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n_out == 128) return launch_flat<128>(mx, mw, p, smem, s, what);
  if (n_out == 64) return launch_flat<64>(mx, mw, p, smem, s, what);
  return launch_flat<32>(mx, mw, p, smem, s, what);
}
