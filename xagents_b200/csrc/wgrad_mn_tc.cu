// Convolution weight (and bias) gradient on the tensor cores straight from the NATURAL NHWC tensors -- no transposed
// copies, no im2col matrix.
//
//   dW[n, (kh, kw, c)] = sum over output pixels (b, y, x) of  dY[b, y, x, n] * X[b, y + kh, x + kw, c]
//
// dY lives on the INPUT pixel grid (zero where y >= OH or x >= OW; the kernels that produce it write it there, see
// xa_conv2d_nhwc_bf16_ex / xa_gemm_bf16_tn_ex) and pixels are flattened as q = (b*H + y)*W + x, so a kernel tap is a
// shift of the pixel index, q -> q + kh*W + kw (where the shift would wrap into the next row or image dY is zero):
//
//   dW_tap[c, n] = sum_q X[q + shift_tap, c] * dYg[q, n]          X [Q, C], dYg [Q, N]  bf16, pixel-major
//
// Both operands are "MN-major" for tcgen05 (the contraction index q is the slow one), which the UMMA descriptors
// support for bf16: a TMA box of 64 pixels x 64 channels (128-B rows, SWIZZLE_128B) IS the canonical MN-major tile, and
// the tap shift is a shift of the box's ROW coordinate, which TMA does not need aligned.  The kw shifts of one kernel row
// are not even loaded separately: one box of 72 rows per kernel row is fetched and the tap is selected by advancing the
// descriptor's start address by kw * 128 B (the 128-B swizzle is a function of the absolute shared-memory address, so
// a row-shifted window of a swizzled tile is still a valid swizzled operand).
//   A (M = 128)  two (tap, 64-channel) groups of X, LBO = their distance in shared memory
//   B (N = n_out) the dYg tile
//   D            TMEM lanes = (group, c), columns = n; one accumulator column block per group pair, all resident
// The bias gradient db[n] = sum_q dYg[q, n] comes from the same pass: the four epilogue warps, idle while the K loop runs,
// add up the columns of every dYg tile in shared memory (a first version multiplied by a group of constant ones on the
// tensor core: one more MMA per K step, and these kernels are bound by the shared-memory reads of their MMA operands --
// 4 KB of A per instruction for N = 32 / 64 -- not by tensor throughput).
// The pixel range is split over the SMs (split-K); fp32 partials are added in split order (deterministic).
// HBM traffic: (C + N) * Q * 2 bytes, each operand read once.
#include "tc_common.cuh"

#include <cstdlib>

namespace {

using namespace xa_tc;

constexpr int kMaxGroups = 20;
constexpr int kMaxBoxes = 20;

struct WgradMnParams {
  float* partial;  // [splits, n_out, ld_p]
  int n_out, n_groups, n_boxes, box_rows, ld_p, db_col;
  int kb_per_split;
  int64_t k_blocks;
  uint32_t stage_bytes, b_off, tmem_cols;
  int box_shift[kMaxBoxes];   // pixel shift of the box's first row
  int box_col[kMaxBoxes];     // first channel of the box
  int box_off[kMaxBoxes];     // byte offset inside the stage
  int group_off[kMaxGroups];  // byte offset of the group's first row inside the stage
  int group_col[kMaxGroups];  // column of the group's channel 0 in a partial row
};

template <int kMB>  // accumulator column blocks = ceil(groups / 2): descriptors of one K block live in registers
__global__ void __launch_bounds__(kThreads) wgrad_mn_kernel(const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_dy,
                                                            const __grid_constant__ WgradMnParams p, int stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* db_red = reinterpret_cast<float*>(smem + static_cast<size_t>(stages) * p.stage_bytes);  // [4 warps][64] column sums
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(db_red) + 8192);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  uint64_t* desc_tab = acc_full + 2;  // [stages][kMB]: A descriptor of (stage, column block) at K step 0

  xa::pdl_trigger();   // chained launch (xa_common.cuh): the on-chip set-up below overlaps the predecessor's tail
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x;
  const int64_t kb0 = static_cast<int64_t>(split) * p.kb_per_split;
  const int64_t kb1 = kb0 + p.kb_per_split < p.k_blocks ? kb0 + p.kb_per_split : p.k_blocks;
  constexpr int n_mblocks = kMB;

  for (int i = threadIdx.x; i < stages * kMB; i += kThreads) {
    const int s = i / kMB, mb = i - s * kMB;
    const uint32_t base = xa::smem_u32(smem + static_cast<size_t>(s) * p.stage_bytes);
    // an odd group count pairs the last group with itself (LBO = 0): lanes 64-127 of that block are a copy the epilogue skips
    const int g0 = 2 * mb, g1 = 2 * mb + 1 < p.n_groups ? 2 * mb + 1 : 2 * mb;
    const uint32_t a0 = base + p.group_off[g0], a1 = base + p.group_off[g1];
    desc_tab[i] = make_smem_desc_mn(a0, a1 - a0, 1024, false);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA's async proxy
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    for (int s = 0; s < stages; ++s) {
      xa::mbar_init(full + s, 1);
      xa::mbar_init(empty + s, 1 + 4);  // the MMAs' commit + the four column-summing warps
    }
    xa::mbar_init(acc_full, 1);
    xa::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xa::smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  xa::pdl_wait();      // first global-memory access below: the predecessor grid has completed

  if (warp == 0) {
    if (elect_one()) {  // ---- TMA producer (no division / modulo per K block: `stages` is a run-time value)
      int s = 0;
      uint32_t round = 0;
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        if (round > 0) mbar_wait_wd(empty + s, (round - 1) & 1);
        uint8_t* dst = smem + static_cast<size_t>(s) * p.stage_bytes;
        xa::mbar_expect_tx(full + s, p.stage_bytes);
        const int q0 = static_cast<int>(kb * kBlockK);
        tma_load_2d(dst + p.b_off, &map_dy, 0, q0, full + s);
        for (int b = 0; b < p.n_boxes; ++b) tma_load_2d(dst + p.box_off[b], &map_x, p.box_col[b], q0 + p.box_shift[b], full + s);
        if (++s == stages) s = 0, ++round;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ---- MMA issuer
      const uint32_t idesc = make_idesc(kBlockM, p.n_out, true, true);
      const bool sw64 = p.n_out == 32;                // dYg rows of 64 B
      const uint32_t b_kstep = sw64 ? 1024u : 2048u;  // 16 pixel rows
      const uint64_t db_proto = make_smem_desc_mn(0, 0, sw64 ? 512 : 1024, sw64);
      int s = 0;
      uint32_t phase = 0;
      const uint32_t b_addr0 = xa::smem_u32(smem) + p.b_off;
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        uint64_t da[kMB];
#pragma unroll
        for (int mb = 0; mb < kMB; ++mb) da[mb] = desc_tab[s * kMB + mb];
        const uint32_t b_addr = b_addr0 + static_cast<uint32_t>(s) * p.stage_bytes;
        mbar_wait_wd(full + s, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // K step outer, accumulator inner: consecutive MMAs write DIFFERENT accumulators, so the tensor pipe never waits
        // for the previous instruction's accumulation (back-to-back MMAs on one accumulator serialise at ~150 cycles)
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {  // 16 pixel rows further: +2048 B in the start-address field
#pragma unroll
          for (int mb = 0; mb < kMB; ++mb)
            umma_bf16(tmem_base + mb * p.n_out, da[mb] + k * (2048u >> 4), db_proto | ((b_addr + k * b_kstep) >> 4), idesc,
                      (kb > kb0) || (k != 0));
        }
        umma_commit(empty + s);
        if (++s == stages) s = 0, phase ^= 1;
      }
      umma_commit(acc_full);
    }
  } else {
    // ---- while the K loop runs: column sums of every dYg tile (the bias gradient).  Warp `quad` takes 16 of the tile's 64
    // pixel rows; a lane reads two adjacent columns (4 bytes) per row, un-swizzling by hand: 128-byte rows (n_out = 64)
    // have their 16-byte chunk index XORed with (row & 7), 64-byte rows (n_out = 32) with ((row >> 1) & 3), lanes 0-15 on
    // the even row of a pair and lanes 16-31 on the odd one.
    const int quad = warp & 3;
    float db0 = 0.0f, db1 = 0.0f;
    {
      const bool wide = p.n_out == 64;
      const uint32_t b_addr0 = xa::smem_u32(smem) + p.b_off;
      const int lane_row = wide ? 0 : lane >> 4;            // which row of a pair (narrow tiles)
      const uint32_t chunk = wide ? lane >> 2 : (lane & 15) >> 2, within = (lane & 3) * 4u;
      int s = 0;
      uint32_t phase = 0;
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        mbar_wait_backoff(full + s, phase);
        const uint32_t tile = b_addr0 + static_cast<uint32_t>(s) * p.stage_bytes;
        if (wide) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t row = quad * 16 + i;
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tile + row * 128u + ((chunk ^ (row & 7u)) << 4) + within));
            const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
            db0 += f.x, db1 += f.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t row = quad * 16 + 2 * i + lane_row;
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(tile + row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4) + within));
            const float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
            db0 += f.x, db1 += f.y;
          }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xa::smem_u32(empty + s)) : "memory");
        if (++s == stages) s = 0, phase ^= 1;
      }
      if (!wide) {  // the odd rows' sums join the even rows' (fixed order)
        db0 += __shfl_xor_sync(0xffffffffu, db0, 16);
        db1 += __shfl_xor_sync(0xffffffffu, db1, 16);
      }
      const int cols2 = p.n_out / 2;  // lanes holding a column pair
      if (lane < cols2) db_red[quad * 64 + 2 * lane] = db0, db_red[quad * 64 + 2 * lane + 1] = db1;
    }
    // ---- epilogue: lanes = (group, channel), columns = n  ->  partial[split][n][group_col + channel]
    mbar_wait_wd(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int mb = 0; mb < n_mblocks; ++mb) {
      const int g = 2 * mb + (quad >> 1);
      const int c = (quad & 1) * 32 + lane;
      const bool valid = g < p.n_groups;
      float* dst = p.partial + static_cast<int64_t>(split) * p.n_out * p.ld_p + (valid ? p.group_col[g] : 0) + c;
#pragma unroll 1
      for (int n0 = 0; n0 < p.n_out; n0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + mb * p.n_out + n0, v);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dst[static_cast<int64_t>(n0 + j) * p.ld_p] = __uint_as_float(v[j]);
        }
      }
    }
    // bias gradient of this split: the four warps' column sums added in warp order -> column `db_col` of the partial rows
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
    {
      const int n = (quad * 32 + lane);
      if (n < p.n_out)
        p.partial[(static_cast<int64_t>(split) * p.n_out + n) * p.ld_p + p.db_col] = ((db_red[n] + db_red[64 + n]) + db_red[128 + n]) + db_red[192 + n];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// dw[n, col] = sum_s partial[s][n][col] for col < ld_out; db[n] = the column after them.
// Four lanes per output, each adding a contiguous quarter of the splits in order, combined by a fixed shuffle tree:
// deterministic, and four times the loads in flight of one thread walking all (up to 148) splits.
__global__ void __launch_bounds__(256) wgrad_mn_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                                               int n_out, int ld_out, int ld_p, int splits) {
  const int64_t total = static_cast<int64_t>(n_out) * (ld_out + 1);
  const int part = threadIdx.x & 3;
  const int per = (splits + 3) / 4;
  const int s_begin = part * per, s_end = s_begin + per < splits ? s_begin + per : splits;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 2; i < ((total + 63) / 64) * 64;
       i += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 2) {
    float acc = 0.0f;
    int n = 0, col = 0;
    if (i < total) {
      n = static_cast<int>(i / (ld_out + 1)), col = static_cast<int>(i - static_cast<int64_t>(n) * (ld_out + 1));
      const float* src = partial + static_cast<int64_t>(n) * ld_p + col;
      const int64_t step = static_cast<int64_t>(n_out) * ld_p;
      for (int s = s_begin; s < s_end; ++s) acc += src[s * step];
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);  // (q0 + q1), (q2 + q3)
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);  // (q0 + q1) + (q2 + q3)
    if (part == 0 && i < total) {
      if (col < ld_out)
        dw[static_cast<int64_t>(n) * ld_out + col] = acc;
      else if (db != nullptr)
        db[n] = acc;
    }
  }
}

template <int kMB>
int launch_wgrad_mn(const CUtensorMap& mx, const CUtensorMap& mdy, const WgradMnParams& p, int stages, int splits, size_t smem, cudaStream_t s,
                    const char* what) {
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_mn_kernel<kMB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e != cudaSuccess) {
      xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_dev = dev;
  }
  xa::launch_chained(xa::kChainLarge, wgrad_mn_kernel<kMB>, dim3(splits), dim3(kThreads), smem, s, mx, mdy, p, stages);
  return xa::check_launch(what);
}

int wgrad_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("XA_WGRAD_MODE");
    mode = e ? std::atoi(e) : 0;
  }
  return mode;
}

}  // namespace

extern "C" {

int64_t xa_conv_wgrad_nhwc_workspace_bytes(int n_out, int channels, int kh, int kw) {
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  return static_cast<int64_t>(sms) * n_out * (kh * kw * channels + 8) * static_cast<int64_t>(sizeof(float));
}

int xa_conv_wgrad_nhwc_plan(int n_out, int channels, int kh, int kw, int64_t q_total, int* splits, int* ld_partial) {
  const char* what = "xa_conv_wgrad_nhwc_plan";
  XA_REQUIRE(n_out > 0 && channels > 0 && kh > 0 && kw > 0 && q_total > 0, XA_EINVAL, "%s: non-positive size", what);
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int64_t k_blocks = (q_total + kBlockK - 1) / kBlockK;
  int s = static_cast<int>(k_blocks < sms ? k_blocks : sms);
  const int64_t per = (k_blocks + s - 1) / s;
  s = static_cast<int>((k_blocks + per - 1) / per);
  if (splits) *splits = s;
  if (ld_partial) *ld_partial = kh * kw * channels + 8;
  return XA_OK;
}

static int wgrad_run(const char* what, const void* x, const void* dy_grid, float* dw, float* db, int n_out, int channels, int kh, int kw,
                     int grid_w, int64_t q_total, void* workspace, int64_t workspace_bytes, bool reduce, xa_stream_t stream) {
  XA_REQUIRE(x && dy_grid && (dw || !reduce) && workspace, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE((n_out == 32 || n_out == 64) && channels > 0 && channels % 64 == 0 && kh > 0 && kw > 0 && kw <= 8 && grid_w >= kw, XA_EINVAL,
             "%s: n_out=%d channels=%d kernel %dx%d not supported", what, n_out, channels, kh, kw);
  const int cg = channels / 64;
  const int n_groups = kh * kw * cg;
  const int n_mblocks = (n_groups + 1) / 2;
  XA_REQUIRE(n_groups <= kMaxGroups && n_mblocks * n_out <= 512, XA_EINVAL, "%s: %d taps x %d channels exceed the TMEM accumulator", what,
             kh * kw, channels);
  XA_REQUIRE(q_total > 0 && q_total < (int64_t(1) << 31) - 4096, XA_EOVERFLOW, "%s: q_total out of range", what);
  XA_REQUIRE(xa::aligned(x, 16) && xa::aligned(dy_grid, 16) && (!reduce || xa::aligned(dw, 16)) && xa::aligned(workspace, 16), XA_EALIGN,
             "%s: 16-byte alignment required", what);
  XA_REQUIRE(workspace_bytes >= xa_conv_wgrad_nhwc_workspace_bytes(n_out, channels, kh, kw), XA_ENOSPACE, "%s: workspace too small", what);
  // 0: ONE window per 64-channel block (64 + (kh-1)*W + (kw-1) pixel rows), every tap a descriptor shifted by whole rows --
  //    each input pixel crosses L2 -> shared memory once per K block (these kernels were bound by that traffic, ~8 TB/s
  //    chip-wide, when every kernel row had its own box); 1: one box per kernel row; 2: one box per tap
  const int mode = wgrad_mode();
  WgradMnParams p{};
  p.partial = static_cast<float*>(workspace);
  p.n_out = n_out, p.n_groups = n_groups;
  p.ld_p = kh * kw * channels + 8;
  const bool per_tap = mode == 2;
  const int window_rows = ((64 + (kh - 1) * grid_w + (kw - 1) + 7) / 8) * 8;
  const bool window = mode == 0 && window_rows <= 256;
  p.box_rows = per_tap ? 64 : window ? window_rows : 64 + ((kw - 1 + 7) / 8) * 8;
  const int box_bytes = p.box_rows * 128;
  int nb = 0, ng = 0;
  if (window) {
    for (int g = 0; g < cg; ++g, ++nb) p.box_shift[nb] = 0, p.box_col[nb] = g * 64, p.box_off[nb] = nb * box_bytes;
    for (int i = 0; i < kh; ++i)
      for (int g = 0; g < cg; ++g)
        for (int j = 0; j < kw; ++j, ++ng)
          p.group_off[ng] = p.box_off[g] + (i * grid_w + j) * 128, p.group_col[ng] = (i * kw + j) * channels + g * 64;
  } else {
    for (int i = 0; i < kh; ++i) {
      for (int g = 0; g < cg; ++g) {
        if (per_tap) {
          for (int j = 0; j < kw; ++j, ++nb, ++ng) {
            p.box_shift[nb] = i * grid_w + j, p.box_col[nb] = g * 64, p.box_off[nb] = nb * box_bytes;
            p.group_off[ng] = p.box_off[nb], p.group_col[ng] = (i * kw + j) * channels + g * 64;
          }
        } else {
          p.box_shift[nb] = i * grid_w, p.box_col[nb] = g * 64, p.box_off[nb] = nb * box_bytes;
          for (int j = 0; j < kw; ++j, ++ng) p.group_off[ng] = p.box_off[nb] + j * 128, p.group_col[ng] = (i * kw + j) * channels + g * 64;
          ++nb;
        }
      }
    }
  }
  XA_REQUIRE(nb <= kMaxBoxes, XA_EINVAL, "%s: too many boxes", what);
  p.db_col = kh * kw * channels;
  p.n_boxes = nb;
  p.b_off = static_cast<uint32_t>(nb) * box_bytes;
  p.stage_bytes = p.b_off + (n_out == 64 ? 8192u : 4096u);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(n_mblocks * n_out)) cols *= 2;
  p.tmem_cols = cols;
  int stages = static_cast<int>((214 * 1024 - 8192) / p.stage_bytes);
  if (stages > 8) stages = 8;
  XA_REQUIRE(stages >= 2, XA_EINVAL, "%s: a stage of %u bytes does not fit twice in shared memory", what, p.stage_bytes);
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  p.k_blocks = (q_total + kBlockK - 1) / kBlockK;
  int splits = static_cast<int>(p.k_blocks < sms ? p.k_blocks : sms);
  p.kb_per_split = static_cast<int>((p.k_blocks + splits - 1) / splits);
  splits = static_cast<int>((p.k_blocks + p.kb_per_split - 1) / p.kb_per_split);
  CUtensorMap mx, mdy;
  if (int rc = make_map_2d_box(&mx, x, q_total, channels, p.box_rows, 64, what)) return rc;
  if (int rc = make_map_2d_box(&mdy, dy_grid, q_total, n_out, 64, n_out, what)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(stages) * p.stage_bytes + 8192 + 1024 + 1024;
  int rc = XA_EINVAL;
  switch (n_mblocks) {
#define XA_WGRAD_CASE(MB) \
  case MB:                \
    rc = launch_wgrad_mn<MB>(mx, mdy, p, stages, splits, smem, s, what); \
    break;
    XA_WGRAD_CASE(1) XA_WGRAD_CASE(2) XA_WGRAD_CASE(3) XA_WGRAD_CASE(4) XA_WGRAD_CASE(5)
    XA_WGRAD_CASE(6) XA_WGRAD_CASE(7) XA_WGRAD_CASE(8) XA_WGRAD_CASE(9) XA_WGRAD_CASE(10)
#undef XA_WGRAD_CASE
    default:
      xa::set_error("%s: %d accumulator blocks", what, n_mblocks);
  }
  if (rc || !reduce) return rc;
  const int ld_out = kh * kw * channels;
  const int64_t total = static_cast<int64_t>(n_out) * (ld_out + 1);
  wgrad_mn_reduce_kernel<<<static_cast<unsigned>((total * 4 + 255) / 256), 256, 0, s>>>(static_cast<const float*>(workspace), dw, db, n_out, ld_out,
                                                                                     p.ld_p, splits);
  return xa::check_launch(what);
}

int xa_conv_wgrad_nhwc_bf16(const void* x, const void* dy_grid, float* dw, float* db, int n_out, int channels, int kh, int kw,
                            int grid_w, int64_t q_total, void* workspace, int64_t workspace_bytes, xa_stream_t stream) {
  return wgrad_run("xa_conv_wgrad_nhwc_bf16", x, dy_grid, dw, db, n_out, channels, kh, kw, grid_w, q_total, workspace, workspace_bytes, true,
                   stream);
}

int xa_conv_wgrad_nhwc_bf16_partial(const void* x, const void* dy_grid, int n_out, int channels, int kh, int kw, int grid_w, int64_t q_total,
                                    float* partial, int64_t partial_bytes, xa_stream_t stream) {
  return wgrad_run("xa_conv_wgrad_nhwc_bf16_partial", x, dy_grid, nullptr, nullptr, n_out, channels, kh, kw, grid_w, q_total, partial,
                   partial_bytes, false, stream);
}

}  // extern "C"
