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

// Epilogue groups of four warps, one TMEM accumulator each.  The bf16-input kernels run FOUR (all 512 TMEM columns at BN = 128):
// with two, the data gradients and the 64-channel layers were bound by the instruction issue of their eight epilogue warps
// (conv2's data gradient: 5600 epilogue warp-instructions per tile at IPC 1.3 = the 4200 cycles a tile took; DRAM, the tensor pipe
// and shared memory all below 50 %) -- halving their DRAM reads with bit masks changed nothing.  The uint8 first layer keeps two
// (its five converter warps need the registers, and its 32-column epilogue is short).
constexpr int kGroupsU8 = 2;
__host__ __device__ constexpr int groups_bf16(int bn) { return bn >= 128 ? 4 : 2; }   // measured: four groups gain 11 us at BN = 128, lose 2.5 at 64
constexpr int kConvertWarps = 5;                              // uint8 input: five more warps turn raw windows into bf16 operand tiles
__host__ __device__ constexpr int flat_threads(int bn, bool u8) {
  return 64 + (u8 ? kGroupsU8 : groups_bf16(bn)) * 4 * 32 + (u8 ? kConvertWarps * 32 : 0);
}
// One staged row of a 32-column chunk = 64 bytes, unpadded; the 16-byte piece j of row r sits at piece j ^ ((r >> 1) & 3).  That keeps
// BOTH access patterns of the epilogue conflict-free: rows in (lane = row, piece j: a quarter-warp covers eight different 16-byte bank
// groups) and pieces out (four lanes per row, eight rows per access).  The padded 80-byte pitch it replaces was conflict-free only for the
// writes: every piece-wise read cost 8 wavefronts instead of 4, and these kernels are bound by shared-memory bandwidth (conv2's data
// gradient: 1024 cycles of MMA operand reads + ~1500 LSU wavefronts per tile of ~3300 cycles, ncu).
constexpr int kEpiPitch = 32 * 2;
__device__ __forceinline__ uint32_t epi_piece(int row, int piece) { return static_cast<uint32_t>(row * kEpiPitch + ((piece ^ ((row >> 1) & 3)) << 4)); }
__host__ __device__ constexpr int epi_warp_bytes(int bn) { return 32 * kEpiPitch + 512 + 32 * (bn / 32) * 4; }   // per epilogue warp: the chunk's tile + row tables
//                                                                                            (output offsets, mask offsets, mask bits)
constexpr int kRawStages = 4;

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
  // ReLU derivatives as BIT masks (bit j of word i <=> element 32 i + j of the compact [B, OH, OW, N] activation is > 0): a
  // forward layer writes them from its epilogue (bits_out, 1/16 of the activation's bytes), the data gradient of the layer above
  // reads them (bits_in) instead of re-reading the bf16 activation for its sign -- 210 MB less traffic for conv2's data gradient
  uint32_t* bits_out;
  const uint32_t* bits_in;
  int B, H, W, N, OH, OW, PH, PW;
  int relu, out_mode;
  int n_entries;   // taps x 64-channel blocks
  int kc_blocks;   // 64-channel blocks of the input
  int win_rows;    // rows of the window box (multiple of 8)
  int min_shift;   // row shift of the first tap: -(pad_y*W + pad_x)
  int stages;
  int64_t Q;       // B*H*W
  uint32_t w_bytes, stage_bytes;
  uint32_t raw_stage_bytes, raw_tx_bytes;  // uint8 input: bytes of one raw window stage / of one TMA box
  uint32_t div_w_magic;                    // (g * div_w_magic) >> 16 == g / W for every pixel index g inside a raw box
  int store_x1;                            // uint8 input: also write the converted bf16 rows to HBM through map_x1
  // uint8 input read THROUGH A PERMUTATION (the minibatch gather folded into the first layer): frame f of the batch is row
  // frame_idx[f] (env-major sample id when idx_T > 0: row = (id % T) * E + id / T, xagents/base.py:559-564) of `frames_base`
  const uint8_t* frames_base;
  const int32_t* frame_idx;
  int idx_T, idx_E, box_rows;
  uint32_t grid_row_bytes;                 // one row of the space-to-depth grid = 4 image rows
  uint32_t a_units[kMaxEntries];  // (byte offset of the entry's A operand inside a stage) >> 4
  uint32_t b_units[kMaxEntries];  // (byte offset of the entry's weight tile) >> 4
};

// kU8: the input is the raw uint8 frame tensor [B, 4H, 4W, 4] of a 4x4-strided first layer (space-to-depth grid H x W, 64
// "channels" = (dy, dx, c)).  map_x then covers the frames as image rows ([B*4H rows, 16 W bytes]); a box of 4 nY whole image
// rows = nY grid rows arrives as they lie in memory (long TMA rows: a box shaped [pixel][dy][16 B] would be fetched in
// 16-byte requests), and five converter warps pick the 16-byte (dx, c) pieces of every (pixel, dy), scale by 1/255, round
// to bf16 and write the SWIZZLE_128B operand tile the MMA reads -- the 2x larger bf16 space-to-depth tensor is never written to or read from HBM (xagents/base.py:505-506: the
// cast and the division of the image batch, fused here into the first layer).
template <int BN, bool kU8 = false>
__global__ void __launch_bounds__(flat_threads(BN, kU8)) conv_flat_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                 const __grid_constant__ CUtensorMap map_w,
                                                                 const __grid_constant__ CUtensorMap map_x1,
                                                                 const __grid_constant__ FlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + p.w_bytes;  // weights first, then the ring of window stages
  uint8_t* raw_ring = ring + static_cast<size_t>(p.stages) * p.stage_bytes;  // kU8 only
  uint8_t* tail = raw_ring + (kU8 ? static_cast<size_t>(kRawStages) * p.raw_stage_bytes : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + 8;
  constexpr int kGroups = kU8 ? kGroupsU8 : groups_bf16(BN);
  uint64_t* acc_full = empty + 8;
  uint64_t* acc_empty = acc_full + 4;
  uint64_t* w_full = acc_empty + 4;
  uint64_t* raw_full = w_full + 1;             // [kRawStages]
  uint64_t* raw_empty = raw_full + kRawStages;  // [kRawStages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(raw_empty + kRawStages);
  uint32_t* s_a = reinterpret_cast<uint32_t*>(tail + 512);  // [kMaxEntries]
  uint32_t* s_b = s_a + kMaxEntries;
  float* s_bias = reinterpret_cast<float*>(tail + 512 + 2 * kMaxEntries * 4);  // [N <= 128]
  uint8_t* epi_stage = tail + 2048;  // 4 kGroups warps x epi_warp_bytes(BN), then (bits_in only) the 4 KB bit-expansion table

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the on-chip set-up below overlaps the predecessor's tail
  for (int i = threadIdx.x; i < p.n_entries; i += blockDim.x) s_a[i] = p.a_units[i], s_b[i] = p.b_units[i];
  if (p.bits_in != nullptr) {  // byte b of a row's mask bits -> the AND masks of its eight bf16 values
    uint32_t* table = reinterpret_cast<uint32_t*>(epi_stage + 4 * kGroups * epi_warp_bytes(BN));
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
      const int b = i >> 2, k = i & 3;
      table[i] = (((b >> (2 * k)) & 1) ? 0x0000FFFFu : 0u) | (((b >> (2 * k + 1)) & 1) ? 0xFFFF0000u : 0u);
    }
  }
  const int n_tiles = static_cast<int>((p.Q + kBlockM - 1) / kBlockM);
  constexpr uint32_t kAccStride = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemCols = kGroups * kAccStride;
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation: a power of two, at most 512 columns");

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      xa::mbar_init(full + s, kU8 ? kConvertWarps : 1);  // filled by TMA, or by one arrival per converter warp
      xa::mbar_init(empty + s, 1);
    }
    for (int a = 0; a < kGroups; ++a) {
      xa::mbar_init(acc_full + a, 1);
      xa::mbar_init(acc_empty + a, 4);
    }
    xa::mbar_init(w_full, 1);
    if (kU8) {
      for (int s = 0; s < kRawStages; ++s) {
        xa::mbar_init(raw_full + s, 1);
        xa::mbar_init(raw_empty + s, kConvertWarps);
      }
    }
    xa::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xa::smem_u32(tmem_slot)), "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  xa::pdl_wait();      // first global-memory access below: the predecessor grid has completed
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) s_bias[i] = p.bias != nullptr ? p.bias[i] : 0.0f;
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
      if (kU8 && p.frames_base != nullptr) {
        // frames addressed one by one: a frame is contiguous (H grid rows of grid_row_bytes), so the box of a tile is one bulk
        // copy of the grid rows [gy0, gy0 + n0) of frame f0 and, where the window runs into the next frame of the batch, a
        // second one of that frame's first rows, placed behind it -- the same shared-memory picture the tensor-map box gives
        // for frames that lie in batch order.  The frame ids of the NEXT tile are fetched while this tile's stage is awaited
        // (a first version used them at once: one L2 round trip per tile in this single thread made the layer 256 us instead of 178).
        const uint32_t frame_bytes = static_cast<uint32_t>(p.H) * p.grid_row_bytes;
        const uint32_t hw = static_cast<uint32_t>(p.H) * p.W;
        auto ids_of = [&](int tile, uint32_t& f0, int32_t& a, int32_t& b) {   // loads only: the arithmetic on them waits a tile
          f0 = static_cast<uint32_t>(tile) * kBlockM / hw;
          a = p.frame_idx != nullptr ? __ldg(p.frame_idx + f0) : static_cast<int32_t>(f0);
          b = static_cast<int>(f0) + 1 < p.B ? (p.frame_idx != nullptr ? __ldg(p.frame_idx + f0 + 1) : static_cast<int32_t>(f0 + 1)) : 0;
        };
        auto row_of = [&](int32_t id) -> int64_t {
          return p.idx_T > 0 ? static_cast<int64_t>(id % p.idx_T) * p.idx_E + id / p.idx_T : static_cast<int64_t>(id);
        };
        int tile = blockIdx.x;
        uint32_t f0 = 0, nf0 = 0;
        int32_t a = 0, b = 0, na = 0, nb = 0;
        if (tile < n_tiles) ids_of(tile, f0, a, b);
        for (; tile < n_tiles; tile += gridDim.x) {
          const int next = tile + gridDim.x;
          if (next < n_tiles) ids_of(next, nf0, na, nb);
          const uint32_t q0 = static_cast<uint32_t>(tile) * kBlockM;
          const int gy0 = static_cast<int>((q0 - f0 * hw) / p.W);
          const int n0 = p.box_rows < p.H - gy0 ? p.box_rows : p.H - gy0;
          const int n1 = static_cast<int>(f0) + 1 < p.B ? p.box_rows - n0 : 0;
          const uint8_t* src0 = p.frames_base + row_of(a) * frame_bytes + static_cast<uint32_t>(gy0) * p.grid_row_bytes;
          const uint8_t* src1 = p.frames_base + row_of(b) * frame_bytes;
          if (round > 0) mbar_wait_wd(raw_empty + s, (round - 1) & 1);
          uint8_t* dst = raw_ring + static_cast<size_t>(s) * p.raw_stage_bytes;
          xa::mbar_expect_tx(raw_full + s, static_cast<uint32_t>(n0 + n1) * p.grid_row_bytes);
          // no L2 hint: the rows a neighbouring tile's window shares with this one are re-read from L2 (evict-first measured the same)
          xa::bulk_g2s_nohint(dst, src0, static_cast<uint32_t>(n0) * p.grid_row_bytes, raw_full + s);
          if (n1 > 0) xa::bulk_g2s_nohint(dst + static_cast<uint32_t>(n0) * p.grid_row_bytes, src1, static_cast<uint32_t>(n1) * p.grid_row_bytes, raw_full + s);
          f0 = nf0, a = na, b = nb;
          if (++s == kRawStages) s = 0, ++round;
        }
      } else if (kU8) {  // raw uint8 windows: the nY whole grid rows that contain pixels [q0, q0 + win_rows)
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
          if (round > 0) mbar_wait_wd(raw_empty + s, (round - 1) & 1);
          xa::mbar_expect_tx(raw_full + s, p.raw_tx_bytes);
          tma_load_2d(raw_ring + static_cast<size_t>(s) * p.raw_stage_bytes, &map_x, 0, 4 * ((tile * kBlockM) / p.W), raw_full + s);
          if (++s == kRawStages) s = 0, ++round;
        }
      } else {
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
          if (round > 0) mbar_wait_wd(empty + s, (round - 1) & 1);
          uint8_t* dst = ring + static_cast<size_t>(s) * p.stage_bytes;
          xa::mbar_expect_tx(full + s, p.stage_bytes);
          const int q0 = tile * kBlockM + p.min_shift;  // may be negative: rows before the tensor read as zeros
          for (int c = 0; c < p.kc_blocks; ++c) tma_load_2d(dst + c * blk_bytes, &map_x, c * kBlockK, q0, full + s);
          if (++s == p.stages) s = 0, ++round;
        }
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
        const uint32_t acc = lt % kGroups, use = lt / kGroups;
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
  } else if (kU8 && warp >= 2 + 4 * kGroupsU8) {
    // ---- converters (uint8 input): raw window stage -> bf16 operand stage.  A thread owns one window pixel: its four
    // 16-byte pieces (the (dx, c) bytes of dy = 0..3, one image row apart) become the eight 16-byte chunks of the pixel's
    // 128-byte operand row, stored where SWIZZLE_128B puts them (chunk ^ (row & 7): stage bases are 1024-byte aligned, so
    // the row index supplies address bits 7-9).  Consecutive threads take consecutive pixels: reads walk an image row, the
    // stores of 8 threads hit 8 different chunks.  The kernel is bound by instruction issue (ncu: 4600 warp instructions
    // per tile before this layout, 58 % of all issue slots), so the address arithmetic is done once per pixel, not per piece.
    // byte -> float without the (quarter-rate) integer conversion unit: PRMT builds 2^23 + x, one FMA does
    // (2^23 + x) * m - 2^23 * m = fl(x * m) with m = fl(1/255) -- the product the space-to-depth kernel rounds to bf16
    // (bit-identical to x / 255 for every byte value, tests/test_gpu_conv.py).
    // Warp w converts window rows [32 w, 32 w + 32).  With store_x1 the tile's own 128 rows (warps 0-3) also go to HBM as the
    // [Q, 64] bf16 space-to-depth tensor the weight-gradient kernel reads in the backward pass: each warp's elected lane
    // TMA-stores its 32 rows from the operand stage (SWIZZLE_128B undone by the store's tensor map) and, before writing
    // that stage again, waits until the store has read it.
    const int cw = warp - (2 + 4 * kGroupsU8);
    const float mul = 1.0f / 255.0f, bias23 = -8388608.0f * mul;
    const uint32_t ring_u32 = xa::smem_u32(ring), raw_u32 = xa::smem_u32(raw_ring);
    const uint32_t dy_pitch = static_cast<uint32_t>(p.W) * 16u;  // one image row
    const bool storing = p.store_x1 != 0 && cw < kBlockM / 32;
    int rs = 0, bs = 0;
    uint32_t raw_phase = 0, b_round = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int q0 = tile * kBlockM;
      const int w0 = q0 - (q0 / p.W) * p.W;  // first window pixel inside the first grid row of the box
      mbar_wait_backoff(raw_full + rs, raw_phase);
      if (b_round > 0) mbar_wait_backoff(empty + bs, (b_round - 1) & 1);
      if (storing) {  // at most the two most recent of my stores may still be reading their stages (stages >= 3)
        if (lane == 0) xa::bulk_wait_read<2>();
        __syncwarp();
      }
      const uint32_t src0 = raw_u32 + static_cast<uint32_t>(rs) * p.raw_stage_bytes;
      const uint32_t dst0 = ring_u32 + static_cast<uint32_t>(bs) * p.stage_bytes;
#pragma unroll 1
      for (int row = cw * 32 + lane; row < p.win_rows; row += kConvertWarps * 32) {
        const uint32_t g = static_cast<uint32_t>(w0 + row);
        const uint32_t gy = (g * p.div_w_magic) >> 16, gx = g - gy * p.W;
        const uint32_t src = src0 + (gy * 4u * p.W + gx) * 16u;
        uint4 in[4];
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) in[dy] = lds_u4(src + dy * dy_pitch);
        const uint32_t row_u32 = dst0 + static_cast<uint32_t>(row) * 128u;
        const uint32_t sw = static_cast<uint32_t>(row & 7) << 4;
#pragma unroll
        for (int dy = 0; dy < 4; ++dy) {
          const uint32_t words[4] = {in[dy].x, in[dy].y, in[dy].z, in[dy].w};
          __nv_bfloat162 out[8];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float f0 = fmaf(__uint_as_float(__byte_perm(words[k], 0x4B000000u, 0x7650)), mul, bias23);
            const float f1 = fmaf(__uint_as_float(__byte_perm(words[k], 0x4B000000u, 0x7651)), mul, bias23);
            const float f2 = fmaf(__uint_as_float(__byte_perm(words[k], 0x4B000000u, 0x7652)), mul, bias23);
            const float f3 = fmaf(__uint_as_float(__byte_perm(words[k], 0x4B000000u, 0x7653)), mul, bias23);
            out[2 * k] = __floats2bfloat162_rn(f0, f1);
            out[2 * k + 1] = __floats2bfloat162_rn(f2, f3);
          }
          sts_u4(row_u32 + ((32u * dy) ^ sw), *reinterpret_cast<uint4*>(out));
          sts_u4(row_u32 + ((32u * dy + 16u) ^ sw), *reinterpret_cast<uint4*>(out + 4));
        }
      }
      xa::fence_proxy_async();  // generic-proxy stores -> visible to the async proxy (tensor core, TMA store)
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xa::smem_u32(full + bs)) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xa::smem_u32(raw_empty + rs)) : "memory");
        if (storing) {  // rows past the end of the tensor are clipped by the tensor map
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_x1), "r"(0),
                       "r"(q0 + cw * 32), "r"(dst0 + static_cast<uint32_t>(cw) * 4096u)
                       : "memory");
          xa::bulk_commit();
        }
      }
      if (++rs == kRawStages) rs = 0, raw_phase ^= 1;
      if (++bs == p.stages) bs = 0, ++b_round;
    }
    if (storing && lane == 0) xa::bulk_wait_all<0>();  // shared memory must outlive the stores
  } else {
    // ---- epilogue: group `grp` (warps 2 + 4 grp ..) owns accumulator `grp` = the CTA's tiles lt with lt % kGroups == grp.
    // tcgen05.ld hands every thread one ROW (32 consecutive columns): stored directly, a warp-wide 16-byte access touches
    // 32 different lines, and those load/store wavefronts -- not the tensor pipe, not HBM -- bounded the data-gradient
    // layers (ncu: LSU 80 % busy).  So each warp transposes through its own padded shared-memory tile, one 32-column chunk at a
    // time: phase 1 writes the chunk's rows as they come out of TMEM (thread = row), phase 2 moves 16-byte pieces with four
    // consecutive lanes on the 64 bytes of a row, so that mask loads and output stores are coalesced.  Chunk by chunk the
    // tile costs 2.5 KB of shared memory per warp whatever BN is, which is what lets four groups fit.
    const int quad = warp & 3;
    const uint32_t grp = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const uint32_t hw = static_cast<uint32_t>(p.H) * p.W;
    constexpr int kWords = BN / 32;                             // mask words per row
    uint8_t* my_stage = epi_stage + static_cast<size_t>(warp - 2) * epi_warp_bytes(BN);
    // explicit shared-state-space accesses: through the aligned-up base pointer the compiler only sees generic addresses
    const uint32_t my_stage_u32 = xa::smem_u32(my_stage), s_bias_u32 = xa::smem_u32(s_bias);
    const uint32_t row_out_u32 = my_stage_u32 + 32 * kEpiPitch;  // [32] int64 output offset of the row, -1 = dropped
    const uint32_t row_msk_u32 = row_out_u32 + 256;              // [32] int64 mask offset of the row
    const uint32_t row_bits_u32 = row_out_u32 + 512;             // [32][kWords] this tile's mask bits, row by row (bits_in)
    const uint32_t table_u32 = xa::smem_u32(epi_stage + 4 * kGroups * epi_warp_bytes(BN));   // [256][4]: byte of mask bits -> four bf16x2 AND masks
    // phase-2 coordinates: iteration `it` moves 32 consecutive 16-byte pieces = 8 rows of the chunk; a thread keeps its piece
    const int piece = lane & 3, row0 = lane >> 2;
    const uint32_t my_piece_u32 = my_stage_u32 + epi_piece(row0, piece);   // rows it * 8 + row0 share (row >> 1) & 3: the piece index holds for all four
    // my row's pixel (image ob, position rem inside it) advances by a constant from one of this group's tiles to the next:
    // two divisions here instead of three per tile (the epilogue's instruction count is what these short-K layers wait for)
    const uint32_t step_px = static_cast<uint32_t>(kGroups) * gridDim.x * kBlockM, step_b = step_px / hw, step_rem = step_px - step_b * hw;
    int64_t q = (static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(grp) * gridDim.x) * kBlockM + r;
    uint32_t ob_u = static_cast<uint32_t>(q / hw), rem = static_cast<uint32_t>(q - static_cast<int64_t>(ob_u) * hw);
    const bool plain = p.bias == nullptr && !p.relu && p.bits_out == nullptr && (p.mask == nullptr || p.bits_in != nullptr);
    for (uint32_t lt = grp;; lt += kGroups, q += step_px) {
      const int tile = blockIdx.x + static_cast<int>(lt) * static_cast<int>(gridDim.x);
      if (tile >= n_tiles) break;
      const uint32_t acc = grp;
      int64_t my_out_off = -1;
      uint4 my_bits = make_uint4(0, 0, 0, 0);   // my row's mask bits (bits_in), words c0 / 32 = 0..3
      {  // phase 0: where does my row go?
        const int ob = static_cast<int>(ob_u);
        const uint32_t oy_u = p.div_w_magic != 0 ? (rem * p.div_w_magic) >> 16 : rem / p.W;
        const int oy = static_cast<int>(oy_u), ox = static_cast<int>(rem - oy_u * p.W);
        ob_u += step_b, rem += step_rem;
        if (rem >= hw) rem -= hw, ++ob_u;
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
        my_out_off = out_off;
        if (p.bits_in != nullptr) {  // my row's mask bits (kWords consecutive words), parked in shared memory for phase 2
          const uint32_t* src = p.bits_in + (mask_off >> 5);
          const uint32_t dst = row_bits_u32 + lane * (kWords * 4);
          if constexpr (kWords == 4) {
            const uint4 v = valid ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0, 0, 0, 0);
            if (!plain) sts_u4(dst, v);   // the plain form masks from registers: storing here would wait for the load (ncu: 9 % of all stalls)
            my_bits = v;
          } else if constexpr (kWords == 2) {
            const uint2 v = valid ? __ldg(reinterpret_cast<const uint2*>(src)) : make_uint2(0, 0);
            if (!plain) asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(dst), "r"(v.x), "r"(v.y) : "memory");
            my_bits.x = v.x, my_bits.y = v.y;
          } else {
            const uint32_t v = valid ? __ldg(src) : 0u;
            if (!plain) asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst), "r"(v) : "memory");
            my_bits.x = v;
          }
        }
      }
      __syncwarp();
      mbar_wait_backoff(acc_full + acc, (lt / kGroups) & 1);  // polling eight warps took a sixth of the SM's issue slots (ncu)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float lo = p.relu ? 0.0f : -INFINITY;  // ReLU as one max per element, no per-element branch on the flag
      // data gradients (no bias, no ReLU, no bits to write; the mask, if any, as bits): the epilogue is what these layers wait
      // for (conv2's data gradient: 1024 MMA cycles per tile against ~3700 of epilogue), so this form does nothing it does not
      // need -- no bias loads / adds / maxima (72 of ~150 instructions per chunk), and the mask is applied HERE, in the thread = row
      // registers, from the row's own bit words (phase 2 then only moves 16-byte pieces)
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (plain) {  // phase 1, plain form: TMEM -> bf16 pairs -> (mask bits) -> the chunk's rows in shared memory
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride + c0, v);
          const uint32_t dst = my_stage_u32 + lane * kEpiPitch;
          uint32_t h[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
            h[e] = *reinterpret_cast<const uint32_t*>(&t);
          }
          if (p.bits_in != nullptr) {
            // bit 2e / 2e + 1 of my row's word -> keep / clear the low / high half of pair e.  Arithmetic, not a table: the multiply
            // moves four bits into the sign bits of four bytes (no two partial products meet, so no carries), PRMT's sign-replicate mode
            // turns each into 0x00 / 0xFF bytes.  (A 4 KB byte -> mask table in shared memory cost 42 wavefronts per chunk and warp --
            // random rows, 2-3-way bank conflicts -- in a kernel bound by shared-memory bandwidth.)
            const uint32_t word = c0 == 0 ? my_bits.x : (c0 == 32 ? my_bits.y : (c0 == 64 ? my_bits.z : my_bits.w));
#pragma unroll
            for (int e = 0; e < 8; ++e) {   // nibble e = values 4 e .. 4 e + 3 = pairs 2 e, 2 e + 1
              const uint32_t x = ((word >> (4 * e)) & 15u) * 0x10204080u;
              uint32_t m0, m1;
              asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(m0) : "r"(x));
              asm("prmt.b32 %0, %1, %1, 0xBBAA;" : "=r"(m1) : "r"(x));
              h[2 * e] &= m0, h[2 * e + 1] &= m1;
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) sts_u4(dst + ((j ^ ((lane >> 1) & 3)) << 4), make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]));
        } else {  // phase 1: TMEM -> (+bias, ReLU) -> the chunk's bf16 rows in shared memory
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride + c0, v);
          const uint32_t dst = my_stage_u32 + lane * kEpiPitch;
          uint32_t bits = 0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 b0 = lds_f4(s_bias_u32 + (c0 + 8 * j) * 4), b1 = lds_f4(s_bias_u32 + (c0 + 8 * j + 4) * 4);
            float f[8];
            f[0] = fmaxf(__uint_as_float(v[8 * j + 0]) + b0.x, lo), f[1] = fmaxf(__uint_as_float(v[8 * j + 1]) + b0.y, lo);
            f[2] = fmaxf(__uint_as_float(v[8 * j + 2]) + b0.z, lo), f[3] = fmaxf(__uint_as_float(v[8 * j + 3]) + b0.w, lo);
            f[4] = fmaxf(__uint_as_float(v[8 * j + 4]) + b1.x, lo), f[5] = fmaxf(__uint_as_float(v[8 * j + 5]) + b1.y, lo);
            f[6] = fmaxf(__uint_as_float(v[8 * j + 6]) + b1.z, lo), f[7] = fmaxf(__uint_as_float(v[8 * j + 7]) + b1.w, lo);
            if (p.bits_out != nullptr) {  // ReLU output (>= +0): positive <=> its bit pattern, negated as an integer, has the sign bit set;
#pragma unroll                            // one funnel shift per value collects the flags (first value in the top bit: reversed below)
              for (int e = 0; e < 8; ++e) bits = __funnelshift_l(0u - __float_as_uint(f[e]), bits, 1);
            }
            __nv_bfloat162 h[4];
            h[0] = __floats2bfloat162_rn(f[0], f[1]), h[1] = __floats2bfloat162_rn(f[2], f[3]);
            h[2] = __floats2bfloat162_rn(f[4], f[5]), h[3] = __floats2bfloat162_rn(f[6], f[7]);
            sts_u4(dst + ((j ^ ((lane >> 1) & 3)) << 4), *reinterpret_cast<uint4*>(h));
          }
          if (p.bits_out != nullptr && my_out_off >= 0) p.bits_out[(my_out_off >> 5) + (c0 >> 5)] = __brev(bits);
        }
        if (c0 + 32 >= BN) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the accumulator is drained ...
        __syncwarp();
        if (c0 + 32 >= BN && elect_one())                                                      // ... hand it back before the global stores
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xa::smem_u32(acc_empty + acc)) : "memory");   // (lane == 0 made ptxas re-read %tid per chunk)
        // phase 2: coalesced mask application and stores of the chunk.  Element offset of my piece inside the output row:
        int col_delta = c0 + piece * 8;
        if (p.out_mode == 2) {  // (dy, dx, c) channel blocks land on pixels (2y+dy, 2x+dx): see xa_conv2d_nhwc_bf16_ex
          constexpr int kN4 = BN / 4;
          const int sub = col_delta / kN4;
          col_delta = ((sub >> 1) * p.PW + (sub & 1)) * kN4 + (col_delta - sub * kN4);
        }
        if (plain) {  // masked in phase 1: all eight shared-memory loads first (four dependent load -> test -> load -> store chains were
                      // half of this kernel's stall samples), then the stores
          int64_t off[4];
          uint4 val[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) off[it] = lds_i64(row_out_u32 + (it * 8 + row0) * 8), val[it] = lds_u4(my_piece_u32 + it * (8 * kEpiPitch));
#pragma unroll
          for (int it = 0; it < 4; ++it)
            if (off[it] >= 0) *reinterpret_cast<uint4*>(p.y + off[it] + col_delta) = val[it];
        } else
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = it * 8 + row0;
          const int64_t off0 = lds_i64(row_out_u32 + row * 8);
          if (off0 >= 0) {
            uint4 val = lds_u4(my_piece_u32 + it * (8 * kEpiPitch));
            if (p.bits_in != nullptr) {  // ReLU derivative from the mask bits: one byte = this piece's eight values
              uint32_t b;
              asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(row_bits_u32 + row * (kWords * 4) + (c0 >> 3) + piece));
              const uint4 mk = lds_u4(table_u32 + b * 16);
              val.x &= mk.x, val.y &= mk.y, val.z &= mk.z, val.w &= mk.w;
            } else if (p.mask != nullptr) {  // ... or from the bf16 activation of the layer below: zero where it was <= 0
              const uint4 mraw = __ldg(reinterpret_cast<const uint4*>(p.mask + lds_i64(row_msk_u32 + row * 8) + c0) + piece);
              const __nv_bfloat162* mk = reinterpret_cast<const __nv_bfloat162*>(&mraw);
              const __nv_bfloat162 zero = __floats2bfloat162_rn(0.0f, 0.0f);
              val.x &= __hgt2_mask(mk[0], zero);
              val.y &= __hgt2_mask(mk[1], zero);
              val.z &= __hgt2_mask(mk[2], zero);
              val.w &= __hgt2_mask(mk[3], zero);
            }
            *reinterpret_cast<uint4*>(p.y + off0 + col_delta) = val;
          }
        }
        __syncwarp();  // the chunk's staging rows are free again
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

template <int BN, bool kU8 = false>
int launch_flat(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mx1, const FlatParams& p, size_t smem, cudaStream_t stream,
                const char* what) {
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_flat_kernel<BN, kU8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_dev = dev;
  }
  const int64_t tiles = (p.Q + kBlockM - 1) / kBlockM;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  xa::launch_chained(tiles <= 16 * sms ? xa::kChainSmall : xa::kChainLarge, conv_flat_kernel<BN, kU8>, dim3(static_cast<unsigned>(tiles < sms ? tiles : sms)), dim3(flat_threads(BN, kU8)), smem, stream, mx, mw, mx1, p);
  return xa::check_launch(what);
}

}  // namespace

// Returns XA_OK after launching, or 1 when the shape does not fit this kernel (the caller falls back to conv_tc.cu's).
int xa_conv_flat_try(const void* x, const void* w, const float* bias, void* y, int batch, int height, int width, int channels, int kh, int kw,
                     int n_out, int pad_y, int pad_x, int relu, int out_mode, const void* relu_mask, int mask_is_bits, uint32_t* relu_bits_out,
                     int OH, int OW, int PH, int PW, xa_stream_t stream) {
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
  // the epilogue warps' transposing tiles + row tables, the bit-expansion table
  const int64_t epi_bytes = 4 * groups_bf16(n_out) * epi_warp_bytes(n_out) + (mask_is_bits ? 4096 : 0);
  const int64_t budget = 227 * 1024 - 1024 /*alignment*/ - 2048 /*barriers, tables, bias*/ - epi_bytes - p.w_bytes;
  int stages = static_cast<int>(budget / p.stage_bytes);
  if (stages < 2) return 1;
  if (stages > 6) stages = 6;
  p.y = static_cast<__nv_bfloat16*>(y), p.bias = bias;
  p.mask = mask_is_bits ? nullptr : static_cast<const __nv_bfloat16*>(relu_mask);
  p.bits_in = mask_is_bits ? static_cast<const uint32_t*>(relu_mask) : nullptr;
  p.bits_out = relu_bits_out;
  XA_REQUIRE(relu_bits_out == nullptr || (relu && (out_mode == 1 || (out_mode == 0 && PH == OH && PW == OW))), XA_EINVAL,
             "%s: ReLU mask bits are written by ReLU layers with a compact output", what);
  p.B = batch, p.H = height, p.W = width, p.N = n_out, p.OH = OH, p.OW = OW, p.PH = PH, p.PW = PW;
  p.relu = relu, p.out_mode = out_mode;
  p.n_entries = n_entries, p.kc_blocks = kc_blocks, p.win_rows = win_rows, p.stages = stages, p.Q = Q;
  p.min_shift = -(pad_y * width + pad_x);
  p.div_w_magic = (65536u + width - 1) / width;  // (rem * magic) >> 16 == rem / W for every pixel position of an image, else 0
  for (uint32_t g = 0; g < static_cast<uint32_t>(height * width) && p.div_w_magic != 0; ++g)
    if (g * static_cast<uint64_t>(p.div_w_magic) >= (uint64_t(1) << 32) || ((g * p.div_w_magic) >> 16) != g / width) p.div_w_magic = 0;
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
  if (n_out == 128) return launch_flat<128>(mx, mw, mx, p, smem, s, what);
  if (n_out == 64) return launch_flat<64>(mx, mw, mx, p, smem, s, what);
  return launch_flat<32>(mx, mw, mx, p, smem, s, what);
}

// First layer straight from the uint8 frames (see conv_flat_kernel's kU8 notes): frames [B, height, width, 4] uint8, a
// kh x kw stride-1 kernel over the 4x4 space-to-depth grid (= a 4kh x 4kw / 4 convolution of the frames), w [n_out, kh*kw*64]
// bf16 with K ordered (kh, kw, dy, dx, c), x/255 applied on the way in.
static int conv_u8_s2d(const char* what, const uint8_t* frames, const int32_t* frame_idx, int64_t n_frames, int idx_n_steps, int idx_n_envs,
                       const void* w, const float* bias, void* y, void* x_s2d_out, uint32_t* relu_bits_out, int batch, int height, int width,
                       int kh, int kw, int n_out, int relu, int out_s2d, xa_stream_t stream) {
  XA_REQUIRE(frames && w && y, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(batch > 0 && height > 0 && width > 0 && height % 4 == 0 && width % 4 == 0 && kh > 0 && kw > 0, XA_EINVAL,
             "%s: batch=%d frames %dx%d kernel %dx%d", what, batch, height, width, kh, kw);
  XA_REQUIRE(n_out == 32 || n_out == 64, XA_EINVAL, "%s: n_out=%d (32 or 64)", what, n_out);
  XA_REQUIRE(xa::aligned(frames, 16) && xa::aligned(w, 16) && xa::aligned(y, 16) && xa::aligned(x_s2d_out, 16), XA_EALIGN,
             "%s: 16-byte alignment required", what);
  const int H = height / 4, W = width / 4;
  const int OH = H - kh + 1, OW = W - kw + 1;
  XA_REQUIRE(OH > 0 && OW > 0 && width <= 256 && (!out_s2d || (OH % 2 == 0 && OW % 2 == 0)), XA_EINVAL, "%s: output %dx%d", what, OH, OW);
  const int n_entries = kh * kw;
  const int win_rows = ((kBlockM + (kh - 1) * W + (kw - 1) + 7) / 8) * 8;
  const int box_rows = (W - 1 + win_rows + W - 1) / W;  // whole grid rows that can hold a window starting anywhere in a row
  XA_REQUIRE(n_entries <= kMaxEntries && win_rows <= 256 && 4 * box_rows <= 256, XA_EINVAL, "%s: kernel %dx%d on a %d-wide grid does not fit", what,
             kh, kw, W);
  const int64_t Q = static_cast<int64_t>(batch) * H * W;
  XA_REQUIRE(Q < (int64_t(1) << 31) - 4096, XA_EOVERFLOW, "%s: batch too large", what);
  FlatParams p{};
  p.w_bytes = static_cast<uint32_t>(n_entries) * n_out * 128u;
  p.stage_bytes = static_cast<uint32_t>(win_rows) * 128u;
  p.raw_tx_bytes = static_cast<uint32_t>(box_rows) * W * 64u;
  p.raw_stage_bytes = ((p.raw_tx_bytes + 1023u) / 1024u) * 1024u;
  const int64_t epi_bytes = 4 * kGroupsU8 * epi_warp_bytes(n_out);
  const int64_t budget = 227 * 1024 - 1024 - 2048 - epi_bytes - p.w_bytes - static_cast<int64_t>(kRawStages) * p.raw_stage_bytes;
  int stages = static_cast<int>(budget / p.stage_bytes);
  XA_REQUIRE(stages >= 2, XA_EINVAL, "%s: shared memory does not hold two operand stages", what);
  if (stages > 6) stages = 6;
  p.y = static_cast<__nv_bfloat16*>(y), p.bias = bias, p.mask = nullptr;
  XA_REQUIRE(relu_bits_out == nullptr || (relu && xa::aligned(relu_bits_out, 16)), XA_EINVAL, "%s: ReLU mask bits need relu = 1 and a 16-byte aligned buffer", what);
  p.bits_out = relu_bits_out, p.bits_in = nullptr;
  p.B = batch, p.H = H, p.W = W, p.N = n_out, p.OH = OH, p.OW = OW, p.PH = OH, p.PW = OW;
  p.relu = relu, p.out_mode = out_s2d ? 1 : 0;
  p.n_entries = n_entries, p.kc_blocks = 1, p.win_rows = win_rows, p.stages = stages, p.Q = Q, p.min_shift = 0;
  int e = 0;
  for (int i = 0; i < kh; ++i)
    for (int j = 0; j < kw; ++j, ++e) {
      p.a_units[e] = (static_cast<uint32_t>(i * W + j) * 128u) >> 4;
      p.b_units[e] = (static_cast<uint32_t>(e) * n_out * 128u) >> 4;
    }
  if (frame_idx != nullptr) {
    XA_REQUIRE(n_frames > 0 && box_rows <= H && xa::aligned(frame_idx, 4), XA_EINVAL, "%s: indexed frames need n_frames > 0 and a window within two frames", what);
    XA_REQUIRE(idx_n_steps <= 0 || static_cast<int64_t>(idx_n_steps) * idx_n_envs == n_frames, XA_EINVAL,
               "%s: n_steps * n_envs = %lld is not the number of stored frames %lld", what,
               static_cast<long long>(idx_n_steps) * idx_n_envs, static_cast<long long>(n_frames));
    p.frames_base = frames, p.frame_idx = frame_idx, p.idx_T = idx_n_steps > 0 ? idx_n_steps : 0, p.idx_E = idx_n_envs;
    p.box_rows = box_rows, p.grid_row_bytes = static_cast<uint32_t>(W) * 64u;
  }
  p.div_w_magic = (65536u + W - 1) / W;
  for (uint32_t g = 0; g < static_cast<uint32_t>((box_rows > H ? box_rows : H) * W); ++g)
    XA_REQUIRE(((g * p.div_w_magic) >> 16) == g / W, XA_EINVAL, "%s: no 16-bit reciprocal for a grid %d wide", what, W);
  CUtensorMap mx, mw;
  {
    EncodeTiledFn fn = encode_fn();
    XA_REQUIRE(fn != nullptr, XA_EINVAL, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
    // the frames as image rows of 32-bit words: [batch * height rows, width words]; a box = 4 * box_rows whole rows
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(frame_idx != nullptr ? n_frames : batch) * height};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(width) * 4};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(width), static_cast<cuuint32_t>(4 * box_rows)};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult r = fn(&mx, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(frames), dims, strides, box, elem,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    XA_REQUIRE(r == CUDA_SUCCESS, XA_EINVAL, "%s: cuTensorMapEncodeTiled(frames) failed with %d", what, static_cast<int>(r));
  }
  if (int rc = make_map_2d(&mw, w, n_out, static_cast<int64_t>(kh) * kw * 64, n_out, what)) return rc;
  CUtensorMap mx1 = mx;
  p.store_x1 = x_s2d_out != nullptr;
  if (p.store_x1) {
    XA_REQUIRE(stages >= 3, XA_EINVAL, "%s: storing the space-to-depth tensor needs three operand stages", what);
    if (int rc = make_map_2d_box(&mx1, x_s2d_out, Q, 64, 32, 64, what)) return rc;
  }
  const size_t smem = 1024 + p.w_bytes + static_cast<size_t>(stages) * p.stage_bytes + static_cast<size_t>(kRawStages) * p.raw_stage_bytes + 2048 +
                      epi_bytes;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return n_out == 64 ? launch_flat<64, true>(mx, mw, mx1, p, smem, s, what) : launch_flat<32, true>(mx, mw, mx1, p, smem, s, what);
}

extern "C" int xa_conv2d_u8_s2d_bf16(const uint8_t* frames, const void* w, const float* bias, void* y, void* x_s2d_out, int batch, int height,
                                     int width, int kh, int kw, int n_out, int relu, int out_s2d, xa_stream_t stream) {
  return conv_u8_s2d("xa_conv2d_u8_s2d_bf16", frames, nullptr, 0, 0, 0, w, bias, y, x_s2d_out, nullptr, batch, height, width, kh, kw, n_out, relu,
                     out_s2d, stream);
}

// The same layer with its options spelled out.  frame_idx != NULL: the `batch` frames are read THROUGH A PERMUTATION of a larger
// frame store: frame f = row frame_idx[f] of frames [n_frames, height, width, 4] (with n_steps > 0 the ids are env-major sample
// ids of a time-major [n_steps, n_envs] rollout, xagents/base.py:559-564) -- get_mini_batches' tf.gather of the states
// (ppo/agent.py:139-155) folded into the network's first layer: the gathered minibatch is never written to or read back from
// HBM.  relu_bits_out != NULL: one bit per output element (> 0), the ReLU derivative the data gradient of the next layer needs.
extern "C" int xa_conv2d_u8_s2d_bf16_ex(const uint8_t* frames, int64_t n_frames, const int32_t* frame_idx, int n_steps, int n_envs,
                                        const void* w, const float* bias, void* y, void* x_s2d_out, uint32_t* relu_bits_out, int batch,
                                        int height, int width, int kh, int kw, int n_out, int relu, int out_s2d, xa_stream_t stream) {
  return conv_u8_s2d("xa_conv2d_u8_s2d_bf16_ex", frames, frame_idx, n_frames, n_steps, n_envs, w, bias, y, x_s2d_out, relu_bits_out, batch, height,
                     width, kh, kw, n_out, relu, out_s2d, stream);
}

extern "C" int xa_conv2d_u8_s2d_bf16_indexed(const uint8_t* frames, int64_t n_frames, const int32_t* frame_idx, int n_steps, int n_envs,
                                             const void* w, const float* bias, void* y, void* x_s2d_out, int batch, int height, int width,
                                             int kh, int kw, int n_out, int relu, int out_s2d, xa_stream_t stream) {
  XA_REQUIRE(frame_idx != nullptr, XA_EINVAL, "xa_conv2d_u8_s2d_bf16_indexed: null frame_idx");
  return conv_u8_s2d("xa_conv2d_u8_s2d_bf16_indexed", frames, frame_idx, n_frames, n_steps, n_envs, w, bias, y, x_s2d_out, nullptr, batch, height,
                     width, kh, kw, n_out, relu, out_s2d, stream);
}
