// bf16 GEMM on the 5th-generation tensor cores: C[M,N] = A[M,K] * B[N,K]^T (+ bias, ReLU), fp32 accumulate.
//
// Row "next" (SURVEY.md 8f-1): the policy/value network's dense contractions -- FC 3136->512 and the actor /
// critic heads built by ModelReader from ppo/models/cnn-actor-critic.cfg:30-42 (xagents/utils/common.py:239-258),
// forward and both backward products -- which the reference leaves to Keras/Eigen.  Both operands are
// K-major (x and W of a Dense layer as they sit in memory), so y = x W^T needs no transposition.
//
// Warp-specialised and persistent: one CTA per SM walks 128 x BN output tiles (M fastest, so CTAs running
// together share the same weight tile in L2); the accumulator is double-buffered in TMEM so the epilogue of
// tile i drains while the tensor core already works on tile i+1.  BN = 256 for wide outputs: a 128 x 128
// tile needs 32 KB of operands per 256 MMA cycles (~128 B/cycle/SM, more than L2 can feed 148 SMs);
// 128 x 256 cuts that by a third.
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles (64 bf16 = one 128-B swizzle row per K block) of A
//            and B into a ring of shared-memory stages, completion counted on "full" mbarriers;
//   warp 1   allocates 2 x BN TMEM columns, then one elected lane issues tcgen05.mma (UMMA 128 x BN x 16,
//            operands read from shared memory through SWIZZLE_128B descriptors, accumulator in TMEM);
//            tcgen05.commit releases each stage ("empty" mbarrier) and signals the finished accumulator;
//   warps 2-5 epilogue: tcgen05.ld their TMEM lane quadrant (32 rows x 32 columns per instruction), add
//            bias, apply ReLU, convert, store, then hand the accumulator back ("acc_empty").
// SASS carries UTCHMMA / UTMALDG / LDTM (B200_PROFILING.md evidence table).
#include <stdlib.h>

#include "tc_common.cuh"

namespace {

using namespace xa_tc;

struct GemmParams {
  void* c;
  const float* bias;
  int64_t m, n, k, ldc;
  int relu;
  int splits;                  // > 1: split-K, fp32 partial tiles go to `partial` [splits, m, n]
  int kb_per_split;
  float* partial;
  const __nv_bfloat16* mask;   // optional [m, mask_ld]: result *= (mask > 0)  (ReLU derivative in a backward product)
  int64_t mask_ld;
  int64_t col_group, col_group_pitch;  // col_group > 0: output column j lands at (j / col_group) * col_group_pitch + j % col_group
  // bf16 output through TMA stores (tma_store != 0; map_c covers c as [m, ldc] in 32 x 32 boxes, SWIZZLE_64B): the epilogue's
  // thread = row registers go to shared memory once and ONE instruction per warp and chunk writes the 32 x 32 box -- no
  // transpose, no per-lane address arithmetic, no per-lane stores.  The ReLU-derivative mask is then a BIT mask (bit j of
  // word i <=> element 32 i + j of the [m, mask_ld] activation is > 0): one word per row and chunk.
  const uint32_t* mask_bits;
  int64_t mask_words;  // words per row of mask_bits
  int tma_store;
};

template <int BN, bool kPair = false>
struct Smem {
  static constexpr int kStageA = kBlockM * kBlockK * 2;
  static constexpr int kStageB = (kPair ? BN / 2 : BN) * kBlockK * 2;   // a CTA pair splits B's rows between its two CTAs
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (208 * 1024) / kStage > 8 ? 8 : (208 * 1024) / kStage;
  static constexpr int kEpiPitch = 32 * 2 + 16;  // bf16 epilogue: a warp's 32 x 32 chunk, rows padded against bank conflicts
  static constexpr int kEpiWarp = 32 * kEpiPitch;  // 2560 B per warp = 5 x 512: also holds the 2 KB, 512-byte-aligned box of the TMA-store path
  static constexpr int kEpiBytes = 8 * kEpiWarp;
  static constexpr int kTableBytes = 4096;       // byte of mask bits -> four bf16x2 AND masks
  static constexpr int kBytes = kStages * kStage + 1024 /*alignment slack*/ + kEpiBytes + kTableBytes + 256 /*barriers*/;
};

// 2 + 8 warps: TMA producer, MMA issuer, two epilogue groups of four warps; group g drains accumulator g (every other
// tile of the CTA), so two epilogues are in flight per SM and their latencies (tcgen05.ld, mask loads) overlap.
constexpr int kGemmThreads = 64 + 8 * 32;

// kPair: launched as clusters of two CTAs (the two SMs of a TPC) that run ONE 256 x BN tile with cta_group::2 MMAs: each CTA
// loads its 128 rows of A and HALF of the B tile, so a stage is 32 KB instead of 48 KB per 128 x 256 x 64 block of products --
// the FC products at 8192 frames were bound by L2 -> SM operand traffic (every CTA pulled 48 KB per 512 MMA cycles, ~28 TB/s
// over 148 SMs).  The leader (cluster rank 0) issues the MMAs and owns the `full` / `acc_empty` barriers; tcgen05.commit
// multicasts `empty` / `acc_full` arrivals to both CTAs; each CTA drains its own 128 accumulator rows.
template <int BN, bool kOutBf16, bool kPair = false>
__global__ void __launch_bounds__(kGemmThreads) gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                 const __grid_constant__ CUtensorMap map_b,
                                                                 const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
  using S = Smem<BN, kPair>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + S::kStages * S::kStage;  // [8 warps][32 rows][kEpiPitch]; 1024-byte aligned (stages are multiples of 1 KB)
  uint32_t* bit_table = reinterpret_cast<uint32_t*>(epi_stage + S::kEpiBytes);   // [256][4]
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_stage + S::kEpiBytes + S::kTableBytes);
  uint64_t* empty = full + S::kStages;
  uint64_t* acc_full = empty + S::kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  xa::pdl_trigger();   // chained launch (xa_common.cuh): the on-chip set-up below overlaps the predecessor's tail
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int k_blocks = static_cast<int>((p.k + kBlockK - 1) / kBlockK);
  constexpr int kTileM = kPair ? 2 * kBlockM : kBlockM;   // rows of one work item (a pair's item is 256 rows, 128 per CTA)
  const int tiles_m = static_cast<int>((p.m + kTileM - 1) / kTileM);
  const int tiles_n = static_cast<int>((p.n + BN - 1) / BN);
  const int n_tiles = tiles_m * tiles_n;
  const int n_items = n_tiles * p.splits;  // (tile, K split) pairs; tile varies fastest
  const uint32_t rank = kPair ? cluster_rank() : 0u;
  const int unit = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);       // this CTA's (pair's) first item ...
  const int n_units = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);       // ... and the stride between its items
  constexpr uint32_t kAccStride = BN < 32 ? 32 : BN;  // the epilogue reads 32 columns at a time
  constexpr uint32_t kTmemCols = 2 * kAccStride;      // two accumulators

  if (kOutBf16 && p.mask_bits != nullptr) {  // byte b of a row's mask bits -> the AND masks of its eight bf16 values
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
      const int b = i >> 2, k = i & 3;
      bit_table[i] = (((b >> (2 * k)) & 1) ? 0x0000FFFFu : 0u) | (((b >> (2 * k + 1)) & 1) ? 0xFFFF0000u : 0u);
    }
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if (kOutBf16 && p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    for (int s = 0; s < S::kStages; ++s) {
      xa::mbar_init(full + s, 1);
      xa::mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      xa::mbar_init(acc_full + a, 1);
      xa::mbar_init(acc_empty + a, kPair ? 8 : 4);  // one arrival per epilogue warp (of both CTAs of a pair, on the leader's barrier)
    }
    xa::fence_barrier_init();
  }
  if (warp == 1) {  // TMEM allocation is warp-wide; the same warp frees it at the end
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xa::smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xa::smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (kPair) cluster_sync_all();   // the peer's barriers must be initialised before anything arrives on them
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  xa::pdl_wait();      // first global-memory access below: the predecessor grid has completed

  if (warp == 0) {
    if (elect_one()) {  // ---- TMA producer: one ring of stages shared by all of this CTA's tiles
      int s = 0;
      uint32_t round = 0;
      for (int item = unit; item < n_items; item += n_units) {
        const int tile = item % n_tiles, split = item / n_tiles;
        const int tile_m = tile % tiles_m, tile_n = tile / tiles_m;
        const int kb0 = split * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (round > 0) mbar_wait_wd(empty + s, (round - 1) & 1);
          uint8_t* a_dst = smem + s * S::kStage;
          uint8_t* b_dst = a_dst + S::kStageA;
          if (kPair) {  // the leader's barrier counts the bytes of both CTAs' boxes; each CTA loads its rows of A and its half of B
            if (rank == 0) xa::mbar_expect_tx(full + s, 2 * S::kStage);
            tma_load_2d_pair(a_dst, &map_a, kb * kBlockK, tile_m * kTileM + static_cast<int>(rank) * kBlockM, full + s);
            tma_load_2d_pair(b_dst, &map_b, kb * kBlockK, tile_n * BN + static_cast<int>(rank) * (BN / 2), full + s);
          } else {
            xa::mbar_expect_tx(full + s, S::kStage);
            tma_load_2d(a_dst, &map_a, kb * kBlockK, tile_m * kBlockM, full + s);
            tma_load_2d(b_dst, &map_b, kb * kBlockK, tile_n * BN, full + s);
          }
          if (++s == S::kStages) s = 0, ++round;
        }
      }
    }
  } else if (warp == 1) {
    if ((!kPair || rank == 0) && elect_one()) {  // ---- MMA issuer (a pair's leader only)
      constexpr uint32_t idesc = make_idesc(kTileM, BN);
      uint32_t lt = 0, phase = 0;
      int s = 0;
      const uint64_t desc0 = make_smem_desc(smem);  // stage 0; stage s adds s * kStage to the address field
      for (int item = unit; item < n_items; item += n_units, ++lt) {
        const int split = item / n_tiles;
        const int kb0 = split * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
        const uint32_t acc = lt & 1, use = lt >> 1;
        if (use > 0) {  // the epilogue must have drained this accumulator
          mbar_wait_wd(acc_empty + acc, (use - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_wd(full + s, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = desc0 + static_cast<uint64_t>(s * (S::kStage >> 4));
          const uint64_t db = da + (S::kStageA >> 4);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advance 16 bf16 = 32 B inside the 128-B swizzle row: +2 in the (address >> 4) field
            if (kPair) umma_bf16_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0) || (k != 0));
            else umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0) || (k != 0));
          }
          if (kPair) umma_commit_pair(empty + s);   // both CTAs' producers may refill the stage once these MMAs have read it
          else umma_commit(empty + s);
          if (++s == S::kStages) s = 0, phase ^= 1;
        }
        if (kPair) umma_commit_pair(acc_full + acc);  // accumulator complete: both CTAs' epilogue groups
        else umma_commit(acc_full + acc);
      }
    }
  } else {
    // ---- epilogue: warp w may only touch TMEM lanes [32*(w%4), +32); group (w-2)/4 owns accumulator (w-2)/4
    const int quad = warp & 3;
    const uint32_t grp = (warp - 2) >> 2;
    const bool vec_ok = (p.ldc % (kOutBf16 ? 8 : 4)) == 0 && xa::aligned(p.c, 16) && (p.col_group % 8) == 0 && (p.col_group_pitch % 8) == 0;
    const bool bias_vec = xa::aligned(p.bias, 16);
    constexpr int kChunks = BN < 32 ? 1 : BN / 32;
    for (uint32_t lt = grp;; lt += 2) {
      const int item = unit + static_cast<int>(lt) * n_units;
      if (item >= n_items) break;
      const int tile = item % n_tiles, split = item / n_tiles;
      const int tile_n = tile / tiles_m;
      const int tile_m = kPair ? 2 * (tile % tiles_m) + static_cast<int>(rank) : tile % tiles_m;   // in units of this CTA's 128 rows
      const uint32_t acc = grp;
      const int64_t row = static_cast<int64_t>(tile_m) * kBlockM + quad * 32 + lane;
      // The ReLU-derivative mask does not depend on the accumulator: chunks 0 and 1 are fetched before the wait, chunk
      // ci + 2 while chunk ci is processed (three rotating register sets; the chunk loop stays rolled -- unrolled, the
      // eight epilogue warps run ~200 KB of code and their instruction-cache misses stall the producer / issuer warps).
      const int64_t colt = static_cast<int64_t>(tile_n) * BN;
      const bool mvec = p.mask != nullptr && p.splits == 1 && row < p.m && (p.mask_ld % 8) == 0 && xa::aligned(p.mask, 16);
      const uint4* mrow = reinterpret_cast<const uint4*>(p.mask + (mvec ? row * p.mask_ld + colt : 0));
      uint4 m0[4], m1[4], m2[4];
      auto load_mask = [&](uint4 (&m)[4], int chunk) {
        if (!kOutBf16 && mvec && chunk < kChunks && colt + chunk * 32 + 32 <= p.n) {
#pragma unroll
          for (int q = 0; q < 4; ++q) m[q] = __ldg(mrow + chunk * 4 + q);
        }
      };
      load_mask(m0, 0);
      load_mask(m1, 1);
      // bf16 output: the mask is fetched in the layout of the transposed stores (lane -> row it*8 + lane/4, 16-byte piece
      // lane%4 of the chunk's 64 bytes), one chunk ahead
      const bool fast_mask = kOutBf16 && p.mask != nullptr && p.splits == 1 && (p.mask_ld % 8) == 0 && xa::aligned(p.mask, 16);
      uint4 cmask[4], cnext[4];
      auto load_cmask = [&](uint4 (&m)[4], int chunk) {
        if (fast_mask && chunk < kChunks && colt + chunk * 32 + 32 <= p.n) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int64_t grow = static_cast<int64_t>(tile_m) * kBlockM + quad * 32 + it * 8 + (lane >> 2);
            if (grow < p.m) m[it] = __ldg(reinterpret_cast<const uint4*>(p.mask + grow * p.mask_ld + colt + chunk * 32) + (lane & 3));
          }
        }
      };
      load_cmask(cmask, 0);
      mbar_wait_backoff(acc_full + acc, (lt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ++ci) {
        const int c0 = ci * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccStride + c0, v);
        load_mask(m2, ci + 2);
        load_cmask(cnext, ci + 1);
        const int64_t col0 = colt + c0;
        if (p.splits > 1) {  // raw fp32 partial tile; bias / ReLU / conversion happen in the reduction pass
          if (row < p.m && col0 < p.n) {
            float* dstp = p.partial + (static_cast<int64_t>(split) * p.m + row) * p.n + col0;
            if (col0 + 32 <= p.n && (p.n % 4) == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(dstp)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            } else {
              for (int j = 0; j < 32 && col0 + j < p.n; ++j) dstp[j] = __uint_as_float(v[j]);
            }
          }
        } else if (kOutBf16 && p.tma_store && col0 + 32 <= p.n && bias_vec && p.mask == nullptr) {  // warp-uniform
          // TMA-store path: thread = row all the way.  The short-K products are bound by the instruction count of this epilogue
          // (ncu: the FC data gradient ran 12 400 warp-instructions per 128 x 256 tile against 4096 MMA cycles, eight epilogue
          // warps 74 % busy); staged through a shared-memory transpose a chunk cost ~190 instructions per warp, here ~70.
          const int64_t ocol0 = p.col_group > 0 ? (col0 / p.col_group) * p.col_group_pitch + col0 % p.col_group : col0;
          uint32_t h[16];
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b = __ldg(bp + q);
              float x0 = __uint_as_float(v[4 * q]) + b.x, x1 = __uint_as_float(v[4 * q + 1]) + b.y;
              float x2 = __uint_as_float(v[4 * q + 2]) + b.z, x3 = __uint_as_float(v[4 * q + 3]) + b.w;
              if (p.relu) x0 = fmaxf(x0, 0.0f), x1 = fmaxf(x1, 0.0f), x2 = fmaxf(x2, 0.0f), x3 = fmaxf(x3, 0.0f);
              const __nv_bfloat162 lo2 = __floats2bfloat162_rn(x0, x1), hi2 = __floats2bfloat162_rn(x2, x3);
              h[2 * q] = *reinterpret_cast<const uint32_t*>(&lo2), h[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&hi2);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              float x0 = __uint_as_float(v[2 * q]), x1 = __uint_as_float(v[2 * q + 1]);
              if (p.relu) x0 = fmaxf(x0, 0.0f), x1 = fmaxf(x1, 0.0f);
              const __nv_bfloat162 t = __floats2bfloat162_rn(x0, x1);
              h[q] = *reinterpret_cast<const uint32_t*>(&t);
            }
          }
          if (p.mask_bits != nullptr) {  // ReLU derivative of the layer below: my row's 32 bits of this chunk
            const uint32_t bits = row < p.m ? __ldg(p.mask_bits + row * p.mask_words + (col0 >> 5)) : 0u;
            const uint32_t tb = xa::smem_u32(bit_table);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 mk;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(mk.x), "=r"(mk.y), "=r"(mk.z), "=r"(mk.w) : "r"(tb + ((bits >> (8 * q)) & 255u) * 16u));
              h[4 * q] &= mk.x, h[4 * q + 1] &= mk.y, h[4 * q + 2] &= mk.z, h[4 * q + 3] &= mk.w;
            }
          }
          const uint32_t st = xa::smem_u32(epi_stage + (warp - 2) * S::kEpiWarp);   // this warp's 32 x 64-byte box (SWIZZLE_64B)
          if (lane == 0) xa::bulk_wait_read<0>();   // the previous chunk's store has read the box
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)   // 16-byte piece j of my row sits at piece j ^ ((row >> 1) & 3): conflict-free, and the tensor map's swizzle
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(st + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)), "r"(h[4 * j]), "r"(h[4 * j + 1]),
                         "r"(h[4 * j + 2]), "r"(h[4 * j + 3])
                         : "memory");
          xa::fence_proxy_async();   // my writes -> visible to the TMA unit
          __syncwarp();
          if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_c), "r"(static_cast<int>(ocol0)),
                         "r"(tile_m * kBlockM + quad * 32), "r"(st)
                         : "memory");
            xa::bulk_commit();
          }
        } else if ((kOutBf16 || row < p.m) && col0 + 32 <= p.n && vec_ok && bias_vec && (p.mask == nullptr || fast_mask)) {  // warp-uniform for bf16
          // Fast path (a full chunk of 32 columns, vector-aligned): the epilogue of the short-K products is bounded by its
          // instruction count, so no per-element flag tests here -- bias by vector loads, ReLU as one max, the mask as a
          // packed bf16 compare that yields a bit mask.
          const int64_t ocol0 = p.col_group > 0 ? (col0 / p.col_group) * p.col_group_pitch + col0 % p.col_group : col0;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias != nullptr) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 b = __ldg(bp + q);
              f[4 * q] += b.x, f[4 * q + 1] += b.y, f[4 * q + 2] += b.z, f[4 * q + 3] += b.w;
            }
          }
          const float lo = p.relu ? 0.0f : -INFINITY;
          if (kOutBf16) {
            // Thread = row out of TMEM; stored like that, one warp-wide 16-byte access touches 32 different lines, and those
            // load/store wavefronts (not HBM, not the tensor pipe) set the pace of the short-K products: 8 chunks x 8
            // accesses x 32 wavefronts per tile and warp against 4096 MMA cycles.  So the chunk goes through a padded
            // shared-memory tile: rows in, 16-byte pieces out with four consecutive lanes on one row's 64 bytes -- mask
            // loads and output stores then touch 8 rows x 64 contiguous bytes per access.
            uint8_t* my_stage = epi_stage + (warp - 2) * (32 * S::kEpiPitch);
            const uint32_t st = xa::smem_u32(my_stage);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162 h[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(fmaxf(f[8 * j + 2 * q], lo), fmaxf(f[8 * j + 2 * q + 1], lo));
              const uint4 o = *reinterpret_cast<uint4*>(h);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(st + lane * S::kEpiPitch + j * 16), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w)
                           : "memory");
            }
            __syncwarp();
            const int piece = lane & 3, r0 = lane >> 2;
            const __nv_bfloat162 zero = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int rr = it * 8 + r0;
              const int64_t grow = static_cast<int64_t>(tile_m) * kBlockM + quad * 32 + rr;
              uint4 o;
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(o.x), "=r"(o.y), "=r"(o.z), "=r"(o.w) : "r"(st + rr * S::kEpiPitch + piece * 16));
              if (grow < p.m) {
                if (p.mask != nullptr) {
                  const uint4 mk4 = cmask[it];
                  const __nv_bfloat162* mk = reinterpret_cast<const __nv_bfloat162*>(&mk4);
                  o.x &= __hgt2_mask(mk[0], zero), o.y &= __hgt2_mask(mk[1], zero);
                  o.z &= __hgt2_mask(mk[2], zero), o.w &= __hgt2_mask(mk[3], zero);
                }
                *(reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.c) + grow * p.ldc + ocol0) + piece) = o;
              }
            }
            __syncwarp();
          } else {
            float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.c) + row * p.ldc + ocol0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(fmaxf(f[4 * j], lo), fmaxf(f[4 * j + 1], lo), fmaxf(f[4 * j + 2], lo), fmaxf(f[4 * j + 3], lo));
          }
        } else if (row < p.m && col0 < p.n) {
          float f[32];
          // grouped output columns (a [B, h*w*c] gradient written onto a zero-bordered [B, H, W, c] grid)
          const int64_t ocol0 = p.col_group > 0 ? (col0 / p.col_group) * p.col_group_pitch + col0 % p.col_group : col0;
          const bool mv_ok = !kOutBf16 && mvec && col0 + 32 <= p.n;
          const __nv_bfloat16* mv = reinterpret_cast<const __nv_bfloat16*>(m0);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(v[j]);
            if (p.bias != nullptr && col0 + j < p.n) x += __ldg(p.bias + col0 + j);
            if (p.relu) x = fmaxf(x, 0.0f);
            if (p.mask != nullptr && col0 + j < p.n) {
              const float mval = mv_ok ? __bfloat162float(mv[j]) : __bfloat162float(p.mask[row * p.mask_ld + col0 + j]);
              if (!(mval > 0.0f)) x = 0.0f;
            }
            f[j] = x;
          }
          if (vec_ok && col0 + 32 <= p.n) {
            if (kOutBf16) {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.c) + row * p.ldc + ocol0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[8 * j + 2 * q], f[8 * j + 2 * q + 1]);
                dst[j] = *reinterpret_cast<uint4*>(h);
              }
            } else {
              float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.c) + row * p.ldc + ocol0);
#pragma unroll
              for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < p.n; ++j) {
              if (kOutBf16)
                static_cast<__nv_bfloat16*>(p.c)[row * p.ldc + ocol0 + j] = __float2bfloat16_rn(f[j]);
              else
                static_cast<float*>(p.c)[row * p.ldc + ocol0 + j] = f[j];
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) m0[q] = m1[q], m1[q] = m2[q], cmask[q] = cnext[q];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (kPair) mbar_arrive_leader(acc_empty + acc);   // the issuing CTA waits for both CTAs' epilogues
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xa::smem_u32(acc_empty + acc)) : "memory");
      }
    }
  }
  if (kOutBf16 && p.tma_store && warp >= 2 && lane == 0) xa::bulk_wait_all<0>();   // shared memory must outlive the stores
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (kPair) cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other still arrives on its barriers / reads its operands
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host
}  // namespace
// XA_GEMM_TMA_STORE=0 keeps the bf16 epilogue on the shared-memory transpose + per-lane stores (A/B switch; read once)
bool xa::gemm_tma_store_enabled() {
  static const bool on = [] {
    const char* e = getenv("XA_GEMM_TMA_STORE");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}
namespace {
// split-K second pass: C = sum_s partial[s] (+ bias) (ReLU), in split order -> deterministic
template <bool kOutBf16>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmParams p) {
  const int64_t total = p.m * p.n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / p.n, col = i - row * p.n;
    float acc = 0.0f;
    for (int s = 0; s < p.splits; ++s) acc += p.partial[static_cast<int64_t>(s) * total + i];
    if (p.bias != nullptr) acc += p.bias[col];
    if (p.relu) acc = fmaxf(acc, 0.0f);
    if (kOutBf16)
      static_cast<__nv_bfloat16*>(p.c)[row * p.ldc + col] = __float2bfloat16_rn(acc);
    else
      static_cast<float*>(p.c)[row * p.ldc + col] = acc;
  }
}

// CTA pairs: clusters of two CTAs, at most one pair per TPC.  The launch attribute makes the hardware co-schedule the two
// CTAs of a pair on the two SMs of one TPC (what cta_group::2 needs).
template <int BN, bool kOutBf16>
int launch_pair(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const GemmParams& p, cudaStream_t stream, const char* what) {
  auto kernel = gemm_bf16_tn_kernel<BN, kOutBf16, true>;
  using S = Smem<BN, true>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
    if (e != cudaSuccess) {
      xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_dev = dev;
  }
  const int64_t items = ((p.m + 2 * kBlockM - 1) / (2 * kBlockM)) * ((p.n + BN - 1) / BN) * p.splits;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int64_t pairs = items < sms / 2 ? items : sms / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * pairs)), cfg.blockDim = dim3(kGemmThreads), cfg.dynamicSmemBytes = S::kBytes, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, ma, mb, mc, p);
  return xa::check_launch(what);
}

// XA_GEMM_PAIR=0 in the environment keeps every product on single CTAs (A/B switch; read once)
static bool gemm_pairs_enabled() {
  static const bool on = [] {
    const char* e = getenv("XA_GEMM_PAIR");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}


template <int BN, bool kOutBf16>
int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const GemmParams& p, cudaStream_t stream, const char* what,
           bool reduce = true) {
  auto kernel = gemm_bf16_tn_kernel<BN, kOutBf16>;
  static thread_local int configured_dev = -1;  // opt-in shared memory is a per-device attribute of the kernel
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<BN>::kBytes);
    if (e != cudaSuccess) {
      xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_dev = dev;
  }
  const int64_t items = ((p.m + kBlockM - 1) / kBlockM) * ((p.n + BN - 1) / BN) * p.splits;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);  // persistent: at most one CTA per SM
  xa::launch_chained(items * ((p.k + kBlockK - 1) / kBlockK) <= 16 * sms ? xa::kChainSmall : xa::kChainLarge, kernel, dim3(grid), dim3(kGemmThreads), Smem<BN>::kBytes, stream, ma, mb, mc, p);
  if (int rc = xa::check_launch(what)) return rc;
  if (p.splits > 1 && reduce) {
    const int64_t want = (p.m * p.n + 255) / 256;
    splitk_reduce_kernel<kOutBf16><<<static_cast<unsigned>(want < 4096 ? want : 4096), 256, 0, stream>>>(p);
    return xa::check_launch(what);
  }
  return XA_OK;
}

}  // namespace

static int gemm_tile_n(int64_t n) { return n >= 256 ? 256 : (n > 64 ? 128 : (n > 16 ? 64 : 16)); }

// K splits for shapes that would otherwise leave most SMs idle (few output tiles, long K: the weight-gradient
// products).  0 = no split.
static int gemm_auto_splits(int64_t m, int64_t n, int64_t k) {
  const int bn = gemm_tile_n(n);
  const int64_t tiles = ((m + kBlockM - 1) / kBlockM) * ((n + bn - 1) / bn);
  const int64_t k_blocks = (k + kBlockK - 1) / kBlockK;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  if (tiles * 2 > sms || k_blocks < 32) return 1;
  int64_t splits = sms / tiles;  // one wave: tiles * splits <= SMs
  if (splits > k_blocks / 8) splits = k_blocks / 8;
  return splits < 2 ? 1 : static_cast<int>(splits);
}

extern "C" int64_t xa_gemm_workspace_bytes(int64_t m, int64_t n, int64_t k) {
  const int splits = gemm_auto_splits(m, n, k);
  return splits > 1 ? static_cast<int64_t>(splits) * m * n * static_cast<int64_t>(sizeof(float)) : 0;
}

extern "C" int xa_gemm_bf16_tn(const void* a, const void* b, void* c, const float* bias, int64_t m, int64_t n, int64_t k,
                               int64_t ldc, int out_bf16, int relu, const void* relu_mask, void* workspace,
                               int64_t workspace_bytes, xa_stream_t stream) {
  return xa_gemm_bf16_tn_ex(a, b, c, bias, m, n, k, ldc, out_bf16, relu, relu_mask, ldc, 0, 0, workspace, workspace_bytes, stream);
}

static int gemm_impl(const char* what, const void* a, const void* b, void* c, const float* bias, int64_t m, int64_t n, int64_t k, int64_t ldc,
                     int out_bf16, int relu, const void* relu_mask, int mask_is_bits, int64_t mask_ld, int64_t col_group, int64_t col_group_pitch,
                     void* workspace, int64_t workspace_bytes, xa_stream_t stream);

extern "C" int xa_gemm_bf16_tn_ex(const void* a, const void* b, void* c, const float* bias, int64_t m, int64_t n, int64_t k,
                                  int64_t ldc, int out_bf16, int relu, const void* relu_mask, int64_t mask_ld, int64_t col_group,
                                  int64_t col_group_pitch, void* workspace, int64_t workspace_bytes, xa_stream_t stream) {
  return gemm_impl("xa_gemm_bf16_tn", a, b, c, bias, m, n, k, ldc, out_bf16, relu, relu_mask, 0, mask_ld, col_group, col_group_pitch, workspace,
                   workspace_bytes, stream);
}

// The same product with the ReLU-derivative mask as BITS (bit j of word i <=> element 32 i + j of the [m, mask_ld] activation is > 0;
// written by the forward epilogue of the layer below, xa_conv2d_nhwc_bf16_ex bits_out): bf16 output only, mask_ld a multiple of 32.
extern "C" int xa_gemm_bf16_tn_maskbits(const void* a, const void* b, void* c, int64_t m, int64_t n, int64_t k, int64_t ldc, const uint32_t* mask_bits,
                                        int64_t mask_ld, int64_t col_group, int64_t col_group_pitch, xa_stream_t stream) {
  const char* what = "xa_gemm_bf16_tn_maskbits";
  XA_REQUIRE(mask_bits != nullptr && mask_ld % 32 == 0 && xa::aligned(mask_bits, 4), XA_EINVAL, "%s: mask_bits null or mask_ld=%lld not a multiple of 32", what,
             static_cast<long long>(mask_ld));
  return gemm_impl(what, a, b, c, nullptr, m, n, k, ldc, 1, 0, mask_bits, 1, mask_ld, col_group, col_group_pitch, nullptr, 0, stream);
}

static int gemm_impl(const char* what, const void* a, const void* b, void* c, const float* bias, int64_t m, int64_t n, int64_t k, int64_t ldc,
                     int out_bf16, int relu, const void* relu_mask, int mask_is_bits, int64_t mask_ld, int64_t col_group, int64_t col_group_pitch,
                     void* workspace, int64_t workspace_bytes, xa_stream_t stream) {
  XA_REQUIRE(a && b && c, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(col_group >= 0 && (col_group == 0 || (col_group % 32 == 0 && col_group_pitch >= col_group &&
                                                   ((n + col_group - 1) / col_group - 1) * col_group_pitch + col_group <= ldc)),
             XA_EINVAL, "%s: col_group=%lld (multiple of 32) / col_group_pitch=%lld do not fit ldc=%lld", what,
             static_cast<long long>(col_group), static_cast<long long>(col_group_pitch), static_cast<long long>(ldc));
  XA_REQUIRE(relu_mask == nullptr || mask_ld >= n, XA_EINVAL, "%s: mask_ld=%lld < n", what, static_cast<long long>(mask_ld));
  XA_REQUIRE(m > 0 && n > 0 && k > 0 && ldc >= n, XA_EINVAL, "%s: m=%lld n=%lld k=%lld ldc=%lld", what, static_cast<long long>(m),
             static_cast<long long>(n), static_cast<long long>(k), static_cast<long long>(ldc));
  XA_REQUIRE(k % 8 == 0, XA_EALIGN, "%s: k=%lld must be a multiple of 8 (16-byte row pitch for TMA)", what, static_cast<long long>(k));
  XA_REQUIRE(xa::aligned(a, 16) && xa::aligned(b, 16), XA_EALIGN, "%s: a and b must be 16-byte aligned", what);
  XA_REQUIRE(m < (int64_t(1) << 31) && n < (int64_t(1) << 31) && k < (int64_t(1) << 31), XA_EOVERFLOW, "%s: dimension too large", what);
  const int bn = gemm_tile_n(n);
  CUtensorMap ma, mb;
  if (int rc = make_map_2d(&ma, a, m, k, kBlockM, what)) return rc;
  if (int rc = make_map_2d(&mb, b, n, k, bn, what)) return rc;
  GemmParams p{};
  p.c = c, p.bias = bias, p.m = m, p.n = n, p.k = k, p.ldc = ldc, p.relu = relu;
  p.mask = mask_is_bits ? nullptr : static_cast<const __nv_bfloat16*>(relu_mask);
  p.mask_bits = mask_is_bits ? static_cast<const uint32_t*>(relu_mask) : nullptr;
  p.mask_words = mask_ld / 32;
  p.mask_ld = mask_ld, p.col_group = col_group, p.col_group_pitch = col_group_pitch;
  const int64_t k_blocks = (k + kBlockK - 1) / kBlockK;
  int splits = workspace != nullptr ? gemm_auto_splits(m, n, k) : 1;
  if (splits > 1 && relu_mask == nullptr && col_group == 0 && workspace_bytes >= static_cast<int64_t>(splits) * m * n * 4) {
    p.kb_per_split = static_cast<int>((k_blocks + splits - 1) / splits);
    p.splits = static_cast<int>((k_blocks + p.kb_per_split - 1) / p.kb_per_split);  // every split owns >= 1 block
    p.partial = static_cast<float*>(workspace);
  } else {
    p.splits = 1;
    p.kb_per_split = static_cast<int>(k_blocks);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // bf16 output through TMA stores: c as a [m, ldc] tensor in 32 x 32 boxes (bf16 masks keep the transposed per-lane path)
  const bool store_ok = out_bf16 && bn >= 32 && p.splits == 1 && p.mask == nullptr && ldc % 8 == 0 && xa::aligned(c, 16) && xa::aligned(bias, 16) &&
                        (col_group == 0 || col_group_pitch % 8 == 0) && (!mask_is_bits || (n % 32 == 0 && mask_ld % 32 == 0));
  XA_REQUIRE(!mask_is_bits || store_ok, XA_EALIGN, "%s: the bit-mask form needs ldc %% 8 == 0, n %% 32 == 0 and a 16-byte aligned c", what);
  CUtensorMap mc = ma;
  if (store_ok && (mask_is_bits || xa::gemm_tma_store_enabled())) {
    if (int rc = make_map_2d_box(&mc, c, m, ldc, 32, 32, what)) return rc;
    p.tma_store = 1;
  }
  if (bn == 256 && p.splits == 1 && m >= 512 && gemm_pairs_enabled()) {  // CTA pairs: each CTA loads half of the B tile
    CUtensorMap mb2;
    if (int rc = make_map_2d(&mb2, b, n, k, bn / 2, what)) return rc;
    return out_bf16 ? launch_pair<256, true>(ma, mb2, mc, p, s, what) : launch_pair<256, false>(ma, mb2, mc, p, s, what);
  }
  if (bn == 256) return out_bf16 ? launch<256, true>(ma, mb, mc, p, s, what) : launch<256, false>(ma, mb, mc, p, s, what);
  if (bn == 128) return out_bf16 ? launch<128, true>(ma, mb, mc, p, s, what) : launch<128, false>(ma, mb, mc, p, s, what);
  if (bn == 64) return out_bf16 ? launch<64, true>(ma, mb, mc, p, s, what) : launch<64, false>(ma, mb, mc, p, s, what);
  return out_bf16 ? launch<16, true>(ma, mb, mc, p, s, what) : launch<16, false>(ma, mb, mc, p, s, what);
}

// The split-K product WITHOUT its reduction pass: fp32 partial tiles [splits, m, n] in `workspace`, for a consumer that adds
// them itself (xa_heads_forward_partial_bf16: at rollout batch sizes the reduction kernel was one launch in eight of an
// environment step).  *splits_out = 1 and nothing launched when the shape is not split: the caller takes the plain path.
extern "C" int xa_gemm_bf16_tn_partial(const void* a, const void* b, int64_t m, int64_t n, int64_t k, void* workspace, int64_t workspace_bytes,
                                       int* splits_out, xa_stream_t stream) {
  const char* what = "xa_gemm_bf16_tn_partial";
  XA_REQUIRE(a && b && splits_out, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(m > 0 && n > 0 && k > 0 && k % 8 == 0, XA_EINVAL, "%s: m=%lld n=%lld k=%lld (k a multiple of 8)", what, static_cast<long long>(m),
             static_cast<long long>(n), static_cast<long long>(k));
  XA_REQUIRE(xa::aligned(a, 16) && xa::aligned(b, 16), XA_EALIGN, "%s: a and b must be 16-byte aligned", what);
  XA_REQUIRE(m < (int64_t(1) << 31) && n < (int64_t(1) << 31) && k < (int64_t(1) << 31), XA_EOVERFLOW, "%s: dimension too large", what);
  const int splits = gemm_auto_splits(m, n, k);
  *splits_out = 1;
  if (splits <= 1 || workspace == nullptr || workspace_bytes < static_cast<int64_t>(splits) * m * n * 4) return XA_OK;
  const int bn = gemm_tile_n(n);
  CUtensorMap ma, mb;
  if (int rc = make_map_2d(&ma, a, m, k, kBlockM, what)) return rc;
  if (int rc = make_map_2d(&mb, b, n, k, bn, what)) return rc;
  GemmParams p{};
  p.m = m, p.n = n, p.k = k, p.ldc = n;
  const int64_t k_blocks = (k + kBlockK - 1) / kBlockK;
  p.kb_per_split = static_cast<int>((k_blocks + splits - 1) / splits);
  p.splits = static_cast<int>((k_blocks + p.kb_per_split - 1) / p.kb_per_split);
  p.partial = static_cast<float*>(workspace);
  if (p.splits <= 1) return XA_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  if (bn == 256) rc = launch<256, false>(ma, mb, ma, p, s, what, false);
  else if (bn == 128) rc = launch<128, false>(ma, mb, ma, p, s, what, false);
  else if (bn == 64) rc = launch<64, false>(ma, mb, ma, p, s, what, false);
  else rc = launch<16, false>(ma, mb, ma, p, s, what, false);
  if (rc == XA_OK) *splits_out = p.splits;
  return rc;
}

// ---------------------------------------------------------------------------------------------- layout helper
// fp32|bf16 [rows, cols] -> bf16, optionally transposed into [cols, ld_dst] (ld_dst >= rows; the padding the
// GEMM's K % 8 rule may need is the caller's zero-filled buffer).  32x32 shared-memory tiles, coalesced both ways.
namespace {

template <typename TIn>
__global__ void __launch_bounds__(256) to_bf16_kernel(const TIn* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                                                       int64_t cols, int64_t ld_dst, int transpose) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 32, c0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int64_t r = r0 + j, c = c0 + tx;
    if (r < rows && c < cols) tile[j][tx] = static_cast<float>(src[r * cols + c]);
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    if (transpose) {
      const int64_t c = c0 + j, r = r0 + tx;  // dst[c, r]
      if (r < rows && c < cols) dst[c * ld_dst + r] = __float2bfloat16_rn(tile[tx][j]);
    } else {
      const int64_t r = r0 + j, c = c0 + tx;
      if (r < rows && c < cols) dst[r * ld_dst + c] = __float2bfloat16_rn(tile[j][tx]);
    }
  }
}

}  // namespace

extern "C" int xa_to_bf16(const void* src, int src_is_f32, void* dst, int64_t rows, int64_t cols, int64_t ld_dst, int transpose,
                          xa_stream_t stream) {
  XA_REQUIRE(src && dst, XA_EINVAL, "xa_to_bf16: null pointer");
  XA_REQUIRE(rows > 0 && cols > 0 && ld_dst >= (transpose ? rows : cols), XA_EINVAL, "xa_to_bf16: rows=%lld cols=%lld ld_dst=%lld",
             static_cast<long long>(rows), static_cast<long long>(cols), static_cast<long long>(ld_dst));
  const dim3 grid(static_cast<unsigned>((rows + 31) / 32), static_cast<unsigned>((cols + 31) / 32));  // rows may be millions
  XA_REQUIRE(grid.y <= 65535, XA_EOVERFLOW, "xa_to_bf16: too many columns for one launch (%lld)", static_cast<long long>(cols));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (src_is_f32)
    to_bf16_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst), rows, cols, ld_dst, transpose);
  else
    to_bf16_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), rows, cols,
                                                        ld_dst, transpose);
  return xa::check_launch("xa_to_bf16");
}

// dst[i] = src[map[i]] (0 where map[i] < 0), as bf16 or fp32: every weight layout the tensor-core kernels read is a
// permutation (+ zero padding) of the fp32 parameters, so one launch of this kernel re-derives all of them after an
// optimiser step (agents/tc_operands.py builds the map once by running the re-layout code on parameter indices).
namespace {
__global__ void __launch_bounds__(256) gather_cast_kernel(const float* __restrict__ src, const int32_t* __restrict__ map, void* __restrict__ dst,
                                                           int64_t n, int out_bf16, int vec_ok) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  // eight outputs per thread and pass: two 16-byte loads of map entries, eight independent parameter loads in flight, one 16-byte
  // (bf16) or two 16-byte (fp32) stores (one output per thread left the kernel waiting on its two dependent loads: 16 us for 3.4 M
  // elements, 13.8 us with four per thread)
  const uint64_t keep = xa::policy_evict_last();   // the parameters stay in L2 for the optimiser's next pass
  const int64_t n8 = vec_ok ? n / 8 : 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < n8; q += stride) {
    const int4 ja = __ldg(reinterpret_cast<const int4*>(map) + 2 * q), jb = __ldg(reinterpret_cast<const int4*>(map) + 2 * q + 1);
    const int32_t j[8] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y, jb.z, jb.w};
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = xa::ld_keep(src + (j[e] >= 0 ? j[e] : 0), keep);   // unconditional loads: all eight issue before the first use
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = j[e] >= 0 ? v[e] : 0.0f;
    if (out_bf16) {
      __nv_bfloat162 h[4] = {__floats2bfloat162_rn(v[0], v[1]), __floats2bfloat162_rn(v[2], v[3]), __floats2bfloat162_rn(v[4], v[5]),
                             __floats2bfloat162_rn(v[6], v[7])};
      reinterpret_cast<uint4*>(dst)[q] = *reinterpret_cast<uint4*>(h);
    } else {
      reinterpret_cast<float4*>(dst)[2 * q] = make_float4(v[0], v[1], v[2], v[3]);
      reinterpret_cast<float4*>(dst)[2 * q + 1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  for (int64_t i = n8 * 8 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int32_t j = __ldg(map + i);
    const float v = j >= 0 ? __ldg(src + j) : 0.0f;
    if (out_bf16)
      static_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(v);
    else
      static_cast<float*>(dst)[i] = v;
  }
}
}  // namespace

extern "C" int xa_gather_cast_f32(const float* src, const int32_t* map, void* dst, int64_t n, int out_bf16, xa_stream_t stream) {
  XA_REQUIRE(n >= 0, XA_EINVAL, "xa_gather_cast_f32: n=%lld", static_cast<long long>(n));
  if (n == 0) return XA_OK;
  XA_REQUIRE(src && map && dst, XA_EINVAL, "xa_gather_cast_f32: null pointer");
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int vec_ok = xa::aligned(map, 16) && xa::aligned(dst, 16);
  const int64_t want = ((vec_ok ? (n + 7) / 8 : n) + 255) / 256, cap = static_cast<int64_t>(sms) * 16;
  xa::launch_chained(xa::kChainElementwise, gather_cast_kernel, dim3(static_cast<unsigned>(want < cap ? want : cap)), dim3(256), 0, static_cast<cudaStream_t>(stream), src, map, dst, n, out_bf16, vec_ok);
  return xa::check_launch("xa_gather_cast_f32");
}
