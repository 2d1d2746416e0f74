// Shared pieces of the tcgen05 kernels (GEMM and implicit-GEMM convolution): mbarrier wait with a watchdog,
// TMA tile loads, UMMA shared-memory / instruction descriptors, tcgen05.mma / commit / ld wrappers and the
// host-side cuTensorMapEncodeTiled entry point.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "xa_common.cuh"

namespace xa_tc {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int kUmmaK = 16;
constexpr int kThreads = 192;
constexpr uint32_t kWatchdog = 1u << 28;


__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = xa::smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (++spins > kWatchdog) __trap();  // a descriptor / barrier bug must fault, not hang the GPU
  }
}

// One lane of a converged warp (elect.sync): unlike `lane == 0`, ptxas then knows the branch is single-threaded and
// issues tcgen05 / TMA instructions straight from uniform registers instead of wrapping each in an elect-broadcast loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// The same wait for warps that wait LONG (epilogue warps waiting for a whole tile's MMAs): back off between polls so
// that eight polling warps do not compete with the producer / issuer for issue slots and the mbarrier unit.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = xa::smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(128);
    if (++spins > (kWatchdog >> 4)) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          xa::smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(xa::smem_u32(bar))
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell):
// rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(const void* tile) {
  const uint64_t addr = xa::smem_u32(tile);
  uint64_t d = 0;
  d |= (addr >> 4) & 0x3FFF;              // start address            [0,14)
  d |= static_cast<uint64_t>(1) << 16;    // leading byte offset (unused for swizzled K-major) [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;  // stride byte offset: 8 rows * 128 B  [32,46)
  d |= static_cast<uint64_t>(1) << 46;    // descriptor version 1       [46,48)
  d |= static_cast<uint64_t>(2) << 61;    // layout type SWIZZLE_128B   [61,64)
  return d;
}

// MN-major shared-memory matrix descriptor: the operand is stored [K rows][MN contiguous], rows of 128 B
// (SWIZZLE_128B, 64 bf16) or 64 B (SWIZZLE_64B, 32 bf16).  In 16-byte units the canonical layout is
// ((8|4, n), (8, k)) : ((1, LBO), (8|4, SBO)): LBO = distance between 64(32)-element groups along MN, SBO = distance
// between 8-row groups along K (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, bool sw64,
                                                      uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= (addr >> 4) & 0x3FFF;
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(sw64 ? 4 : 2) << 61;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: bf16 x bf16 -> f32; bits 15 / 16 select MN-major A / B
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) /*c = f32*/ | (1u << 7) /*a = bf16*/ | (1u << 10) /*b = bf16*/ | (a_mn ? (1u << 15) : 0u) |
         (b_mn ? (1u << 16) : 0u) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(xa::smem_u32(bar)) : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster -- the two SMs of a TPC -- run ONE 256-row MMA.  Each holds its own 128
// rows of A and HALF of B's rows in shared memory (same offsets in both CTAs: the leader's descriptors address both), each
// accumulates its 128 rows in its own TMEM; the leader (cluster rank 0) issues.  Bit 24 of a shared::cluster address selects
// the peer CTA of the pair: cleared, the address names the LEADER's copy of a barrier from either CTA.
constexpr uint32_t kLeaderMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// this CTA's box -> its own shared memory; the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          xa::smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(xa::smem_u32(bar) & kLeaderMask)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(xa::smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(xa::smem_u32(bar) & kLeaderMask) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          xa::smem_u32(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(xa::smem_u32(bar))
      : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult status;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) == cudaSuccess &&
        status == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    else
      cudaGetLastError();
  }
  return fn;
}

// 2-D bf16 row-major [rows, k] -> tiles of box_rows x 64, 128-B swizzled; out-of-bounds reads give zeros
inline int make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t k, int box_rows, const char* what) {
  EncodeTiledFn fn = encode_fn();
  XA_REQUIRE(fn != nullptr, XA_EINVAL, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XA_REQUIRE(r == CUDA_SUCCESS, XA_EINVAL, "%s: cuTensorMapEncodeTiled failed with %d (rows=%lld k=%lld)", what, static_cast<int>(r),
             static_cast<long long>(rows), static_cast<long long>(k));
  return XA_OK;
}

// 2-D bf16 row-major [rows, cols] -> boxes of box_rows x box_cols (box_cols * 2 = the swizzle span: 128 or 64 bytes)
inline int make_map_2d_box(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int box_rows, int box_cols, const char* what) {
  EncodeTiledFn fn = encode_fn();
  XA_REQUIRE(fn != nullptr, XA_EINVAL, "%s: cuTensorMapEncodeTiled is not available from this driver", what);
  XA_REQUIRE(box_cols == 64 || box_cols == 32, XA_EINVAL, "%s: box of %d columns", what, box_cols);
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  XA_REQUIRE(r == CUDA_SUCCESS, XA_EINVAL, "%s: cuTensorMapEncodeTiled failed with %d (rows=%lld cols=%lld)", what, static_cast<int>(r),
             static_cast<long long>(rows), static_cast<long long>(cols));
  return XA_OK;
}

}  // namespace xa_tc
