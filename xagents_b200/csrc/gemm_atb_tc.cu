// C[M, N] = A^T B for row-major A [K, M] and B [K, N] (bf16, fp32 accumulate): the weight-gradient product of a Dense layer,
// dW = dY^T X with dY [batch, out] and X [batch, in] exactly as the forward / data-gradient kernels leave them
// (Keras Dense backward, utils/common.py:239-258 via tape.gradient, ppo/agent.py:135).
//
// Both operands have the contraction index as their SLOW dimension, i.e. they are "MN-major" for tcgen05, which the UMMA
// shared-memory descriptors support for bf16: a TMA box of 64 k-rows x 64 columns (128-B rows, SWIZZLE_128B) is the
// canonical MN-major tile, LBO = the distance between 64-column groups, SBO = 1024 B between 8-row groups.  So no
// transposed copy of either operand is ever written (the first version of the backward spent more time transposing
// than multiplying).  One CTA per (128 x BN output tile, K split); fp32 partials of a split-K launch are added in split
// order by a second pass (deterministic).
#include "tc_common.cuh"

namespace {

using namespace xa_tc;

struct AtbParams {
  float* c;
  float* partial;  // [splits, m, n] when splits > 1
  int64_t m, n, k, ldc;
  int splits, kb_per_split;
};

template <int BN>
struct AtbSmem {
  static constexpr int kStageA = 2 * 64 * 128;           // two 64-column groups of 64 k-rows
  static constexpr int kStageB = (BN / 64) * 64 * 128;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (200 * 1024) / kStage > 8 ? 8 : (200 * 1024) / kStage;
  static constexpr int kBytes = kStages * kStage + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(kThreads) gemm_atb_kernel(const __grid_constant__ CUtensorMap map_a,
                                                            const __grid_constant__ CUtensorMap map_b, const AtbParams p) {
  using S = AtbSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStage);
  uint64_t* empty = full + S::kStages;
  uint64_t* acc_full = empty + S::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  xa::pdl_trigger();   // chained launch (xa_common.cuh): the on-chip set-up below overlaps the predecessor's tail
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_m = static_cast<int>((p.m + kBlockM - 1) / kBlockM);
  const int tiles_n = static_cast<int>((p.n + BN - 1) / BN);
  const int tile = blockIdx.x % (tiles_m * tiles_n), split = blockIdx.x / (tiles_m * tiles_n);
  const int tile_m = tile % tiles_m, tile_n = tile / tiles_m;
  const int k_blocks = static_cast<int>((p.k + kBlockK - 1) / kBlockK);
  const int kb0 = split * p.kb_per_split, kb1 = min(k_blocks, kb0 + p.kb_per_split);
  constexpr uint32_t kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // a power of two

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < S::kStages; ++s) {
      xa::mbar_init(full + s, 1);
      xa::mbar_init(empty + s, 1);
    }
    xa::mbar_init(acc_full, 1);
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
  xa::pdl_wait();      // first global-memory access below: the predecessor grid has completed

  if (warp == 0) {
    if (elect_one()) {  // ---- TMA producer: 64 k-rows of 128 A columns and BN B columns per stage (columns past M / N read as zero)
      uint32_t it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % S::kStages;
        const uint32_t round = it / S::kStages;
        if (round > 0) mbar_wait_wd(empty + s, (round - 1) & 1);
        uint8_t* dst = smem + s * S::kStage;
        xa::mbar_expect_tx(full + s, S::kStage);
#pragma unroll
        for (int g = 0; g < 2; ++g) tma_load_2d(dst + g * 8192, &map_a, tile_m * kBlockM + g * 64, kb * kBlockK, full + s);
#pragma unroll
        for (int g = 0; g < BN / 64; ++g) tma_load_2d(dst + S::kStageA + g * 8192, &map_b, tile_n * BN + g * 64, kb * kBlockK, full + s);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ---- MMA issuer
      constexpr uint32_t idesc = make_idesc(kBlockM, BN, true, true);
      uint32_t it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % S::kStages;
        mbar_wait_wd(full + s, (it / S::kStages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = xa::smem_u32(smem + s * S::kStage);
        const uint64_t da = make_smem_desc_mn(base, 8192, 1024, false);
        const uint64_t db = make_smem_desc_mn(base + S::kStageA, 8192, 1024, false);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k)  // 16 k-rows further: +2048 B
          umma_bf16(tmem_base, da + k * (2048u >> 4), db + k * (2048u >> 4), idesc, (kb > kb0) || (k != 0));
        umma_commit(empty + s);
      }
      umma_commit(acc_full);
    }
  } else {
    // ---- epilogue: fp32 tile (or split partial)
    const int quad = warp & 3;
    mbar_wait_wd(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t row = static_cast<int64_t>(tile_m) * kBlockM + quad * 32 + lane;
    float* out = p.splits > 1 ? p.partial + (static_cast<int64_t>(split) * p.m + row) * p.n : p.c + row * p.ldc;
    const int64_t ld_ok = p.splits > 1 ? p.n : p.ldc;
    const bool vec = (ld_ok % 4) == 0 && xa::aligned(p.splits > 1 ? static_cast<const void*>(p.partial) : static_cast<const void*>(p.c), 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + c0, v);
      const int64_t col0 = static_cast<int64_t>(tile_n) * BN + c0;
      if (row < p.m && col0 < p.n) {
        if (vec && col0 + 32 <= p.n) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4*>(out + col0)[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        } else {
          for (int j = 0; j < 32 && col0 + j < p.n; ++j) out[col0 + j] = __uint_as_float(v[j]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

__global__ void __launch_bounds__(256) atb_reduce_kernel(const AtbParams p) {
  const int64_t total = p.m * p.n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = 0.0f;
    for (int s = 0; s < p.splits; ++s) acc += p.partial[static_cast<int64_t>(s) * total + i];
    const int64_t row = i / p.n;
    p.c[row * p.ldc + (i - row * p.n)] = acc;
  }
}

// Tile width and K split by a small cost model (cycles of the slowest CTA x waves): a 128 x BN x 16 MMA with both operands in
// shared memory costs max(BN / 2, 32 + BN / 4) cycles (math vs operand fetch, scripts/umma_microbench.cu), a CTA adds ~3000
// cycles of prologue + epilogue, and a launch should fill the SMs once.  The FC weight gradient of the network
// (512 x 3136 x 8192) lands on BN = 192 with 2 splits: 136 CTAs of 64 K blocks instead of 100 CTAs of 128.
struct AtbPlan {
  int bn, splits;
};
AtbPlan atb_plan(int64_t m, int64_t n, int64_t k, bool may_split) {
  const int64_t k_blocks = (k + kBlockK - 1) / kBlockK;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  AtbPlan best{n > 64 ? 128 : 64, 1};
  double best_cost = 1e30;
  const int widths[3] = {64, 128, 192};
  for (int bn : widths) {
    if (bn > 64 && n <= bn - 64) continue;   // a narrower tile already covers n
    const int64_t tiles = ((m + kBlockM - 1) / kBlockM) * ((n + bn - 1) / bn);
    const int max_splits = may_split ? static_cast<int>(k_blocks / 8 > 16 ? 16 : (k_blocks / 8 < 1 ? 1 : k_blocks / 8)) : 1;
    for (int sp = 1; sp <= max_splits; ++sp) {
      const int64_t per = (k_blocks + sp - 1) / sp;
      const int64_t ctas = tiles * ((k_blocks + per - 1) / per);
      const int64_t waves = (ctas + sms - 1) / sms;
      const double mma = bn / 2 > 32 + bn / 4 ? bn / 2 : 32 + bn / 4;
      const double cost = static_cast<double>(waves) * (static_cast<double>(per) * 4.0 * mma + 3000.0) + (sp > 1 ? 1500.0 + 40.0 * sp : 0.0);
      if (cost < best_cost) best_cost = cost, best = AtbPlan{bn, static_cast<int>((k_blocks + per - 1) / per)};
    }
  }
  return best;
}

template <int BN>
int launch_atb(const CUtensorMap& ma, const CUtensorMap& mb, const AtbParams& p, cudaStream_t s, const char* what, bool reduce) {
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(gemm_atb_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, AtbSmem<BN>::kBytes);
    if (e != cudaSuccess) {
      xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_dev = dev;
  }
  const int64_t tiles = ((p.m + kBlockM - 1) / kBlockM) * ((p.n + BN - 1) / BN);
  xa::launch_chained(xa::kChainLarge, gemm_atb_kernel<BN>, dim3(static_cast<unsigned>(tiles * p.splits)), dim3(kThreads), AtbSmem<BN>::kBytes, s, ma, mb, p);
  if (int rc = xa::check_launch(what)) return rc;
  if (p.splits > 1 && reduce) {
    const int64_t want = (p.m * p.n + 255) / 256;
    atb_reduce_kernel<<<static_cast<unsigned>(want < 4096 ? want : 4096), 256, 0, s>>>(p);
    return xa::check_launch(what);
  }
  return XA_OK;
}

}  // namespace

extern "C" {

int64_t xa_gemm_atb_workspace_bytes(int64_t m, int64_t n, int64_t k) {
  const int splits = atb_plan(m, n, k, true).splits;
  return splits > 1 ? static_cast<int64_t>(splits) * m * n * static_cast<int64_t>(sizeof(float)) : 0;
}

int xa_gemm_atb_plan(int64_t m, int64_t n, int64_t k, int* splits) {
  XA_REQUIRE(m > 0 && n > 0 && k > 0, XA_EINVAL, "xa_gemm_atb_plan: non-positive size");
  if (splits) *splits = atb_plan(m, n, k, true).splits;
  return XA_OK;
}

static int atb_run(const char* what, const void* a, const void* b, float* c, int64_t m, int64_t n, int64_t k, int64_t ldc, void* workspace,
                   int64_t workspace_bytes, bool reduce, xa_stream_t stream) {
  XA_REQUIRE(a && b && c, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(m > 0 && n > 0 && k > 0 && ldc >= n, XA_EINVAL, "%s: m=%lld n=%lld k=%lld ldc=%lld", what, static_cast<long long>(m),
             static_cast<long long>(n), static_cast<long long>(k), static_cast<long long>(ldc));
  XA_REQUIRE(m % 8 == 0 && n % 8 == 0, XA_EALIGN, "%s: m=%lld and n=%lld must be multiples of 8 (16-byte row pitch for TMA)", what,
             static_cast<long long>(m), static_cast<long long>(n));
  XA_REQUIRE(xa::aligned(a, 16) && xa::aligned(b, 16) && xa::aligned(c, 4), XA_EALIGN, "%s: a and b must be 16-byte aligned", what);
  XA_REQUIRE(m < (int64_t(1) << 31) && n < (int64_t(1) << 31) && k < (int64_t(1) << 31), XA_EOVERFLOW, "%s: dimension too large", what);
  CUtensorMap ma, mb;
  if (int rc = make_map_2d_box(&ma, a, k, m, 64, 64, what)) return rc;
  if (int rc = make_map_2d_box(&mb, b, k, n, 64, 64, what)) return rc;
  AtbParams p{};
  p.c = c, p.m = m, p.n = n, p.k = k, p.ldc = ldc;
  const int64_t k_blocks = (k + kBlockK - 1) / kBlockK;
  AtbPlan plan = atb_plan(m, n, k, workspace != nullptr);
  if (plan.splits > 1 && workspace_bytes < static_cast<int64_t>(plan.splits) * m * n * 4) plan = atb_plan(m, n, k, false);
  if (plan.splits > 1) {
    p.kb_per_split = static_cast<int>((k_blocks + plan.splits - 1) / plan.splits);
    p.splits = static_cast<int>((k_blocks + p.kb_per_split - 1) / p.kb_per_split);
    p.partial = static_cast<float*>(workspace);
  } else {
    p.splits = 1, p.kb_per_split = static_cast<int>(k_blocks);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (plan.bn) {
    case 192:
      return launch_atb<192>(ma, mb, p, s, what, reduce);
    case 128:
      return launch_atb<128>(ma, mb, p, s, what, reduce);
    default:
      return launch_atb<64>(ma, mb, p, s, what, reduce);
  }
}

int xa_gemm_bf16_atb(const void* a, const void* b, float* c, int64_t m, int64_t n, int64_t k, int64_t ldc, void* workspace,
                     int64_t workspace_bytes, xa_stream_t stream) {
  return atb_run("xa_gemm_bf16_atb", a, b, c, m, n, k, ldc, workspace, workspace_bytes, true, stream);
}

/* The same product left as split partials [splits, m, n] (splits from xa_gemm_atb_plan; with splits == 1 the product itself
 * lands in `partial`): the caller's own reduction (xa_grad_finalize_f32) adds them. */
int xa_gemm_bf16_atb_partial(const void* a, const void* b, int64_t m, int64_t n, int64_t k, float* partial, int64_t partial_bytes,
                             xa_stream_t stream) {
  const char* what = "xa_gemm_bf16_atb_partial";
  XA_REQUIRE(partial != nullptr, XA_EINVAL, "%s: null pointer", what);
  const int splits = atb_plan(m > 0 ? m : 1, n > 0 ? n : 1, k > 0 ? k : 1, true).splits;
  XA_REQUIRE(partial_bytes >= static_cast<int64_t>(splits) * m * n * 4, XA_ENOSPACE, "%s: partial buffer too small", what);
  return atb_run(what, a, b, partial, m, n, k, n, splits > 1 ? partial : nullptr, partial_bytes, false, stream);
}

}  // extern "C"
