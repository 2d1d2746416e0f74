// Shared helpers for the xagents_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "xagents_b200.h"

namespace xa {

// ---- host side: thread-local error string ------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> positive code + message, else XA_OK
int sm_count();                      // cached multiProcessorCount of the current device (<=0: none)

#define XA_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::xa::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

// ---- chained launches (programmatic dependent launch) -----------------------------------------------------------
// The rollout step, the network pass and the loss -> optimiser chain are sequences of SHORT kernels on one stream (8-60 us
// each at rollout batch sizes): between two of them the GPU idles for the launch latency plus the next kernel's on-chip
// set-up (barrier initialisation, TMEM allocation, descriptor prefetch).  Kernels launched through `launch_chained` may
// start while their predecessor on the stream is still draining; each such kernel executes `xa::pdl_wait()` before its
// first global-memory access (it returns once the predecessor grid has completed and its writes are visible) and
// `xa::pdl_trigger()` to let its own successor be scheduled early.  Both are no-ops for a kernel launched the plain way.
// Measured (profiles/r2_chained_launches.md): the rollout's 256-frame forward pass 49 -> 41 us, a 128-step rollout graph 6.6 ->
// 5.7 ms, the loss -> optimiser chain of the benchmarked step +1 %; device-filling launches (the 8192-frame network kernels)
// LOSE ~1 us each, so every launch site says whether its launch is `small` (tensor-core kernels: a few tiles per CTA).
// XA_PDL in the environment: 0 = never (plain stream order), 1 = small launches (default), 2 = every launch, 3 = small
// tensor-core launches only.
int pdl_level();
bool gemm_tma_store_enabled();   // csrc/gemm_tc.cu: bf16 GEMM epilogues through TMA stores (XA_GEMM_TMA_STORE=0 turns them off)
enum : int { kChainSmall = 1, kChainLarge = 0, kChainElementwise = 2 };
template <typename... P, typename... A>
inline void launch_chained(int kind, void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  const int level = pdl_level();
  const bool chain = level == 2 || (level == 1 && kind != kChainLarge) || (level == 3 && kind == kChainSmall);
  cfg.attrs = &attr, cfg.numAttrs = chain ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);   // a failure stays in cudaGetLastError for check_launch
}

__host__ __device__ static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// ---- device side ------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// env-major flat sample id b = e*T + t  ->  row of the time-major [T*E] buffer (base.py:559-564)
__device__ __forceinline__ int64_t sample_row(int32_t b, int n_steps, int n_envs) {
  if (n_steps <= 0) return b;
  const int e = b / n_steps;
  const int t = b - e * n_steps;
  return static_cast<int64_t>(t) * n_envs + e;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + bulk-copy (TMA, non-tensor) primitives
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "XA_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra XA_DONE;\n\t"
      "bra XA_WAIT;\n\t"
      "XA_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// 16-byte global accesses that ask L2 to KEEP the line (optimiser state re-used every minibatch while the gather streams
// gigabytes through the same L2 with evict-first)
__device__ __forceinline__ float4 ld_keep(const float4* p, uint64_t policy) {
  float4 r;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void st_keep(float4* p, const float4& v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}
__device__ __forceinline__ float ld_keep(const float* p, uint64_t policy) {
  float r;
  asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(policy));
  return r;
}
__device__ __forceinline__ void st_keep(float* p, float v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}
// global -> shared, completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_nohint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* gdst, const void* smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// streaming 128-bit accesses (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// block-wide sum of doubles; result valid in thread 0 (and broadcast through smem by callers)
template <int kWarps>
__device__ __forceinline__ double block_sum(double v, double* scratch /*[kWarps]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double total = 0.0;
  if (warp == 0) {
    total = lane < kWarps ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  }
  __syncthreads();
  return total;
}

}  // namespace xa
