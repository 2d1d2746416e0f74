// Issue-rate microbenchmark for tcgen05.mma (cta_group::1, kind::f16, bf16, M = 128): cycles per MMA as a function of N,
// operand majorness, how many accumulators the instruction stream alternates between, and how often it commits.
// Operands are whatever shared memory holds (zeros); only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Ixagents_b200/csrc -o scripts/_build/umma_microbench scripts/umma_microbench.cu
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

using namespace xa_tc;

// Second experiment: `issuers` warps issue concurrently (own accumulators, own barriers), descriptors precomputed and
// the K-step loop unrolled, to separate per-thread issue cost from a limit of the tensor pipe itself.
__global__ void __launch_bounds__(192) bench2(int n, int issuers, int per_commit, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4][9];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int w = 0; w < 4; ++w)
      for (int i = 0; i < 9; ++i) xa::mbar_init(&bars[w][i], 1);
    xa::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xa::smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  if (warp >= 1 && warp <= issuers) {
    if (elect_one()) {
      const int w = warp - 1;
      const uint32_t idesc = make_idesc(128, n, false, false);
      const uint64_t da = make_smem_desc(smem), db = make_smem_desc(smem + 32768);
      const uint32_t acc = tmem_base + w * 128;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        for (int j = 0; j < per_commit; j += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc, da + 2 * k, db + 2 * k, idesc, 1);
        }
        umma_commit(&bars[w][it & 7]);
      }
      umma_commit(&bars[w][8]);
      mbar_wait_wd(&bars[w][8], 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0 && w == 0) out[0] = t1 - t0;
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

__global__ void __launch_bounds__(192) bench(int n, int mn_major, int n_acc, int per_commit, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[9];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 9; ++i) xa::mbar_init(bars + i, 1);
    xa::fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(xa::smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, n, mn_major, mn_major);
      const uint32_t a_base = xa::smem_u32(smem), b_base = a_base + 32768;
      const long long t0 = clock64();
      int m = 0;
      for (int it = 0; it < iters; ++it) {
        const uint32_t stage = (it & 1) * 65536u;   // alternate two operand stages like a real pipeline
        for (int j = 0; j < per_commit; ++j, ++m) {
          const int k = j & 3;
          uint64_t da, db;
          if (mn_major) {
            da = make_smem_desc_mn(a_base + stage + k * 2048, 8192, 1024, false);
            db = make_smem_desc_mn(b_base + stage + k * 2048, 8192, 1024, false);
          } else {
            da = make_smem_desc(smem + stage) + 2 * k;
            db = make_smem_desc(smem + stage + 32768) + 2 * k;
          }
          umma_bf16(tmem_base + (m % n_acc) * n, da, db, idesc, 1);
        }
        umma_commit(bars + (it & 7));
      }
      umma_commit(bars + 8);
      mbar_wait_wd(bars + 8, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

int main() {
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("| N | operands | accumulators | MMAs per commit | cycles per MMA | math floor (N/2) |\n|---|---|---|---|---|---|\n");
  const int ns[] = {32, 64, 128, 256};
  for (int mn = 0; mn < 2; ++mn)
    for (int n : ns)
      for (int n_acc : {1})
        for (int per_commit : {4, 16}) {
          if (n_acc * n > 512) continue;
          const int iters = 4096 / per_commit;
          for (int rep = 0; rep < 2; ++rep) bench<<<148, 192, 200 * 1024>>>(n, mn, n_acc, per_commit, iters, out);
          long long cyc = 0;
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(e));
            return 1;
          }
          cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
          printf("| %d | %s | %d | %d | %.1f | %d |\n", n, mn ? "MN-major" : "K-major", n_acc, per_commit, double(cyc) / 4096.0, n / 2);
        }
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("\n| N | issuing warps | MMAs per commit | cycles per MMA per issuer | aggregate cycles per MMA |\n|---|---|---|---|---|\n");
  for (int n : {32, 64, 128})
    for (int issuers : {1, 2, 4})
      for (int per_commit : {4, 16, 64}) {
        const int iters = 4096 / per_commit;
        for (int rep = 0; rep < 2; ++rep) bench2<<<148, 192, 200 * 1024>>>(n, issuers, per_commit, iters, out);
        long long cyc = 0;
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("error: %s\n", cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
        printf("| %d | %d | %d | %.1f | %.1f |\n", n, issuers, per_commit, double(cyc) / 4096.0, double(cyc) / 4096.0 / issuers);
      }
  return 0;
}
