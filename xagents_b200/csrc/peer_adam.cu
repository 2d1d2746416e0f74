// Collective C1 fused with the optimiser, over NVLink peer memory: ONE kernel per update instead of
// ncclAllReduce + grad_sumsq + clip_adam (SURVEY.md 8e C1 + 8f-2; the reference is single-process: its step is
// tf.clip_by_global_norm + Adam.apply_gradients, xagents/ppo/agent.py:135-137).
//
// Every rank holds a full gradient in a peer-mapped ("symmetric") buffer.  Rank r owns shard r of the flat
// parameter vector (ZeRO-1 style: Adam's m and v exist only for the owned shard):
//
//   phase 0  entry barrier across GPUs (flags in peer memory, release/acquire at system scope): every peer's
//            gradient is complete;
//   phase 1  reduce-scatter by PULL: the owner sums its shard over all ranks' buffers in rank order (P2P loads
//            over NVLink, all G loads of an element group in flight together), keeps the sum in its own buffer
//            and accumulates sum(g^2) in fp64; block partials -> last block (ticket) -> one number per rank,
//            pushed to every peer together with a flag;
//   phase 2  every rank adds the G partial sums in rank order -> the same global norm everywhere -> clip scale;
//   phase 3  Adam on the owned shard (28 B per owned parameter instead of 28 B per parameter), and all-gather
//            by PUSH: the updated weights are stored straight into every peer's parameter buffer;
//   phase 4  exit barrier: all peers' shards have landed here before the kernel completes (the next forward
//            pass reads the weights), and all peers have finished reading this rank's gradient.
//
// Sums are taken in rank order and each element is reduced by exactly one rank, so the result is deterministic
// and the weights are bit-identical on every rank by construction.  The grid is small (the byte count is a few
// MB: NVLink latency, not SM count, bounds it) and launched cooperatively, because blocks wait for each other
// and for the peers: it co-resides with the persistent gather kernel's one CTA per SM.  Every wait is bounded
// (~2 s of clock64) and reports through a status word instead of hanging the GPU.
#include <cooperative_groups.h>
#include <math.h>
#include <string.h>

#include "xa_common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kMaxGrid = 64;
constexpr long long kTimeoutCycles = 4000000000LL;

struct PeerWorkspace {      // local (not peer-mapped) scratch, zero-initialised once
  unsigned int ticket_sumsq;
  unsigned int ticket_exit;
  int status;               // 0 ok; otherwise the phase whose wait timed out
  int pad;
  double partials[kMaxGrid];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {   // peer memory, read once: bypass L1
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// wait until flags[r] >= epoch for every rank r (thread r polls); false on timeout
__device__ __forceinline__ bool wait_all(const uint32_t* flags, int world, uint32_t epoch, int phase, PeerWorkspace* ws) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > kTimeoutCycles) {
        s_ok = 0;
        atomicCAS(&ws->status, 0, phase);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const int ok = s_ok;
  __syncthreads();   // nobody re-arms s_ok (the next wait) before everybody has read it
  return ok != 0;
}

__global__ void __launch_bounds__(kThreads) peer_allreduce_adam_kernel(const xa_peer_adam_args a) {
  __shared__ double scratch[kThreads / 32];
  __shared__ bool s_last;
  __shared__ float s_scale;
  PeerWorkspace* ws = static_cast<PeerWorkspace*>(a.workspace);
  const int me = a.rank, G = a.world;
  uint32_t* my_flags = static_cast<uint32_t*>(a.flags[me]);   // [3][XA_MAX_PEERS]: entry, norm, exit
  const int64_t n4 = a.shard_len / 4;
  const int64_t base4 = static_cast<int64_t>(me) * n4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;

  // ---- phase 0: tell every peer this rank's gradient is complete; wait for theirs --------------------------
  if (blockIdx.x == 0 && threadIdx.x < G)
    st_release_sys(static_cast<uint32_t*>(a.flags[threadIdx.x]) + 0 * XA_MAX_PEERS + me, a.epoch);
  if (!wait_all(my_flags + 0 * XA_MAX_PEERS, G, a.epoch, 1, ws)) return;

  // ---- phase 1: owner pulls its shard from every rank, sums in rank order ---------------------------------
  float4* mine = reinterpret_cast<float4*>(a.grad[me]) + base4;
  const uint64_t keep1 = xa::policy_evict_last();
  double acc = 0.0;
  for (int64_t i = first; i < n4; i += stride) {
    float4 part[XA_MAX_PEERS];
#pragma unroll
    for (int r = 0; r < XA_MAX_PEERS; ++r)
      if (r < G) part[r] = ld_peer(reinterpret_cast<const float4*>(a.grad[r]) + base4 + i);
    float4 s = part[0];
#pragma unroll
    for (int r = 1; r < XA_MAX_PEERS; ++r)
      if (r < G) {
        s.x += part[r].x;
        s.y += part[r].y;
        s.z += part[r].z;
        s.w += part[r].w;
      }
    xa::st_keep(mine + i, s, keep1);   // only the owner reads this region of its own buffer (phase 3: keep it in L2)
    acc += static_cast<double>(s.x) * s.x + static_cast<double>(s.y) * s.y + static_cast<double>(s.z) * s.z +
           static_cast<double>(s.w) * s.w;
  }
  const double block_total = xa::block_sum<kThreads / 32>(acc, scratch);
  if (threadIdx.x == 0) {
    ws->partials[blockIdx.x] = block_total;
    __threadfence();
    s_last = atomicAdd(&ws->ticket_sumsq, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {   // block partials in block order -> this rank's sum(g^2), pushed to every peer with a flag
    __threadfence();
    double t = 0.0;
    if (threadIdx.x == 0) {
      for (unsigned k = 0; k < gridDim.x; ++k) t += const_cast<const volatile double*>(ws->partials)[k];
      scratch[0] = t;
      ws->ticket_sumsq = 0;
    }
    __syncthreads();
    if (threadIdx.x < G) {
      volatile double* slot = static_cast<double*>(a.sumsq[threadIdx.x]) + me;
      *slot = scratch[0];
      st_release_sys(static_cast<uint32_t*>(a.flags[threadIdx.x]) + 1 * XA_MAX_PEERS + me, a.epoch);
    }
  }

  // ---- phase 2: the same global norm on every rank ----------------------------------------------------------
  if (!wait_all(my_flags + 1 * XA_MAX_PEERS, G, a.epoch, 2, ws)) return;
  if (threadIdx.x == 0) {
    double total = 0.0;
    const volatile double* parts = static_cast<const double*>(a.sumsq[me]);
    for (int r = 0; r < G; ++r) total += parts[r];
    float scale = a.inv_world;
    if (a.clip > 0.0f) {   // tf.clip_by_global_norm on the AVERAGED gradient: g * clip * min(1/norm, 1/clip)
      const float norm = a.inv_world * static_cast<float>(sqrt(total));
      scale *= a.clip * fminf(1.0f / norm, 1.0f / a.clip);
    }
    s_scale = scale;
  }
  __syncthreads();
  const float scale = s_scale;

  // ---- phase 3: Adam on the owned shard; push the new weights to every rank ---------------------------------
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  const uint64_t keep = xa::policy_evict_last();
  for (int64_t i = first; i < n4; i += stride) {
    const float4 g = xa::ld_keep(mine + i, keep);
    float4 p = xa::ld_keep(reinterpret_cast<const float4*>(a.param[me]) + base4 + i, keep);
    float4 m = xa::ld_keep(m4 + i, keep), v = xa::ld_keep(v4 + i, keep);   // local optimiser state: stays in L2 under the gather's stream
    const float gx[4] = {g.x * scale, g.y * scale, g.z * scale, g.w * scale};
    float* pp = &p.x;
    float* mm = &m.x;
    float* vv = &v.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      mm[c] = a.beta1 * mm[c] + (1.0f - a.beta1) * gx[c];
      vv[c] = a.beta2 * vv[c] + (1.0f - a.beta2) * gx[c] * gx[c];
      pp[c] -= a.lr_t * mm[c] / (sqrtf(vv[c]) + a.eps);
    }
    xa::st_keep(m4 + i, m, keep);
    xa::st_keep(v4 + i, v, keep);
#pragma unroll
    for (int r = 0; r < XA_MAX_PEERS; ++r)
      if (r < G) reinterpret_cast<float4*>(a.param[r])[base4 + i] = p;
  }

  // ---- phase 4: exit barrier --------------------------------------------------------------------------------
  __threadfence_system();   // this thread's peer stores are visible system-wide before the ticket
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&ws->ticket_exit, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) ws->ticket_exit = 0;
  __threadfence_system();
  if (threadIdx.x < G) st_release_sys(static_cast<uint32_t*>(a.flags[threadIdx.x]) + 2 * XA_MAX_PEERS + me, a.epoch);
  wait_all(my_flags + 2 * XA_MAX_PEERS, G, a.epoch, 3, ws);
}

// Collective C2: the per-minibatch advantage moments of every rank, once per train step -- a few hundred bytes.
// One small CTA: store this rank's block into every peer's receive buffer (P2P stores), raise this rank's flag at every peer,
// wait for every peer's flag here, copy the complete table into the caller's private output.  NCCL's kernels (up to 640 threads
// x 96 registers per CTA) cannot share an SM with a gather CTA and queue until the persistent gather has finished -- measured:
// the whole per-minibatch chain then runs AFTER the gathers instead of under them; this kernel is 128 threads.
// The receive buffer holds two generations (epoch parity): a rank can run at most one all-gather ahead of the slowest peer
// (it cannot finish generation e+1 before every peer has entered it), so generation e+2 never overwrites data still in use.
__global__ void __launch_bounds__(128) peer_allgather_kernel(const xa_peer_gather_args a, const double* __restrict__ src, double* __restrict__ out) {
  PeerWorkspace* ws = nullptr;
  (void)ws;
  const int me = a.rank, G = a.world;
  const int64_t n = a.n_per_rank;
  const int64_t gen = static_cast<int64_t>(a.epoch & 1u) * G * n;
  for (int r = 0; r < G; ++r) {
    double* dst = static_cast<double*>(a.recv[r]) + gen + static_cast<int64_t>(me) * n;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < G) st_release_sys(static_cast<uint32_t*>(a.flags[threadIdx.x]) + me, a.epoch);
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < G) {
    const uint32_t* mine = static_cast<const uint32_t*>(a.flags[me]) + threadIdx.x;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(mine) - a.epoch) < 0) {
      if (clock64() - t0 > kTimeoutCycles) {
        s_ok = 0;
        if (a.status) atomicExch(a.status, 1);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const volatile double* table = static_cast<const double*>(a.recv[me]) + gen;
  for (int64_t i = threadIdx.x; i < n * G; i += blockDim.x) out[i] = table[i];
}

}  // namespace

extern "C" {

int xa_peer_allgather_f64(const xa_peer_gather_args* args, const double* src, double* out, xa_stream_t stream) {
  XA_REQUIRE(args != nullptr && src != nullptr && out != nullptr, XA_EINVAL, "xa_peer_allgather_f64: null pointer");
  const xa_peer_gather_args& a = *args;
  XA_REQUIRE(a.world >= 1 && a.world <= XA_MAX_PEERS && a.rank >= 0 && a.rank < a.world, XA_EINVAL, "xa_peer_allgather_f64: rank %d of %d", a.rank, a.world);
  XA_REQUIRE(a.n_per_rank > 0 && a.epoch > 0, XA_EINVAL, "xa_peer_allgather_f64: n_per_rank=%lld epoch=%u", static_cast<long long>(a.n_per_rank), a.epoch);
  for (int r = 0; r < a.world; ++r)
    XA_REQUIRE(a.recv[r] && a.flags[r] && xa::aligned(a.recv[r], 8) && xa::aligned(a.flags[r], 4), XA_EINVAL, "xa_peer_allgather_f64: bad peer pointer for rank %d", r);
  peer_allgather_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(a, src, out);
  return xa::check_launch("xa_peer_allgather_f64");
}

// Peer-mappable allocations by CUDA IPC (the transport used when torch's symmetric-memory rendezvous is not possible on a
// box): cudaMalloc + cudaIpcGetMemHandle here, cudaIpcOpenMemHandle in the peers (lazy peer-access enabling).
int xa_ipc_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
  XA_REQUIRE(bytes > 0 && ptr && handle64, XA_EINVAL, "xa_ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaMalloc(ptr, static_cast<size_t>(bytes));
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, static_cast<size_t>(bytes));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    xa::set_error("xa_ipc_alloc: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  memcpy(handle64, &h, 64);
  return XA_OK;
}

int xa_ipc_open(const unsigned char* handle64, void** ptr) {
  XA_REQUIRE(handle64 && ptr, XA_EINVAL, "xa_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    xa::set_error("xa_ipc_open: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return XA_OK;
}

int xa_ipc_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) cudaGetLastError();
  return static_cast<int>(e);
}

int xa_ipc_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) cudaGetLastError();
  return static_cast<int>(e);
}

int64_t xa_peer_adam_workspace_bytes(void) { return static_cast<int64_t>(sizeof(PeerWorkspace)); }

int64_t xa_peer_adam_flag_bytes(void) { return 3 * XA_MAX_PEERS * static_cast<int64_t>(sizeof(uint32_t)); }

int xa_peer_allreduce_adam_f32(const xa_peer_adam_args* args, double lr, double beta1, double beta2, double eps, double clip_norm,
                               int64_t step, xa_stream_t stream) {
  XA_REQUIRE(args != nullptr, XA_EINVAL, "xa_peer_allreduce_adam_f32: null args");
  xa_peer_adam_args a = *args;
  XA_REQUIRE(a.world >= 1 && a.world <= XA_MAX_PEERS && a.rank >= 0 && a.rank < a.world, XA_EINVAL,
             "xa_peer_allreduce_adam_f32: rank %d of %d (at most %d peers)", a.rank, a.world, XA_MAX_PEERS);
  XA_REQUIRE(a.shard_len > 0 && a.shard_len % 4 == 0, XA_EINVAL, "xa_peer_allreduce_adam_f32: shard_len=%lld must be a positive multiple of 4",
             static_cast<long long>(a.shard_len));
  XA_REQUIRE(step > 0 && a.epoch > 0, XA_EINVAL, "xa_peer_allreduce_adam_f32: step=%lld epoch=%u must be positive", static_cast<long long>(step), a.epoch);
  XA_REQUIRE(a.m && a.v && a.workspace, XA_EINVAL, "xa_peer_allreduce_adam_f32: null pointer");
  for (int r = 0; r < a.world; ++r) {
    XA_REQUIRE(a.grad[r] && a.param[r] && a.sumsq[r] && a.flags[r], XA_EINVAL, "xa_peer_allreduce_adam_f32: null peer pointer for rank %d", r);
    XA_REQUIRE(xa::aligned(a.grad[r], 16) && xa::aligned(a.param[r], 16) && xa::aligned(a.sumsq[r], 8) && xa::aligned(a.flags[r], 4), XA_EALIGN,
               "xa_peer_allreduce_adam_f32: misaligned peer pointer for rank %d", r);
  }
  XA_REQUIRE(xa::aligned(a.m, 16) && xa::aligned(a.v, 16) && xa::aligned(a.workspace, 8), XA_EALIGN, "xa_peer_allreduce_adam_f32: misaligned local pointer");
  // Keras Adam: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
  a.lr_t = static_cast<float>(lr * sqrt(1.0 - pow(beta2, static_cast<double>(step))) / (1.0 - pow(beta1, static_cast<double>(step))));
  a.beta1 = static_cast<float>(beta1);
  a.beta2 = static_cast<float>(beta2);
  a.eps = static_cast<float>(eps);
  a.clip = static_cast<float>(clip_norm > 0.0 ? clip_norm : 0.0);
  a.inv_world = 1.0f / static_cast<float>(a.world);
  const int64_t want = (a.shard_len / 4 + kThreads - 1) / kThreads;
  const unsigned grid = static_cast<unsigned>(want < 1 ? 1 : (want > kMaxGrid ? kMaxGrid : want));
  void* params[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(peer_allreduce_adam_kernel), dim3(grid), dim3(kThreads), params, 0,
                                              static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    cudaGetLastError();
    xa::set_error("xa_peer_allreduce_adam_f32: cooperative launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return xa::check_launch("xa_peer_allreduce_adam_f32");
}

}  // extern "C"
