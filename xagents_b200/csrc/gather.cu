// Permute-gather of rollout rows into minibatches.
//
// Replaces the five tf.gather calls per minibatch of PPO.get_mini_batches (xagents/ppo/agent.py:149-154)
// and folds in the env-major flatten of BaseAgent.concat_step_batches (xagents/base.py:559-564): the
// source stays in the time-major layout the rollout was written in and the flat env-major sample id is
// remapped to a row on the fly, so the reference's full-rollout transpose copy never happens.
//
// This is pure byte movement, ~99.8 % of the path's traffic (2 * row_bytes per sample per epoch), and
// HBM-bound.  Two data paths:
//
//  * BULK  -- rows are moved by the TMA engine as non-tensor bulk copies: global -> shared
//             (cp.async.bulk ... mbarrier::complete_tx) then shared -> global (cp.async.bulk ... bulk_group).
//             One elected lane per warp drives one stage of a shared-memory ring (4 stages of half a
//             28 224-B frame: ~56 KB in flight per SM, the measured optimum), ~10 instructions per
//             chunk and no register staging.  One more warp gathers the fp32 scalar fields.
//  * VECTOR -- 128-bit LDG/STG (narrower when alignment forces it), 8 independent loads per thread before
//             the stores; the general path for short or unaligned rows.
#include <cstdlib>

#include "xa_common.cuh"

namespace {

constexpr int kMaxStages = 16;
constexpr int kSmemBudget = 227 * 1024 - 1024;  // dynamic shared memory we allow one CTA to take

struct GatherParams {
  const uint8_t* src;
  uint8_t* dst;
  const int32_t* idx;
  int64_t n_idx;
  int64_t row_bytes;
  int n_steps, n_envs;
  // bulk path: a row is cut into chunks_per_row pieces of chunk_bytes (the last one may be shorter)
  int chunks_per_row;
  uint32_t chunk_bytes;
  int stages;
  int load_hint, store_hint;  // L2 evict-first policy on the bulk loads / stores
  // scalar fields riding along
  int64_t n_src_rows;  // host-side checks only
  // fine-grained completion (bulk path): progress[m] += chunks finished, m = the minibatch a destination row belongs to
  uint32_t* progress;
  uint32_t row_offset, rows_per_epoch, mb_rows, mbs_per_epoch;
  // dynamic work distribution (bulk path): work[0] = next item to hand out, work[1] = lanes that have finished; both zero at
  // launch, and the last lane to finish zeroes them again.  nullptr: items are dealt statically (item k -> CTA k mod grid).
  uint32_t* work;
  int n_fields;
  const float* fsrc[XA_MAX_FIELDS];
  float* fdst[XA_MAX_FIELDS];
};

// ------------------------------------------------------------------------------------------ bulk
__global__ void __launch_bounds__(32 * (kMaxStages + 1)) gather_bulk_kernel(const GatherParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * p.chunk_bytes);

  if (warp < stages) {
    if (lane != 0) return;
    uint64_t* bar = bars + warp;
    uint8_t* buf = smem + static_cast<size_t>(warp) * p.chunk_bytes;
    xa::mbar_init(bar, 1);
    xa::fence_barrier_init();
    const uint64_t policy = xa::policy_evict_first();
    const int64_t n_items = p.n_idx * p.chunks_per_row;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * stages;
    // Static dealing makes the launch as slow as its slowest CTA: when other kernels hold SMs at launch time the block
    // scheduler doubles gather CTAs up on the free SMs for the whole persistent launch, and kernels that come and go beside
    // it (losses, optimiser) slow the CTAs they share an SM with.  With a work counter every lane takes the next item when it
    // is ready for one (the ticket for item k+1 is drawn while item k's copy is in flight, so its latency is hidden).
    int64_t k = p.work ? static_cast<int64_t>(atomicAdd(p.work, 1u)) : static_cast<int64_t>(blockIdx.x) * stages + warp;
    uint32_t parity = 0;
    int32_t b = k < n_items ? p.idx[k / p.chunks_per_row] : 0;
    // Fine-grained completion: this warp's items ascend, so the minibatch they belong to only moves forward.  `mb_end` is
    // the first item (row * chunks_per_row, launch-relative) past the current minibatch; when an item crosses it, what the
    // warp finished of the previous minibatch is published one iteration LATER, after `wait_group 1` (every store but the
    // newest has been written) -- by then the older store has long completed, so the wait costs nothing.
    uint32_t cur_mb = 0, cur_count = 0, pend_mb = 0, pend_count = 0;
    int64_t mb_end = 0;
    uint32_t ep = 0, in_ep = 0;  // epoch and minibatch-within-epoch of cur_mb
    if (p.progress) {
      const uint32_t g0 = p.row_offset;
      ep = g0 / p.rows_per_epoch;
      in_ep = (g0 - ep * p.rows_per_epoch) / p.mb_rows;
      cur_mb = ep * p.mbs_per_epoch + in_ep;
      uint32_t end_in_ep = (in_ep + 1) * p.mb_rows;
      if (end_in_ep > p.rows_per_epoch) end_in_ep = p.rows_per_epoch;
      mb_end = (static_cast<int64_t>(ep) * p.rows_per_epoch + end_in_ep - g0) * p.chunks_per_row;
    }
    for (; k < n_items; k += stride) {
      const int64_t i = k / p.chunks_per_row;
      const int64_t off = (k - i * p.chunks_per_row) * static_cast<int64_t>(p.chunk_bytes);
      const int64_t left = p.row_bytes - off;
      const uint32_t bytes = left < static_cast<int64_t>(p.chunk_bytes) ? static_cast<uint32_t>(left) : p.chunk_bytes;
      const int64_t row = xa::sample_row(b, p.n_steps, p.n_envs);
      if (p.progress && k >= mb_end) {
        if (pend_count) {  // a minibatch this warp barely touched: its publication is still pending -- drain and publish now
          xa::bulk_wait_all<0>();
          __threadfence();
          atomicAdd(p.progress + pend_mb, pend_count);
        }
        pend_mb = cur_mb;
        pend_count = cur_count;
        cur_count = 0;
        while (k >= mb_end) {  // advance to the minibatch of item k
          if (++in_ep == p.mbs_per_epoch) {
            in_ep = 0;
            ++ep;
          }
          ++cur_mb;
          uint32_t end_in_ep = (in_ep + 1) * p.mb_rows;
          if (end_in_ep > p.rows_per_epoch) end_in_ep = p.rows_per_epoch;
          mb_end = (static_cast<int64_t>(ep) * p.rows_per_epoch + end_in_ep - p.row_offset) * p.chunks_per_row;
        }
      }
      xa::mbar_expect_tx(bar, bytes);
      if (p.load_hint)
        xa::bulk_g2s(buf, p.src + row * p.row_bytes + off, bytes, bar, policy);
      else
        xa::bulk_g2s_nohint(buf, p.src + row * p.row_bytes + off, bytes, bar);
      const int64_t kn = p.work ? static_cast<int64_t>(atomicAdd(p.work, 1u)) : k + stride;
      if (kn < n_items) b = p.idx[kn / p.chunks_per_row];  // next ticket and next index travel under the copy
      xa::mbar_wait(bar, parity);
      parity ^= 1u;
      if (p.store_hint)
        xa::bulk_s2g_hint(p.dst + i * p.row_bytes + off, buf, bytes, policy);
      else
        xa::bulk_s2g(p.dst + i * p.row_bytes + off, buf, bytes);
      xa::bulk_commit();
      ++cur_count;
      if (pend_count) {
        xa::bulk_wait_all<1>();  // everything but the store just committed has been WRITTEN: the previous minibatch's rows are out
        __threadfence();
        atomicAdd(p.progress + pend_mb, pend_count);
        pend_count = 0;
      }
      xa::bulk_wait_read<0>();  // the engine has read the stage out: it may be refilled
      if (p.work) k = kn - stride;  // the loop increment adds `stride` back: k becomes the ticket drawn above
    }
    xa::bulk_wait_all<0>();
    if (p.progress) {
      __threadfence();
      if (pend_count) atomicAdd(p.progress + pend_mb, pend_count);
      if (cur_count) atomicAdd(p.progress + cur_mb, cur_count);
    }
    if (p.work && atomicAdd(p.work + 1, 1u) == gridDim.x * static_cast<unsigned>(stages) - 1u) {
      p.work[0] = 0;  // every lane has drawn its last ticket: leave the counter ready for the next launch
      p.work[1] = 0;
    }
    return;
  }

  // last warp: scalar fields for this CTA's contiguous slice of samples
  if (p.n_fields > 0) {
    const int64_t per = (p.n_idx + gridDim.x - 1) / gridDim.x;
    const int64_t lo = per * blockIdx.x;
    const int64_t hi = lo + per < p.n_idx ? lo + per : p.n_idx;
    for (int64_t i = lo + lane; i < hi; i += 32) {
      const int64_t row = xa::sample_row(p.idx[i], p.n_steps, p.n_envs);
#pragma unroll
      for (int f = 0; f < XA_MAX_FIELDS; ++f)
        if (f < p.n_fields) p.fdst[f][i] = __ldg(p.fsrc[f] + row);
    }
  }
}

// ------------------------------------------------------------------------------------------ vector
template <typename V>
__device__ __forceinline__ V load_vec(const V* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ int4 load_vec<int4>(const int4* p) {
  return xa::ld_stream(p);
}
template <typename V>
__device__ __forceinline__ void store_vec(V* p, const V& v) {
  *p = v;
}
template <>
__device__ __forceinline__ void store_vec<int4>(int4* p, const int4& v) {
  xa::st_stream(p, v);
}

constexpr int kVecThreads = 256;
constexpr int kVecUnroll = 8;

// threads_per_row (power of two <= 256) lanes cooperate on a row; 256/threads_per_row rows per block pass.  gridDim.y > 1
// cuts every row into that many segments (few long rows: 80 frames of 28 KB would otherwise occupy 80 of 148 SMs with one
// block each and leave the latency of 7 dependent 16-B loads per thread exposed).
template <typename V>
__global__ void __launch_bounds__(kVecThreads) gather_vector_kernel(const GatherParams p, int threads_per_row_log2) {
  const int tpr = 1 << threads_per_row_log2;
  const int rows_per_block = kVecThreads >> threads_per_row_log2;
  const int sub = threadIdx.x >> threads_per_row_log2;
  const int lane = threadIdx.x & (tpr - 1);
  const int64_t n_vec = p.row_bytes / static_cast<int64_t>(sizeof(V));
  const int64_t seg_len = (n_vec + gridDim.y - 1) / gridDim.y;
  const int64_t seg_lo = seg_len * blockIdx.y;
  const int64_t seg_hi = seg_lo + seg_len < n_vec ? seg_lo + seg_len : n_vec;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * rows_per_block + sub; i < p.n_idx;
       i += static_cast<int64_t>(gridDim.x) * rows_per_block) {
    const int64_t row = xa::sample_row(p.idx[i], p.n_steps, p.n_envs);
    const V* s = reinterpret_cast<const V*>(p.src + row * p.row_bytes);
    V* d = reinterpret_cast<V*>(p.dst + i * p.row_bytes);
    for (int64_t base = seg_lo + lane; base < seg_hi; base += static_cast<int64_t>(tpr) * kVecUnroll) {
      V regs[kVecUnroll];
#pragma unroll
      for (int u = 0; u < kVecUnroll; ++u) {
        const int64_t v = base + static_cast<int64_t>(u) * tpr;
        if (v < seg_hi) regs[u] = load_vec(s + v);
      }
#pragma unroll
      for (int u = 0; u < kVecUnroll; ++u) {
        const int64_t v = base + static_cast<int64_t>(u) * tpr;
        if (v < seg_hi) store_vec(d + v, regs[u]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) gather_fields_kernel(const GatherParams p) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n_idx;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = xa::sample_row(p.idx[i], p.n_steps, p.n_envs);
#pragma unroll
    for (int f = 0; f < XA_MAX_FIELDS; ++f)
      if (f < p.n_fields) p.fdst[f][i] = __ldg(p.fsrc[f] + row);
  }
}

// uint8 frame -> fp32 / 255 (base.py:505-506) behind the gather: each thread turns 4 pixels into one float4
template <bool kVec4>
__global__ void __launch_bounds__(256) gather_scaled_kernel(const GatherParams p) {
  float* out = reinterpret_cast<float*>(p.dst);
  for (int64_t i = blockIdx.x; i < p.n_idx; i += gridDim.x) {
    const int64_t row = xa::sample_row(p.idx[i], p.n_steps, p.n_envs);
    const uint8_t* s = p.src + row * p.row_bytes;
    float* d = out + i * p.row_bytes;
    if (kVec4) {
      const int64_t n4 = p.row_bytes >> 2;
      for (int64_t v = threadIdx.x; v < n4; v += blockDim.x) {
        const uchar4 px = __ldg(reinterpret_cast<const uchar4*>(s) + v);
        float4 o;
        o.x = __fdiv_rn(static_cast<float>(px.x), 255.0f);
        o.y = __fdiv_rn(static_cast<float>(px.y), 255.0f);
        o.z = __fdiv_rn(static_cast<float>(px.z), 255.0f);
        o.w = __fdiv_rn(static_cast<float>(px.w), 255.0f);
        __stcs(reinterpret_cast<float4*>(d) + v, o);
      }
    } else {
      for (int64_t v = threadIdx.x; v < p.row_bytes; v += blockDim.x) d[v] = __fdiv_rn(static_cast<float>(s[v]), 255.0f);
    }
  }
}

// ------------------------------------------------------------------------------------------ host
// Tuning knobs of the bulk path (scripts/gather_microbench2.py sweeps them): honoured only when the process was started
// with XA_TUNING=1 (checked once); the product path never calls getenv on a launch.
struct BulkTuning {
  int64_t max_chunk = 16 * 1024;        // rows are cut into equal 16-B-multiple chunks of at most this many bytes
  int64_t target_inflight = 56 * 1024;  // bytes in flight per SM (ring size)
  int stages = 0;                       // 0 = derived from target_inflight
  int load_hint = 1, store_hint = 1;    // L2 evict-first on both sides (hinting only the loads measured 2 % slower)
  int ctas_per_sm = 1;
  int spread = 1;                       // pad the CTA's shared memory so that two gather CTAs never share an SM
  explicit BulkTuning(bool from_env) {
    if (!from_env) return;
    if (const char* e = getenv("XA_GATHER_CHUNK")) max_chunk = atoll(e) > 0 ? atoll(e) : max_chunk;
    if (const char* e = getenv("XA_GATHER_INFLIGHT")) target_inflight = atoll(e) > 0 ? atoll(e) : target_inflight;
    if (const char* e = getenv("XA_GATHER_STAGES")) stages = atoi(e) > 0 && atoi(e) <= kMaxStages ? atoi(e) : 0;
    if (const char* e = getenv("XA_GATHER_LOAD_HINT")) load_hint = atoi(e);
    if (const char* e = getenv("XA_GATHER_STORE_HINT")) store_hint = atoi(e);
    if (const char* e = getenv("XA_GATHER_CTAS")) ctas_per_sm = atoi(e) > 0 ? atoi(e) : 1;
    if (const char* e = getenv("XA_GATHER_SPREAD")) spread = atoi(e);
  }
};

struct GatherEnv {  // the two switches, read once (thread-safe static initialisation)
  bool tuning = false, debug_indices = false;
  GatherEnv() {
    if (const char* e = getenv("XA_TUNING")) tuning = atoi(e) != 0;
    if (const char* e = getenv("XA_DEBUG_INDICES")) debug_indices = atoi(e) != 0;
  }
};

const GatherEnv& gather_env() {
  static const GatherEnv e;
  return e;
}

BulkTuning bulk_tuning() {  // tuning mode re-reads the knobs on every launch (the sweep scripts change them in-process)
  return BulkTuning(gather_env().tuning);
}

// cudaFuncSetAttribute is per device: remember which devices already allow the bulk kernel its dynamic shared memory
int allow_bulk_smem(const char* what) {
  constexpr int kMaxDevices = 64;
  static bool done[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (dev >= 0 && dev < kMaxDevices && done[dev]) return XA_OK;
  cudaError_t e = cudaFuncSetAttribute(gather_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  if (e != cudaSuccess) {
    xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  if (dev >= 0 && dev < kMaxDevices) done[dev] = true;  // racing threads set the same attribute twice: harmless
  return XA_OK;
}

// Debug aid (XA_DEBUG_INDICES=1): one bad index is an out-of-bounds TMA read and a sticky context error, so the check
// runs as its own tiny kernel and the call returns XA_EINVAL instead of launching the gather.  Synchronises the stream.
__global__ void check_indices_kernel(const int32_t* idx, int64_t n, int64_t n_rows, int* bad) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    if (idx[i] < 0 || idx[i] >= n_rows) atomicAdd(bad, 1);
}

int debug_check_indices(const int32_t* idx, int64_t n, int64_t n_rows, cudaStream_t stream, const char* what) {
  int* bad = nullptr;
  if (cudaMallocAsync(&bad, sizeof(int), stream) != cudaSuccess) return XA_OK;  // cannot check: do not block the caller
  cudaMemsetAsync(bad, 0, sizeof(int), stream);
  const int64_t want = (n + 255) / 256;
  check_indices_kernel<<<static_cast<unsigned>(want < 1184 ? want : 1184), 256, 0, stream>>>(idx, n, n_rows, bad);
  int host = 0;
  cudaMemcpyAsync(&host, bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
  cudaStreamSynchronize(stream);
  cudaFreeAsync(bad, stream);
  XA_REQUIRE(host == 0, XA_EINVAL, "%s: %d of %lld indices are outside [0, %lld)", what, host, static_cast<long long>(n),
             static_cast<long long>(n_rows));
  return XA_OK;
}

int widest_vector(const GatherParams& p) {
  const uintptr_t bits = reinterpret_cast<uintptr_t>(p.src) | reinterpret_cast<uintptr_t>(p.dst) |
                         static_cast<uintptr_t>(p.row_bytes);
  if ((bits & 15) == 0) return 16;
  if ((bits & 7) == 0) return 8;
  if ((bits & 3) == 0) return 4;
  if ((bits & 1) == 0) return 2;
  return 1;
}

int log2_ceil_pow2(int64_t n, int cap_log2) {
  int l = 0;
  while ((int64_t(1) << l) < n && l < cap_log2) ++l;
  return l;
}

template <typename V>
void launch_vector(const GatherParams& p, cudaStream_t stream) {
  const int64_t n_vec = p.row_bytes / static_cast<int64_t>(sizeof(V));
  const int tpr_log2 = log2_ceil_pow2((n_vec + kVecUnroll - 1) / kVecUnroll, 8);
  const int rows_per_block = kVecThreads >> tpr_log2;
  const int64_t want = (p.n_idx + rows_per_block - 1) / rows_per_block;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int64_t cap = static_cast<int64_t>(sms) * 8 * 4;  // a few waves of 8 resident blocks per SM
  const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
  // fewer blocks than ~4 per SM: cut the rows into segments (each at least one pass of the block wide, i.e. tpr vectors)
  int64_t segments = 1;
  if (want < static_cast<int64_t>(sms) * 4) {
    segments = (static_cast<int64_t>(sms) * 4 + want - 1) / want;
    const int64_t max_seg = (n_vec + (int64_t(1) << tpr_log2) - 1) >> tpr_log2;
    if (segments > max_seg) segments = max_seg;
    if (segments > 64) segments = 64;
    if (segments < 1) segments = 1;
  }
  gather_vector_kernel<V><<<dim3(grid, static_cast<unsigned>(segments)), kVecThreads, 0, stream>>>(p, tpr_log2);
}

int launch_fields(const GatherParams& p, cudaStream_t stream) {
  if (p.n_fields == 0) return XA_OK;
  const int64_t want = (p.n_idx + 255) / 256;
  const unsigned grid = static_cast<unsigned>(want < 4096 ? want : 4096);
  gather_fields_kernel<<<grid, 256, 0, stream>>>(p);
  return xa::check_launch("xa_gather_fields_f32");
}

int chunks_per_row_for(int64_t row_bytes, int64_t max_chunk) { return static_cast<int>((row_bytes + max_chunk - 1) / max_chunk); }

bool bulk_eligible(const GatherParams& p) {
  return ((reinterpret_cast<uintptr_t>(p.src) | reinterpret_cast<uintptr_t>(p.dst) | static_cast<uintptr_t>(p.row_bytes)) & 15) == 0;
}

int launch_rows(GatherParams p, int mode, cudaStream_t stream, const char* what) {
  if (p.n_idx == 0) return XA_OK;
  if (gather_env().debug_indices)
    if (int rc = debug_check_indices(p.idx, p.n_idx, p.n_src_rows, stream, what)) return rc;
  const bool can_bulk = bulk_eligible(p);
  if (mode == XA_GATHER_BULK)
    XA_REQUIRE(can_bulk, XA_EALIGN, "%s: XA_GATHER_BULK needs 16-byte aligned src, dst and row_bytes (row_bytes=%lld)", what,
               static_cast<long long>(p.row_bytes));
  // AUTO: TMA bulk copies for long aligned rows once there is enough work to fill the ring on every SM; a few rows (the
  // 80-frame env-major reorder of config C2) finish sooner on the segmented vector path
  const bool enough = p.n_idx * p.row_bytes >= (int64_t(16) << 20);
  const bool use_bulk = mode == XA_GATHER_BULK || (mode == XA_GATHER_AUTO && can_bulk && p.row_bytes >= 2048 && enough);
  if (use_bulk) {
    // Bytes in flight per SM decide the rate, and more is NOT better: measured on B200 (scripts/
    // gather_microbench2.py, 32768 rows of 28224 B, random permutation) ~55 KB per SM peaks at 6.45 TB/s
    // (98 % of a plain copy) while 110-226 KB per SM falls to 5.9-6.0 TB/s.  So rows are cut into equal
    // 16-B-multiple chunks of <= 16 KB and the ring holds ~56 KB: 4 stages of half a frame for 84x84x4.
    // (XA_GATHER_* environment variables are tuning knobs for the microbenchmarks, honoured only under XA_TUNING=1.)
    const BulkTuning tune = bulk_tuning();
    p.chunks_per_row = chunks_per_row_for(p.row_bytes, tune.max_chunk);
    int64_t chunk = (p.row_bytes + p.chunks_per_row - 1) / p.chunks_per_row;
    chunk = (chunk + 15) & ~int64_t(15);
    p.chunk_bytes = static_cast<uint32_t>(chunk);
    int stages = static_cast<int>((tune.target_inflight + chunk / 2) / chunk);
    if (stages < 2) stages = 2;
    if (stages > kMaxStages) stages = kMaxStages;
    while (stages > 1 && static_cast<int64_t>(stages) * chunk + 8 * kMaxStages > kSmemBudget) --stages;
    if (tune.stages > 0 && static_cast<int64_t>(tune.stages) * chunk + 8 * kMaxStages <= kSmemBudget) stages = tune.stages;
    p.stages = stages;
    p.load_hint = tune.load_hint;
    p.store_hint = tune.store_hint;
    size_t smem = static_cast<size_t>(stages) * chunk + 8 * kMaxStages;
    // One CTA per SM is the design (grid = SM count), but the block scheduler does not promise it: launched while other kernels
    // are resident, the 148 CTAs were seen packed two per SM on half of the SMs for the whole persistent launch (measured: 3.7
    // instead of 6.5 TB/s -- the per-SM load/store path, not HBM, then bounds it).  Asking for more than half of an SM's shared
    // memory makes a second gather CTA on the same SM impossible.
    const size_t spread_smem = 116 * 1024;
    if (tune.spread && tune.ctas_per_sm <= 1 && smem < spread_smem) smem = spread_smem;
    if (int rc = allow_bulk_smem(what)) return rc;
    const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
    const int ctas_per_sm = static_cast<int>(kSmemBudget / (static_cast<size_t>(stages) * chunk + 8 * kMaxStages + 1024)) > 0
                                ? static_cast<int>(kSmemBudget / (static_cast<size_t>(stages) * chunk + 8 * kMaxStages + 1024))
                                : 1;
    const int64_t items = p.n_idx * p.chunks_per_row;
    int64_t grid = (items + stages - 1) / stages;
    const int cta_cap = tune.ctas_per_sm <= ctas_per_sm ? tune.ctas_per_sm : ctas_per_sm;
    const int64_t cap = static_cast<int64_t>(sms) * cta_cap;
    if (grid > cap) grid = cap;
    gather_bulk_kernel<<<static_cast<unsigned>(grid), 32 * (stages + 1), smem, stream>>>(p);
    return xa::check_launch(what);
  }
  switch (widest_vector(p)) {
    case 16: launch_vector<int4>(p, stream); break;
    case 8: launch_vector<int2>(p, stream); break;
    case 4: launch_vector<int>(p, stream); break;
    case 2: launch_vector<short>(p, stream); break;
    default: launch_vector<char>(p, stream); break;
  }
  if (int rc = xa::check_launch(what)) return rc;
  return launch_fields(p, stream);
}

int fill_common(GatherParams& p, const char* what, const void* src, const int32_t* idx, void* dst, int64_t n_idx,
                int64_t row_bytes, int64_t n_src_rows, int n_steps, int n_envs) {
  XA_REQUIRE(n_idx >= 0 && row_bytes > 0 && n_src_rows > 0, XA_EINVAL, "%s: n_idx=%lld row_bytes=%lld n_src_rows=%lld", what,
             static_cast<long long>(n_idx), static_cast<long long>(row_bytes), static_cast<long long>(n_src_rows));
  XA_REQUIRE(n_idx == 0 || (src && idx && dst), XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(xa::aligned(idx, 4), XA_EALIGN, "%s: idx must be 4-byte aligned", what);
  XA_REQUIRE(n_steps >= 0 && n_envs >= 0, XA_EINVAL, "%s: negative n_steps/n_envs", what);
  if (n_steps > 0)
    XA_REQUIRE(static_cast<int64_t>(n_steps) * n_envs == n_src_rows, XA_EINVAL, "%s: n_steps*n_envs=%lld != n_src_rows=%lld", what,
               static_cast<long long>(n_steps) * n_envs, static_cast<long long>(n_src_rows));
  XA_REQUIRE(n_src_rows <= INT32_MAX, XA_EOVERFLOW, "%s: n_src_rows exceeds int32 indices", what);
  p = GatherParams{};
  p.src = static_cast<const uint8_t*>(src);
  p.dst = static_cast<uint8_t*>(dst);
  p.idx = idx;
  p.n_idx = n_idx;
  p.row_bytes = row_bytes;
  p.n_steps = n_steps;
  p.n_envs = n_envs;
  p.n_src_rows = n_src_rows;
  return XA_OK;
}

int fill_fields(GatherParams& p, const char* what, const float* const* src, float* const* dst, int n_fields) {
  XA_REQUIRE(n_fields >= 0 && n_fields <= XA_MAX_FIELDS, XA_EINVAL, "%s: n_fields=%d not in [0,%d]", what, n_fields, XA_MAX_FIELDS);
  XA_REQUIRE(n_fields == 0 || (src && dst), XA_EINVAL, "%s: null field table", what);
  p.n_fields = n_fields;
  for (int f = 0; f < n_fields; ++f) {
    XA_REQUIRE(src[f] && dst[f], XA_EINVAL, "%s: null pointer for field %d", what, f);
    XA_REQUIRE(xa::aligned(src[f], 4) && xa::aligned(dst[f], 4), XA_EALIGN, "%s: field %d not 4-byte aligned", what, f);
    p.fsrc[f] = src[f];
    p.fdst[f] = dst[f];
  }
  return XA_OK;
}

}  // namespace

extern "C" {

int xa_gather_rows(const void* src, const int32_t* idx, void* dst, int64_t n_idx, int64_t row_bytes, int64_t n_src_rows,
                   int n_steps, int n_envs, int mode, xa_stream_t stream) {
  GatherParams p;
  if (int rc = fill_common(p, "xa_gather_rows", src, idx, dst, n_idx, row_bytes, n_src_rows, n_steps, n_envs)) return rc;
  XA_REQUIRE(mode >= XA_GATHER_AUTO && mode <= XA_GATHER_VECTOR, XA_EINVAL, "xa_gather_rows: unknown mode %d", mode);
  return launch_rows(p, mode, static_cast<cudaStream_t>(stream), "xa_gather_rows");
}

int xa_gather_fields_f32(const float* const* src, float* const* dst, int n_fields, const int32_t* idx, int64_t n_idx, int n_steps,
                         int n_envs, xa_stream_t stream) {
  XA_REQUIRE(n_idx >= 0, XA_EINVAL, "xa_gather_fields_f32: n_idx=%lld", static_cast<long long>(n_idx));
  XA_REQUIRE(n_idx == 0 || idx, XA_EINVAL, "xa_gather_fields_f32: null idx");
  XA_REQUIRE(n_steps >= 0 && n_envs >= 0, XA_EINVAL, "xa_gather_fields_f32: negative n_steps/n_envs");
  GatherParams p{};
  p.idx = idx;
  p.n_idx = n_idx;
  p.n_steps = n_steps;
  p.n_envs = n_envs;
  if (int rc = fill_fields(p, "xa_gather_fields_f32", src, dst, n_fields)) return rc;
  if (n_idx == 0) return XA_OK;
  return launch_fields(p, static_cast<cudaStream_t>(stream));
}

int xa_gather_minibatch(const void* obs_src, void* obs_dst, int64_t row_bytes, int64_t n_src_rows, const float* const* field_src,
                        float* const* field_dst, int n_fields, const int32_t* idx, int64_t n_idx, int n_steps, int n_envs,
                        int mode, xa_stream_t stream) {
  GatherParams p;
  if (int rc = fill_common(p, "xa_gather_minibatch", obs_src, idx, obs_dst, n_idx, row_bytes, n_src_rows, n_steps, n_envs)) return rc;
  XA_REQUIRE(mode >= XA_GATHER_AUTO && mode <= XA_GATHER_VECTOR, XA_EINVAL, "xa_gather_minibatch: unknown mode %d", mode);
  if (int rc = fill_fields(p, "xa_gather_minibatch", field_src, field_dst, n_fields)) return rc;
  return launch_rows(p, mode, static_cast<cudaStream_t>(stream), "xa_gather_minibatch");
}

int xa_gather_progress_units(int64_t row_bytes) {
  if (row_bytes <= 0) return 0;
  return chunks_per_row_for(row_bytes, bulk_tuning().max_chunk);
}

int xa_gather_rows_progress(const void* src, const int32_t* idx, void* dst, int64_t n_idx, int64_t row_bytes, int64_t n_src_rows,
                            int n_steps, int n_envs, uint32_t* progress, int64_t row_offset, int64_t rows_per_epoch, int64_t mb_rows,
                            uint32_t* work, xa_stream_t stream) {
  GatherParams p;
  if (int rc = fill_common(p, "xa_gather_rows_progress", src, idx, dst, n_idx, row_bytes, n_src_rows, n_steps, n_envs)) return rc;
  XA_REQUIRE(progress == nullptr || xa::aligned(progress, 4), XA_EALIGN, "xa_gather_rows_progress: progress must be 4-byte aligned");
  XA_REQUIRE(rows_per_epoch > 0 && mb_rows > 0 && mb_rows <= rows_per_epoch && row_offset >= 0, XA_EINVAL,
             "xa_gather_rows_progress: rows_per_epoch=%lld mb_rows=%lld row_offset=%lld", static_cast<long long>(rows_per_epoch),
             static_cast<long long>(mb_rows), static_cast<long long>(row_offset));
  XA_REQUIRE(row_offset + n_idx < (int64_t(1) << 32), XA_EOVERFLOW, "xa_gather_rows_progress: row_offset + n_idx exceeds 32 bits");
  XA_REQUIRE(work == nullptr || xa::aligned(work, 8), XA_EALIGN, "xa_gather_rows_progress: work must be 8-byte aligned (two uint32 words)");
  XA_REQUIRE(work == nullptr || n_idx * chunks_per_row_for(row_bytes, bulk_tuning().max_chunk) < (int64_t(1) << 31), XA_EOVERFLOW,
             "xa_gather_rows_progress: too many items for a 32-bit work counter");
  p.work = work;
  p.progress = progress;
  p.row_offset = static_cast<uint32_t>(row_offset);
  p.rows_per_epoch = static_cast<uint32_t>(rows_per_epoch);
  p.mb_rows = static_cast<uint32_t>(mb_rows);
  p.mbs_per_epoch = static_cast<uint32_t>((rows_per_epoch + mb_rows - 1) / mb_rows);
  return launch_rows(p, XA_GATHER_BULK, static_cast<cudaStream_t>(stream), "xa_gather_rows_progress");
}

// Device-side wait: a one-warp kernel that spins (nanosleep back-off) until *addr - target >= 0 (cyclic), so that the stream's
// next kernel starts within a microsecond of the producer's publication.  cuStreamWaitValue32 does the same in the front-end,
// but on B200 a wait that lasts more than ~0.5 ms is re-evaluated only every ~3 ms (measured: n_envs=2048, 0.6 ms per
// minibatch -> steps of 9.1 or 12.2 ms), which is useless for a pipeline.  Bounded (~2 s): a producer that never publishes
// sets *status = 1 instead of hanging the GPU.
__global__ void wait_progress_kernel(const uint32_t* addr, uint32_t target, int* status) {
  if (threadIdx.x != 0) return;
  const long long t0 = clock64();
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    if (static_cast<int32_t>(v - target) >= 0) return;
    if (clock64() - t0 > 4000000000LL) {
      if (status) atomicExch(status, 1);
      return;
    }
    __nanosleep(100);
  }
}

// cuStreamWaitValue32(stream, addr, value, GEQ): the stream's later work starts once *addr - value >= 0 (cyclic 32-bit compare).
// Resolved through the runtime (no link-time dependency on libcuda).
int xa_stream_wait_geq_u32(xa_stream_t stream, const uint32_t* addr, uint32_t value) {
  typedef int (*wait_fn)(void*, unsigned long long, unsigned int, unsigned int);
  static wait_fn fn = nullptr;
  static bool looked = false;
  if (!looked) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<wait_fn>(sym);
    else
      cudaGetLastError();
    looked = true;
  }
  XA_REQUIRE(fn != nullptr, XA_EINVAL, "xa_stream_wait_geq_u32: the driver does not export cuStreamWaitValue32");
  XA_REQUIRE(addr != nullptr && xa::aligned(addr, 4), XA_EINVAL, "xa_stream_wait_geq_u32: bad address");
  const int rc = fn(stream, reinterpret_cast<unsigned long long>(addr), value, 1u /* CU_STREAM_WAIT_VALUE_GEQ */);
  if (rc != 0) {
    xa::set_error("xa_stream_wait_geq_u32: cuStreamWaitValue32 failed with CUresult %d", rc);
    return rc;
  }
  return XA_OK;
}

int xa_wait_progress_u32(const uint32_t* addr, uint32_t target, int* status, xa_stream_t stream) {
  XA_REQUIRE(addr != nullptr && xa::aligned(addr, 4), XA_EINVAL, "xa_wait_progress_u32: bad address");
  wait_progress_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(addr, target, status);
  return xa::check_launch("xa_wait_progress_u32");
}

int xa_gather_rows_u8_scaled_f32(const uint8_t* src, const int32_t* idx, float* dst, int64_t n_idx, int64_t row_bytes,
                                 int64_t n_src_rows, int n_steps, int n_envs, xa_stream_t stream) {
  GatherParams p;
  if (int rc = fill_common(p, "xa_gather_rows_u8_scaled_f32", src, idx, dst, n_idx, row_bytes, n_src_rows, n_steps, n_envs)) return rc;
  XA_REQUIRE(xa::aligned(dst, 4), XA_EALIGN, "xa_gather_rows_u8_scaled_f32: dst must be 4-byte aligned");
  if (n_idx == 0) return XA_OK;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int64_t cap = static_cast<int64_t>(sms) * 8 * 4;
  const unsigned grid = static_cast<unsigned>(n_idx < cap ? n_idx : cap);
  const bool vec4 = (row_bytes & 3) == 0 && xa::aligned(src, 4) && xa::aligned(dst, 16);
  if (vec4)
    gather_scaled_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  else
    gather_scaled_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return xa::check_launch("xa_gather_rows_u8_scaled_f32");
}

}  // extern "C"
