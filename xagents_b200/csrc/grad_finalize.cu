// Last kernel of the network's backward pass: every weight / bias gradient goes from where its kernel left it (split
// partial sums, kernel-friendly layouts) to its place in the ONE flat fp32 gradient buffer the fused clip + Adam step reads
// (tape.gradient -> clip_by_global_norm -> apply_gradients, xagents/ppo/agent.py:134-137).
//
//   grad[dest[j]] = sum over s < splits(j) of  src[map[j] + s * stride(j)]        (0 where map[j] < 0; dest = NULL: j itself)
//
// (map, dest) is the permutation between the kernels' operand layouts and the parameter layouts (torch / Keras order),
// built once on the host from index tensors (agents/tc_plan.py); splits / stride are constant over a SEGMENT of
// consecutive j (one segment per parameter tensor).  The caller enumerates each segment in SOURCE order: the partial sums
// are then read with consecutive lanes on consecutive addresses (148 splits x 47 MB in this network) and only the single
// result per output is a scattered 4-byte store -- enumerated in destination order the same kernel took 57 us instead of
// ~12, every 4-byte read pulling its own 32-byte sector.  Splits are added in order with a fixed association, so the result is
// deterministic.  This one launch replaces three split reductions, four layout copies, two torch bias reductions, the
// zero-fill of the gradient buffer and autograd's twelve accumulation kernels.
#include "xa_common.cuh"

namespace {

constexpr int kThreads = 256;

struct FinalizeParams {
  const float* src;
  const int32_t* map;
  const int32_t* dest;
  float* grad;
  int64_t n;
  int n_segments, n_wide;
  int64_t wide_slots;       // lane slots of all wide segments (each padded to whole warps)
  xa_grad_segment_t seg[XA_MAX_GRAD_SEGMENTS];
  int32_t wide_seg[XA_MAX_GRAD_SEGMENTS];         // the w-th wide segment's index in seg
  int64_t wide_begin[XA_MAX_GRAD_SEGMENTS + 1];   // prefix sums of the wide segments' lane slots
  int64_t seg_len[XA_MAX_GRAD_SEGMENTS];
};

__device__ __forceinline__ int find_segment(const FinalizeParams& p, int64_t j) {
  int s = 0;
#pragma unroll 1
  while (s + 1 < p.n_segments && j >= p.seg[s + 1].dest_begin) ++s;
  return s;
}

// blocks [0, narrow_blocks): one thread per output of a segment with few splits; the remaining blocks: `wide` (4, 8, 16 or 32)
// lanes per output of the segments with many splits -- each lane a contiguous share of the splits, combined by a fixed
// shuffle tree.  A segment's lane slots are padded to whole warps, so a warp works on one segment with one lane count.
// More lanes are for segments with few outputs and very many splits (the heads kernel's ~300 per-CTA blocks: at four lanes the
// 74 loads per lane of those 4 600 outputs were the tail of the whole kernel); lanes of one output read different splits, so a
// warp-wide load touches 32 / lanes consecutive outputs per sector: eight lanes keep half of every sector useful.
__global__ void __launch_bounds__(kThreads) grad_finalize_kernel(const __grid_constant__ FinalizeParams p, int narrow_blocks) {
  xa::pdl_trigger();   // chained launch (xa_common.cuh): the successor may be scheduled early;
  xa::pdl_wait();      // the predecessor grid has completed before anything below touches global memory
  if (static_cast<int>(blockIdx.x) < narrow_blocks) {
    // One output per thread and pass, software-pipelined: the (map, dest) entries of the NEXT pass are requested before this
    // pass's partial sums, and up to four splits are loaded before the first add (split order kept).  Unpipelined, every output
    // cost three DRAM latencies in series (map -> partial sums -> dest) and this half -- the FC weight's 1.6 M outputs, five
    // passes per thread -- was 45 % of the kernel (ncu source page).
    const int64_t step = static_cast<int64_t>(narrow_blocks) * kThreads;
    int64_t j = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
    int32_t m_next = -1, d_next = 0;
    int splits_next = -1;                    // splits < 0: nothing to write (a wide segment's output)
    int64_t stride_next = 0;
    auto fetch = [&](int64_t jj) {
      m_next = -1, splits_next = -1;
      if (jj < p.n) {
        const xa_grad_segment_t& sg = p.seg[find_segment(p, jj)];
        if (!sg.wide) {
          m_next = __ldg(p.map + jj);
          d_next = p.dest != nullptr ? __ldg(p.dest + jj) : static_cast<int32_t>(jj);
          splits_next = sg.splits, stride_next = sg.split_stride;
        }
      }
    };
    fetch(j);
    for (; j < p.n; j += step) {
      const int32_t m = m_next, d = d_next;
      const int splits = splits_next;
      const int64_t stride = stride_next;
      fetch(j + step);
      if (splits < 0) continue;
      float acc = 0.0f;
      if (m >= 0) {
        const float* src = p.src + m;
        for (int k0 = 0; k0 < splits; k0 += 4) {
          float v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = k0 + i < splits ? __ldg(src + (k0 + i) * stride) : 0.0f;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (k0 + i < splits) acc += v[i];
        }
      }
      p.grad[d] = acc;
    }
    return;
  }
  const int64_t first = static_cast<int64_t>(blockIdx.x - narrow_blocks) * kThreads + threadIdx.x;
  const int64_t step = static_cast<int64_t>(gridDim.x - narrow_blocks) * kThreads;
  for (int64_t t = first; t < p.wide_slots; t += step) {   // wide_slots is a multiple of 32: whole warps reach the shuffles
    int w = 0;
    while (t >= p.wide_begin[w + 1]) ++w;
    const int si = p.wide_seg[w];
    const xa_grad_segment_t& sg = p.seg[si];
    const int lanes = sg.wide;
    const int64_t local = t - p.wide_begin[w];
    const int64_t i = local / lanes;
    const int part = static_cast<int>(local - i * lanes);
    float acc = 0.0f;
    int64_t j = -1;
    if (i < p.seg_len[si]) {
      j = sg.dest_begin + i;
      const int32_t m = p.map[j];
      if (m >= 0) {
        const int per = (sg.splits + lanes - 1) / lanes;
        const int k0 = part * per, k1 = k0 + per < sg.splits ? k0 + per : sg.splits;
        const float* src = p.src + m;
#pragma unroll 8
        for (int k = k0; k < k1; ++k) acc += src[k * sg.split_stride];   // loads do not depend on acc: eight in flight
      }
    }
    for (int o = 1; o < lanes; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // ((q0 + q1) + (q2 + q3)) + ...
    if (part == 0 && j >= 0) p.grad[p.dest != nullptr ? p.dest[j] : j] = acc;
  }
}

}  // namespace

extern "C" int xa_grad_finalize_f32(const float* src, const int32_t* map, const int32_t* dest, const xa_grad_segment_t* segments,
                                    int n_segments, float* grad, int64_t n, xa_stream_t stream) {
  const char* what = "xa_grad_finalize_f32";
  XA_REQUIRE(src && map && segments && grad, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(n > 0 && n_segments > 0 && n_segments <= XA_MAX_GRAD_SEGMENTS, XA_EINVAL, "%s: n=%lld n_segments=%d (at most %d)", what,
             static_cast<long long>(n), n_segments, XA_MAX_GRAD_SEGMENTS);
  FinalizeParams p{};
  p.src = src, p.map = map, p.dest = dest, p.grad = grad, p.n = n, p.n_segments = n_segments;
  int64_t slots = 0;
  int n_wide = 0;
  for (int s = 0; s < n_segments; ++s) {
    const xa_grad_segment_t& sg = segments[s];   // host memory
    const int64_t end = s + 1 < n_segments ? segments[s + 1].dest_begin : n;
    XA_REQUIRE(sg.dest_begin >= 0 && sg.dest_begin <= end && (s > 0 || sg.dest_begin == 0) && sg.splits >= 1 && sg.split_stride >= 0, XA_EINVAL,
               "%s: segment %d (dest_begin=%lld splits=%d) out of order or empty", what, s, static_cast<long long>(sg.dest_begin), sg.splits);
    XA_REQUIRE(sg.wide == 0 || sg.wide == 1 || sg.wide == 4 || sg.wide == 8 || sg.wide == 16 || sg.wide == 32, XA_EINVAL,
               "%s: segment %d: wide=%d (lanes per output: 0, 4, 8, 16 or 32; 1 = 4)", what, s, sg.wide);
    p.seg[s] = sg;
    p.seg_len[s] = end - sg.dest_begin;
    if (sg.wide) {
      if (sg.wide == 1) p.seg[s].wide = 4;
      p.wide_seg[n_wide] = s;
      p.wide_begin[n_wide++] = slots;
      slots += (((end - sg.dest_begin) * p.seg[s].wide + 31) / 32) * 32;
    }
  }
  p.n_wide = n_wide;
  p.wide_begin[n_wide] = slots;
  for (int w = n_wide + 1; w <= XA_MAX_GRAD_SEGMENTS; ++w) p.wide_begin[w] = INT64_MAX;
  p.wide_slots = slots;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  int64_t nb = (n + kThreads - 1) / kThreads;
  if (nb > 8 * sms) nb = 8 * sms;
  int64_t wb = (slots + kThreads - 1) / kThreads;
  if (wb > 16 * sms) wb = 16 * sms;
  xa::launch_chained(xa::kChainElementwise, grad_finalize_kernel, dim3(static_cast<unsigned>(nb + wb)), dim3(kThreads), 0, static_cast<cudaStream_t>(stream), p, static_cast<int>(nb));
  return xa::check_launch(what);
}
