// Convolution weight gradient as an implicit GEMM on the tensor cores, with no im2col matrix.
//
//   dW[n, (kh, kw, c)] = sum over output pixels (b, y, x) of  dY[b, y, x, n] * X[b, y + kh, x + kw, c]
//
// Put dY on the INPUT grid (zero where y >= OH or x >= OW) and flatten pixels as q = (b*H + y)*W + x.  A kernel tap is
// then nothing but a shift of the flattened pixel index, q -> q + kh*W + kw (whenever the shift would wrap into the
// next row or image, dY is zero), so with both operands transposed to pixel-contiguous form
//
//   dW_tap[n, c] = sum_q dYt[n, q] * Xt[c, q + shift_tap]        dYt [N, Q], Xt [C, Q]   (bf16, K-major)
//
// every tap is the same K-major GEMM over q with a shifted TMA window.  TMA wants the window start 16-byte aligned, so
// the grid rows are padded to a multiple of 8 pixels (kh*W is then aligned) and the kw part of the shift is carried
// by KW pre-shifted copies of the small operand (dYt_kw[n, q] = dYt[n, q - kw], written by xa_place_on_grid_t_bf16):
//   dW[n, (kh,kw,c)] = sum_q dYt_kw[n, q] * Xt[c, q + kh*W].  One CTA
// owns a slice of the q range (split-K over the SMs) and keeps the accumulators of ALL taps in TMEM (taps*C <= 512
// columns), so dYt is fetched once per slice and the overlapping Xt windows of the taps come out of L2.  Per slice the
// fp32 partial [N, taps*C] goes to a workspace; a second pass adds the slices in order (deterministic).
// HBM traffic: (N + C) * Q * 2 bytes, against 2 * taps * C * Q * 2 for a materialised im2col operand.
#include "tc_common.cuh"

namespace {

using namespace xa_tc;

struct WgradParams {
  float* partial;  // [splits, n_out, ld_out]
  int n_out, C, KW, kh0, n_kh, ld_out;  // this launch covers kernel rows kh0 .. kh0+n_kh-1, all KW columns
  int grid_w;                           // padded grid width (multiple of 8): row shift of one kernel row
  int64_t q_total;
  int kb_per_split, splits;
};

__global__ void __launch_bounds__(kThreads) wgrad_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                         const __grid_constant__ CUtensorMap map_x, const WgradParams p, int stages) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // The KW shifted copies of dYt are stacked row-wise ([KW*n_out, ld]), so one 128-row box holds 128/n_out of them
  // and ONE MMA (M = 128) produces the gradients of that many kernel columns at once.
  const int n_abox = (p.KW * p.n_out + kBlockM - 1) / kBlockM;
  const uint32_t tile_a = kBlockM * kBlockK * 2;
  const uint32_t tile_b = static_cast<uint32_t>(p.C) * kBlockK * 2;   // Xt tile of one kernel row: C rows
  const uint32_t stage_bytes = n_abox * tile_a + p.n_kh * tile_b;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* acc_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t k_blocks = (p.q_total + kBlockK - 1) / kBlockK;
  const int split = blockIdx.x;
  const int64_t kb0 = static_cast<int64_t>(split) * p.kb_per_split;
  const int64_t kb1 = kb0 + p.kb_per_split < k_blocks ? kb0 + p.kb_per_split : k_blocks;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < stages; ++s) {
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

  if (warp == 0) {
    if (elect_one()) {  // ---- TMA producer
      uint32_t it = 0;
      for (int64_t kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % stages;
        const uint32_t round = it / stages;
        if (round > 0) mbar_wait_wd(empty + s, (round - 1) & 1);
        uint8_t* a_dst = smem + static_cast<size_t>(s) * stage_bytes;
        xa::mbar_expect_tx(full + s, stage_bytes);
        const int q0 = static_cast<int>(kb * kBlockK);
        for (int a = 0; a < n_abox; ++a) tma_load_2d(a_dst + a * tile_a, &map_dy, q0, a * kBlockM, full + s);
        for (int j = 0; j < p.n_kh; ++j)
          tma_load_2d(a_dst + n_abox * tile_a + j * tile_b, &map_x, q0 + (p.kh0 + j) * p.grid_w, 0, full + s);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // ---- MMA issuer: one accumulator column block per tap
      const uint32_t idesc = make_idesc(kBlockM, p.C);
      uint32_t it = 0;
      for (int64_t kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % stages;
        mbar_wait_wd(full + s, (it / stages) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint8_t* base = smem + static_cast<size_t>(s) * stage_bytes;
        for (int j = 0; j < p.n_kh; ++j) {
          const uint64_t db = make_smem_desc(base + n_abox * tile_a + j * tile_b);
          for (int a = 0; a < n_abox; ++a) {
            const uint64_t da = make_smem_desc(base + a * tile_a);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16(tmem_base + (a * p.n_kh + j) * p.C, da + 2 * k, db + 2 * k, idesc, (kb > kb0) || (k != 0));
          }
        }
        umma_commit(empty + s);
      }
      umma_commit(acc_full);
    }
  } else {
    // ---- epilogue: fp32 partial of this slice
    const int quad = warp & 3;
    mbar_wait_wd(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int a = 0; a < n_abox; ++a) {
      const int r = a * kBlockM + quad * 32 + lane;  // stacked row = (kw, n)
      const int kw = r / p.n_out, n = r - kw * p.n_out;
      const bool valid = kw < p.KW;
      for (int j = 0; j < p.n_kh; ++j) {
        float* dst = p.partial + (static_cast<int64_t>(split) * p.n_out + n) * p.ld_out +
                     (static_cast<int64_t>(p.kh0 + j) * p.KW + kw) * p.C;
#pragma unroll 1
        for (int c0 = 0; c0 < p.C; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (a * p.n_kh + j) * p.C + c0, v);
          if (valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              reinterpret_cast<float4*>(dst + c0)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                    __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          }
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

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int64_t total,
                                                            int splits) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float acc = 0.0f;
    for (int s = 0; s < splits; ++s) acc += partial[static_cast<int64_t>(s) * total + i];
    dw[i] = acc;
  }
}

// Place an NHWC tensor [B, OH, OW, kN] (rows in natural (b, y, x) order, or in the 2x2 space-to-depth order
// (b, y/2, x/2, y%2, x%2) of a layer written with out_s2d) on a [B, H, W] pixel grid, transposed: out[c, q] with
// q = (b*H + y)*W + x holds the value of source pixel (y, x - x_off), zero outside the source.  `copies` outputs
// (blockIdx.y) are written back to back with x_off = 0, 1, ...
// No shared memory: warp w owns channels 8w..8w+7, lane l owns grid pixels 8l..8l+7 of the block's 256; each thread
// loads eight 16-B vectors (its 8 pixels x 8 channels), transposes the 8x8 block of bf16 in registers (byte_perm) and
// stores eight 16-B vectors, so a warp writes 512 contiguous bytes per channel row; the eight warps of a block read
// the same 128-B pixel rows, which L1 serves.
template <int kN>
__global__ void __launch_bounds__(4 * kN) place_on_grid_t_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ out,
                                                                  int B, int H, int W, int OH, int OW, int64_t ld, int s2d_order) {
  const int lane = threadIdx.x & 31, vec = threadIdx.x >> 5;  // kN / 8 warps
  const int x_off = blockIdx.y;
  out += static_cast<int64_t>(blockIdx.y) * kN * ld;
  const int64_t Q = static_cast<int64_t>(B) * H * W;
  const int64_t q0 = static_cast<int64_t>(blockIdx.x) * 256 + lane * 8;
  uint4 rows[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t q = q0 + j;
    rows[j] = make_uint4(0, 0, 0, 0);
    if (q < Q) {
      const int x = static_cast<int>(q % W) - x_off, y = static_cast<int>((q / W) % H);
      const int64_t b = q / (static_cast<int64_t>(W) * H);
      if (y < OH && x >= 0 && x < OW) {
        const int64_t r = s2d_order ? ((b * (OH / 2) + y / 2) * (OW / 2) + x / 2) * 4 + (y & 1) * 2 + (x & 1)
                                    : (b * OH + y) * OW + x;
        rows[j] = __ldg(reinterpret_cast<const uint4*>(src + r * kN) + vec);
      }
    }
  }
  if (q0 >= ld) return;  // ld is a multiple of 8: a group of 8 pixels is all-in or all-out
#pragma unroll
  for (int u = 0; u < 8; ++u) {  // channel 8*vec + u: one bf16 from each of the 8 pixel rows
    const uint32_t sel = (u & 1) ? 0x7632u : 0x5410u;
    uint4 o;
    const uint32_t* w0 = reinterpret_cast<const uint32_t*>(&rows[0]);
    (void)w0;
    auto word = [&](int j) { return reinterpret_cast<const uint32_t*>(&rows[j])[u >> 1]; };
    o.x = __byte_perm(word(0), word(1), sel);
    o.y = __byte_perm(word(2), word(3), sel);
    o.z = __byte_perm(word(4), word(5), sel);
    o.w = __byte_perm(word(6), word(7), sel);
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(vec * 8 + u) * ld + q0) = o;
  }
}

template <int kN>
void launch_place(dim3 grid, cudaStream_t s, const __nv_bfloat16* in, __nv_bfloat16* o, int batch, int grid_h, int grid_w, int src_h,
                  int src_w, int64_t ld, int s2d_order) {
  place_on_grid_t_kernel<kN><<<grid, 4 * kN, 0, s>>>(in, o, batch, grid_h, grid_w, src_h, src_w, ld, s2d_order);
}

}  // namespace

extern "C" {

int xa_place_on_grid_t_bf16(const void* src, void* out, int channels, int batch, int grid_h, int grid_w, int src_h, int src_w,
                            int64_t ld, int s2d_order, int copies, xa_stream_t stream) {
  const char* what = "xa_place_on_grid_t_bf16";
  XA_REQUIRE(src && out, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(channels == 32 || channels == 64 || channels == 128, XA_EINVAL, "%s: channels=%d (32, 64 or 128 supported)", what, channels);
  XA_REQUIRE(batch > 0 && src_h > 0 && src_w > 0 && src_h <= grid_h && src_w + copies - 1 <= grid_w && copies > 0 && copies <= 8, XA_EINVAL,
             "%s: bad shape", what);
  const int64_t Q = static_cast<int64_t>(batch) * grid_h * grid_w;
  XA_REQUIRE(ld >= Q && ld % 8 == 0, XA_EINVAL, "%s: ld=%lld must be a multiple of 8 and >= %lld", what, static_cast<long long>(ld),
             static_cast<long long>(Q));
  XA_REQUIRE(!s2d_order || (src_h % 2 == 0 && src_w % 2 == 0), XA_EINVAL, "%s: s2d order needs an even source size", what);
  XA_REQUIRE(xa::aligned(src, 16) && xa::aligned(out, 16), XA_EALIGN, "%s: 16-byte alignment required", what);
  const dim3 grid(static_cast<unsigned>((ld + 255) / 256), static_cast<unsigned>(copies));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(src);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (channels == 128)
    launch_place<128>(grid, s, in, o, batch, grid_h, grid_w, src_h, src_w, ld, s2d_order);
  else if (channels == 64)
    launch_place<64>(grid, s, in, o, batch, grid_h, grid_w, src_h, src_w, ld, s2d_order);
  else
    launch_place<32>(grid, s, in, o, batch, grid_h, grid_w, src_h, src_w, ld, s2d_order);
  return xa::check_launch(what);
}

int64_t xa_conv_wgrad_workspace_bytes(int n_out, int channels, int kh, int kw) {
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  return static_cast<int64_t>(sms) * n_out * kh * kw * channels * static_cast<int64_t>(sizeof(float));
}

int xa_conv_wgrad_bf16(const void* dyt, const void* xt, float* dw, int n_out, int channels, int kh, int kw, int grid_w, int64_t q_total,
                       int64_t ld, void* workspace, int64_t workspace_bytes, xa_stream_t stream) {
  const char* what = "xa_conv_wgrad_bf16";
  XA_REQUIRE(dyt && xt && dw && workspace, XA_EINVAL, "%s: null pointer", what);
  XA_REQUIRE(n_out > 0 && n_out <= kBlockM && (channels == 64 || channels == 128) && kh > 0 && kw > 0 && kw <= 8 && ((kw * n_out + 127) / 128) * channels <= 512,
             XA_EINVAL, "%s: n_out=%d channels=%d kernel %dx%d not supported", what, n_out, channels, kh, kw);
  XA_REQUIRE(grid_w % 8 == 0, XA_EALIGN, "%s: grid_w=%d must be a multiple of 8 (TMA windows start 16-byte aligned)", what, grid_w);
  XA_REQUIRE(q_total > 0 && ld >= q_total && ld % 8 == 0, XA_EINVAL, "%s: bad pitch", what);
  XA_REQUIRE(xa::aligned(dyt, 16) && xa::aligned(xt, 16) && xa::aligned(dw, 16) && xa::aligned(workspace, 16), XA_EALIGN,
             "%s: 16-byte alignment required", what);
  XA_REQUIRE(workspace_bytes >= xa_conv_wgrad_workspace_bytes(n_out, channels, kh, kw), XA_ENOSPACE, "%s: workspace too small", what);
  const int ld_out = kh * kw * channels;
  const int sms = xa::sm_count() > 0 ? xa::sm_count() : 148;
  const int64_t k_blocks = (q_total + kBlockK - 1) / kBlockK;
  int splits = static_cast<int>(k_blocks < sms ? k_blocks : sms);
  const int kb_per = static_cast<int>((k_blocks + splits - 1) / splits);
  splits = static_cast<int>((k_blocks + kb_per - 1) / kb_per);
  CUtensorMap mdy, mx;
  if (int rc = make_map_2d(&mdy, dyt, static_cast<int64_t>(kw) * n_out, ld, kBlockM, what)) return rc;  // KW shifted copies, stacked
  if (int rc = make_map_2d(&mx, xt, channels, ld, channels, what)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static thread_local int configured_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_dev != dev) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e != cudaSuccess) {
      xa::set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    configured_dev = dev;
  }
  const int n_abox = (kw * n_out + kBlockM - 1) / kBlockM;
  const int kh_per_launch = 512 / (n_abox * channels);  // TMEM: n_abox * n_kh * C accumulator columns <= 512
  for (int kh0 = 0; kh0 < kh; kh0 += kh_per_launch) {
    WgradParams p{};
    p.partial = static_cast<float*>(workspace);
    p.n_out = n_out, p.C = channels, p.KW = kw, p.kh0 = kh0, p.ld_out = ld_out, p.grid_w = grid_w, p.q_total = q_total;
    p.n_kh = kh - kh0 < kh_per_launch ? kh - kh0 : kh_per_launch;
    p.kb_per_split = kb_per, p.splits = splits;
    const size_t stage = static_cast<size_t>(n_abox) * kBlockM * kBlockK * 2 + static_cast<size_t>(p.n_kh) * channels * kBlockK * 2;
    int stages = static_cast<int>((200 * 1024) / stage);
    if (stages > 6) stages = 6;
    XA_REQUIRE(stages >= 2, XA_EINVAL, "%s: stage of %zu bytes does not fit twice in shared memory", what, stage);
    const size_t smem = stages * stage + 1024 + 256;
    wgrad_kernel<<<splits, kThreads, smem, s>>>(mdy, mx, p, stages);
    if (int rc = xa::check_launch(what)) return rc;
  }
  const int64_t total = static_cast<int64_t>(n_out) * ld_out;
  wgrad_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(static_cast<const float*>(workspace), dw, total, splits);
  return xa::check_launch(what);
}

}  // extern "C"
