// Weight gradient of a dilated SAME convolution on the 5th-gen tensor cores (training, bf16 operands).
//
// Reference op: the backward of tf.nn.atrous_conv2d w.r.t. its filter inside MomentumOptimizer.minimize
// (/root/reference/isprs_dilated_random.py:710, 1687).
//
//   dW[tap*Ci + c, o] = sum_m  X[pix(m) + tap_offset, c] * dZ[m, o]          (zero outside the image)
//
// As a GEMM the reduction runs over the pixels m, so BOTH operands are "MN-major" for tcgen05.mma: the
// reduction index is the strided one.  That is exactly what TMA delivers without any transposition:
//   A^T  an im2col load of 64 consecutive output pixels x 64 input channels of one filter tap
//        -> shared memory [64 px][128 B], 128B-swizzled: rows = K (pixels), 128-byte row = 64 M values
//   B    a tiled load of the same 64 pixels x 64 output channels of dZ -> same shape, N values per row
// One UMMA tile is M = 128 = two "row blocks" (tap, 64-channel block) -- e.g. two channel blocks of one
// tap (Ci >= 128) or two taps (Ci = 64) -- and N = Co.  A CTA keeps T such tiles in TMEM (T*Co <= 512
// columns) so one dZ tile feeds T MMAs, and walks a contiguous range of pixel chunks (split-K over
// pixels).  Partial sums go to part[split][K][Co] in fp32 and are reduced in split order by
// reduce_partials_kernel: deterministic, no atomics.
//
// Warp roles as in conv_tc.cuh: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4..7 epilogue (TMEM -> registers -> global fp32).
#pragma once
#include "conv_tc.cuh"

struct WgradTcParams {
  int M_total;         // pixels
  int crop;
  int ksize, rate, pad_b;
  int ci, in_coff;     // input channels (multiple of 64), offset inside the input buffer
  int co, dy_coff;     // N (multiple of 64, <= 256), offset inside the dZ buffer
  int n_rb;            // row blocks = taps * ci / 64
  int n_tiles;         // ceil(n_rb / 2)
  int T;               // tiles per work item
  int n_groups;        // ceil(n_tiles / T)
  int splits;          // pixel splits
  int chunks_per_split;
  int n_chunks;        // ceil(M / 64)
  int stages;
  int acc_stride;      // TMEM columns per tile
  int tmem_cols;
  int smem_needed, smem_provided;
  uint32_t idesc;
  float* part;         // [splits][taps*ci][co]
  uint32_t* diag;
};

constexpr int WG_BPX = 64;   // pixels per pipeline stage

__global__ void __launch_bounds__(CONV_TC_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const WgradTcParams p) {
  constexpr int RB_BYTES = WG_BPX * 64 * 2;   // one [64 px][64 ch] box = 8 KB
  constexpr int MAX_STAGES = 8;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  if (pad + static_cast<uint32_t>(p.smem_needed) > static_cast<uint32_t>(p.smem_provided)) {
    if (threadIdx.x == 0 && p.diag) { p.diag[0] = 0xBAD00002u; p.diag[1] = raw_addr; __threadfence_system(); }
    __trap();
  }
  const int nb_boxes = p.co / 64;
  const int a_bytes = p.T * 2 * RB_BYTES;            // T tiles x 2 row blocks
  const int b_bytes = nb_boxes * RB_BYTES;
  const int stage_bytes = a_bytes + b_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler (see conv_tc.cuh)
  const int lane = threadIdx.x & 31;
  const int n_items = p.n_groups * p.splits;
  const int cb_per_tap = p.ci / 64;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmX);
    ptx::prefetch_tensormap(&tmDY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::mbar_init(tmem_empty, 4);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_holder, static_cast<uint32_t>(p.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tcgen05_fence_before();
  __syncthreads();
  ptx::tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_sync();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (convergent warp, elected lane
    // issues: see conv_tc.cuh)
    {
      const bool leader = ptx::elect_one_sync();
      int stage = 0;
      uint32_t phase = 0;
      const int cc = p.crop * p.crop;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int g = item / p.splits, sp = item - g * p.splits;
        const int q0 = sp * p.chunks_per_split;
        const int q1 = min(p.n_chunks, q0 + p.chunks_per_split);
        const int tile0 = g * p.T;
        for (int q = q0; q < q1; ++q) {
          const int m0 = q * WG_BPX;
          const int n_img = m0 / cc;
          const int rem = m0 - n_img * cc;
          const int py = rem / p.crop;
          const int px = rem - py * p.crop;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.diag, 0x500 + stage);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          if (leader) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
            int rb = tile0 * 2;
            int tap = rb / cb_per_tap, cb = rb - tap * cb_per_tap;
            int ky = tap / p.ksize, kx = tap - ky * p.ksize;
            for (int i = 0; i < 2 * p.T; ++i) {
              // (the padding block of an odd last tile repeats the last row block: its rows are discarded)
              ptx::tma_load_im2col_4d(sa + i * RB_BYTES, &tmX, &full_bar[stage], p.in_coff + cb * 64, px - p.pad_b, py - p.pad_b,
                                      n_img, static_cast<uint16_t>(kx * p.rate), static_cast<uint16_t>(ky * p.rate));
              if (rb + 1 < p.n_rb) {
                ++rb;
                if (++cb == cb_per_tap) {
                  cb = 0;
                  if (++kx == p.ksize) { kx = 0; ++ky; }
                }
              }
            }
            for (int nb = 0; nb < nb_boxes; ++nb)
              ptx::tma_load_2d(sb + nb * RB_BYTES, &tmDY, &full_bar[stage], p.dy_coff + nb * 64, m0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, elected lane issues)
    {
      const bool leader = ptx::elect_one_sync();
      int stage = 0;
      uint32_t phase = 0, aphase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int g = item / p.splits, sp = item - g * p.splits;
        const int q0 = sp * p.chunks_per_split;
        const int q1 = min(p.n_chunks, q0 + p.chunks_per_split);
        ptx::mbar_wait(tmem_empty, aphase ^ 1, p.diag, 0x600);
        ptx::tcgen05_fence_after();
        for (int q = q0; q < q1; ++q) {
          ptx::mbar_wait(&full_bar[stage], phase, p.diag, 0x700 + stage);
          ptx::tcgen05_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          // MN-major, 128B swizzle: LBO = distance between 64-element M/N blocks, SBO = 8 K-rows (1024 B)
          const uint64_t bdesc = ptx::make_smem_desc(sb, RB_BYTES, 1024, 2u);
          if (leader) {
            for (int t = 0; t < p.T; ++t) {
              const uint64_t adesc = ptx::make_smem_desc(sa + t * 2 * RB_BYTES, RB_BYTES, 1024, 2u);
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(t * p.acc_stride);
#pragma unroll
              for (int j = 0; j < WG_BPX / 16; ++j) {
                // 16 pixels = 2048 bytes along K: +128 in the (>>4) start-address field
                ptx::umma_f16(d_tmem, adesc + static_cast<uint64_t>(j * 128), bdesc + static_cast<uint64_t>(j * 128), p.idesc,
                              static_cast<uint32_t>((q != q0) || (j != 0)));
              }
            }
            ptx::umma_commit(&empty_bar[stage]);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (leader) ptx::umma_commit(tmem_full);
        aphase ^= 1;
        (void)g;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> fp32 partials
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int Ktot = p.ksize * p.ksize * p.ci;
    uint32_t aphase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int g = item / p.splits, sp = item - g * p.splits;
      ptx::mbar_wait(tmem_full, aphase, p.diag, 0x800);
      ptx::tcgen05_fence_after();
      for (int t = 0; t < p.T; ++t) {
        const int tile = g * p.T + t;
        const int kk = tile * 128 + row;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(t * p.acc_stride);
        float* dst = p.part + ((size_t)sp * Ktot + kk) * p.co;
        for (int c0 = 0; c0 < p.co; c0 += 32) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_row + c0, v);     // warp-collective: every lane participates
          ptx::tmem_wait_ld();
          if (tile < p.n_tiles && kk < Ktot) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                     __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty);
      aphase ^= 1;
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct WgradTcArgs {
  const void* x;        // NHWC [B,crop,crop,in_cstride] bf16
  int in_cstride, in_coff, ci;
  const void* dy;       // [M, dy_cstride] bf16
  int dy_cstride, dy_coff, co;
  int B, crop, k, rate, pad_b;
  float* dw;            // [k*k*ci][co] fp32 (HWIO)
  float* part;          // workspace for split partials
  size_t part_capacity; // floats
  int out_rows = 0;     // > 0: only the first out_rows rows of dW are wanted (conv1 through its materialised im2col)
};

// conv1 (5x5, Ci <= 5 image channels): its 25*Ci <= 125 reduction rows are far too few for the kernel above as a 25-tap
// layer (Ci must be a multiple of 64), and the CUDA-core kernel that served it (wgrad_conv1_kernel, 72 + 15 us at batch 64,
// crop 37) sat at the very end of the backward where nothing overlaps it.  Instead the im2col matrix is materialised once
// per step -- xcol[m][tap*Ci + c], 128 bf16 per pixel, zero padded -- and the filter gradient is the 1x1 "layer" Ci = 128
// of the tensor-core kernel: dW[0 .. 25*Ci) = xcol^T dZ.
template <typename T>
__global__ void __launch_bounds__(256)
im2col_conv1_kernel(const T* __restrict__ x8, T* __restrict__ xcol, int ci, int crop, int64_t M) {
  pdl_sync();
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * 16) return;
  const int64_t m = gid >> 4;
  const int g = (int)(gid & 15);           // 16-byte group of the 256-byte row: elements 8g .. 8g+7
  const int cc = crop * crop;
  const int r = (int)(m % cc);
  const int y = r / crop, x = r - y * crop;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (ci == 4) {
    // two taps per group, one 8-byte load each (x8 rows are 16 bytes: channels 0..3 are the low half)
    uint2 v[2] = {make_uint2(0u, 0u), make_uint2(0u, 0u)};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int tap = 2 * g + j;
      const int ky = tap / 5, kx = tap - ky * 5;
      const int yy = y + ky - 2, xx = x + kx - 2;
      if (tap < 25 && yy >= 0 && yy < crop && xx >= 0 && xx < crop)
        v[j] = *reinterpret_cast<const uint2*>(x8 + (m + (int64_t)(ky - 2) * crop + (kx - 2)) * 8);
    }
    o = make_uint4(v[0].x, v[0].y, v[1].x, v[1].y);
  } else {
    alignas(16) T e8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = g * 8 + e;
      const int tap = k / ci, c = k - tap * ci;
      const int ky = tap / 5, kx = tap - ky * 5;
      const int yy = y + ky - 2, xx = x + kx - 2;
      T val = from_f32<T>(0.0f);
      if (tap < 25 && yy >= 0 && yy < crop && xx >= 0 && xx < crop) val = x8[(m + (int64_t)(ky - 2) * crop + (kx - 2)) * 8 + c];
      e8[e] = val;
    }
    o = *reinterpret_cast<const uint4*>(e8);
  }
  *reinterpret_cast<uint4*>(xcol + m * 128 + g * 8) = o;
}

// Channel counts of 32 (the first layers of the ICPR nets, isprs:791-1033) run as 64: the TMA boxes are 64 channels wide
// whatever the tensor holds -- past the end of a 32-channel tensor the box is zero-filled, inside the dense nets' 448-wide
// concat buffer it picks up 32 foreign (finite) channels -- and the rows / columns of dW that belong to the padding are
// dropped by the reduction (reduce_partials_remap_kernel).  Twice the MMA work of a tiny layer, instead of the CUDA-core
// kernel that held up the whole step (CUPTI timeline, DenseDilated6 batch 16 crop 49: 695 us of a 1455 us step).
static inline bool wgrad_tc_supported(int ci, int co) {
  return (ci % 64 == 0 || ci == 32) && (co % 64 == 0 || co == 32) && co <= 256 && !(getenv("DRS_NO_WGRAD_PAD") && (ci == 32 || co == 32));
}

// out[(tap*ci + c)*co + o] = sum_s part[s][(tap*ci_pad + c)*co_pad + o]   (splits in ascending order: deterministic)
__global__ void __launch_bounds__(256)
reduce_partials_remap_kernel(const float* __restrict__ part, float* __restrict__ out, int S, int64_t stride, int rows_out, int ci,
                             int ci_pad, int co, int co_pad) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_out * co) return;
  const int r = i / co, o = i - r * co;
  const int tap = r / ci, c = r - tap * ci;
  const float* src = part + ((int64_t)tap * ci_pad + c) * co_pad + o;
  float a = 0.0f;
#pragma unroll 8
  for (int k = 0; k < S; ++k) a += __ldcs(src + (int64_t)k * stride);
  out[i] = a;
}


static void launch_wgrad_tc(Handle* h, const WgradTcArgs& a) {
  DRS_CHECK(wgrad_tc_supported(a.ci, a.co), "wgrad_tc: unsupported Ci=%d Co=%d", a.ci, a.co);
  DRS_CHECK(a.in_cstride % 8 == 0 && a.dy_cstride % 8 == 0 && a.in_coff % 8 == 0 && a.dy_coff % 8 == 0, "wgrad_tc: alignment");
  const int64_t M = (int64_t)a.B * a.crop * a.crop;
  const int taps = a.k * a.k;
  const int ci_pad = (a.ci + 63) / 64 * 64, co_pad = (a.co + 63) / 64 * 64;    // 32 -> 64 (see wgrad_tc_supported)
  const int Ktot = taps * ci_pad;
  alignas(64) CUtensorMap tmX, tmDY;
  encode_im2col(h, &tmX, ET_BF16, a.x, a.in_cstride, a.crop, a.B, a.pad_b, 64, WG_BPX, 128);
  encode_tiled_2d(h, &tmDY, ET_BF16, a.dy, (uint64_t)a.dy_cstride, (uint64_t)M, (uint64_t)a.dy_cstride * 2, 64, WG_BPX, 128);

  WgradTcParams p;
  p.M_total = (int)M; p.crop = a.crop; p.ksize = a.k; p.rate = a.rate; p.pad_b = a.pad_b;
  p.ci = ci_pad; p.in_coff = a.in_coff; p.co = co_pad; p.dy_coff = a.dy_coff;
  p.n_rb = taps * (ci_pad / 64);
  p.n_tiles = (p.n_rb + 1) / 2;
  p.acc_stride = co_pad <= 64 ? 64 : co_pad <= 128 ? 128 : 256;
  // tiles per work item: bounded by TMEM (512 columns) and by a >= 3-stage shared-memory pipeline
  int T = 512 / p.acc_stride;
  const int budget = 227 * 1024 - 2048;
  while (T > 1 && 3 * (T * 2 * 8192 + (co_pad / 64) * 8192) > budget) --T;
  if (T > p.n_tiles) T = p.n_tiles;
  p.T = T;
  p.tmem_cols = 32;
  while (p.tmem_cols < T * p.acc_stride) p.tmem_cols *= 2;
  p.n_groups = (p.n_tiles + T - 1) / T;
  p.n_chunks = (int)ceil_div(M, WG_BPX);
  // pixel splits: one work item per SM (every extra split costs a K*Co fp32 partial write + read), keep >= 8 chunks per
  // split, fit the workspace
  int splits = std::max(1, h->sm_count / p.n_groups);
  splits = std::min<int>(splits, std::max(1, p.n_chunks / 8));
  splits = std::min<int64_t>(splits, (int64_t)(a.part_capacity / ((size_t)Ktot * co_pad)));
  DRS_CHECK(splits >= 1, "wgrad_tc: workspace too small");
  p.chunks_per_split = (int)ceil_div(p.n_chunks, splits);
  splits = (int)ceil_div(p.n_chunks, p.chunks_per_split);
  p.splits = splits;
  const int stage_bytes = T * 2 * 8192 + (co_pad / 64) * 8192;
  int stages = budget / stage_bytes;
  if (stages > 8) stages = 8;
  DRS_CHECK(stages >= 2, "wgrad_tc: stage does not fit shared memory");
  p.stages = stages;
  p.smem_needed = stages * stage_bytes + (2 * 8 + 2) * 8 + 16;
  const int smem_bytes = std::min(p.smem_needed + 1024, 227 * 1024);
  p.smem_provided = smem_bytes;
  p.idesc = make_idesc_f16(128, co_pad, 1, 1, 1, 1);   // bf16 x bf16, both operands MN-major
  p.part = a.part;
  p.diag = h->diag_dev;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int n_items = p.n_groups * p.splits;
  const int grid = std::min(n_items, h->sm_count);
  launch_pdl(h, wgrad_tc_kernel, dim3(grid), dim3(CONV_TC_THREADS), (size_t)smem_bytes, tmX, tmDY, p);
  LAUNCH_CHECK(h);
  const int rows_out = a.out_rows > 0 ? a.out_rows : taps * a.ci;
  const int64_t n = (int64_t)rows_out * a.co;
  if (ci_pad != a.ci || co_pad != a.co) {
    launch_pdl(h, reduce_partials_remap_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), 0, (const float*)a.part, a.dw, splits,
               (int64_t)Ktot * co_pad, rows_out, a.ci, ci_pad, a.co, co_pad);
  } else {
    launch_pdl(h, reduce_partials_kernel, dim3(reduce_partials_grid(n)), dim3(reduce_partials_block(n, splits)), 0, (const float*)a.part, a.dw, n, splits,
               (int64_t)Ktot * a.co);
  }
  LAUNCH_CHECK(h);
}
