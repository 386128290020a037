// conv1 (5x5, rate 1, Ci = 3..5 input channels, isprs:766 / 966 / 918) on the tensor cores for 16-bit inference.
//
// The generic kernel (conv_tc.cuh) needs Ci to be a multiple of 32; padding 5 channels to 32 makes the im2col
// traffic through L2 6.4x larger than the data and the layer L2-bound.  Here the input is padded to 8 channels
// (one 16-byte im2col row per pixel and tap) and the operands use the NON-swizzled K-major canonical layout, whose
// unit is the 8-row x 16-byte core matrix: a TMA im2col box of 128 pixels x 8 channels lands as 16 consecutive core
// matrices, and two taps (two boxes, 2048 bytes apart = the descriptor's leading byte offset) form one K = 16 MMA.
//   K = 26 taps x 8 (tap 25 is a zero slot so that K is a multiple of 16) = 13 MMAs per 128-pixel tile
//   B = [Co][208] packed once on the host side in core-matrix order, copied to shared memory at kernel start
// One pipeline stage = one whole tile (25 TMA loads on one mbarrier).  Epilogue as in conv_tc.cuh.
#pragma once
#include "conv_tc.cuh"

constexpr int C1_TAPS = 25, C1_SLOTS = 26, C1_SLOT_BYTES = CONV_TC_BM * 16, C1_STAGE_BYTES = C1_SLOTS * C1_SLOT_BYTES;
constexpr int C1_KP = C1_SLOTS * 8;   // 208

struct Conv1TcParams {
  int M_total, crop, num_tiles;
  int co, out_coff;
  int stages;
  int acc_stride, tmem_cols;
  int act;
  int smem_needed, smem_provided;
  uint32_t idesc;
  const void* wpack;    // [26][Co/8][8][8] 16-bit (core-matrix order)
  const float* scale;
  const float* shift;
  uint32_t* diag;
  BnFinish stats;       // training forward: batch statistics of the rounded output reduced in the epilogue (acc == nullptr: off)
};

// W HWIO [5,5,ci,co] fp32 -> core-matrix order: out[((kc*(co/8) + n/8)*8 + n%8)*8 + c] = (kc < 25 && c < ci) ? W[kc][c][n] : 0
template <typename T>
__global__ void pack_conv1_kernel(const float* __restrict__ w, T* __restrict__ out, int ci, int co) {
  pdl_sync();
  const int n_total = C1_SLOTS * co * 8;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  const int c = i & 7;
  const int r = i >> 3;
  const int nl = r & 7;
  const int g = r >> 3;
  const int ng = g % (co / 8), kc = g / (co / 8);
  const int n = ng * 8 + nl;
  out[i] = from_f32<T>((kc < C1_TAPS && c < ci) ? w[((int64_t)kc * ci + c) * co + n] : 0.0f);
}

// x [M][C] fp32 -> [M][8] 16-bit, zero padded
template <typename T>
__global__ void pad_cast8_kernel(const float* __restrict__ x, T* __restrict__ out, int C, int64_t M) {
  pdl_sync();
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  alignas(16) T v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = from_f32<T>(e < C ? x[m * C + e] : 0.0f);
  *reinterpret_cast<uint4*>(out + m * 8) = *reinterpret_cast<const uint4*>(v);
}

template <int EPI_C, typename OutT>
__global__ void __launch_bounds__(CONV_TC_THREADS, 1)
conv1_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC, const Conv1TcParams p) {
  constexpr int STG_BYTES = CONV_TC_BM * EPI_C * 2;
  constexpr int MAX_STAGES = 4;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  if (pad + static_cast<uint32_t>(p.smem_needed) > static_cast<uint32_t>(p.smem_provided)) {
    if (threadIdx.x == 0 && p.diag) { p.diag[0] = 0xBAD00003u; p.diag[1] = raw_addr; __threadfence_system(); }
    __trap();
  }
  const int b_bytes = p.co * C1_KP * 2;
  uint8_t* s_b = smem + p.stages * C1_STAGE_BYTES;
  uint8_t* stg = s_b + ((b_bytes + 1023) & ~1023);
  float* s_scale = reinterpret_cast<float*>(stg + 2 * STG_BYTES);
  float* s_shift = s_scale + 64;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_shift + 64);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler (see conv_tc.cuh)
  const int lane = threadIdx.x & 31;

  pdl_sync();      // the packed weights below are written by the kernel right in front of this one
  // weights (core-matrix order) and the zero slot of every stage: generic-proxy writes, fenced for the async proxy
  for (int i = threadIdx.x; i < b_bytes / 16; i += CONV_TC_THREADS)
    reinterpret_cast<uint4*>(s_b)[i] = reinterpret_cast<const uint4*>(p.wpack)[i];
  for (int s = 0; s < p.stages; ++s)
    for (int i = threadIdx.x; i < C1_SLOT_BYTES / 16; i += CONV_TC_THREADS)
      reinterpret_cast<uint4*>(smem + s * C1_STAGE_BYTES + C1_TAPS * C1_SLOT_BYTES)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < p.co; i += CONV_TC_THREADS) {
    s_scale[i] = p.scale[i];
    s_shift[i] = p.shift[i];
  }
  ptx::fence_proxy_async_smem();
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 4);
    }
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: 25 taps per tile
    // (convergent warp, elected lane issues: see conv_tc.cuh)
    {
      const bool leader = ptx::elect_one_sync();
      int stage = 0;
      uint32_t phase = 0;
      const int cc = p.crop * p.crop;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = tile * CONV_TC_BM;
        const int n_img = m0 / cc;
        const int rem = m0 - n_img * cc;
        const int py = rem / p.crop;
        const int px = rem - py * p.crop;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1, p.diag, 0x900 + stage);
        uint8_t* sa = smem + stage * C1_STAGE_BYTES;
        if (leader) {
          ptx::mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(C1_TAPS * C1_SLOT_BYTES));
#pragma unroll
          for (int t = 0; t < C1_TAPS; ++t)
            ptx::tma_load_im2col_4d(sa + t * C1_SLOT_BYTES, &tmA, &full_bar[stage], 0, px - 2, py - 2, n_img,
                                    static_cast<uint16_t>(t % 5), static_cast<uint16_t>(t / 5));
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: 13 x (M128, N=Co, K16)
    {
      const bool leader = ptx::elect_one_sync();
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t sb = ptx::smem_u32(s_b);
      const uint32_t b_lbo = static_cast<uint32_t>(p.co / 8) * 128u;      // next K core matrix of B
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[as], aphase ^ 1, p.diag, 0xA00 + as);
        ptx::mbar_wait(&full_bar[stage], phase, p.diag, 0xB00 + stage);
        ptx::tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * p.acc_stride);
        const uint32_t sa = ptx::smem_u32(smem + stage * C1_STAGE_BYTES);
        if (leader) {
#pragma unroll
          for (int j = 0; j < C1_SLOTS / 2; ++j) {
            // no swizzle, K-major: LBO = distance between the two K core matrices, SBO = next 8 rows (128 B)
            const uint64_t adesc = ptx::make_smem_desc(sa + j * 2 * C1_SLOT_BYTES, C1_SLOT_BYTES, 128, 0u);
            const uint64_t bdesc = ptx::make_smem_desc(sb + j * 2 * b_lbo, b_lbo, 128, 0u);
            ptx::umma_f16(d_tmem, adesc, bdesc, p.idesc, static_cast<uint32_t>(j != 0));
          }
          ptx::umma_commit(&empty_bar[stage]);
          ptx::umma_commit(&tmem_full[as]);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (same as conv_tc.cuh)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int epi_tid = threadIdx.x - 128;
    // one elected lane of warp 4 owns the TMA stores (bulk groups are per thread: elect.sync picks the same lane for the
    // same member mask every time); elected at each site so that ptxas sees a single active thread
    constexpr int CHUNK16 = EPI_C / 8;
    const int sw = (EPI_C == 64) ? (row & 7) : ((row >> 1) & 3);
    const int n_chunks = p.co / EPI_C;
    int as = 0;
    uint32_t aphase = 0;
    int sbuf = 0;
    float st0 = 0.0f, st1 = 0.0f;          // this thread's (channel, row group) sums of z and z^2 (statistics mode)
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_full[as], aphase, p.diag, 0xC00 + as);
      ptx::tcgen05_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(as * p.acc_stride);
      for (int ch = 0; ch < n_chunks; ++ch) {
        uint32_t v[EPI_C];
        ptx::tmem_ld_32x32b_x32(t_row + ch * EPI_C, v);
        if (EPI_C == 64) ptx::tmem_ld_32x32b_x32(t_row + ch * EPI_C + 32, v + (EPI_C == 64 ? 32 : 0));
        ptx::tmem_wait_ld();
        if (warp == 4 && ptx::elect_one_sync()) ptx::tma_store_wait_read<1>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint8_t* srow = stg + sbuf * STG_BYTES + row * (EPI_C * 2);
        const float* sc = s_scale + ch * EPI_C;
        const float* sh = s_shift + ch * EPI_C;
#pragma unroll
        for (int j = 0; j < CHUNK16; ++j) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c0 = j * 8 + e * 2;
            float a = apply_act(fmaf(__uint_as_float(v[c0]), sc[c0], sh[c0]), p.act);
            float b = apply_act(fmaf(__uint_as_float(v[c0 + 1]), sc[c0 + 1], sh[c0 + 1]), p.act);
            o[e] = pack2<OutT>(a, b);
          }
          *reinterpret_cast<uint4*>(srow + ((j ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 4 && ptx::elect_one_sync()) {
          ptx::tma_store_2d(&tmC, stg + sbuf * STG_BYTES, p.out_coff + ch * EPI_C, tile * CONV_TC_BM);
          ptx::tma_store_commit();
        }
        if (p.stats.acc) {
          // column sums of the staged (already rounded) tile, as in conv_tc.cuh: thread = (channel, row group); EPI_C == Co
          // here, so there is a single chunk and one register pair per thread
          constexpr int ROWS = EPI_C;
          const int c = epi_tid % EPI_C, grp = epi_tid / EPI_C;
          const int rows_valid = min(CONV_TC_BM, p.M_total - tile * CONV_TC_BM);
          const uint8_t* col = stg + sbuf * STG_BYTES + (c & 7) * 2;
          const int c16 = c >> 3;
          float a0 = 0.0f, a1 = 0.0f;
          const int r_end = min(rows_valid, (grp + 1) * ROWS);
          for (int r = grp * ROWS; r < r_end; ++r) {
            const int swr = (EPI_C == 64) ? (r & 7) : ((r >> 1) & 3);
            const float f = to_f32(*reinterpret_cast<const OutT*>(col + r * (EPI_C * 2) + ((c16 ^ swr) << 4)));
            a0 += f;
            a1 = fmaf(f, f, a1);
          }
          st0 += a0;
          st1 += a1;
        }
        sbuf ^= 1;
      }
      ptx::tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (warp == 4 && ptx::elect_one_sync()) ptx::tma_store_wait_all();
    if (p.stats.acc) {
      // CTA partial (row groups in fixed order) -> 64-bit fixed point -> integer atomics; the last CTA finalises (conv_tc.cuh)
      constexpr int GROUPS = CONV_TC_BM / EPI_C;
      float* s_stat = reinterpret_cast<float*>(stg);            // the staging buffers are free: every TMA store has drained
      asm volatile("bar.sync 1, 128;" ::: "memory");
      {
        const int c = epi_tid % EPI_C, grp = epi_tid / EPI_C;
        s_stat[(grp * 2) * p.co + c] = st0;
        s_stat[(grp * 2 + 1) * p.co + c] = st1;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int i = epi_tid; i < 2 * p.co; i += 128) {
        const int which = i / p.co, c = i - which * p.co;
        float a = 0.0f;
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) a += s_stat[(g * 2 + which) * p.co + c];
        const long long q = __double2ll_rn((double)a * p.stats.fx_scale);
        atomicAdd(bn_acc_mine(p.stats) + i, static_cast<unsigned long long>(q));
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint32_t* s_flag = reinterpret_cast<uint32_t*>(s_stat);
      if (epi_tid == 0) *s_flag = (atomicAdd(p.stats.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*s_flag) {
        __threadfence();
        const double inv_scale = 1.0 / p.stats.fx_scale;
        const int C = p.co;
        for (int i = epi_tid; i < C; i += 128) {
          const double a = (double)bn_acc_take(p.stats, i) * inv_scale;
          const double b = (double)bn_acc_take(p.stats, C + i) * inv_scale;
          p.stats.sums[i] = (float)a;
          p.stats.sums[C + i] = (float)b;
          if (p.stats.mean) {
            const double mu = (double)(float)a / p.stats.count;
            double var = (double)(float)b / p.stats.count - mu * mu;
            if (var < 0.0) var = 0.0;
            p.stats.mean[i] = (float)mu;
            p.stats.inv_std[i] = (float)(1.0 / sqrt(var + (double)p.stats.eps));
            const double var_ema = p.stats.unbiased_ema ? var * (p.stats.count / fmax(p.stats.count - 1.0, 1.0)) : var;
            p.stats.mov_mean[i] = p.stats.decay * p.stats.mov_mean[i] + (1.0f - p.stats.decay) * (float)mu;
            p.stats.mov_var[i] = p.stats.decay * p.stats.mov_var[i] + (1.0f - p.stats.decay) * (float)var_ema;
          }
        }
        if (epi_tid == 0) *p.stats.counter = 0u;
      }
    }
  }

  ptx::tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tcgen05_fence_after();
    ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

struct Conv1TcArgs {
  const void* x8;       // [B,crop,crop,8] fp16/bf16 (channels >= C are zero)
  const void* wpack;    // pack_conv1_kernel output
  void* out;            // NHWC [M, out_cstride]
  int out_cstride, out_coff, co;
  int B, crop;
  const float* scale;
  const float* shift;
  int act, etype;
  const BnFinish* stats = nullptr;     // training forward: fused batch statistics
};

template <int EPI_C, typename OutT>
static void launch_conv1_tc_t(Handle* h, const Conv1TcArgs& a) {
  alignas(64) CUtensorMap tmA, tmC;
  const int64_t M = (int64_t)a.B * a.crop * a.crop;
  encode_im2col(h, &tmA, a.etype, a.x8, 8, a.crop, a.B, 2, 8, CONV_TC_BM, 0);
  encode_tiled_2d(h, &tmC, a.etype, a.out, (uint64_t)a.out_cstride, (uint64_t)M, (uint64_t)a.out_cstride * 2, EPI_C, CONV_TC_BM,
                  EPI_C * 2);
  Conv1TcParams p;
  p.M_total = (int)M; p.crop = a.crop; p.num_tiles = (int)ceil_div(M, CONV_TC_BM);
  p.co = a.co; p.out_coff = a.out_coff;
  p.acc_stride = a.co <= 32 ? 32 : 64;
  p.tmem_cols = 2 * p.acc_stride;
  p.act = a.act;
  p.idesc = make_idesc_f16(128, a.co, a.etype == ET_BF16, a.etype == ET_BF16, 0, 0);
  p.wpack = a.wpack; p.scale = a.scale; p.shift = a.shift; p.diag = h->diag_dev;
  if (a.stats) p.stats = *a.stats; else memset(&p.stats, 0, sizeof(p.stats));
  const int b_bytes = ((a.co * C1_KP * 2) + 1023) & ~1023;
  const int fixed = b_bytes + 2 * CONV_TC_BM * EPI_C * 2 + 2 * 64 * 4 + (2 * 4 + 4) * 8 + 16;
  const int budget = 227 * 1024;
  int stages = (budget - 1024 - fixed) / C1_STAGE_BYTES;
  if (stages > 4) stages = 4;
  DRS_CHECK(stages >= 2, "conv1_tc: does not fit shared memory");
  p.stages = stages;
  p.smem_needed = fixed + stages * C1_STAGE_BYTES;
  const int smem_bytes = std::min(p.smem_needed + 1024, budget);
  p.smem_provided = smem_bytes;
  auto kern = conv1_tc_kernel<EPI_C, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, budget));
    attr_set = true;
  }
  const int grid = p.num_tiles < h->sm_count ? p.num_tiles : h->sm_count;
  launch_pdl(h, kern, dim3(grid), dim3(CONV_TC_THREADS), (size_t)smem_bytes, tmA, tmC, p);
  LAUNCH_CHECK(h);
}

static inline bool conv1_tc_supported(int k, int rate, int ci, int co) { return k == 5 && rate == 1 && ci <= 8 && (co == 64 || co == 32); }

static void launch_conv1_tc(Handle* h, const Conv1TcArgs& a) {
  DRS_CHECK(a.co == 64 || a.co == 32, "conv1_tc: Co=%d", a.co);
  DRS_CHECK(a.out_cstride % 8 == 0 && a.out_coff % 8 == 0, "conv1_tc: output alignment");
  if (a.etype == ET_F16) {
    if (a.co == 64) launch_conv1_tc_t<64, __half>(h, a); else launch_conv1_tc_t<32, __half>(h, a);
  } else {
    if (a.co == 64) launch_conv1_tc_t<64, __nv_bfloat16>(h, a); else launch_conv1_tc_t<32, __nv_bfloat16>(h, a);
  }
}
