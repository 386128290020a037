// HBM-bound layers of the path: 3x3 stride-1 SAME max-pool (fwd/bwd), batch-norm without gamma/beta
// (statistics, apply, backward), 1x1 classifier (fwd/bwd), softmax cross-entropy (+grad, argmax),
// momentum update.  All reductions are two-stage with a fixed order (no float atomics), so every
// result is run-to-run deterministic.  Activations are NHWC with a channel stride/offset so that the
// dense net's concat (isprs:921-948) is a view, never a copy.
#pragma once
#include "drs_common.cuh"

template <typename T>
struct alignas(16) Vec8 {
  T v[8];
};

// ------------------------------------------------------------------------------------------------
// _max_pool(kernel 3x3, stride 1, SAME)  isprs:745-750.  Padding never wins (window clipped).
// idx (optional, training): position 0..8 of the first maximum in row-major window order.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool3_fwd_kernel(const T* __restrict__ in, int in_cs, int in_co, T* __restrict__ out, int out_cs,
                                    int out_co, uint8_t* __restrict__ idx, int C, int64_t M, int crop) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int64_t m = gid / cv;
  const int c0 = (int)(gid - m * cv) << 3;
  const int cc = crop * crop;
  const int r = (int)(m % cc);
  const int y = r / crop, x = r - y * crop;
  float best[8];
  int bi[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; bi[e] = 0; }
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yy = y + dy;
    if (yy < 0 || yy >= crop) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= crop) continue;
      const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(in + (m + dy * crop + dx) * in_cs + in_co + c0);
      const int code = (dy + 1) * 3 + (dx + 1);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float f = to_f32(v.v[e]);
        if (f > best[e]) { best[e] = f; bi[e] = code; }
      }
    }
  }
  Vec8<T> o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = from_f32<T>(best[e]);
  *reinterpret_cast<Vec8<T>*>(out + m * out_cs + out_co + c0) = o;
  if (idx) {
    uint2 pk;
    pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
    pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
    *reinterpret_cast<uint2*>(idx + m * C + c0) = pk;
  }
}

// dIn[q] = sum over the (<=9) windows o that contain q and whose recorded maximum is q, of dOut[o]
template <typename T>
__global__ void maxpool3_bwd_kernel(const T* __restrict__ dout, int do_cs, int do_co, const uint8_t* __restrict__ idx,
                                    T* __restrict__ din, int di_cs, int di_co, int C, int64_t M, int crop) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int64_t m = gid / cv;
  const int c0 = (int)(gid - m * cv) << 3;
  const int cc = crop * crop;
  const int r = (int)(m % cc);
  const int y = r / crop, x = r - y * crop;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int yo = y - dy;            // output window centre such that q = o + (dy, dx)
    if (yo < 0 || yo >= crop) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xo = x - dx;
      if (xo < 0 || xo >= crop) continue;
      const int64_t mo = m - dy * crop - dx;
      const uint2 pk = *reinterpret_cast<const uint2*>(idx + mo * C + c0);
      const Vec8<T> g = *reinterpret_cast<const Vec8<T>*>(dout + mo * do_cs + do_co + c0);
      const int code = (dy + 1) * 3 + (dx + 1);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int id = ((e < 4 ? pk.x : pk.y) >> ((e & 3) * 8)) & 0xff;
        if (id == code) acc[e] += to_f32(g.v[e]);
      }
    }
  }
  Vec8<T> o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = from_f32<T>(acc[e]);
  *reinterpret_cast<Vec8<T>*>(din + m * di_cs + di_co + c0) = o;
}

// ------------------------------------------------------------------------------------------------
// _batch_norm (isprs:655-663): tf.contrib.layers.batch_norm(center=False, scale=False)
// ------------------------------------------------------------------------------------------------
constexpr int BN_ROWS_PER_BLOCK = 256;

// part[blk][0][c] = sum_m a, part[blk][1][c] = sum_m a*b over the block's rows.
//   MODE 0 (forward statistics):  a = z,            b = z
//   MODE 1 (backward sums):       a = g = dA*act'(xh), b = xh        (xh = (z-mean)*inv_std)
template <typename TZ, typename TG, int MODE>
__global__ void bn_partial_kernel(const TZ* __restrict__ z, int z_cs, int z_co, const TG* __restrict__ dA, int g_cs,
                                  int g_co, const float* __restrict__ mean, const float* __restrict__ inv_std, int act,
                                  float* __restrict__ part, int C, int64_t M) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const int64_t r0 = (int64_t)blockIdx.x * BN_ROWS_PER_BLOCK;
  const int64_t r1 = min(M, r0 + BN_ROWS_PER_BLOCK);
  float s0 = 0.0f, s1 = 0.0f;
  float mu = 0.0f, is = 1.0f;
  if (MODE == 1) { mu = mean[c]; is = inv_std[c]; }
  for (int64_t m = r0; m < r1; ++m) {
    const float zv = to_f32(z[m * z_cs + z_co + c]);
    if (MODE == 0) {
      s0 += zv;
      s1 = fmaf(zv, zv, s1);
    } else {
      const float xh = (zv - mu) * is;
      float g = to_f32(dA[m * g_cs + g_co + c]);
      if (act == ACT_RELU) g = xh > 0.0f ? g : 0.0f;
      else if (act == ACT_LRELU) g = xh > 0.0f ? g : 0.1f * g;
      s0 += g;
      s1 = fmaf(g, xh, s1);
    }
  }
  part[((int64_t)blockIdx.x * 2 + 0) * C + c] = s0;
  part[((int64_t)blockIdx.x * 2 + 1) * C + c] = s1;
}

// sums[0][c], sums[1][c] = fixed-order (double) reduction over blocks
__global__ void bn_reduce_kernel(const float* __restrict__ part, float* __restrict__ sums, int C, int nblk) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < nblk; ++k) {
    a += (double)part[((int64_t)k * 2 + 0) * C + c];
    b += (double)part[((int64_t)k * 2 + 1) * C + c];
  }
  sums[c] = (float)a;
  sums[C + c] = (float)b;
}

// mean / inv_std from (possibly all-reduced) sums; moving-average update (decay 0.999, isprs:658 defaults)
__global__ void bn_finalize_kernel(const float* __restrict__ sums, float* __restrict__ mean, float* __restrict__ inv_std,
                                   float* __restrict__ mov_mean, float* __restrict__ mov_var, int C, double count,
                                   float eps, float decay, int unbiased_ema) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mu = (double)sums[c] / count;
  double var = (double)sums[C + c] / count - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  inv_std[c] = (float)(1.0 / sqrt(var + (double)eps));
  const double var_ema = unbiased_ema ? var * (count / fmax(count - 1.0, 1.0)) : var;
  mov_mean[c] = decay * mov_mean[c] + (1.0f - decay) * (float)mu;
  mov_var[c] = decay * mov_var[c] + (1.0f - decay) * (float)var_ema;
}

// out = act((z - mean) * inv_std)     (training-mode normalise + activation)
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ z, int z_cs, int z_co, const float* __restrict__ mean,
                                const float* __restrict__ inv_std, int act, T* __restrict__ out, int o_cs, int o_co,
                                int C, int64_t M) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int64_t m = gid / cv;
  const int c0 = (int)(gid - m * cv) << 3;
  const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(z + m * z_cs + z_co + c0);
  Vec8<T> o;
#pragma unroll
  for (int e = 0; e < 8; ++e)
    o.v[e] = from_f32<T>(apply_act((to_f32(v.v[e]) - mean[c0 + e]) * inv_std[c0 + e], act));
  *reinterpret_cast<Vec8<T>*>(out + m * o_cs + o_co + c0) = o;
}

// dZ = inv_std * (g - s0/M - xh * s1/M),  g = dA * act'(xh)      (no gamma/beta: SURVEY F5)
template <typename TZ, typename TG>
__global__ void bn_bwd_apply_kernel(const TZ* __restrict__ z, int z_cs, int z_co, const TG* __restrict__ dA, int g_cs,
                                    int g_co, const float* __restrict__ mean, const float* __restrict__ inv_std,
                                    const float* __restrict__ sums, double inv_count, int act, TG* __restrict__ dZ,
                                    int d_cs, int d_co, int C, int64_t M) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int64_t m = gid / cv;
  const int c0 = (int)(gid - m * cv) << 3;
  const Vec8<TZ> zv = *reinterpret_cast<const Vec8<TZ>*>(z + m * z_cs + z_co + c0);
  const Vec8<TG> gv = *reinterpret_cast<const Vec8<TG>*>(dA + m * g_cs + g_co + c0);
  Vec8<TG> o;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = c0 + e;
    const float is = inv_std[c];
    const float xh = (to_f32(zv.v[e]) - mean[c]) * is;
    float g = to_f32(gv.v[e]);
    if (act == ACT_RELU) g = xh > 0.0f ? g : 0.0f;
    else if (act == ACT_LRELU) g = xh > 0.0f ? g : 0.1f * g;
    const float m0 = (float)((double)sums[c] * inv_count);
    const float m1 = (float)((double)sums[C + c] * inv_count);
    o.v[e] = from_f32<TG>(is * (g - m0 - xh * m1));
  }
  *reinterpret_cast<Vec8<TG>*>(dZ + m * d_cs + d_co + c0) = o;
}

// dst[:, coff:coff+C] += src (dense-net gradient accumulation into the concat gradient buffer)
template <typename T>
__global__ void add_slice_kernel(T* __restrict__ dst, int d_cs, int d_co, const T* __restrict__ src, int s_cs, int s_co,
                                 int C, int64_t M) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int64_t m = gid / cv;
  const int c0 = (int)(gid - m * cv) << 3;
  Vec8<T> a = *reinterpret_cast<const Vec8<T>*>(dst + m * d_cs + d_co + c0);
  const Vec8<T> b = *reinterpret_cast<const Vec8<T>*>(src + m * s_cs + s_co + c0);
#pragma unroll
  for (int e = 0; e < 8; ++e) a.v[e] = from_f32<T>(to_f32(a.v[e]) + to_f32(b.v[e]));
  *reinterpret_cast<Vec8<T>*>(dst + m * d_cs + d_co + c0) = a;
}

// ------------------------------------------------------------------------------------------------
// conv_classifier (1x1, Ci -> K, bias, no BN/act)  isprs:779-786  + tf.argmax(logits, 3) isprs:1690
// One warp per pixel: lanes stride the channels (coalesced), shuffle-reduce K partial sums.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_CLASSES = 8;

template <typename T>
__global__ void classifier_fwd_kernel(const T* __restrict__ x, int x_cs, int x_co, int Ci, const float* __restrict__ w,
                                      const float* __restrict__ b, int K, float* __restrict__ logits,
                                      uint8_t* __restrict__ pred, int64_t M) {
  extern __shared__ float s_w[];   // [Ci][K]
  for (int i = threadIdx.x; i < Ci * K; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int64_t m = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); m < M;
       m += (int64_t)gridDim.x * warps_per_block) {
    float acc[MAX_CLASSES];
#pragma unroll
    for (int k = 0; k < MAX_CLASSES; ++k) acc[k] = 0.0f;
    const T* row = x + m * x_cs + x_co;
    for (int c0 = lane * 8; c0 < Ci; c0 += 256) {
      const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(row + c0);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float f = to_f32(v.v[e]);
        const float* wr = s_w + (c0 + e) * K;
#pragma unroll
        for (int k = 0; k < MAX_CLASSES; ++k)
          if (k < K) acc[k] = fmaf(f, wr[k], acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < MAX_CLASSES; ++k) {
      if (k < K) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      }
    }
    if (lane == 0) {
      float best = -INFINITY;
      int bi = 0;
#pragma unroll
      for (int k = 0; k < MAX_CLASSES; ++k) {
        if (k < K) {
          const float v = acc[k] + b[k];
          if (logits) logits[m * K + k] = v;
          if (v > best) { best = v; bi = k; }      // first maximum (Appendix B.7)
        }
      }
      if (pred) pred[m] = (uint8_t)bi;
    }
  }
}

// dX[m][c] = sum_k dl[m][k] * W[c][k]
template <typename TG>
__global__ void classifier_bwd_data_kernel(const float* __restrict__ dl, const float* __restrict__ w, int K,
                                           TG* __restrict__ dx, int dx_cs, int dx_co, int Ci, int64_t M) {
  extern __shared__ float s_w[];
  for (int i = threadIdx.x; i < Ci * K; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int cv = Ci >> 3;
  for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < M * cv;
       gid += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = gid / cv;
    const int c0 = (int)(gid - m * cv) << 3;
    float g[MAX_CLASSES];
#pragma unroll
    for (int k = 0; k < MAX_CLASSES; ++k) g[k] = k < K ? dl[m * K + k] : 0.0f;
    Vec8<TG> o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < MAX_CLASSES; ++k)
        if (k < K) s = fmaf(g[k], s_w[(c0 + e) * K + k], s);
      o.v[e] = from_f32<TG>(s);
    }
    *reinterpret_cast<Vec8<TG>*>(dx + m * dx_cs + dx_co + c0) = o;
  }
}

// part[blk][c][k] = sum over the block's rows of x[m][c]*dl[m][k];  part_b[blk][k] = sum dl[m][k]
constexpr int CLS_ROWS_PER_BLOCK = 512;
template <typename T>
__global__ void classifier_bwd_weight_kernel(const T* __restrict__ x, int x_cs, int x_co, int Ci,
                                             const float* __restrict__ dl, int K, float* __restrict__ part,
                                             float* __restrict__ part_b, int64_t M) {
  const int c = threadIdx.x;          // blockDim.x >= max(Ci, K)
  const int64_t r0 = (int64_t)blockIdx.x * CLS_ROWS_PER_BLOCK;
  const int64_t r1 = min(M, r0 + CLS_ROWS_PER_BLOCK);
  float acc[MAX_CLASSES];
#pragma unroll
  for (int k = 0; k < MAX_CLASSES; ++k) acc[k] = 0.0f;
  float accb = 0.0f;
  __shared__ float s_dl[MAX_CLASSES];
  for (int64_t m = r0; m < r1; ++m) {
    __syncthreads();
    if (threadIdx.x < K) s_dl[threadIdx.x] = dl[m * K + threadIdx.x];
    __syncthreads();
    if (c < Ci) {
      const float xv = to_f32(x[m * x_cs + x_co + c]);
#pragma unroll
      for (int k = 0; k < MAX_CLASSES; ++k)
        if (k < K) acc[k] = fmaf(xv, s_dl[k], acc[k]);
    }
    if (c < K) accb += s_dl[c];
  }
  if (c < Ci)
    for (int k = 0; k < K; ++k) part[((int64_t)blockIdx.x * Ci + c) * K + k] = acc[k];
  if (c < K) part_b[(int64_t)blockIdx.x * K + c] = accb;
}

// ------------------------------------------------------------------------------------------------
// loss_def (isprs:1089-1099, contest:881-901): per-pixel softmax cross-entropy, mean over (masked) pixels.
// One thread per pixel; block partial sums in fixed order.  dlogits = (softmax - onehot) * mask * inv_count.
// labels are the float class ids the scripts feed (isprs:1654, cast at 1091).
// ------------------------------------------------------------------------------------------------
constexpr int CE_THREADS = 256;
__global__ void __launch_bounds__(CE_THREADS)
ce_fwd_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ labels_f, const uint8_t* __restrict__ mask,
                  int K, int64_t M, float inv_count, float* __restrict__ dlogits, float* __restrict__ part_loss,
                  uint8_t* __restrict__ labels_u8) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.0f;
  if (m < M) {
    float z[MAX_CLASSES];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < MAX_CLASSES; ++k) {
      z[k] = k < K ? logits[m * K + k] : -INFINITY;
      mx = fmaxf(mx, z[k]);
    }
    float se = 0.0f;
#pragma unroll
    for (int k = 0; k < MAX_CLASSES; ++k)
      if (k < K) se += expf(z[k] - mx);
    const float lse = mx + logf(se);
    const int y = (int)labels_f[m];
    if (labels_u8) labels_u8[m] = (uint8_t)y;
    const bool on = mask ? (mask[m] != 0) : true;
    if (on && y >= 0 && y < K) loss = lse - z[y];
    if (dlogits) {
#pragma unroll
      for (int k = 0; k < MAX_CLASSES; ++k)
        if (k < K) {
          const float pk = expf(z[k] - lse);
          dlogits[m * K + k] = on ? (pk - (k == y ? 1.0f : 0.0f)) * inv_count : 0.0f;
        }
    }
  }
  __shared__ float s_red[CE_THREADS];
  s_red[threadIdx.x] = loss;
  __syncthreads();
  for (int o = CE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part_loss[blockIdx.x] = s_red[0];
}

// count of mask != 0 (contest): block partials -> single value
__global__ void mask_count_kernel(const uint8_t* __restrict__ mask, int64_t M, unsigned int* __restrict__ out) {
  unsigned int c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (int64_t)gridDim.x * blockDim.x)
    c += mask[i] != 0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);   // integer: order-independent
}

// out[0] = sum of n floats in fixed order (single block)
__global__ void sum_fixed_kernel(const float* __restrict__ in, int n, float* __restrict__ out, float scale) {
  __shared__ double s[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += (double)in[i];
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)(s[0] * (double)scale);
}

// ------------------------------------------------------------------------------------------------
// calc_accuracy_by_crop (isprs:510-531) / scene confusion (isprs:1289-1296): K x K counts + #correct
// ------------------------------------------------------------------------------------------------
__global__ void confusion_kernel(const uint8_t* __restrict__ truth, const uint8_t* __restrict__ pred,
                                 const uint8_t* __restrict__ mask, int64_t n, int K, int ignore_label,
                                 unsigned int* __restrict__ cm /* K*K+1 */) {
  __shared__ unsigned int s_cm[MAX_CLASSES * MAX_CLASSES + 1];
  for (int i = threadIdx.x; i < K * K + 1; i += blockDim.x) s_cm[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (mask && !mask[i]) continue;
    const int t = truth[i], p = pred[i];
    if (t == ignore_label || t >= K || p >= K) continue;
    atomicAdd(&s_cm[t * K + p], 1u);
    if (t == p) atomicAdd(&s_cm[K * K], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K + 1; i += blockDim.x)
    if (s_cm[i]) atomicAdd(&cm[i], s_cm[i]);
}

// ------------------------------------------------------------------------------------------------
// MomentumOptimizer (isprs:1685-1687): accum = mom*accum + g ; var -= lr*accum   (Appendix B.6)
// g already contains the CE gradient; the L2 term wd*W is added here for `weights` variables, and the
// loss term wd*sum(W^2)/2 (isprs:640-652) is reduced per block from the pre-update weights.
// ------------------------------------------------------------------------------------------------
__global__ void momentum_update_kernel(float* __restrict__ w, float* __restrict__ g, float* __restrict__ a, int64_t n,
                                       const uint8_t* __restrict__ is_weight, float wd, float lr, float mom,
                                       float grad_scale, float* __restrict__ part_l2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float l2 = 0.0f;
  if (i < n) {
    const float wv = w[i];
    float gv = g[i] * grad_scale;
    if (is_weight[i]) {
      gv = fmaf(wd, wv, gv);
      l2 = wv * wv;
    }
    g[i] = gv;                       // total gradient (kept for drs_get_gradient)
    const float av = fmaf(mom, a[i], gv);
    a[i] = av;
    w[i] = wv - lr * av;
  }
  __shared__ float s_red[256];
  s_red[threadIdx.x] = l2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part_l2[blockIdx.x] = s_red[0];
}

// eval-mode BN folded into the conv epilogue: y = act(conv*scale + shift),
// scale = rsqrt(mv+eps), shift = (bias - mm)*scale        (Appendix B.3)
__global__ void fold_bn_kernel(const float* __restrict__ bias, const float* __restrict__ mm, const float* __restrict__ mv,
                               float eps, float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = 1.0f / sqrtf(mv[c] + eps);
  scale[c] = s;
  shift[c] = (bias[c] - mm[c]) * s;
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = from_f32<TO>(to_f32(in[i]));
}
__global__ void fill_kernel(float* __restrict__ p, float v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// strided 2-D slice -> dense fp32 (debug taps)
template <typename T>
__global__ void slice_to_f32_kernel(const T* __restrict__ in, int cs, int co, int C, int64_t M, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const int64_t m = i / C;
  const int c = (int)(i - m * C);
  out[i] = to_f32(in[m * cs + co + c]);
}
