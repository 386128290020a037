// Layer types of the structural net variants that the four north-star graphs do not use:
//   _avg_pool(k x k, stride 1, SAME)            isprs:753-758   (dilated_icpr_rate6_avgpool, 5x5 / 7x7)
//   _squeeze_excitation_layer(ratio 4)          isprs:682-697   (dilated_icpr_rate6_SE): global mean -> FC -> ReLU -> FC ->
//                                                               sigmoid -> per-(image, channel) gate
// All HBM-bound / tiny; written for clarity and a fixed summation order (run-to-run identical), not tuned: these variants
// are not on the benchmarked path.  Included by drs_api.cu.
#pragma once
#include "ops.cuh"

// ------------------------------------------------------------------------------------------------
// average pooling, SAME, stride 1: TF divides by the number of in-image elements of the window
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ in, int in_cs, int in_co, T* __restrict__ out, int out_cs, int out_co, int C,
                                   int64_t M, int crop, int k) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int cg = (int)(gid % cv);
  const int64_t m = gid / cv;
  const int x = (int)(m % crop), y = (int)((m / crop) % crop);
  const int64_t img0 = m - (int64_t)y * crop - x;
  const int pad = (k - 1) / 2;
  const int y0 = max(0, y - pad), y1 = min(crop - 1, y + pad), x0 = max(0, x - pad), x1 = min(crop - 1, x + pad);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  for (int yy = y0; yy <= y1; ++yy)
    for (int xx = x0; xx <= x1; ++xx) {
      const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(in + (img0 + (int64_t)yy * crop + xx) * in_cs + in_co + cg * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += to_f32(v.v[e]);
    }
  const float inv = 1.0f / (float)((y1 - y0 + 1) * (x1 - x0 + 1));
  Vec8<T> o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = from_f32<T>(acc[e] * inv);
  *reinterpret_cast<Vec8<T>*>(out + m * out_cs + out_co + cg * 8) = o;
}

// dIn[p] = sum over the windows w that contain p of dOut[w] / count(w)
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dout, int do_cs, int do_co, T* __restrict__ din, int di_cs, int di_co, int C,
                                   int64_t M, int crop, int k) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int cg = (int)(gid % cv);
  const int64_t m = gid / cv;
  const int x = (int)(m % crop), y = (int)((m / crop) % crop);
  const int64_t img0 = m - (int64_t)y * crop - x;
  const int pad = (k - 1) / 2;
  const int y0 = max(0, y - pad), y1 = min(crop - 1, y + pad), x0 = max(0, x - pad), x1 = min(crop - 1, x + pad);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  for (int yy = y0; yy <= y1; ++yy) {
    const int ny = min(crop - 1, yy + pad) - max(0, yy - pad) + 1;
    for (int xx = x0; xx <= x1; ++xx) {
      const int nx = min(crop - 1, xx + pad) - max(0, xx - pad) + 1;
      const float inv = 1.0f / (float)(ny * nx);
      const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(dout + (img0 + (int64_t)yy * crop + xx) * do_cs + do_co + cg * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(to_f32(v.v[e]), inv, acc[e]);
    }
  }
  Vec8<T> o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = from_f32<T>(acc[e]);
  *reinterpret_cast<Vec8<T>*>(din + m * di_cs + di_co + cg * 8) = o;
}

template <typename T>
static void launch_avgpool_fwd(Handle* h, const T* in, int in_cs, int in_co, T* out, int out_cs, int out_co, int C, int B, int crop, int k) {
  const int64_t M = (int64_t)B * crop * crop;
  avgpool_fwd_kernel<T><<<nblk(M * (C / 8), 256), 256, 0, h->stream>>>(in, in_cs, in_co, out, out_cs, out_co, C, M, crop, k);
  LAUNCH_CHECK(h);
}
template <typename T>
static void launch_avgpool_bwd(Handle* h, const T* dout, int do_cs, int do_co, T* din, int di_cs, int di_co, int C, int B, int crop, int k) {
  const int64_t M = (int64_t)B * crop * crop;
  avgpool_bwd_kernel<T><<<nblk(M * (C / 8), 256), 256, 0, h->stream>>>(dout, do_cs, do_co, din, di_cs, di_co, C, M, crop, k);
  LAUNCH_CHECK(h);
}

// ------------------------------------------------------------------------------------------------
// squeeze-and-excitation
// ------------------------------------------------------------------------------------------------
// out[b][c] = sum over the image's pixels of a (MODE 0) or of a * g (MODE 1).  Block = (image, 64-channel slab): 8 channel
// groups x 32 pixel lanes, per-thread partial sums, then a fixed-order reduction over the lanes.
template <typename T, int MODE>
__global__ void __launch_bounds__(256)
se_sum_kernel(const T* __restrict__ a, int a_cs, int a_co, const T* __restrict__ g, int g_cs, int g_co, int C, int pixels,
              float* __restrict__ out) {
  __shared__ float s_red[8][32][9];
  const int b = blockIdx.x, slab = blockIdx.y;
  const int cgi = threadIdx.x & 7, lane = threadIdx.x >> 3;
  const int c0 = slab * 64 + cgi * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
  if (c0 < C) {
    const int64_t base = (int64_t)b * pixels;
    for (int p = lane; p < pixels; p += 32) {
      const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(a + (base + p) * a_cs + a_co + c0);
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += to_f32(v.v[e]);
      } else {
        const Vec8<T> w = *reinterpret_cast<const Vec8<T>*>(g + (base + p) * g_cs + g_co + c0);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(to_f32(v.v[e]), to_f32(w.v[e]), acc[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_red[cgi][lane][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int cg2 = threadIdx.x >> 3, e = threadIdx.x & 7;
    const int c = slab * 64 + cg2 * 8 + e;
    if (c < C) {
      float s = 0.0f;
      for (int l = 0; l < 32; ++l) s += s_red[cg2][l][e];
      out[(int64_t)b * C + c] = s;
    }
  }
}

// One block per image: s = sum / pixels; hid = relu(s W1 + b1); e = sigmoid(hid W2 + b2).   W1 [C][R], W2 [R][C] (TF layout)
__global__ void se_fc_fwd_kernel(const float* __restrict__ sums, float inv_pixels, const float* __restrict__ w1,
                                 const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, int C, int R,
                                 float* __restrict__ s_out, float* __restrict__ h_out, float* __restrict__ e_out) {
  __shared__ float s_s[256], s_h[64];
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = sums[(int64_t)b * C + c] * inv_pixels;
    s_s[c] = v;
    s_out[(int64_t)b * C + c] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < R; j += blockDim.x) {
    float a = b1[j];
    for (int c = 0; c < C; ++c) a = fmaf(s_s[c], w1[c * R + j], a);
    a = fmaxf(a, 0.0f);
    s_h[j] = a;
    h_out[(int64_t)b * R + j] = a;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = b2[c];
    for (int j = 0; j < R; ++j) a = fmaf(s_h[j], w2[j * C + c], a);
    e_out[(int64_t)b * C + c] = 1.0f / (1.0f + expf(-a));
  }
}

// out = a * e[image][channel]
template <typename T>
__global__ void se_scale_fwd_kernel(const T* __restrict__ a, int a_cs, int a_co, const float* __restrict__ e, T* __restrict__ out,
                                    int o_cs, int o_co, int C, int64_t M, int pixels) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int cg = (int)(gid % cv);
  const int64_t m = gid / cv;
  const float* eb = e + (m / pixels) * C + cg * 8;
  const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(a + m * a_cs + a_co + cg * 8);
  Vec8<T> o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = from_f32<T>(to_f32(v.v[i]) * eb[i]);
  *reinterpret_cast<Vec8<T>*>(out + m * o_cs + o_co + cg * 8) = o;
}

// One block per image.  de = sum_px dOut * A (from se_sum_kernel<MODE 1>); back through sigmoid, FC2, ReLU, FC1 and the mean.
// Writes ds[b][c] (already divided by the pixel count: the gradient every pixel of the channel receives through the mean) and
// this image's contribution to the four parameter gradients, part[b][...] in the order W1, b1, W2, b2 (summed over images by
// reduce_partials_kernel in fixed order).
__global__ void se_fc_bwd_kernel(const float* __restrict__ de, const float* __restrict__ s_in, const float* __restrict__ h_in,
                                 const float* __restrict__ e_in, const float* __restrict__ w1, const float* __restrict__ w2, int C, int R,
                                 float inv_pixels, float* __restrict__ ds_out, float* __restrict__ part) {
  __shared__ float s_dz2[256], s_dz1[64];
  const int b = blockIdx.x;
  const int64_t np = (int64_t)2 * C * R + R + C;
  float* pw1 = part + (int64_t)b * np;
  float* pb1 = pw1 + (int64_t)C * R;
  float* pw2 = pb1 + R;
  float* pb2 = pw2 + (int64_t)R * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float ev = e_in[(int64_t)b * C + c];
    const float d = de[(int64_t)b * C + c] * ev * (1.0f - ev);
    s_dz2[c] = d;
    pb2[c] = d;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < R; j += blockDim.x) {
    float a = 0.0f;
    for (int c = 0; c < C; ++c) a = fmaf(s_dz2[c], w2[j * C + c], a);
    const float hv = h_in[(int64_t)b * R + j];
    const float d = hv > 0.0f ? a : 0.0f;
    s_dz1[j] = d;
    pb1[j] = d;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R * C; i += blockDim.x) {
    const int j = i / C, c = i - j * C;
    pw2[i] = h_in[(int64_t)b * R + j] * s_dz2[c];
  }
  for (int i = threadIdx.x; i < C * R; i += blockDim.x) {
    const int c = i / R, j = i - c * R;
    pw1[i] = s_in[(int64_t)b * C + c] * s_dz1[j];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.0f;
    for (int j = 0; j < R; ++j) a = fmaf(s_dz1[j], w1[c * R + j], a);
    ds_out[(int64_t)b * C + c] = a * inv_pixels;
  }
}

// dA = dOut * e + ds
template <typename T>
__global__ void se_scale_bwd_kernel(const T* __restrict__ dout, int do_cs, int do_co, const float* __restrict__ e,
                                    const float* __restrict__ ds, T* __restrict__ dA, int d_cs, int d_co, int C, int64_t M, int pixels) {
  const int cv = C >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= M * cv) return;
  const int cg = (int)(gid % cv);
  const int64_t m = gid / cv;
  const int64_t off = (m / pixels) * C + cg * 8;
  const Vec8<T> v = *reinterpret_cast<const Vec8<T>*>(dout + m * do_cs + do_co + cg * 8);
  Vec8<T> o;
#pragma unroll
  for (int i = 0; i < 8; ++i) o.v[i] = from_f32<T>(fmaf(to_f32(v.v[i]), e[off + i], ds[off + i]));
  *reinterpret_cast<Vec8<T>*>(dA + m * d_cs + d_co + cg * 8) = o;
}
