// HBM-bound helper kernels of the DPT hot path: im2col gathers, token assembly, LayerNorm, bilinear / bicubic
// resampling, ReLU copy. All activations are channels-last; 16-bit type T is __nv_bfloat16 or __half.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include "ptx.cuh"

namespace dpt {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

template <typename T> struct __align__(16) Vec8 { T v[8]; };  // 16 bytes

// ---------------------------------------------------------------------------------------------------------------
// Patch-embed im2col (patch_embed.py:92-97): img [B,3,H,W] -> A [B*gh*gw, kpad], k = c*P*P + ky*P + kx, zero padded.
// One thread per PAIR of horizontally adjacent image pixels: the reads are fully coalesced 4-byte loads of the NCHW
// image; the two pixels land in the same patch row (P is even), 4 bytes apart in A. The zero padding of columns
// [3*P*P, kpad) is written by the tail of the same grid-stride loop.
template <typename T>
__global__ void im2col_patch_kernel(const T* __restrict__ img, T* __restrict__ A, int B, int Cin, int H, int W, int P,
                                    int gh, int gw, int kpad) {
  const int kreal = Cin * P * P;
  const int W2 = W >> 1;
  const long long npairs = (long long)B * Cin * H * W2;
  const long long M = (long long)B * gh * gw;
  const int padw = (kpad - kreal) >> 1;  // zero pairs per row of A (kreal and kpad are even)
  const long long total = npairs + M * padw;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    if (idx < npairs) {
      const int x = (int)(idx % W2) * 2;
      long long t = idx / W2;
      const int y = (int)(t % H);
      t /= H;
      const int c = (int)(t % Cin);
      const int b = (int)(t / Cin);
      const int px = x / P, kx = x - px * P, py = y / P, ky = y - py * P;
      const uint32_t v = *reinterpret_cast<const uint32_t*>(img + idx * 2);
      *reinterpret_cast<uint32_t*>(A + (((long long)b * gh + py) * gw + px) * kpad + (c * P + ky) * P + kx) = v;
    } else {
      const long long j = idx - npairs;
      const long long m = j / padw;
      *reinterpret_cast<uint32_t*>(A + m * kpad + kreal + (int)(j - m * padw) * 2) = 0u;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 stride-2 pad-1 im2col on NHWC (reassembly_model.py:302-309): in [B,H,W,C] -> A [B*(H/2)*(W/2), 9*cpad],
// k = tap*cpad + c, tap = ky*3 + kx. 8 channels (16 B) per thread.
template <typename T>
__global__ void im2col_3x3s2_kernel(const T* __restrict__ in, T* __restrict__ A, int B, int H, int W, int C, int cpad) {
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const int cv = cpad / 8;
  const long long total = (long long)B * OH * OW * 9 * cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % cv);
    long long t = idx / cv;
    const int tap = (int)(t % 9);
    t /= 9;
    const int ox = (int)(t % OW);
    t /= OW;
    const int oy = (int)(t % OH);
    const int b = (int)(t / OH);
    const int iy = oy * 2 - 1 + tap / 3, ix = ox * 2 - 1 + tap % 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W && c8 * 8 < C)
      v = *reinterpret_cast<const uint4*>(in + (((long long)b * H + iy) * W + ix) * C + c8 * 8);
    *reinterpret_cast<uint4*>(A + ((((long long)b * OH + oy) * OW + ox) * 9 + tap) * cpad + c8 * 8) = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Position table (position_encoder.py:55-76,108-143): pos[0,:] = cls_token + cls_embedding;
// pos[1 + y*gw + x, :] = bicubic(base[bh,bw,F]) at (y,x), align_corners=False, A=-0.75, fp32.
__device__ __forceinline__ float cubic_w1(float x, float A) { return ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f; }
__device__ __forceinline__ float cubic_w2(float x, float A) { return ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A; }

__global__ void pos_table_kernel(const float* __restrict__ base, const float* __restrict__ cls_tok,
                                 const float* __restrict__ cls_emb, float* __restrict__ pos, int bh, int bw, int gh,
                                 int gw, int F) {
  const int row = blockIdx.x;  // 0 .. gh*gw
  if (row == 0) {
    for (int f = threadIdx.x; f < F; f += blockDim.x) pos[f] = cls_tok[f] + cls_emb[f];
    return;
  }
  const int y = (row - 1) / gw, x = (row - 1) % gw;
  const float A = -0.75f;
  const float sy = (float)bh / (float)gh, sx = (float)bw / (float)gw;
  const float fy = sy * (y + 0.5f) - 0.5f, fx = sx * (x + 0.5f) - 0.5f;
  const int iy = (int)floorf(fy), ix = (int)floorf(fx);
  const float ty = fy - iy, tx = fx - ix;
  float wy[4] = {cubic_w2(ty + 1.0f, A), cubic_w1(ty, A), cubic_w1(1.0f - ty, A), cubic_w2(2.0f - ty, A)};
  float wx[4] = {cubic_w2(tx + 1.0f, A), cubic_w1(tx, A), cubic_w1(1.0f - tx, A), cubic_w2(2.0f - tx, A)};
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float acc = 0.0f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), bh - 1);
      float racc = 0.0f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int xx = min(max(ix - 1 + c, 0), bw - 1);
        racc += wx[c] * base[((long long)yy * bw + xx) * F + f];
      }
      acc += wy[a] * racc;
    }
    pos[(long long)row * F + f] = acc;
  }
}

// x32[b,0,:] = pos[0,:]; x32[b,1+p,:] = tok[b,p,:] + pos[1+p,:]   (image_encoder_model.py:83-84)
// pos_rows == N: full table; pos_rows == 1: only the cls row exists (BEiT: no position embedding on patches)
template <typename T>
__global__ void assemble_tokens_kernel(const T* __restrict__ tok, const float* __restrict__ pos, float* __restrict__ x,
                                       int B, int N, int F, int pos_rows) {
  const long long total = (long long)B * N * F / 4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx * 4;
    const int f = (int)(e % F);
    const long long row = e / F;
    const int n = (int)(row % N);
    const int b = (int)(row / N);
    float4 pv = (n < pos_rows) ? *reinterpret_cast<const float4*>(pos + (long long)n * F + f)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    if (n > 0) {
      const T* t = tok + ((long long)b * (N - 1) + (n - 1)) * F + f;
      pv.x += to_f32(t[0]); pv.y += to_f32(t[1]); pv.z += to_f32(t[2]); pv.w += to_f32(t[3]);
    }
    *reinterpret_cast<float4*>(x + e) = pv;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (biased variance), fp32 in -> 16-bit out. One warp per row, values held in registers.
template <typename T, typename TIN>
__global__ void layernorm_kernel(const TIN* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bia,
                                 T* __restrict__ y, long long M, int F, float eps) {
  constexpr int MAXV = 12;  // F <= 1536
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const TIN* xr = x + row * F;
  float4 v[MAXV];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = i * 128 + lane * 4;
    if (f < F) {
      if constexpr (sizeof(TIN) == 4) {
        v[i] = *reinterpret_cast<const float4*>(xr + f);
      } else {
        v[i] = make_float4(to_f32(xr[f]), to_f32(xr[f + 1]), to_f32(xr[f + 2]), to_f32(xr[f + 3]));
      }
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)F;
  float var = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = i * 128 + lane * 4;
    if (f < F) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      var += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / (float)F + eps);
  T* yr = y + row * F;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = i * 128 + lane * 4;
    if (f < F) {
      const float4 ww = *reinterpret_cast<const float4*>(w + f);
      const float4 bb = *reinterpret_cast<const float4*>(bia + f);
      T o[4];
      o[0] = from_f32<T>((v[i].x - mean) * rstd * ww.x + bb.x);
      o[1] = from_f32<T>((v[i].y - mean) * rstd * ww.y + bb.y);
      o[2] = from_f32<T>((v[i].z - mean) * rstd * ww.z + bb.z);
      o[3] = from_f32<T>((v[i].w - mean) * rstd * ww.w + bb.w);
      *reinterpret_cast<uint2*>(yr + f) = *reinterpret_cast<uint2*>(o);
    }
  }
}

// Row statistics for the LayerNorm folded into the next GEMM (gemm_tc.cuh): stats[row] = {(sum, sum of squares), (0, 0)} of the
// fp32 row (two parts: the consumer reads pairs of parts), y = 16-bit copy of the row. One warp per row. Only the first block of an encoder needs this kernel; later
// rows statistics come out of the epilogue of the GEMM that updates the residual stream.
template <typename T>
__global__ void row_stats_cast_kernel(const float* __restrict__ x, T* __restrict__ y, float* __restrict__ stats,
                                      long long M, int F) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * F;
  T* yr = y + row * F;
  float sum = 0.0f, sq = 0.0f;
  for (int f = lane * 4; f < F; f += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + f);
    sum += (v.x + v.y) + (v.z + v.w);
    sq += fmaf(v.x, v.x, v.y * v.y) + fmaf(v.z, v.z, v.w * v.w);
    T o[4] = {from_f32<T>(v.x), from_f32<T>(v.y), from_f32<T>(v.z), from_f32<T>(v.w)};
    *reinterpret_cast<uint2*>(yr + f) = *reinterpret_cast<uint2*>(o);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if (lane == 0) reinterpret_cast<float4*>(stats)[row] = make_float4(sum, sq, 0.0f, 0.0f);  // two parts, second empty
}

// ---------------------------------------------------------------------------------------------------------------
// Bilinear resize, align_corners=True (misc_helpers.py:39-42), NHWC, 8 channels per thread.
// src coordinate = dst * (in-1)/(out-1); matches ATen's area_pixel_compute_scale for align_corners.
// A thread owns one (output column, 8-channel group) and walks down RESIZE_ROWS output rows, keeping the two
// horizontally interpolated source rows h(y0), h(y0+1) in registers: consecutive output rows of an up-sample share
// them, so a ~x2 resize issues ~1.1 16-byte loads per 16-byte store instead of 4 (the kernel was LSU/L2-bound at
// 1.9 TB/s of algorithmic bytes). Output is written with streaming stores (read next by TMA through L2 only once).
constexpr int RESIZE_ROWS = 8;

// 16-bit pair <-> packed fp32 pair
template <typename T> __device__ __forceinline__ float2 resize_unpack2(uint32_t u);
template <> __device__ __forceinline__ float2 resize_unpack2<__nv_bfloat16>(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 resize_unpack2<__half>(uint32_t u) {
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
template <typename T> __device__ __forceinline__ uint32_t resize_pack2(float2 v);
template <> __device__ __forceinline__ uint32_t resize_pack2<__nv_bfloat16>(float2 v) {
  __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t resize_pack2<__half>(float2 v) {
  __half2 h = __floats2half2_rn(v.x, v.y);
  return *reinterpret_cast<uint32_t*>(&h);
}

// horizontally interpolated source row: 8 channels as four packed fp32 pairs, h = hx * in[x0] + lx * in[x1]
template <typename T>
__device__ __forceinline__ void resize_hrow(const T* __restrict__ p0, const T* __restrict__ p1, float2 hx2, float2 lx2,
                                            float2 (&h)[4]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p0));
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(p1));
  const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    h[k] = __ffma2_rn(lx2, resize_unpack2<T>(ub[k]), __fmul2_rn(hx2, resize_unpack2<T>(ua[k])));
}

template <typename T>
__device__ __forceinline__ void resize_emit(T* __restrict__ dst, const float2 (&lo)[4], const float2 (&hi)[4], float2 hy2,
                                            float2 ly2) {
  uint4 o;
  o.x = resize_pack2<T>(__ffma2_rn(ly2, hi[0], __fmul2_rn(hy2, lo[0])));
  o.y = resize_pack2<T>(__ffma2_rn(ly2, hi[1], __fmul2_rn(hy2, lo[1])));
  o.z = resize_pack2<T>(__ffma2_rn(ly2, hi[2], __fmul2_rn(hy2, lo[2])));
  o.w = resize_pack2<T>(__ffma2_rn(ly2, hi[3], __fmul2_rn(hy2, lo[3])));
  __stcs(reinterpret_cast<uint4*>(dst), o);
}

// Round 2, second pass: the kernel was issue-bound, not DRAM-bound (ncu: 78 % issue-active, ALU pipe 60 %, 4.9 of
// 6.5 TB/s) - packed fp32x2 arithmetic (half the FP instructions), two-instruction 16-bit unpack, and the two held source
// rows swap ROLES when the source row advances instead of being copied register by register.
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_ac_kernel(const T* __restrict__ in, T* __restrict__ out, int B,
                                                                 int IH, int IW, int OH, int OW, int C) {
  // grid = (ceil(OW * C/8 / blockDim), ceil(OH / RESIZE_ROWS), B); 32-bit index math inside one image
  pdl_wait();
  pdl_launch_dependents();
  const int cv = C >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= OW * cv) return;
  const int ox = i / cv, c8 = i - ox * cv;
  const int oy0 = blockIdx.y * RESIZE_ROWS, b = blockIdx.z;
  const float sy = OH > 1 ? (float)(IH - 1) / (float)(OH - 1) : 0.0f;
  const float sx = OW > 1 ? (float)(IW - 1) / (float)(OW - 1) : 0.0f;
  const float fx = sx * ox;
  const int x0 = min((int)fx, IW - 1);
  const int x1 = min(x0 + 1, IW - 1);
  const float lx = fx - x0, hx = 1.0f - lx;
  const float2 lx2 = make_float2(lx, lx), hx2 = make_float2(hx, hx);
  const T* base = in + (size_t)b * IH * IW * C + c8 * 8;
  const T* col0 = base + x0 * C;
  const T* col1 = base + x1 * C;
  const int rs = IW * C;  // source row stride (elements)
  T* obase = out + (((size_t)b * OH + oy0) * OW + ox) * C + c8 * 8;
  const int ors = OW * C;
  float2 hA[4], hB[4];
  bool sw = false;  // false: (lower, upper) source rows = (hA, hB); true: (hB, hA)
  int cy = -2;      // lower source row currently held
#pragma unroll
  for (int r = 0; r < RESIZE_ROWS; ++r) {
    const int oy = oy0 + r;
    if (oy >= OH) break;
    const float fy = sy * oy;
    const int y0 = min((int)fy, IH - 1);
    const float ly = fy - y0, hy = 1.0f - ly;
    if (y0 != cy) {  // block-uniform
      const int y1 = min(y0 + 1, IH - 1);
      if (y0 == cy + 1) {  // the upper row becomes the lower one: load the new upper row over the old lower one
        if (!sw) resize_hrow<T>(col0 + y1 * rs, col1 + y1 * rs, hx2, lx2, hA);
        else resize_hrow<T>(col0 + y1 * rs, col1 + y1 * rs, hx2, lx2, hB);
        sw = !sw;
      } else {
        resize_hrow<T>(col0 + y0 * rs, col1 + y0 * rs, hx2, lx2, hA);
        resize_hrow<T>(col0 + y1 * rs, col1 + y1 * rs, hx2, lx2, hB);
        sw = false;
      }
      cy = y0;
    }
    const float2 hy2 = make_float2(hy, hy), ly2 = make_float2(ly, ly);
    if (!sw) resize_emit<T>(obase + r * ors, hA, hB, hy2, ly2);
    else resize_emit<T>(obase + r * ors, hB, hA, hy2, ly2);
  }
}

// out = relu(in), 8 elements per thread
template <typename T>
__global__ void relu_copy_kernel(const T* __restrict__ in, T* __restrict__ out, long long n8) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n8;
       idx += (long long)gridDim.x * blockDim.x) {
    Vec8<T> v = reinterpret_cast<const Vec8<T>*>(in)[idx];
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] = from_f32<T>(fmaxf(to_f32(v.v[i]), 0.0f));
    reinterpret_cast<Vec8<T>*>(out)[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------
// fp32 -> 16-bit cast (BEiT taps: the encoder has no output norm), 4 elements per thread
template <typename T>
__global__ void cast_f32_kernel(const float* __restrict__ in, T* __restrict__ out, long long n4) {
  pdl_wait();
  pdl_launch_dependents();
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n4;
       idx += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(in)[idx];
    T o[4] = {from_f32<T>(v.x), from_f32<T>(v.y), from_f32<T>(v.z), from_f32<T>(v.w)};
    reinterpret_cast<uint2*>(out)[idx] = *reinterpret_cast<uint2*>(o);
  }
}

// 16-bit -> fp32 (SwinV2: the residual stream starts from the normalised patch tokens), 4 elements per thread
template <typename T>
__global__ void cast_to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, long long n4) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n4;
       idx += (long long)gridDim.x * blockDim.x) {
    const T* p = in + idx * 4;
    reinterpret_cast<float4*>(out)[idx] = make_float4(to_f32(p[0]), to_f32(p[1]), to_f32(p[2]), to_f32(p[3]));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BEiT relative position bias (v31_beit/components/relative_positional_encoder.py:117-309):
// out[h, i, j] = lut'[index(i, j), h], lut' = bilinear resize (align_corners=False) of the reference table
// [(2bh-1)*(2bw-1) + 3, H] to (2gh-1, 2gw-1), the three extra rows being cls->token, token->cls, cls->cls.
// out is [H, N, ldb] 16-bit, columns >= N zero. One block per (i, h).
template <typename T>
__global__ void beit_bias_table_kernel(const float* __restrict__ table, T* __restrict__ out, int heads, int bh, int bw,
                                       int gh, int gw, int ldb) {
  const int i = blockIdx.x, h = blockIdx.y;
  const int N = gh * gw + 1;
  const int rh = 2 * bh - 1, rw = 2 * bw - 1, nh = 2 * gh - 1, nw = 2 * gw - 1;
  const float* cls_rows = table + (long long)rh * rw * heads;
  const float sy = (float)rh / (float)nh, sx = (float)rw / (float)nw;
  T* orow = out + ((long long)h * N + i) * ldb;
  const int yi = (i - 1) / gw, xi = (i - 1) % gw;
  for (int j = threadIdx.x; j < ldb; j += blockDim.x) {
    float v = 0.0f;
    if (j < N) {
      if (i == 0 && j == 0) v = cls_rows[2 * heads + h];
      else if (i == 0) v = cls_rows[0 * heads + h];
      else if (j == 0) v = cls_rows[1 * heads + h];
      else {
        const int yj = (j - 1) / gw, xj = (j - 1) % gw;
        const int ry = yi - yj + gh - 1, rx = xi - xj + gw - 1;  // position in the resized (nh x nw) table
        const float fy = fmaxf(sy * (ry + 0.5f) - 0.5f, 0.0f), fx = fmaxf(sx * (rx + 0.5f) - 0.5f, 0.0f);
        const int y0 = min((int)fy, rh - 1), x0 = min((int)fx, rw - 1);
        const int y1 = min(y0 + 1, rh - 1), x1 = min(x0 + 1, rw - 1);
        const float ly = fy - y0, lx = fx - x0;
        const float t00 = table[((long long)y0 * rw + x0) * heads + h], t01 = table[((long long)y0 * rw + x1) * heads + h];
        const float t10 = table[((long long)y1 * rw + x0) * heads + h], t11 = table[((long long)y1 * rw + x1) * heads + h];
        v = (1.0f - ly) * ((1.0f - lx) * t00 + lx * t01) + ly * ((1.0f - lx) * t10 + lx * t11);
      }
    }
    orow[j] = from_f32<T>(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// BEiT readout (v31_beit/components/readout_projection.py:71-81): the cls half of Linear(2F, F) is constant per image:
// u[b, n] = bias[n] + sum_k W2[n, k] * cls[b, k], cls = tap[b, 0, :]. One warp per output.
template <typename T>
__global__ void readout_vec_kernel(const T* __restrict__ tap, const float* __restrict__ w2, const float* __restrict__ bias,
                                   float* __restrict__ u, int B, int N, int F) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * F) return;
  const int b = warp / F, n = warp % F;
  const T* cls = tap + (long long)b * N * F;
  const float* wr = w2 + (long long)n * F;
  float acc = 0.0f;
  for (int k = lane; k < F; k += 32) acc = fmaf(wr[k], to_f32(cls[k]), acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) u[warp] = acc + bias[n];
}

// ---------------------------------------------------------------------------------------------------------------
// Pre-processing (v2_depthanything/patch_embed.py:103-145): BGR uint8 HWC -> RGB, antialiased bilinear resize,
// (v/255 - mean) * inv_std, NCHW 16-bit. Weights follow ATen's separable antialias kernel
// (aten/native/cpu/UpSampleKernel.cpp, _compute_indices_min_size_weights_aa with the triangle filter):
// scale = in/out, support = scale >= 1 ? scale : 1, center = scale * (o + 0.5), taps [xmin, xmin + xsize) with
// xmin = max(0, int(center - support + 0.5)), xsize = min(in, int(center + support + 0.5)) - xmin,
// w_j = tri((j + xmin - center + 0.5) / max(scale, 1)) normalised to sum 1. One thread per output pixel: the 2-D
// weight is the product of the two 1-D weights (horizontal then vertical pass of the reference, done at once in fp32).
__device__ __forceinline__ void aa_taps(int o, int in, int out, int& xmin, int& xsize, float& center, float& invscale,
                                        float& total) {
  const float scale = (float)in / (float)out;
  const float support = scale >= 1.0f ? scale : 1.0f;
  invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  center = scale * (o + 0.5f);
  xmin = max(0, (int)(center - support + 0.5f));
  xsize = min(in, (int)(center + support + 0.5f)) - xmin;
  total = 0.0f;
  for (int j = 0; j < xsize; ++j) total += fmaxf(0.0f, 1.0f - fabsf((j + xmin - center + 0.5f) * invscale));
}

template <typename T>
__global__ void prepare_image_kernel(const uint8_t* __restrict__ bgr, T* __restrict__ out, int IH, int IW, int OH, int OW,
                                     float3 mean, float3 inv_std) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
  if (ox >= OW) return;
  int x0, xn, y0, yn;
  float cx, isx, tx, cy, isy, ty;
  aa_taps(ox, IW, OW, x0, xn, cx, isx, tx);
  aa_taps(oy, IH, OH, y0, yn, cy, isy, ty);
  float r = 0.0f, g = 0.0f, b = 0.0f;
  for (int jy = 0; jy < yn; ++jy) {
    const float wy = fmaxf(0.0f, 1.0f - fabsf((jy + y0 - cy + 0.5f) * isy)) / ty;
    const uint8_t* row = bgr + ((size_t)(y0 + jy) * IW + x0) * 3;
    float rr = 0.0f, gg = 0.0f, bb = 0.0f;
    for (int jx = 0; jx < xn; ++jx) {
      const float wx = fmaxf(0.0f, 1.0f - fabsf((jx + x0 - cx + 0.5f) * isx)) / tx;
      bb += wx * (float)row[3 * jx + 0];
      gg += wx * (float)row[3 * jx + 1];
      rr += wx * (float)row[3 * jx + 2];
    }
    r += wy * rr;
    g += wy * gg;
    b += wy * bb;
  }
  const size_t plane = (size_t)OH * OW, o = (size_t)oy * OW + ox;
  out[o] = from_f32<T>((r / 255.0f - mean.x) * inv_std.x);
  out[plane + o] = from_f32<T>((g / 255.0f - mean.y) * inv_std.y);
  out[2 * plane + o] = from_f32<T>((b / 255.0f - mean.z) * inv_std.z);
}

// ---------------------------------------------------------------------------------------------------------------
// Post-processing (demo_helpers/postprocess.py): bilinear resize of the prediction (align_corners=False, no
// antialias: F.interpolate default), global min / max of the SCALED tensor, then floor(255 * (v - min) / (max - min)).
// All arithmetic in fp32 on the 16-bit prediction (the oracle is the reference's fp32 CPU arithmetic on the same values).
template <typename T>
__device__ __forceinline__ float scaled_prediction_at(const T* __restrict__ d, int H, int W, int OH, int OW, int oy, int ox) {
  const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
  const float fy = fmaxf(sy * (oy + 0.5f) - 0.5f, 0.0f), fx = fmaxf(sx * (ox + 0.5f) - 0.5f, 0.0f);
  const int y0 = min((int)fy, H - 1), x0 = min((int)fx, W - 1);
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = fy - y0, lx = fx - x0;
  const float v = (1.0f - ly) * ((1.0f - lx) * to_f32(d[(size_t)y0 * W + x0]) + lx * to_f32(d[(size_t)y0 * W + x1])) +
                  ly * ((1.0f - lx) * to_f32(d[(size_t)y1 * W + x0]) + lx * to_f32(d[(size_t)y1 * W + x1]));
  return v;
}

__global__ void minmax_init_kernel(float* mm) {
  mm[0] = INFINITY;
  mm[1] = -INFINITY;
}

// float atomic min / max through the ordered-integer trick (valid for any mix of signs, no NaNs expected)
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}

template <typename T>
__global__ void post_minmax_kernel(const T* __restrict__ depth, float* __restrict__ mm, int B, int H, int W, int OH, int OW) {
  const long long total = (long long)B * OH * OW;
  float lo = INFINITY, hi = -INFINITY;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW);
    const long long t = idx / OW;
    const int oy = (int)(t % OH), b = (int)(t / OH);
    const float v = scaled_prediction_at(depth + (size_t)b * H * W, H, W, OH, OW, oy, ox);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomic_min_f(mm, lo);
    atomic_max_f(mm + 1, hi);
  }
}

template <typename T>
__global__ void post_u8_kernel(const T* __restrict__ depth, const float* __restrict__ mm, uint8_t* __restrict__ out, int B,
                               int H, int W, int OH, int OW) {
  const long long total = (long long)B * OH * OW;
  const float lo = mm[0], range = mm[1] - mm[0];  // (data - min) / (max - min), 255.0 * x, .byte() (truncation)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % OW);
    const long long t = idx / OW;
    const int oy = (int)(t % OH), b = (int)(t / OH);
    const float v = scaled_prediction_at(depth + (size_t)b * H * W, H, W, OH, OW, oy, ox);
    const float s = 255.0f * ((v - lo) / range);
    out[idx] = (uint8_t)(int)fminf(fmaxf(s, 0.0f), 255.0f);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Debug / experiments: the attention probabilities softmax(scale * q k^T + bias) of one block,
// materialised as [Bt, H, N, N] so that nn.Softmax forward hooks (demo_helpers/model_capture.py:15-61,
// experiments/attention_visualization.py:324-332) receive what the reference's manual attention path hands them
// (transformer_block.py:127-132). Plain CUDA-core kernel, never on the forward path: one CTA per (query row, head,
// batch); the row's logits are recomputed in each of the three passes (max, sum, write).
template <typename T>
__global__ void attn_probs_kernel(const T* __restrict__ qkv, const T* __restrict__ bias, long long ldb, int wmod,
                                  T* __restrict__ probs, int N, int H, int HD, float scale) {
  extern __shared__ float att_q[];  // [HD]
  __shared__ float red[32];
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int F = H * HD;
  const T* base = qkv + (long long)b * N * 3 * F;
  for (int d = threadIdx.x; d < HD; d += blockDim.x) att_q[d] = to_f32(base[(long long)i * 3 * F + h * HD + d]) * scale;
  __syncthreads();
  const T* brow = bias ? bias + (((long long)(b % wmod) * H + h) * N + i) * ldb : nullptr;
  auto logit = [&](int j) {
    const T* k = base + (long long)j * 3 * F + F + h * HD;
    float acc = 0.f;
    for (int d = 0; d < HD; ++d) acc = fmaf(att_q[d], to_f32(k[d]), acc);
    return brow ? acc + to_f32(brow[j]) : acc;
  };
  auto block_reduce = [&](float v, bool is_max) {
    for (int o = 16; o > 0; o >>= 1) {
      const float w = __shfl_xor_sync(0xffffffffu, v, o);
      v = is_max ? fmaxf(v, w) : v + w;
    }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = is_max ? -INFINITY : 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    __syncthreads();
    return r;
  };
  float m = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) m = fmaxf(m, logit(j));
  m = block_reduce(m, true);
  float l = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) l += __expf(logit(j) - m);
  l = block_reduce(l, false);
  const float inv = 1.0f / l;
  T* out = probs + (((long long)b * H + h) * N + i) * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) out[j] = from_f32<T>(__expf(logit(j) - m) * inv);
}

}  // namespace dpt
