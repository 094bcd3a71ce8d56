// SwinV2-specific helper kernels (v31_swinv2/*): window partition with cyclic shift, continuous-position-bias tables
// (+ shift mask), post-norm LayerNorm with residual scatter (+ the next GEMM's 16-bit operand), patch merging gather.
// (The cosine-attention q/k normalisation lives in the QKV GEMM's epilogue, gemm_tc.cuh qk_logit.)
#pragma once
#include "kernels_misc.cuh"

namespace dpt {

struct SwinWin {
  int gh, gw;   // token grid of this stage
  int wh, ww;   // window size actually used (adjust_window_and_shift_sizes, windowed_attention.py:345-388)
  int sh, sw;   // cyclic shift applied by this block (0 when the block does not shift)
};

// window-major token index (b, wy, wx, ty, tx) -> pixel (y, x) of the un-rolled image:
// torch.roll(x, -s)[i] = x[(i + s) mod n]  (windowed_attention.py:194,298-299; reverse :226,336-337)
__device__ __forceinline__ long long swin_src_pixel(const SwinWin& w, long long tok, int& b_out) {
  const int A = w.wh * w.ww, nwx = w.gw / w.ww, nwy = w.gh / w.wh;
  const int t = (int)(tok % A);
  long long win = tok / A;
  const int wx = (int)(win % nwx);
  win /= nwx;
  const int wy = (int)(win % nwy);
  const int b = (int)(win / nwy);
  const int ty = t / w.ww, tx = t % w.ww;
  const int y = (wy * w.wh + ty + w.sh) % w.gh, x = (wx * w.ww + tx + w.sw) % w.gw;
  b_out = b;
  return ((long long)b * w.gh + y) * w.gw + x;
}

// x32 [B, gh*gw, C] -> xw [B*nW, A, C] 16-bit (roll + partition), 4 channels per thread
template <typename T>
__global__ void swin_window_gather_kernel(const float* __restrict__ x, T* __restrict__ xw, SwinWin w, int B, int C) {
  const int cv = C / 4;
  const long long total = (long long)B * w.gh * w.gw * cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % cv);
    const long long tok = idx / cv;
    int b;
    const long long pix = swin_src_pixel(w, tok, b);
    const float4 v = *reinterpret_cast<const float4*>(x + pix * C + c4 * 4);
    T o[4] = {from_f32<T>(v.x), from_f32<T>(v.y), from_f32<T>(v.z), from_f32<T>(v.w)};
    *reinterpret_cast<uint2*>(xw + tok * C + c4 * 4) = *reinterpret_cast<uint2*>(o);
  }
}

// Continuous position bias table (relative_positional_encoder.py:60-93,121-283):
// table[e, h] = 16 * sigmoid( W2[h,:] . relu(W1 . coords(e) + b1) ), e over the (2wh-1)*(2ww-1) relative offsets,
// coords = sign(d) * log2(8|d| + 1) / log2(8), d = offset / (pretrained_window - 1)  (window - 1 when none).
__global__ void swin_cpb_table_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                      const float* __restrict__ w2, float* __restrict__ table, int wh, int ww, int heads,
                                      float div_h, float div_w) {
  __shared__ float hid[512];
  const int e = blockIdx.x;
  const int ny = 2 * ww - 1;
  const float dy = (float)(e / ny - (wh - 1)) / div_h, dx = (float)(e % ny - (ww - 1)) / div_w;
  const float inv_log2_8 = 1.0f / 3.0f;
  const float cy = (dy > 0.f ? 1.f : (dy < 0.f ? -1.f : 0.f)) * log2f(fabsf(dy * 8.0f) + 1.0f) * inv_log2_8;
  const float cx = (dx > 0.f ? 1.f : (dx < 0.f ? -1.f : 0.f)) * log2f(fabsf(dx * 8.0f) + 1.0f) * inv_log2_8;
  for (int j = threadIdx.x; j < 512; j += blockDim.x) hid[j] = fmaxf(w1[2 * j] * cy + w1[2 * j + 1] * cx + b1[j], 0.0f);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int h = warp; h < heads; h += nwarps) {
    float acc = 0.0f;
    for (int j = lane; j < 512; j += 32) acc = fmaf(w2[h * 512 + j], hid[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) table[(long long)e * heads + h] = 16.0f / (1.0f + __expf(-acc));
  }
}

// Shift-mask region ids (make_shift_mask, windowed_attention.py:394-439): the reference paints nine slice products in
// order, the last one painting a cell wins; slices are passed as [start, stop) pairs resolved on the host.
struct SwinMaskSlices {
  int h0[3], h1[3], w0[3], w1[3];
};
__device__ __forceinline__ int swin_region(const SwinMaskSlices& s, int y, int x) {
  int id = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
      if (y >= s.h0[a] && y < s.h1[a] && x >= s.w0[b] && x < s.w1[b]) id = a * 3 + b;
  return id;
}

// Attention bias for one layer: out[(wm*H + h), i, j] = table[rel(i,j), h] + (shifted ? (region_i != region_j ? -100 : 0) : 0)
// out rows padded to ldb (zeros). wm runs over the windows of ONE image when shifted (n_wm = nW), else n_wm = 1.
// grid = (A, H, n_wm)
template <typename T>
__global__ void swin_bias_kernel(const float* __restrict__ table, T* __restrict__ out, SwinWin w, SwinMaskSlices ms,
                                 int heads, int shifted, int ldb) {
  const int i = blockIdx.x, h = blockIdx.y, wm = blockIdx.z;
  const int A = w.wh * w.ww;
  const int iy = i / w.ww, ix = i % w.ww;
  const int nwx = w.gw / w.ww;
  const int wy = wm / nwx, wx = wm % nwx;
  const int ri = shifted ? swin_region(ms, wy * w.wh + iy, wx * w.ww + ix) : 0;
  T* orow = out + (((long long)wm * heads + h) * A + i) * ldb;
  for (int j = threadIdx.x; j < ldb; j += blockDim.x) {
    float v = 0.0f;
    if (j < A) {
      const int jy = j / w.ww, jx = j % w.ww;
      const int e = (iy - jy + w.wh - 1) * (2 * w.ww - 1) + (ix - jx + w.ww - 1);
      v = table[(long long)e * heads + h];
      if (shifted && swin_region(ms, wy * w.wh + jy, wx * w.ww + jx) != ri) v += -100.0f;
    }
    orow[j] = from_f32<T>(v);
  }
}

// image pixel row (b, y, x) -> window-major token index under window config w (inverse of swin_src_pixel)
__device__ __forceinline__ long long swin_dst_token(const SwinWin& w, long long pix) {
  const int x0 = (int)(pix % w.gw);
  long long t = pix / w.gw;
  const int y0 = (int)(t % w.gh);
  const int b = (int)(t / w.gh);
  const int y = (y0 - w.sh + w.gh) % w.gh, x = (x0 - w.sw + w.gw) % w.gw;  // position after roll(-shift)
  const int nwx = w.gw / w.ww, nwy = w.gh / w.wh;
  const int wy = y / w.wh, ty = y % w.wh, wx = x / w.ww, tx = x % w.ww;
  return (((long long)b * nwy + wy) * nwx + wx) * (w.wh * w.ww) + ty * w.ww + tx;
}

// Post-norm residual (image_encoder_model.py:213-225): x32[dst(i), :] (+)= LayerNorm(y[i, :]) * g + b, eps 1e-5.
// MAP: 0 identity, 1 window-major -> image (reverse partition + roll back). One warp per row; the residual row and the
// LayerNorm parameters are fetched before the two reductions so that all of a row's memory traffic is in flight at once.
// OUT16: 0 none; 1 = also write the new residual row as 16 bits at the same (image-order) row: the A operand of fc1;
//        2 = also write it as 16 bits at its window-major position under `wn`: the A operand of the NEXT block's QKV
//            GEMM (window partition + cyclic shift as pure addressing, windowed_attention.py:182-228,269-339).
// LPR lanes share one row (8 / 16 / 32: narrow stages put 4 / 2 rows in a warp so the per-row index arithmetic and the
// reductions are amortised), each lane owns up to MAXV float4 of it; a block of 256 threads handles 8 * 32 / LPR rows.
template <typename T, int MAP, bool ADD, int OUT16, int LPR, int MAXV>
__global__ void __launch_bounds__(256) swin_ln_residual_kernel(const T* __restrict__ y, const float* __restrict__ g,
                                                               const float* __restrict__ bia, float* __restrict__ x,
                                                               T* __restrict__ x16, long long M, int F, float eps,
                                                               SwinWin w, SwinWin wn) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const long long row = (blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool live = row < M;  // (dead lanes still take part in the shuffles)
  const long long rowc = live ? row : M - 1;
  const T* yr = y + rowc * F;
  long long dst = rowc;
  if (MAP == 1) {
    int b;
    dst = swin_src_pixel(w, rowc, b);
  }
  float* xr = x + dst * F;
  float4 v[MAXV], r[MAXV];
  uint2 raw[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = (i * LPR + sub) * 4;
    if (f < F) {
      raw[i] = *reinterpret_cast<const uint2*>(yr + f);
      if (ADD) r[i] = *reinterpret_cast<const float4*>(xr + f);
    }
  }
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = (i * LPR + sub) * 4;
    if (f < F) {
      const T* e = reinterpret_cast<const T*>(&raw[i]);
      v[i] = make_float4(to_f32(e[0]), to_f32(e[1]), to_f32(e[2]), to_f32(e[3]));
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)F;
  float var = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = (i * LPR + sub) * 4;
    if (f < F) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      var += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  if (!live) return;
  const float rstd = rsqrtf(var / (float)F + eps);
  T* x16r = nullptr;
  if (OUT16 == 1) x16r = x16 + dst * F;
  if (OUT16 == 2) x16r = x16 + swin_dst_token(wn, dst) * F;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int f = (i * LPR + sub) * 4;
    if (f < F) {
      const float4 ww4 = __ldg(reinterpret_cast<const float4*>(g + f));
      const float4 bb = __ldg(reinterpret_cast<const float4*>(bia + f));
      float4 o;
      o.x = (v[i].x - mean) * rstd * ww4.x + bb.x;
      o.y = (v[i].y - mean) * rstd * ww4.y + bb.y;
      o.z = (v[i].z - mean) * rstd * ww4.z + bb.z;
      o.w = (v[i].w - mean) * rstd * ww4.w + bb.w;
      if (ADD) {
        o.x += r[i].x; o.y += r[i].y; o.z += r[i].z; o.w += r[i].w;
      }
      *reinterpret_cast<float4*>(xr + f) = o;
      if (OUT16 != 0) {
        T h4[4] = {from_f32<T>(o.x), from_f32<T>(o.y), from_f32<T>(o.z), from_f32<T>(o.w)};
        *reinterpret_cast<uint2*>(x16r + f) = *reinterpret_cast<uint2*>(h4);
      }
    }
  }
}

// host launcher: picks the lanes-per-row split from the row length (F <= 1536, a multiple of 32)
template <typename T, int MAP, bool ADD, int OUT16>
cudaError_t launch_swin_ln_residual(const T* y, const float* g, const float* bia, float* x, T* x16, long long M, int F,
                                    float eps, SwinWin w, SwinWin wn, cudaStream_t s) {
  const int f4 = F / 4;
  auto blocks = [&](int rows_per_block) { return (unsigned)((M + rows_per_block - 1) / rows_per_block); };
  if (f4 <= 8 * 6) swin_ln_residual_kernel<T, MAP, ADD, OUT16, 8, 6><<<blocks(32), 256, 0, s>>>(y, g, bia, x, x16, M, F, eps, w, wn);
  else if (f4 <= 16 * 6) swin_ln_residual_kernel<T, MAP, ADD, OUT16, 16, 6><<<blocks(16), 256, 0, s>>>(y, g, bia, x, x16, M, F, eps, w, wn);
  else if (f4 <= 32 * 6) swin_ln_residual_kernel<T, MAP, ADD, OUT16, 32, 6><<<blocks(8), 256, 0, s>>>(y, g, bia, x, x16, M, F, eps, w, wn);
  else swin_ln_residual_kernel<T, MAP, ADD, OUT16, 32, 12><<<blocks(8), 256, 0, s>>>(y, g, bia, x, x16, M, F, eps, w, wn);
  return cudaGetLastError();
}

// PatchMerge gather (patch_merge.py:81-101): x32 [B, gh, gw, C] -> [B, gh/2 * gw/2, 4C] 16-bit, channel blocks in
// the order TL(0,0), BL(1,0), TR(0,1), BR(1,1). 4 channels per thread.
template <typename T>
__global__ void swin_patch_merge_gather_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int gh, int gw,
                                               int C) {
  const int cv = C / 4;
  const int oh = gh / 2, ow = gw / 2;
  const long long total = (long long)B * oh * ow * 4 * cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % cv);
    long long t = idx / cv;
    const int q = (int)(t % 4);  // 0 TL, 1 BL, 2 TR, 3 BR
    t /= 4;
    const int ox = (int)(t % ow);
    t /= ow;
    const int oy = (int)(t % oh);
    const int b = (int)(t / oh);
    const int y = 2 * oy + (q & 1), xx = 2 * ox + (q >> 1);
    const float4 v = *reinterpret_cast<const float4*>(x + (((long long)b * gh + y) * gw + xx) * C + c4 * 4);
    T o[4] = {from_f32<T>(v.x), from_f32<T>(v.y), from_f32<T>(v.z), from_f32<T>(v.w)};
    *reinterpret_cast<uint2*>(out + ((((long long)b * oh + oy) * ow + ox) * 4 + q) * C + c4 * 4) = *reinterpret_cast<uint2*>(o);
  }
}

}  // namespace dpt
