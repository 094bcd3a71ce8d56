// 3x3 / stride 1 / pad 1 convolution to 32 channels with the depth-head epilogue (head_model.py:80-106:
// ReLU -> 1x1 conv 32->1 -> ReLU | Sigmoid), the second convolution of MonocularDepthHead.
//
// The generic spatial GEMM (gemm_tc.cuh) issues nine shifted TMA loads of the same 128-pixel patch, i.e. it pulls the
// activation through L2 nine times. For wide outputs that traffic is amortised over BLOCK_N = 256 columns; for this
// N = 32 convolution it is the bound (22 GB of L2 -> SM traffic per ViT-L forward at batch 32). Here an M-tile is
// 8 x 16 output pixels and its (8+2) x (16+2) input halo is loaded ONCE per 64-channel chunk, as a 10 x 18 pixel TMA
// box (128 B per pixel). The A operand of tap (dy, dx) is then a shifted VIEW of that tile:
//     row m = ty * 8 + tx  ->  halo pixel (ty + dy, tx + dx)  ->  byte ((ty + dy) * 10 + tx + dx) * 128
// i.e. a K-major SWIZZLE_128B operand whose 8-row groups (the eight tx of one ty: 1024 contiguous bytes) lie
// SBO = 1280 B apart and whose start address is ((dy * 10 + dx) * 128) past the 1024-aligned tile base. Neither is a
// multiple of 1024 - that is fine: measured on B200, the 128-byte swizzle of both the TMA write and the tcgen05 operand
// read is a pure function of the shared-memory ADDRESS bits (16-byte chunk index ^= address bits [7,10)), so any
// 128-byte-aligned row-shifted view of a TMA-written tile reads back consistently (matrix-descriptor base_offset = 0;
// setting it to the start's row phase gives wrong results). All nine taps' weights (9 x C x 32) stay resident in
// shared memory for the life of the persistent CTA.
//
// FUSED_RESIZE: the convolution's input is the bilinear (align_corners=True) up-sampling of a smaller map
// (head_model.py:95-101: conv -> interpolate x1.75 / x2 -> conv). Instead of reading the up-sampled map, four extra
// "halo builder" warps interpolate each halo pixel from the source map (same arithmetic and 16-bit rounding as
// resize_bilinear_ac_kernel) and write it into the stage with the swizzle applied by hand (16-byte chunk index
// ^= pixel row & 7); the up-sampled map (2 GB written + read per ViT-L forward at batch 32) is never materialised.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "gemm_tc.cuh"  // ACT_* , rcp_approx, pack2 / unpack2

namespace dpt {

constexpr int HALO_TW = 8, HALO_TH = 16;    // output pixels per tile (x, y): 128 accumulator rows
constexpr int HALO_PW = HALO_TW + 2, HALO_PH = HALO_TH + 2;   // halo box in pixels
constexpr int HALO_A_BYTES = (HALO_PW * HALO_PH * 128 + 1023) / 1024 * 1024;  // one 64-channel chunk, stage-aligned: 23 552 B
constexpr int HALO_STAGES = 6;
constexpr int HALO_N = 32;
constexpr int HALO_W_TILE_BYTES = HALO_N * 128;  // weights of one (tap, 64-channel chunk): 4 096 B
constexpr int HALO_MAX_KCHUNKS = 2;              // C <= 128
constexpr int HALO_THREADS = 192;                // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int HALO_BUILDERS = 256;               // FUSED_RESIZE: warps 6..13 build the halo tiles
constexpr int HALO_THREADS_FUSED = HALO_THREADS + HALO_BUILDERS;
constexpr int HALO_SMEM_BYTES = HALO_STAGES * HALO_A_BYTES + 9 * HALO_MAX_KCHUNKS * HALO_W_TILE_BYTES + 256;

struct __align__(64) HaloParams {
  CUtensorMap tmA;  // 4-D (C, W, H, B), box (64, 10, 18, 1), 128B swizzle
  CUtensorMap tmB;  // 2-D (9 * kpad, 32), box (64, 32), 128B swizzle
  int W, H, B;
  int tiles_x, tiles_y;
  int kchunks;      // 64-channel chunks (1 or 2)
  int is_bf16;
  const float* bias;  // [32]
  float head_w[32];
  float head_b;
  int head_act;     // ACT_RELU or ACT_SIGMOID
  void* out;        // [B, H, W] 16-bit
  // FUSED_RESIZE: source map [B, IH, IW, C] 16-bit; the convolution runs on its bilinear resize to H x W
  const void* src;
  int IH, IW, C;
};

// K-major SWIZZLE_128B descriptor with explicit stride between 8-row groups and swizzle phase of the first row
DPT_DEVICE uint64_t make_smem_desc_sw128_ex(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_offset & 7) << 49;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool BF16, bool FUSED_RESIZE = false>
__global__ void __launch_bounds__(FUSED_RESIZE ? HALO_THREADS_FUSED : HALO_THREADS, 1)
    conv3x3_halo_head_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sW = smem + HALO_STAGES * HALO_A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + 9 * HALO_MAX_KCHUNKS * HALO_W_TILE_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + HALO_STAGES;      // [STAGES]
  uint64_t* w_full = bars + 2 * HALO_STAGES;
  uint64_t* tmem_full = w_full + 1;              // [2]
  uint64_t* tmem_empty = w_full + 3;             // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_full + 5);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.B * p.tiles_y * p.tiles_x;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("dpt conv_halo: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    if (!FUSED_RESIZE) prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    for (int i = 0; i < HALO_STAGES; ++i) {
      mbar_init(&full_bar[i], FUSED_RESIZE ? HALO_BUILDERS : 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_ptr_smem, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp_idx == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      // weights (do not depend on the previous kernel's output, but come after pdl_wait for simplicity)
      mbar_arrive_expect_tx(w_full, 9 * p.kchunks * HALO_W_TILE_BYTES);
      for (int tap = 0; tap < 9; ++tap)
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_2d(sW + (tap * HALO_MAX_KCHUNKS + kc) * HALO_W_TILE_BYTES, &p.tmB, w_full,
                      (tap * p.kchunks + kc) * 64, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; !FUSED_RESIZE && tile < total_tiles; tile += gridDim.x) {
        int t = tile;
        const int tx = t % p.tiles_x;
        t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int b = t / p.tiles_y;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], HALO_PW * HALO_PH * 128);
          // halo origin (x0 - 1, y0 - 1): negative / past-the-edge coordinates are zero-filled = zero padding
          tma_load_4d(sA + s * HALO_A_BYTES, &p.tmA, &full_bar[s], kc * 64, tx * HALO_TW - 1, ty * HALO_TH - 1, b);
          if (++s == HALO_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer =====================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, HALO_N, BF16, false, false);
      mbar_wait(w_full, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * HALO_N;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + s * HALO_A_BYTES);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap % 3;  // halo offsets (0..2) = conv offsets + 1
            const uint64_t a_desc = make_smem_desc_sw128_ex(a_base + (dy * HALO_PW + dx) * 128, HALO_PW * 128, 0);
            const uint64_t b_desc = make_smem_desc_sw128(smem_u32(sW + (tap * HALO_MAX_KCHUNKS + kc) * HALO_W_TILE_BYTES));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | tap | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == HALO_STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full[as]);
      }
    }
    __syncwarp();
  } else if (FUSED_RESIZE && warp_idx >= 6) {
    // ===================================== halo builders =====================================
    const int bt = threadIdx.x - HALO_THREADS;  // 0..127
    const int ch = bt & 7;                      // 16-byte chunk (8 channels) of a pixel's 128-byte row
    const float sy = p.H > 1 ? (float)(p.IH - 1) / (float)(p.H - 1) : 0.0f;
    const float sx = p.W > 1 ? (float)(p.IW - 1) / (float)(p.W - 1) : 0.0f;
    const uint16_t* src = reinterpret_cast<const uint16_t*>(p.src);
    constexpr int is_bf16 = BF16 ? 1 : 0;
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      int t = tile;
      const int tx = t % p.tiles_x;
      t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int b = t / p.tiles_y;
      const size_t img_off = (size_t)b * p.IH * p.IW * p.C;
      {
        // The source map was written by the previous kernel and is larger than L2: first touches are DRAM latency.
        // Pull the source footprint of the tile this CTA builds HALO_PF tiles from now into L2 (one line per thread).
        constexpr int HALO_PF = 3;
        const int tn = tile + HALO_PF * (int)gridDim.x;
        if (tn < total_tiles) {
          int t2 = tn;
          const int ntx = t2 % p.tiles_x;
          t2 /= p.tiles_x;
          const int nty = t2 % p.tiles_y;
          const int nb = t2 / p.tiles_y;
          const int ys = max((int)(sy * (nty * HALO_TH - 1)), 0), xs = max((int)(sx * (ntx * HALO_TW - 1)), 0);
          const int ye = min((int)(sy * (nty * HALO_TH + HALO_TH)) + 1, p.IH - 1);
          const int xe = min((int)(sx * (ntx * HALO_TW + HALO_TW)) + 1, p.IW - 1);
          const int wpx = xe - xs + 1, npx = wpx * (ye - ys + 1);
          const int lines_per_px = (p.C * 2 + 127) / 128;
          for (int i = bt; i < npx * lines_per_px; i += HALO_BUILDERS) {
            const int px = i / lines_per_px, ln = i - px * lines_per_px;
            const int yy = ys + px / wpx, xx = xs + px % wpx;
            const uint16_t* a = src + (size_t)nb * p.IH * p.IW * p.C + ((size_t)yy * p.IW + xx) * p.C + ln * 64;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
        }
      }
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        const uint32_t stage = smem_u32(sA + s * HALO_A_BYTES);
        const int c0 = kc * 64 + ch * 8;
        // 180 pixels x 8 chunks = 1440 tasks per stage, 256 threads: rounds of NB tasks per thread with all 4*NB source
        // loads of a round in flight before the first use (the source footprint of a tile is ~10 KB: L1 hits)
        constexpr int NB = 3;
        constexpr int PIX_STEP = HALO_BUILDERS / 8;  // 32 pixels between a thread's consecutive tasks
#pragma unroll 1
        for (int pix0 = bt >> 3; pix0 < HALO_PW * HALO_PH; pix0 += NB * PIX_STEP) {
          uint4 v[NB][4];
          float ly[NB], lx[NB];
          bool ok[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int pix = pix0 + i * PIX_STEP;
            const int py = pix / HALO_PW, px = pix - py * HALO_PW;
            const int oy = ty * HALO_TH - 1 + py, ox = tx * HALO_TW - 1 + px;
            ok[i] = pix < HALO_PW * HALO_PH && oy >= 0 && oy < p.H && ox >= 0 && ox < p.W && c0 < p.C;
            if (ok[i]) {
              const float fy = sy * oy, fx = sx * ox;
              const int y0 = min((int)fy, p.IH - 1), x0 = min((int)fx, p.IW - 1);
              const int y1 = min(y0 + 1, p.IH - 1), x1 = min(x0 + 1, p.IW - 1);
              ly[i] = fy - y0;
              lx[i] = fx - x0;
              const uint16_t* base = src + img_off + c0;
              v[i][0] = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * p.IW + x0) * p.C));
              v[i][1] = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y0 * p.IW + x1) * p.C));
              v[i][2] = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y1 * p.IW + x0) * p.C));
              v[i][3] = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)y1 * p.IW + x1) * p.C));
            }
          }
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int pix = pix0 + i * PIX_STEP;
            if (pix >= HALO_PW * HALO_PH) break;
            uint4 o = make_uint4(0u, 0u, 0u, 0u);  // zero padding outside the map / past the channel count
            if (ok[i]) {
              const float hy = 1.0f - ly[i], hx = 1.0f - lx[i];
              const uint32_t* a = reinterpret_cast<const uint32_t*>(&v[i][0]);
              const uint32_t* bq = reinterpret_cast<const uint32_t*>(&v[i][1]);
              const uint32_t* cq = reinterpret_cast<const uint32_t*>(&v[i][2]);
              const uint32_t* d = reinterpret_cast<const uint32_t*>(&v[i][3]);
              uint32_t r[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f00 = unpack2(a[k], is_bf16), f01 = unpack2(bq[k], is_bf16);
                const float2 f10 = unpack2(cq[k], is_bf16), f11 = unpack2(d[k], is_bf16);
                // same expression as resize_bilinear_ac_kernel: hy * (hx v00 + lx v01) + ly * (hx v10 + lx v11)
                const float rx = hy * (hx * f00.x + lx[i] * f01.x) + ly[i] * (hx * f10.x + lx[i] * f11.x);
                const float ry = hy * (hx * f00.y + lx[i] * f01.y) + ly[i] * (hx * f10.y + lx[i] * f11.y);
                r[k] = pack2(rx, ry, is_bf16);
              }
              o = make_uint4(r[0], r[1], r[2], r[3]);
            }
            // 128-byte swizzle by hand: the chunk index is XORed with address bits [7,10) = pixel row & 7
            const uint32_t dst = stage + pix * 128 + ((ch ^ (pix & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(&full_bar[s]);
        if (++s == HALO_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const int q = warp_idx & 3;                 // TMEM lane quarter of this warp
    const int r = q * 32 + lane;                // accumulator row = tile pixel ty * 8 + tx
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    float bias_r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias_r[j] = p.bias != nullptr ? __ldg(p.bias + j) : 0.0f;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      int t = tile;
      const int tx = t % p.tiles_x;
      t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int b = t / p.tiles_y;
      const int x = tx * HALO_TW + (r & 7), y = ty * HALO_TH + (r >> 3);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + lane_addr + as * HALO_N, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      float acc = p.head_b;
#pragma unroll
      for (int j = 0; j < 32; ++j) acc = fmaf(fmaxf(__uint_as_float(v[j]) + bias_r[j], 0.0f), p.head_w[j], acc);
      acc = p.head_act == ACT_SIGMOID ? rcp_approx(1.0f + __expf(-acc)) : fmaxf(acc, 0.0f);
      if (x < p.W && y < p.H) {
        const size_t o = ((size_t)b * p.H + y) * p.W + x;
        if (BF16) reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(acc);
        else reinterpret_cast<__half*>(p.out)[o] = __float2half_rn(acc);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace dpt
