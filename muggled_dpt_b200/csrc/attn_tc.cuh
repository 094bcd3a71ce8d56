// Flash-style self-attention forward for the ViT encoders, head_dim = 64, no mask.
//
//   O[b, i, h*64 + :] = softmax_j( scale * Q[b,i,h,:] . K[b,j,h,:] (+ bias[h,i,j]) ) @ V[b,j,h,:]
//
// Q/K/V are read in place from the fused QKV GEMM output [B, N, 3F] (row order [3][H][64], the reference's
// reshape(B,N,3,H,d).permute(2,0,3,1,4) - transformer_block.py:160) through one 3-D TMA tensor map (3F, N, B):
// rows past N are zero-filled by TMA and masked to -inf before the softmax.
//
// One CTA = one 128-row query tile of one (batch, head); two CTAs are co-resident per SM so one CTA's tensor-core work
// overlaps the other's softmax. Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..5 softmax
// (one query row per thread).
//   S = Q K_j^T        tcgen05.mma M=128 N=128 K=64  -> TMEM cols [0,128)
//   P = exp2(S*c - m)  two passes over S in TMEM (row max, then exp/sum), P written 16-bit to swizzled smem
//   O_j = P V_j        tcgen05.mma M=128 N=64 K=128 (V is the MN-major B operand, straight from the TMA tile)
//                      -> TMEM cols [128 + 64*(j&1), +64); accumulated in registers as O = O*alpha_j + O_j.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "gemm_tc.cuh"  // pack2

namespace dpt {

constexpr int ATT_THREADS = 192;
constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_BN = 128;   // kv rows per step
constexpr int ATT_D = 64;
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
// smem: Q | K[2] | V[2] | P (2 chunks) | barriers
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES + 2) + 256;
constexpr int ATT_TMEM_COLS = 256;

struct __align__(64) AttnParams {
  CUtensorMap tmQKV;  // 3-D (3F, N, B), box (64, 128, 1), 128B swizzle
  int N, H, B, F;
  int is_bf16;
  float scale_log2;   // softmax scale * log2(e)
  void* out;          // [B, N, F] 16-bit
  const void* bias;   // optional additive bias [H, N, ldb] 16-bit (BEiT relative position bias), shared over batch
  long long ldb;      // row stride of bias in elements: a multiple of 128 (whole kv tiles stay in bounds)
};

// 32 consecutive 16-bit bias values (64 B, 16-byte aligned) -> fp32, pre-multiplied by log2(e)
DPT_DEVICE void load_bias32(const uint16_t* src, float (&bf)[32], int is_bf16) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 u = __ldg(s4 + k);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack2(w[t], is_bf16);
      bf[8 * k + 2 * t] = f.x * 1.4426950408889634f;
      bf[8 * k + 2 * t + 1] = f.y * 1.4426950408889634f;
    }
  }
}

// o_acc = o_acc * alpha + O_tile (64 fp32 columns of this thread's TMEM lane)
DPT_DEVICE void fold_o(uint32_t o_addr, float2 (&o_acc)[ATT_D / 2], float alpha) {
  const float2 a2 = make_float2(alpha, alpha);
#pragma unroll
  for (int cc = 0; cc < ATT_D; cc += 32) {
    uint32_t v[32];
    tmem_ld32(o_addr + cc, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      o_acc[cc / 2 + i] = __ffma2_rn(o_acc[cc / 2 + i], a2,
                                     make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])));
  }
}

template <bool HAS_BIAS, bool BF16>
__global__ void __launch_bounds__(ATT_THREADS, 2) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;
  uint8_t* sP = sV + ATT_KV_STAGES * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * ATT_TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_ready = bars + 10;
  uint64_t* o_full = bars + 11;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.N + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("dpt attn: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    prefetch_tmap(&p.tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&o_full[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    tmem_alloc(tmem_ptr_smem, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp_idx == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &p.tmQKV, q_full, h * ATT_D, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        tma_load_3d(sK + s * ATT_TILE_BYTES, &p.tmQKV, &k_full[s], p.F + h * ATT_D, j * ATT_BN, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_3d(sV + s * ATT_TILE_BYTES, &p.tmQKV, &v_full[s], 2 * p.F + h * ATT_D, j * ATT_BN, b);
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer =====================================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_f16(128, ATT_BN, BF16, false, false);
      const uint32_t idesc_o = make_idesc_f16(128, ATT_D, BF16, false, true);  // V: MN-major B operand
      const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ));
      const uint64_t p_desc0 = make_smem_desc_sw128(smem_u32(sP));
      const uint64_t p_desc1 = make_smem_desc_sw128(smem_u32(sP + ATT_TILE_BYTES));
      mbar_wait(q_full, 0);
      // S_0
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      {
        const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
        umma_commit(&k_empty[0]);
        umma_commit(s_full);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        // P_j is in smem and S_j has been fully read out of TMEM
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        if (j + 1 < n_kv) {
          const int s1 = (j + 1) & 1;
          const uint32_t ph1 = ((j + 1) >> 1) & 1;
          mbar_wait(&k_full[s1], ph1);
          tc_fence_after();
          const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK + s1 * ATT_TILE_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          umma_commit(&k_empty[s1]);
          umma_commit(s_full);
        }
        mbar_wait(&v_full[s], ph);
        tc_fence_after();
        const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV + s * ATT_TILE_BYTES));
        const uint32_t d_o = tmem_O + (j & 1) * ATT_D;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t a_desc = (kk < 4 ? p_desc0 : p_desc1) + 2 * (kk & 3);
          // V rows kk*16.. : 16 rows * 128 B = 2048 B -> +128 in the (addr >> 4) field
          umma_f16_ss(d_o, a_desc, v_desc + 128 * kk, idesc_o, kk != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(&o_full[j & 1]);
      }
    }
    __syncwarp();
  } else {
    // ===================================== softmax / output =====================================
    const int q = warp_idx & 3;          // TMEM lane quarter
    const int r = q * 32 + lane;         // query row within the tile
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    constexpr int is_bf16 = BF16 ? 1 : 0;
    const float c = p.scale_log2;
    const float2 c2 = make_float2(c, c);
    float m_run = -INFINITY;   // running max, in exp2 units (score * scale * log2e [+ bias * log2e])
    float2 l_run2 = make_float2(0.0f, 0.0f);
    float alpha_prev = 1.0f;
    float2 o_acc[ATT_D / 2];
#pragma unroll
    for (int i = 0; i < ATT_D / 2; ++i) o_acc[i] = make_float2(0.0f, 0.0f);
    const int qrow = q0 + r;
    const uint16_t* bias_row = nullptr;
    if constexpr (HAS_BIAS) {
      // rows past N read row N-1 (their results are never stored); ldb is a multiple of ATT_BN so whole tiles are
      // in bounds
      bias_row = reinterpret_cast<const uint16_t*>(p.bias) + ((long long)h * p.N + min(qrow, p.N - 1)) * p.ldb;
    }

    uint32_t vbuf[2][32];  // software-pipelined TMEM reads: chunk i+1 is in flight while chunk i is processed
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv0 = j * ATT_BN;
      const bool tail = (kv0 + ATT_BN > p.N);
      const uint32_t s_addr = tmem_S + lane_addr;
      // ---- pass 1: row max
      float m_tile = -INFINITY;
      tmem_ld32(s_addr, vbuf[0]);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        uint32_t(&v)[32] = vbuf[ci & 1];
        const int cc = ci * 32;
        tmem_ld_wait_dep(v);
        tmem_ld32(s_addr + ((ci + 1) & 3) * 32, vbuf[(ci + 1) & 1]);  // after chunk 3: chunk 0 again, for pass 2
        if constexpr (HAS_BIAS) {
          float bf[32];
          load_bias32(bias_row + kv0 + cc, bf, is_bf16);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaf(__uint_as_float(v[i]), c, bf[i]));
        }
        if (tail) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kv0 + cc + i >= p.N) v[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          m_tile = fmaxf(m_tile, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
      }
      if constexpr (!HAS_BIAS) m_tile *= c;  // max(c*s) = c*max(s), c > 0
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = ex2_approx(m_run - m_new);  // first tile: exp2(-inf) = 0
      const float2 neg_m2 = make_float2(-m_new, -m_new);
      // ---- the previous P@V must have consumed the P buffer before it is overwritten
      if (j > 0) mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
      const uint32_t o_prev_addr = tmem_O + lane_addr + ((j - 1) & 1) * ATT_D;
      // ---- pass 2: P = exp2(s*c - m_new), row sum, 16-bit P -> swizzled smem
      float2 l_tile2 = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        uint32_t(&v)[32] = vbuf[ci & 1];
        const int cc = ci * 32;
        tmem_ld_wait_dep(v);
        if (ci + 1 < 4) tmem_ld32(s_addr + (ci + 1) * 32, vbuf[(ci + 1) & 1]);
        else if (j > 0) tmem_ld32(o_prev_addr, vbuf[0]);  // first half of O_{j-1}, folded after the arrive below
        float bf[HAS_BIAS ? 32 : 1];
        if constexpr (HAS_BIAS) load_bias32(bias_row + kv0 + cc, *reinterpret_cast<float(*)[32]>(&bf), is_bf16);
        uint8_t* chunk_base = sP + (cc >> 6) * ATT_TILE_BYTES + r * 128;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // 16 columns at a time keeps the live register set small
          float2 pf[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 sv = make_float2(__uint_as_float(v[16 * hh + 2 * i]), __uint_as_float(v[16 * hh + 2 * i + 1]));
            if constexpr (HAS_BIAS)
              pf[i] = __fadd2_rn(__ffma2_rn(sv, c2, make_float2(bf[16 * hh + 2 * i], bf[16 * hh + 2 * i + 1])), neg_m2);
            else
              pf[i] = __ffma2_rn(sv, c2, neg_m2);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            pf[i].x = ex2_approx(pf[i].x);
            pf[i].y = ex2_approx(pf[i].y);
          }
          if (tail) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (kv0 + cc + 16 * hh + 2 * i >= p.N) pf[i].x = 0.0f;
              if (kv0 + cc + 16 * hh + 2 * i + 1 >= p.N) pf[i].y = 0.0f;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) l_tile2 = __fadd2_rn(l_tile2, pf[i]);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const int phys = (((cc & 63) >> 3) + 2 * hh + ch) ^ (r & 7);
            uint4 o;
            o.x = pack2(pf[4 * ch + 0].x, pf[4 * ch + 0].y, is_bf16);
            o.y = pack2(pf[4 * ch + 1].x, pf[4 * ch + 1].y, is_bf16);
            o.z = pack2(pf[4 * ch + 2].x, pf[4 * ch + 2].y, is_bf16);
            o.w = pack2(pf[4 * ch + 3].x, pf[4 * ch + 3].y, is_bf16);
            *reinterpret_cast<uint4*>(chunk_base + phys * 16) = o;
          }
        }
      }
      l_run2 = __ffma2_rn(l_run2, make_float2(alpha, alpha), l_tile2);
      m_run = m_new;
      // make P visible to the tensor core (async proxy) and release S
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- fold O_{j-1} (scaled by the alpha of step j-1) while the tensor core works on S_{j+1}, P_j V_j
      if (j > 0) {
        const float2 a2 = make_float2(alpha_prev, alpha_prev);
        tmem_ld_wait_dep(vbuf[0]);
        tmem_ld32(o_prev_addr + 32, vbuf[1]);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          o_acc[i] = __ffma2_rn(o_acc[i], a2, make_float2(__uint_as_float(vbuf[0][2 * i]), __uint_as_float(vbuf[0][2 * i + 1])));
        tmem_ld_wait_dep(vbuf[1]);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          o_acc[16 + i] = __ffma2_rn(o_acc[16 + i], a2, make_float2(__uint_as_float(vbuf[1][2 * i]), __uint_as_float(vbuf[1][2 * i + 1])));
      }
      alpha_prev = alpha;
    }
    // last O tile
    {
      const int j = n_kv - 1;
      mbar_wait(&o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      fold_o(tmem_O + lane_addr + (j & 1) * ATT_D, o_acc, alpha_prev);
    }
    if (qrow < p.N) {
      const float inv_l = 1.0f / (l_run2.x + l_run2.y);
      uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + ((long long)b * p.N + qrow) * p.F + h * ATT_D;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 o;
        o.x = pack2(o_acc[4 * ch + 0].x * inv_l, o_acc[4 * ch + 0].y * inv_l, is_bf16);
        o.y = pack2(o_acc[4 * ch + 1].x * inv_l, o_acc[4 * ch + 1].y * inv_l, is_bf16);
        o.z = pack2(o_acc[4 * ch + 2].x * inv_l, o_acc[4 * ch + 2].y * inv_l, is_bf16);
        o.w = pack2(o_acc[4 * ch + 3].x * inv_l, o_acc[4 * ch + 3].y * inv_l, is_bf16);
        *reinterpret_cast<uint4*>(orow + 8 * ch) = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace dpt
