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
  long long ldb;      // row stride of bias in elements
};

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
      const uint32_t idesc_s = make_idesc_f16(128, ATT_BN, p.is_bf16 != 0, false, false);
      const uint32_t idesc_o = make_idesc_f16(128, ATT_D, p.is_bf16 != 0, false, true);  // V: MN-major B operand
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
    const int is_bf16 = p.is_bf16;
    const float c = p.scale_log2;
    float m_run = -INFINITY;   // running max of raw scores (scaled by c lazily)
    float l_run = 0.0f;
    float alpha_prev = 1.0f;
    float o_acc[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o_acc[i] = 0.0f;
    const int qrow = q0 + r;
    const uint16_t* bias_row = nullptr;
    if (p.bias != nullptr && qrow < p.N)
      bias_row = reinterpret_cast<const uint16_t*>(p.bias) + ((long long)h * p.N + qrow) * p.ldb;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv0 = j * ATT_BN;
      const bool tail = (kv0 + ATT_BN > p.N);
      // ---- pass 1: row max (in units of score*c, bias*log2e folded in)
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < ATT_BN; cc += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float sv = __uint_as_float(v[i]) * c;
          if (bias_row != nullptr && kv0 + cc + i < p.N) {
            const uint16_t raw = bias_row[kv0 + cc + i];
            const float bv = is_bf16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&raw))
                                     : __half2float(*reinterpret_cast<const __half*>(&raw));
            sv = fmaf(bv, 1.4426950408889634f, sv);
          }
          if (tail && kv0 + cc + i >= p.N) sv = -INFINITY;
          m_tile = fmaxf(m_tile, sv);
        }
      }
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = exp2f(m_run - m_new);  // first tile: exp2(-inf) = 0
      // ---- wait until the previous P@V has consumed the P buffer, then fold O_{j-1} later
      if (j > 0) mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
      // ---- pass 2: P = exp2(s - m_new), row sum, 16-bit P -> swizzled smem
      float l_tile = 0.0f;
#pragma unroll 1
      for (int cc = 0; cc < ATT_BN; cc += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_addr + cc, v);
        tmem_ld_wait();
        float pf[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float sv = __uint_as_float(v[i]) * c;
          if (bias_row != nullptr && kv0 + cc + i < p.N) {
            const uint16_t raw = bias_row[kv0 + cc + i];
            const float bv = is_bf16 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&raw))
                                     : __half2float(*reinterpret_cast<const __half*>(&raw));
            sv = fmaf(bv, 1.4426950408889634f, sv);
          }
          float pv = exp2f(sv - m_new);
          if (tail && kv0 + cc + i >= p.N) pv = 0.0f;
          pf[i] = pv;
          l_tile += pv;
        }
        uint8_t* chunk_base = sP + (cc >> 6) * ATT_TILE_BYTES + r * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int lch = ((cc & 63) >> 3) + ch;
          const int phys = lch ^ (r & 7);
          uint4 o;
          o.x = pack2(pf[8 * ch + 0], pf[8 * ch + 1], is_bf16);
          o.y = pack2(pf[8 * ch + 2], pf[8 * ch + 3], is_bf16);
          o.z = pack2(pf[8 * ch + 4], pf[8 * ch + 5], is_bf16);
          o.w = pack2(pf[8 * ch + 6], pf[8 * ch + 7], is_bf16);
          *reinterpret_cast<uint4*>(chunk_base + phys * 16) = o;
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // make P visible to the tensor core (async proxy) and release S
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- fold O_{j-1} (scaled by the alpha of step j-1) while the tensor core works on S_{j+1}, P_j V_j
      if (j > 0) {
        tc_fence_after();
        const uint32_t o_addr = tmem_O + lane_addr + ((j - 1) & 1) * ATT_D;
#pragma unroll
        for (int cc = 0; cc < ATT_D; cc += 32) {
          uint32_t v[32];
          tmem_ld32(o_addr + cc, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[cc + i] = fmaf(o_acc[cc + i], alpha_prev, __uint_as_float(v[i]));
        }
      }
      alpha_prev = alpha;
    }
    // last O tile
    {
      const int j = n_kv - 1;
      mbar_wait(&o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t o_addr = tmem_O + lane_addr + (j & 1) * ATT_D;
#pragma unroll
      for (int cc = 0; cc < ATT_D; cc += 32) {
        uint32_t v[32];
        tmem_ld32(o_addr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[cc + i] = fmaf(o_acc[cc + i], alpha_prev, __uint_as_float(v[i]));
      }
    }
    if (qrow < p.N) {
      const float inv_l = 1.0f / l_run;
      uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + ((long long)b * p.N + qrow) * p.F + h * ATT_D;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 o;
        o.x = pack2(o_acc[8 * ch + 0] * inv_l, o_acc[8 * ch + 1] * inv_l, is_bf16);
        o.y = pack2(o_acc[8 * ch + 2] * inv_l, o_acc[8 * ch + 3] * inv_l, is_bf16);
        o.z = pack2(o_acc[8 * ch + 4] * inv_l, o_acc[8 * ch + 5] * inv_l, is_bf16);
        o.w = pack2(o_acc[8 * ch + 6] * inv_l, o_acc[8 * ch + 7] * inv_l, is_bf16);
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
