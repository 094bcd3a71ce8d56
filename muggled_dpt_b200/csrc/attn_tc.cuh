// Flash-style self-attention forward for the ViT encoders, head_dim = 64, no mask.
//
//   O[b, i, h*64 + :] = softmax_j( scale * Q[b,i,h,:] . K[b,j,h,:] (+ bias[h,i,j]) ) @ V[b,j,h,:]
//
// Q/K/V are read in place from the fused QKV GEMM output [B, N, 3F] (row order [3][H][64], the reference's
// reshape(B,N,3,H,d).permute(2,0,3,1,4) - transformer_block.py:160) through one 3-D TMA tensor map (3F, N, B):
// rows past N are zero-filled by TMA and masked to -inf before the softmax.
//
// One CTA = one 128-row query tile of one (batch, head); two CTAs are co-resident per SM so one CTA's tensor-core work
// overlaps the other's softmax. Warp roles: warps 0..3 softmax (one query row per thread, 208 registers after
// setmaxnreg), warp 4 TMA producer, warp 5 MMA issuer (+TMEM alloc).
//   S = Q K_j^T        tcgen05.mma (SS) M=128 N=128 K=d  -> TMEM cols [0,128)
//   softmax thread     reads its whole S row (128 fp32) into registers in ONE TMEM pass and releases S at once
//                      (S_{j+1} is computed while P_j is still being exponentiated); exact row max; the stabiliser mu
//                      only moves when the max grows by more than 2^8 ("lazy rescale"), so P <= 256 and the
//                      accumulator rarely needs touching; P = exp2(S*c - mu), packed to 16 bits and written straight
//                      back to TMEM cols [128,192) with tcgen05.st (row = lane, two kv columns per 32-bit column)
//   O += P V_j         tcgen05.mma (TS: A = P from TMEM, B = V from smem, MN-major, straight from the TMA tile)
//                      M=128 N=64 K=128 accumulating in TMEM cols [192,256). On the rare rescale a warp multiplies its
//                      32 accumulator rows by exp2(mu_old - mu_new) through tcgen05.ld / tcgen05.st first.
// Keeping P out of shared memory matters: with P staged in smem the tile's smem traffic (STS P 32 KB + MMA reads of
// Q,K 32 KB and P,V 48 KB + TMA writes 32 KB = 144 KB per kv step, ~1150 clk at 128 B/clk) exceeded the MUFU time
// (16384 ex2 / 16 per clk = 1024 clk) that should bound this kernel at d = 64; P in TMEM takes 64 KB of that away.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"
#include <type_traits>
#include "gemm_tc.cuh"  // pack2

namespace dpt {

// -DATT_TRACE (tools/attn_trace.py, never in the shipped library): lane 0 of each softmax warp of one CTA stamps
// clock64 at the phase boundaries of every kv step; `dep` ties the stamp to the last value the phase produced.
#ifdef ATT_TRACE
__device__ long long g_att_trace[4][16][10];
#define ATT_T(ph, dep)                                                                        \
  if (trace_on && j < 16) {                                                                   \
    long long t_;                                                                             \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) : "r"(dep) : "memory");                  \
    g_att_trace[q][j][ph] = t_;                                                               \
  }
#else
#define ATT_T(ph, dep)
#endif

#ifndef ATT_POLY_PAIRS
#define ATT_POLY_PAIRS 2  // of every 8 element pairs, how many take exp2 on the FMA pipe (ex2_poly2) instead of the MUFU
#endif
constexpr int ATT_THREADS = 256;  // warpgroup 0 = softmax (4 warps), warpgroup 1 = TMA producer, MMA issuer, 2 idle
constexpr int ATT_BM = 128;   // query rows per CTA
constexpr int ATT_BN = 128;   // kv rows per step
constexpr int ATT_D = 64;
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;  // 16 KB
// smem: Q | K[stages] | V[stages] | barriers
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES * (1 + 2 * ATT_KV_STAGES) + 256;
constexpr int ATT_TMEM_COLS = 256;

struct __align__(64) AttnParams {
  CUtensorMap tmQKV;  // 3-D (3F, N, B), box (64, 128, 1), 128B swizzle
  int N, H, B, F;
  int is_bf16;
  float scale_log2;   // softmax scale * log2(e)
  void* out;          // [B, N, F] 16-bit
  const void* bias;   // optional additive bias [H, N, ldb] 16-bit (BEiT relative position bias), shared over batch
  long long ldb;      // row stride of bias in elements: a multiple of 128 (whole kv tiles stay in bounds)
  int bias_wmod;      // bias table index = (batch % bias_wmod) * H + h  (SwinV2: per-window shift masks; else 1)
  CUtensorMap tmBias; // attn64_tc.cuh: the same bias tables as a 3-D map (ldb, N, bias_wmod * H), box (64, 128, 1), 128B swizzle
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

// Rare path of the lazy rescale: this thread's accumulator row (64 fp32 TMEM columns) *= beta. Whole warp calls it.
__device__ __noinline__ void rescale_accumulator(uint32_t o_addr, float beta) {
  const float2 b2 = make_float2(beta, beta);
#pragma unroll
  for (int cc = 0; cc < ATT_D; cc += 32) {
    uint32_t ov[32];
    tmem_ld32(o_addr + cc, ov);
    tmem_ld_wait_dep(ov);
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 t = __fmul2_rn(make_float2(__uint_as_float(ov[i]), __uint_as_float(ov[i + 1])), b2);
      ov[i] = __float_as_uint(t.x);
      ov[i + 1] = __float_as_uint(t.y);
    }
    tmem_st32(o_addr + cc, ov);
  }
  tmem_st_wait();
  tc_fence_before();
}

// HD = features per head: 64 (ViT / BEiT) or 32 (SwinV2). For HD = 32 the 64-column TMA boxes start at the head's
// first column, so the head occupies the first half of every tile: QK^T contracts over K = 32 only, P@V still runs at
// N = 64 and the upper 32 accumulator columns are ignored.
template <bool HAS_BIAS, bool BF16, int HD>
__global__ void __launch_bounds__(ATT_THREADS, 2) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + ATT_TILE_BYTES;
  uint8_t* sV = sK + ATT_KV_STAGES * ATT_TILE_BYTES;
  constexpr int ST = ATT_KV_STAGES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * ATT_TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;           // [ST]
  uint64_t* k_empty = k_full + ST;       // [ST]
  uint64_t* v_full = k_empty + ST;       // [ST]
  uint64_t* v_empty = v_full + ST;       // [ST]
  uint64_t* s_full = v_empty + ST;
  uint64_t* p_ready = s_full + 1;
  uint64_t* o_full = s_full + 2;         // [2]
  uint64_t* s_free = s_full + 4;         // S_j has been read out of TMEM for the last time (S_{j+1} may overwrite it)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 5);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.N + ATT_BN - 1) / ATT_BN;
  // The last kv step holds N - (n_kv-1)*128 rows: only its first `last_chunks` 32-column chunks are computed at all
  // (S = Q K^T at N = 32*last_chunks, exponentials on those chunks only, P@V at K = 32*last_chunks).
  const int last_chunks = (p.N - (n_kv - 1) * ATT_BN + 31) >> 5;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("dpt attn: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    prefetch_tmap(&p.tmQKV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(&o_full[0], 1);
    mbar_init(&o_full[1], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(s_free, 128);
    fence_barrier_init();
  }
  if (warp_idx == 5) {
    tmem_alloc(tmem_ptr_smem, ATT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_P = tmem_base + 128;  // 16-bit P [128 x 128] = 64 columns
  const uint32_t tmem_O = tmem_base + 192;  // single fp32 accumulator [128 x 64]
  pdl_wait();                // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();

  // Register rebalancing (setmaxnreg is per warpgroup): the kernel launches at 128 registers/thread so that two CTAs
  // fit an SM; the producer/MMA warpgroup gives most of its registers back and the softmax warpgroup, which keeps a
  // whole 128-column S row per thread, takes them.
  if (warp_idx >= 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp_idx == 4) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_3d(sQ, &p.tmQKV, q_full, h * HD, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % ST;
        const uint32_t ph = (j / ST) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
        tma_load_3d(sK + s * ATT_TILE_BYTES, &p.tmQKV, &k_full[s], p.F + h * HD, j * ATT_BN, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
        tma_load_3d(sV + s * ATT_TILE_BYTES, &p.tmQKV, &v_full[s], 2 * p.F + h * HD, j * ATT_BN, b);
      }
    }
    __syncwarp();
  } else if (warp_idx == 5) {
    // ===================================== MMA issuer =====================================
    if (elect_one()) {
      const uint32_t idesc_s = make_idesc_f16(128, ATT_BN, BF16, false, false);
      const uint32_t idesc_s_last = make_idesc_f16(128, 32 * last_chunks, BF16, false, false);  // tail: fewer kv columns
      const uint32_t idesc_o = make_idesc_f16(128, ATT_D, BF16, false, true);  // V: MN-major B operand
      const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ));
      mbar_wait(q_full, 0);
      // S_0
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      {
        const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, n_kv == 1 ? idesc_s_last : idesc_s, k != 0);
        umma_commit(&k_empty[0]);
        umma_commit(s_full);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % ST;
        const uint32_t ph = (j / ST) & 1;
        // S_{j+1} as soon as S_j has been read for the last time (the softmax warps are still exponentiating)
        if (j + 1 < n_kv) {
          const int s1 = (j + 1) % ST;
          const uint32_t ph1 = ((j + 1) / ST) & 1;
          mbar_wait(s_free, j & 1);
          mbar_wait(&k_full[s1], ph1);
          tc_fence_after();
          const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK + s1 * ATT_TILE_BYTES));
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, j + 2 == n_kv ? idesc_s_last : idesc_s, k != 0);
          umma_commit(&k_empty[s1]);
          umma_commit(s_full);
        }
        // P_j is in TMEM
        mbar_wait(p_ready, j & 1);
        mbar_wait(&v_full[s], ph);
        tc_fence_after();
        const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV + s * ATT_TILE_BYTES));
        const int kk_end = (j + 1 == n_kv) ? 2 * last_chunks : 8;  // tail: only the 32-column chunks that hold kv rows
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // P columns kk*16.. = 8 TMEM columns; V rows kk*16.. : 16 rows * 128 B = 2048 B -> +128 in the (addr >> 4) field
          if (kk < kk_end) umma_f16_ts(tmem_O, tmem_P + 8 * kk, v_desc + 128 * kk, idesc_o, (j | kk) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(&o_full[j & 1]);
      }
    }
    __syncwarp();
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ===================================== softmax / output =====================================
    const int q = warp_idx & 3;          // TMEM lane quarter
    const int r = q * 32 + lane;         // query row within the tile
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    constexpr int is_bf16 = BF16 ? 1 : 0;
    const float c = p.scale_log2;
    const float2 c2 = make_float2(c, c);
    float mu = 0.0f;       // stabiliser in exp2 units (score * scale * log2e [+ bias * log2e]); >= row max - 8
    float2 l2[4];          // row sum of P, four independent packed accumulators
#pragma unroll
    for (int i = 0; i < 4; ++i) l2[i] = make_float2(0.0f, 0.0f);
    const int qrow = q0 + r;
    const uint16_t* bias_row = nullptr;
    if constexpr (HAS_BIAS) {
      // rows past N read row N-1 (never stored); ldb is a multiple of ATT_BN so whole kv tiles are in bounds
      bias_row = reinterpret_cast<const uint16_t*>(p.bias) +
                 (((long long)(b % p.bias_wmod) * p.H + h) * p.N + min(qrow, p.N - 1)) * p.ldb;
    }
    const uint32_t s_addr = tmem_S + lane_addr;
    const uint32_t o_addr = tmem_O + lane_addr;

    const uint32_t p_addr = tmem_P + lane_addr;

    // one kv step; MASKED = this step holds columns >= N (only the last one can)
#ifdef ATT_TRACE
    const bool trace_on = lane == 0 && blockIdx.x == 3 && blockIdx.y == 1 && blockIdx.z == 1;
#endif
    auto step = [&](const int j, auto nch_tag) {
      constexpr int NCH_TAG = decltype(nch_tag)::value;  // 0: full unmasked step; 1..4: tail step with that many chunks
      constexpr bool MASKED = NCH_TAG != 0;
      ATT_T(0, j);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      ATT_T(1, j);
      const int kv0 = j * ATT_BN;
      // ---- the whole S row -> registers, then S is free for the next QK^T
      // MASKED step: chunks [nch, 4) hold no kv rows at all - never loaded, exponentiated or fed to P@V
      constexpr int nch = MASKED ? NCH_TAG : 4;
      uint32_t sv[4][32];
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
        if (!MASKED || ci < nch) tmem_ld32(s_addr + ci * 32, sv[ci]);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
        if (!MASKED || ci < nch) tmem_ld_wait_dep(sv[ci]);
      tc_fence_before();
      mbar_arrive(s_free);
      ATT_T(2, sv[3][31]);
      if constexpr (HAS_BIAS) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          if (MASKED && ci >= nch) continue;
          float bf[32];
          load_bias32(bias_row + kv0 + ci * 32, bf, is_bf16);
#pragma unroll
          for (int i = 0; i < 32; ++i) sv[ci][i] = __float_as_uint(fmaf(__uint_as_float(sv[ci][i]), c, bf[i]));
        }
      }
      if constexpr (MASKED) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          if (ci >= nch) continue;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kv0 + ci * 32 + i >= p.N) sv[ci][i] = 0xff800000u;  // -inf
        }
      }
      // ---- exact row max (four independent chains)
      float m_t[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (MASKED && ci >= nch) continue;
#pragma unroll
        for (int i = 0; i < 32; i += 8)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            m_t[k] = fmaxf(m_t[k], fmaxf(__uint_as_float(sv[ci][i + 2 * k]), __uint_as_float(sv[ci][i + 2 * k + 1])));
      }
      float m_tile = fmaxf(fmaxf(m_t[0], m_t[1]), fmaxf(m_t[2], m_t[3]));
      if constexpr (!HAS_BIAS) m_tile *= c;  // max(c*s) = c*max(s), c > 0
      // ---- lazy rescale decision: move the stabiliser only when the row max outgrew it by more than 2^8
      const bool need = m_tile > mu + 8.0f;
      const bool warp_rescale = (j > 0) && __any_sync(0xffffffffu, need);
      float beta = 1.0f;
      if (j == 0) {
        mu = m_tile;  // finite: the first step always holds at least one unmasked column
      } else if (warp_rescale) {
        const float mu_new = need ? m_tile : mu;
        beta = ex2_approx(mu - mu_new);  // 1 for rows that keep their stabiliser
        mu = mu_new;
        const float2 b2 = make_float2(beta, beta);
#pragma unroll
        for (int i = 0; i < 4; ++i) l2[i] = __fmul2_rn(l2[i], b2);
      }
      // ---- P = exp2(s*c - mu) (<= 256) in place, in three sweeps (scale/shift, ex2, sum + pack) so that the MUFU
      //      results are consumed long after they are issued; kept packed in registers while P_{j-1} V_{j-1} finishes
      ATT_T(3, __float_as_uint(mu));
      const float2 neg_mu2 = make_float2(-mu, -mu);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (MASKED && ci >= nch) continue;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1]));
          float2 e;
          if constexpr (HAS_BIAS) e = __fadd2_rn(x, neg_mu2);  // bias and scale already applied
          else e = __ffma2_rn(x, c2, neg_mu2);
          sv[ci][i] = __float_as_uint(e.x);
          sv[ci][i + 1] = __float_as_uint(e.y);
        }
      }
      ATT_T(4, sv[3][31]);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (MASKED && ci >= nch) continue;
        // of every 8 pairs, the last ATT_POLY_PAIRS go to the FMA-pipe polynomial, the rest to the MUFU
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if (((i >> 1) & 7) >= 8 - ATT_POLY_PAIRS) {
            const float2 e = ex2_poly2<BF16 ? 3 : 4>(make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1])));
            sv[ci][i] = __float_as_uint(e.x);
            sv[ci][i + 1] = __float_as_uint(e.y);
          } else {
            sv[ci][i] = __float_as_uint(ex2_approx(__uint_as_float(sv[ci][i])));  // ex2(-inf) = 0
            sv[ci][i + 1] = __float_as_uint(ex2_approx(__uint_as_float(sv[ci][i + 1])));
          }
        }
      }
      ATT_T(5, sv[3][31]);
      uint32_t pk[2][32];  // 128 kv columns, two per register: the K-major A operand of P@V, 64 TMEM columns
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        if (MASKED && ci >= nch) {  // never read by the K = 32*nch P@V
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[ci >> 1][(ci & 1) * 16 + i] = 0u;
          continue;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 pf = make_float2(__uint_as_float(sv[ci][2 * i]), __uint_as_float(sv[ci][2 * i + 1]));
          l2[i & 3] = __fadd2_rn(l2[i & 3], pf);
          pk[ci >> 1][(ci & 1) * 16 + i] = pack2(pf.x, pf.y, is_bf16);
        }
      }
      ATT_T(6, pk[1][31]);
      // ---- the previous P@V must be complete: it reads P and writes the accumulator
      if (j > 0) {
        mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        if (warp_rescale) rescale_accumulator(o_addr, beta);  // out of line: rare
      }
      ATT_T(7, j);
      // ---- 16-bit P -> TMEM; P_j V_j is issued once all four warps have arrived
      tmem_st32(p_addr, pk[0]);
      if (!MASKED || nch > 2) tmem_st32(p_addr + 32, pk[1]);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      ATT_T(8, j);
    };
    if (q0 + q * 32 >= p.N) {
      // All 32 query rows of this warp lie past N (tail q tile): no softmax work, only the barrier protocol, in
      // lockstep with the live warps (S_j produced -> released; P@V_{j-1} done -> "P_j ready"). The accumulator rows
      // of this warp take whatever the P columns of TMEM happen to hold and are never stored.
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full, j & 1);
        mbar_arrive(s_free);
        if (j > 0) mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
        mbar_arrive(p_ready);
      }
    } else {
    for (int j = 0; j + 1 < n_kv; ++j) step(j, std::integral_constant<int, 0>{});
    switch (last_chunks) {
      case 1: step(n_kv - 1, std::integral_constant<int, 1>{}); break;
      case 2: step(n_kv - 1, std::integral_constant<int, 2>{}); break;
      case 3: step(n_kv - 1, std::integral_constant<int, 3>{}); break;
      default: step(n_kv - 1, std::integral_constant<int, 4>{}); break;
    }
    // ---- epilogue: O / l
    {
      const int j = n_kv - 1;
      mbar_wait(&o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const float2 ls = __fadd2_rn(__fadd2_rn(l2[0], l2[1]), __fadd2_rn(l2[2], l2[3]));
      const float inv_l = 1.0f / (ls.x + ls.y);
      uint32_t ov[2][32];
      tmem_ld32(o_addr, ov[0]);
      tmem_ld32(o_addr + 32, ov[1]);
      tmem_ld_wait_dep(ov[0]);
      tmem_ld_wait_dep(ov[1]);
      if (qrow < p.N) {
        uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + ((long long)b * p.N + qrow) * p.F + h * HD;
#pragma unroll
        for (int ch = 0; ch < HD / 8; ++ch) {
          const uint32_t* w = &ov[ch >> 2][(ch & 3) * 8];
          uint4 o;
          o.x = pack2(__uint_as_float(w[0]) * inv_l, __uint_as_float(w[1]) * inv_l, is_bf16);
          o.y = pack2(__uint_as_float(w[2]) * inv_l, __uint_as_float(w[3]) * inv_l, is_bf16);
          o.z = pack2(__uint_as_float(w[4]) * inv_l, __uint_as_float(w[5]) * inv_l, is_bf16);
          o.w = pack2(__uint_as_float(w[6]) * inv_l, __uint_as_float(w[7]) * inv_l, is_bf16);
          *reinterpret_cast<uint4*>(orow + 8 * ch) = o;
        }
      }
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace dpt
