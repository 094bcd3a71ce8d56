// Flash-style self-attention forward, second generation: 64-column kv steps, three CTAs per SM.
//
//   O[b, i, h*HD + :] = softmax_j( scale * Q[b,i,h,:] . K[b,j,h,:] (+ bias[h,i,j]) ) @ V[b,j,h,:]
//
// Why this shape (measured on B200: tools/microbench/softmax_lab.cu, sync_lab.cu and the ncu captures under profiles/):
// the softmax of a 128 x 128 score tile costs an SM sub-partition ~770 clk of MUFU time (ex2 at 4 lanes/clk, a quarter
// of the pairs on an FMA-pipe polynomial), ~450 clk of FMA-pipe and ~350 clk of ALU-pipe work against 512 clk of
// tensor-pipe time. One in-order softmax warp issues an instruction every 3-4 clk in that sweep (fixed-latency
// dependencies), so the pipes only fill when a sub-partition has several warps that are COMPUTING at the same time -
// and every hop of the S -> softmax -> P -> O pipeline is long (mbarrier hop ~130 clk, commit -> wake-up ~200 clk,
// four MMAs issue -> complete ~600 clk, CTA prologue ~3000 clk). The first-generation kernel (attn_tc.cuh: a thread
// owns a whole 128-column row = 200 registers -> two softmax warps per sub-partition, 65 % of their time computing)
// ran at ~1600 clk per tile per SM. Here a thread owns a 64-column row per step (<= 120 registers), a CTA needs
// 128 + 32 TMEM columns (S [0,64) fp32, O [64,128) fp32, 16-bit P in its own 32 columns) and 64 KB of shared memory,
// so THREE CTAs are resident per SM, and S_{j+1} is computed while the softmax of step j runs (S is released as soon
// as it has been read into registers), so a softmax warp only ever waits for an MMA that was issued a step earlier.
//
// Per CTA (one 128-row Q tile of one (batch, head)), per kv step j:
//   MMA warp   : S_{j+1} = Q K_{j+1}^T (SS, M128 N64 K=HD) -> TMEM S, as soon as the softmax warps released S_j
//   softmax    : S_j row -> registers, release S; (+bias); exact row max; lazy stabiliser (moves only when the max
//                grows by more than 2^8); P = exp2(S*c - mu) in one fused sweep (scale, ex2, row sum and 16-bit pack of
//                each pair adjacent in program order so the MUFU / FMA / ALU pipes overlap); wait P_{j-1} V_{j-1};
//                P -> TMEM; arrive p_ready
//   MMA warp   : O += P_j V_j   (TS: A = P from TMEM, B = V_j MN-major straight from its TMA tile, M128 N64 K64)
// Q/K/V are read in place from the fused QKV GEMM output [B, N, 3F] (row order [3][H][HD], the reference's
// reshape(B,N,3,H,d).permute(2,0,3,1,4) - transformer_block.py:160) through one 3-D TMA map with 64 x 64 boxes; rows
// past N are zero-filled by TMA and masked to -inf. HD = 32 (SwinV2): the 64-column boxes start at the head's first
// column, QK^T contracts over K = 32 only, P@V runs at N = 64 and the upper 32 accumulator columns are ignored.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <type_traits>
#include "attn_tc.cuh"  // AttnParams, load_bias32, ATT_POLY_PAIRS
#include "ptx.cuh"

namespace dpt {

// -DA2_ABLATE=<bits> (tools/attn_ablate.sh, never in the shipped library): drop parts of the work to see what binds.
//   1 = no exponentials (P = the shifted score), 2 = no MMAs issued (barrier protocol only), 4 = no row max,
//   8 = no row sum, 16 = no 16-bit pack
#ifndef A2_ABLATE
#define A2_ABLATE 0
#endif
constexpr int A2_THREADS = 256;  // warpgroup 0 = softmax (4 warps), warpgroup 1 = TMA producer, MMA issuer, 2 idle
constexpr int A2_BM = 128;       // query rows per CTA
constexpr int A2_BN = 64;        // kv rows per step
constexpr int A2_STAGES = 3;       // K/V stages without bias; with bias one of them becomes the bias tile (same 64 KB)
constexpr int A2_BIAS_BYTES = A2_BM * A2_BN * 2;  // 16 KB: bias[q0 .. q0+127][kv0 .. kv0+63], rows of 128 B, 128B swizzle
constexpr int A2_Q_BYTES = A2_BM * 64 * 2;   // 16 KB
constexpr int A2_KV_BYTES = A2_BN * 64 * 2;  // 8 KB
constexpr int A2_SMEM_BYTES = A2_Q_BYTES + 2 * A2_STAGES * A2_KV_BYTES + 256;  // 64.25 KB -> 3 CTAs per SM
// TMEM: two allocations per CTA, 128 (S | O) + 32 (P) columns -> 3 x 160 = 480 of 512. Residency is capped at three
// CTAs per SM by registers (launch bound) and shared memory, so the second allocation can never deadlock.
constexpr int A2_TMEM_COLS = 128, A2_TMEM_COLS_P = 32;
// launch: 80 registers / thread; softmax warpgroup -> 120, producer / MMA warpgroup -> 40 (3 * 256 * 80 = 61440)

DPT_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
DPT_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
         "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

DPT_DEVICE float fmax3(float a, float b, float c) {  // FMNMX3: one instruction, half the issue slots of two FMNMX
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Rare path of the lazy rescale: this thread's accumulator row (64 fp32 TMEM columns) *= beta. Whole warp calls it.
__device__ __noinline__ void a2_rescale_accumulator(uint32_t o_addr, float beta) {
  const float2 b2 = make_float2(beta, beta);
#pragma unroll 1
  for (int cc = 0; cc < 64; cc += 16) {
    uint32_t ov[16];
    tmem_ld16(o_addr + cc, ov);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const float2 t = __fmul2_rn(make_float2(__uint_as_float(ov[i]), __uint_as_float(ov[i + 1])), b2);
      ov[i] = __float_as_uint(t.x);
      ov[i + 1] = __float_as_uint(t.y);
    }
    tmem_st16(o_addr + cc, ov);
  }
  tmem_st_wait();
  tc_fence_before();
}

// CTA -> (q tile, head, batch). Full 128-row q tiles come first, q fastest within a (batch, head) so that the tiles
// sharing one K/V run together (L2); the partial tail tiles (N % 128 rows: most of their softmax warps idle) are
// dealt last, where they fill the final, partial wave of the grid with cheap work.
DPT_DEVICE void a2_tile_of_cta(const AttnParams& p, int id, int& qt, int& h, int& b) {
  const int n_full = p.N / A2_BM;
  const int bh_all = p.H * p.B;
  int bh;
  if (id < n_full * bh_all) {
    qt = id % n_full;
    bh = id / n_full;
  } else {
    qt = n_full;
    bh = id - n_full * bh_all;
  }
  h = bh % p.H;
  b = bh / p.H;
}

template <bool HAS_BIAS, bool BF16, int HD>
__global__ void __launch_bounds__(A2_THREADS, 3) attn64_tc_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int ST = HAS_BIAS ? A2_STAGES - 1 : A2_STAGES;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + A2_Q_BYTES;
  uint8_t* sV = sK + ST * A2_KV_BYTES;
  uint8_t* sBias = sV + ST * A2_KV_BYTES;  // HAS_BIAS only (the space of the third K/V stage)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A2_Q_BYTES + 2 * A2_STAGES * A2_KV_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;       // [ST]
  uint64_t* k_empty = k_full + ST;   // [ST]
  uint64_t* v_full = k_empty + ST;   // [ST]
  uint64_t* v_empty = v_full + ST;   // [ST]
  uint64_t* s_full = v_empty + ST;   // S_j complete
  uint64_t* s_free = s_full + 1;     // all 128 softmax threads hold their S_j row in registers
  uint64_t* p_ready = s_full + 2;    // all 128 softmax threads wrote their P_j row
  uint64_t* pv_done = s_full + 3;    // P_j V_j complete (P may be overwritten, O may be read)
  uint64_t* bias_full = s_full + 4;  // the bias tile of step j has landed (TMA)
  uint64_t* bias_empty = s_full + 5; // all 128 softmax threads have read it
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 6);  // [2]

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int qt, h, b;
  a2_tile_of_cta(p, (int)blockIdx.x, qt, h, b);
  const int q0 = qt * A2_BM;
  const int n_kv = (p.N + A2_BN - 1) / A2_BN;
  // the last kv step holds N - (n_kv-1)*64 rows: only its first `last_chunks` 32-column chunks are computed at all
  const int last_chunks = (p.N - (n_kv - 1) * A2_BN + 31) >> 5;  // 1 or 2

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("dpt attn64: dynamic smem base not 1024-aligned\n");
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
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(pv_done, 1);
    mbar_init(bias_full, 1);
    mbar_init(bias_empty, 128);
    if constexpr (HAS_BIAS) prefetch_tmap(&p.tmBias);
    fence_barrier_init();
  }
  if (warp_idx == 5) {
    tmem_alloc(tmem_ptr_smem, A2_TMEM_COLS);
    tmem_alloc(tmem_ptr_smem + 1, A2_TMEM_COLS_P);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr_smem[0];
  const uint32_t tmem_S = tmem_base;       // fp32 S [128 x 64]
  const uint32_t tmem_O = tmem_base + 64;  // fp32 accumulator [128 x 64]
  const uint32_t tmem_P = tmem_ptr_smem[1];  // 16-bit P [128 x 64] = 32 columns
  pdl_wait();  // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();

  if (warp_idx >= 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp_idx == 4) {
      // ===================================== TMA producer =====================================
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, A2_Q_BYTES);
        tma_load_3d(sQ, &p.tmQKV, q_full, h * HD, q0, b);
        tma_load_3d(sQ + A2_Q_BYTES / 2, &p.tmQKV, q_full, h * HD, q0 + 64, b);
        for (int j = 0; j < n_kv; ++j) {
          const int s = j % ST;
          const uint32_t ph = (j / ST) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], A2_KV_BYTES);
          tma_load_3d(sK + s * A2_KV_BYTES, &p.tmQKV, &k_full[s], p.F + h * HD, j * A2_BN, b);
          mbar_wait(&v_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[s], A2_KV_BYTES);
          tma_load_3d(sV + s * A2_KV_BYTES, &p.tmQKV, &v_full[s], 2 * p.F + h * HD, j * A2_BN, b);
          if constexpr (HAS_BIAS) {
            // bias tile of step j (single buffer: it is consumed at the very start of the step's softmax, so the next
            // load has almost a whole step to land). Rows / columns past N are zero-filled by TMA and masked later.
            mbar_wait(bias_empty, (j & 1) ^ 1);
            mbar_arrive_expect_tx(bias_full, A2_BIAS_BYTES);
            tma_load_3d(sBias, &p.tmBias, bias_full, j * A2_BN, q0, (b % p.bias_wmod) * p.H + h);
          }
        }
      }
      __syncwarp();
    } else if (warp_idx == 5) {
      // ===================================== MMA issuer =====================================
      if (elect_one()) {
        const uint32_t idesc_s = make_idesc_f16(128, A2_BN, BF16, false, false);
        const uint32_t idesc_s_last = make_idesc_f16(128, 32 * last_chunks, BF16, false, false);  // tail: fewer kv columns
        const uint32_t idesc_o = make_idesc_f16(128, 64, BF16, false, true);                      // V: MN-major B operand
        const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ));
        mbar_wait(q_full, 0);
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        {
          const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            if (!(A2_ABLATE & 2)) umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, n_kv == 1 ? idesc_s_last : idesc_s, k != 0);
          umma_commit(&k_empty[0]);
          umma_commit(s_full);
        }
        for (int j = 0; j < n_kv; ++j) {
          const int s = j % ST;
          const uint32_t ph = (j / ST) & 1;
          // S_{j+1} as soon as S_j sits in the softmax warps' registers (they are still exponentiating)
          if (j + 1 < n_kv) {
            const int s1 = (j + 1) % ST;
            const uint32_t ph1 = ((j + 1) / ST) & 1;
            mbar_wait(s_free, j & 1);
            mbar_wait(&k_full[s1], ph1);
            tc_fence_after();
            const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK + s1 * A2_KV_BYTES));
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              if (!(A2_ABLATE & 2)) umma_f16_ss(tmem_S, q_desc + 2 * k, k_desc + 2 * k, j + 2 == n_kv ? idesc_s_last : idesc_s, k != 0);
            umma_commit(&k_empty[s1]);
            umma_commit(s_full);
          }
          mbar_wait(p_ready, j & 1);
          mbar_wait(&v_full[s], ph);
          tc_fence_after();
          const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV + s * A2_KV_BYTES));
          const int kk_end = (j + 1 == n_kv) ? 2 * last_chunks : 4;  // K = 16 kv rows per MMA
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // P columns kk*16.. = 8 TMEM columns; V rows kk*16.. : 16 rows * 128 B = 2048 B -> +128 in the (addr >> 4) field
            if (kk < kk_end && !(A2_ABLATE & 2)) umma_f16_ts(tmem_O, tmem_P + 8 * kk, v_desc + 128 * kk, idesc_o, (j | kk) != 0);
          }
          umma_commit(&v_empty[s]);
          umma_commit(pv_done);
        }
      }
      __syncwarp();
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
    // ===================================== softmax / output =====================================
    const int q = warp_idx & 3;   // TMEM lane quarter
    const int r = q * 32 + lane;  // query row within the tile
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    constexpr int is_bf16 = BF16 ? 1 : 0;
    const float c = p.scale_log2;
    const float2 c2 = make_float2(c, c);
    float mu = 0.0f;  // stabiliser in exp2 units (score * scale * log2e [+ bias * log2e]); >= row max - 8
    float2 l2[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};  // row sum of P, two packed accumulators
    const int qrow = q0 + r;
    // my row of the bias tile: 128 B, its eight 16-byte chunks XOR-swizzled by (row & 7) (TMA SWIZZLE_128B), which
    // also makes the 32 lanes' 16-byte reads conflict-free
    const uint32_t bias_row_smem = smem_u32(sBias) + r * 128;
    const uint32_t s_addr = tmem_S + lane_addr;
    const uint32_t p_addr = tmem_P + lane_addr;
    const uint32_t o_addr = tmem_O + lane_addr;

    // one kv step; NCH_TAG = 0: full step (two 32-column chunks, no masking); 1 / 2: last step with that many chunks
    auto step = [&](const int j, auto nch_tag) {
      constexpr int NCH_TAG = decltype(nch_tag)::value;
      constexpr bool MASKED = NCH_TAG != 0;
      constexpr int nch = MASKED ? NCH_TAG : 2;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv0 = j * A2_BN;
      uint32_t sv[2][32];
#pragma unroll
      for (int ci = 0; ci < nch; ++ci) tmem_ld32(s_addr + ci * 32, sv[ci]);
#pragma unroll
      for (int ci = 0; ci < nch; ++ci) tmem_ld_wait_dep(sv[ci]);
      tc_fence_before();
      mbar_arrive(s_free);  // S_{j+1} may overwrite S now
      if constexpr (HAS_BIAS) {
        mbar_wait(bias_full, j & 1);
#pragma unroll
        for (int ci = 0; ci < nch; ++ci) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int chunk = ci * 4 + k;  // columns chunk*8 .. +7 of the tile
            const float4 raw = lds_f4(bias_row_smem + (uint32_t)((chunk ^ (r & 7)) << 4));
            const uint32_t w[4] = {__float_as_uint(raw.x), __float_as_uint(raw.y), __float_as_uint(raw.z), __float_as_uint(raw.w)};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 bf = unpack2(w[t], is_bf16);
              const int i = k * 8 + 2 * t;
              sv[ci][i] = __float_as_uint(fmaf(__uint_as_float(sv[ci][i]), c, bf.x * 1.4426950408889634f));
              sv[ci][i + 1] = __float_as_uint(fmaf(__uint_as_float(sv[ci][i + 1]), c, bf.y * 1.4426950408889634f));
            }
          }
        }
        mbar_arrive(bias_empty);
      }
      if constexpr (MASKED) {
#pragma unroll
        for (int ci = 0; ci < nch; ++ci)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kv0 + ci * 32 + i >= p.N) sv[ci][i] = 0xff800000u;  // -inf
      }
      // ---- exact row max (four independent chains of 3-input max)
      float m_t[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (A2_ABLATE & 4) m_t[0] = __uint_as_float(sv[0][0]);
#pragma unroll
      for (int ci = 0; ci < ((A2_ABLATE & 4) ? 0 : nch); ++ci)
#pragma unroll
        for (int i = 0; i < 32; i += 8)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            m_t[k] = fmax3(m_t[k], __uint_as_float(sv[ci][i + 2 * k]), __uint_as_float(sv[ci][i + 2 * k + 1]));
      float m_tile = fmaxf(fmaxf(m_t[0], m_t[1]), fmaxf(m_t[2], m_t[3]));
      if constexpr (!HAS_BIAS) m_tile *= c;  // max(c*s) = c*max(s), c > 0
      // ---- lazy rescale decision: move the stabiliser only when the row max outgrew it by more than 2^8
      const bool need = m_tile > mu + 8.0f;
      const bool warp_rescale = (j > 0) && __any_sync(0xffffffffu, need);
      if (j == 0) {
        mu = m_tile;  // finite: the first step always holds at least one unmasked column
      } else if (warp_rescale) {
        const float mu_new = need ? m_tile : mu;
        const float beta = ex2_approx(mu - mu_new);  // 1 for rows that keep their stabiliser
        mu = mu_new;
        const float2 b2 = make_float2(beta, beta);
        l2[0] = __fmul2_rn(l2[0], b2);
        l2[1] = __fmul2_rn(l2[1], b2);
        mbar_wait(pv_done, (j - 1) & 1);  // P_{j-1} V_{j-1} must be complete before the accumulator is touched
        tc_fence_after();
        a2_rescale_accumulator(o_addr, beta);
      }
      // ---- P = exp2(s*c - mu) (<= 256): ONE fused sweep - scale/shift, ex2, row sum and 16-bit pack of a pair sit next
      //      to each other in program order, so the scheduler has MUFU, FMA-pipe and ALU-pipe work to interleave.
      //      Of every 8 pairs the last ATT_POLY_PAIRS take exp2 on the FMA pipe (ex2_poly2) instead of the MUFU.
      const float2 neg_mu2 = make_float2(-mu, -mu);
      uint32_t pk[32];  // 64 kv columns, two per register: the K-major A operand of P@V, 32 TMEM columns
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        if (ci >= nch) {  // never read by the K = 32*nch P@V
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[ci * 16 + i] = 0u;
          continue;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 x = make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1]));
          float2 e;
          if constexpr (HAS_BIAS) e = __fadd2_rn(x, neg_mu2);  // bias and scale already applied
          else e = __ffma2_rn(x, c2, neg_mu2);
          float2 pf;
          if (A2_ABLATE & 1) {
            pf = e;
          } else if (((i >> 1) & 7) >= 8 - ATT_POLY_PAIRS) {
            pf = ex2_poly2<BF16 ? 3 : 4>(e);
          } else {
            pf.x = ex2_approx(e.x);  // ex2(-inf) = 0
            pf.y = ex2_approx(e.y);
          }
          if (!(A2_ABLATE & 8)) l2[(i >> 1) & 1] = __fadd2_rn(l2[(i >> 1) & 1], pf);
          if (A2_ABLATE & 16) pk[ci * 16 + (i >> 1)] = __float_as_uint(pf.x) ^ __float_as_uint(pf.y);
          else pk[ci * 16 + (i >> 1)] = pack2(pf.x, pf.y, is_bf16);
        }
      }
      // ---- 16-bit P -> TMEM once P_{j-1} V_{j-1} has consumed the previous P
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      if constexpr (nch == 2) {
        tmem_st32(p_addr, pk);
      } else {
        uint32_t lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) lo[i] = pk[i];
        tmem_st16(p_addr, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
    };

    if (q0 + q * 32 >= p.N) {
      // All 32 query rows of this warp lie past N (tail q tile): no softmax work, only the barrier protocol. The
      // accumulator rows of this warp take whatever the P columns of TMEM happen to hold and are never stored.
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(s_full, j & 1);
        mbar_arrive(s_free);
        if constexpr (HAS_BIAS) {
          mbar_wait(bias_full, j & 1);
          mbar_arrive(bias_empty);
        }
        if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
        mbar_arrive(p_ready);
      }
    } else {
      for (int j = 0; j + 1 < n_kv; ++j) step(j, std::integral_constant<int, 0>{});
      if (last_chunks == 1) step(n_kv - 1, std::integral_constant<int, 1>{});
      else step(n_kv - 1, std::integral_constant<int, 2>{});
      // ---- epilogue: O / l
      mbar_wait(pv_done, (n_kv - 1) & 1);
      tc_fence_after();
      const float2 ls = __fadd2_rn(l2[0], l2[1]);
      const float inv_l = 1.0f / (ls.x + ls.y);
      uint16_t* orow = reinterpret_cast<uint16_t*>(p.out) + ((long long)b * p.N + qrow) * p.F + h * HD;
#pragma unroll
      for (int half = 0; half < HD / 32; ++half) {
        uint32_t ov[32];
        tmem_ld32(o_addr + half * 32, ov);
        tmem_ld_wait_dep(ov);
        if (qrow < p.N) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint32_t* w = &ov[ch * 8];
            uint4 o;
            o.x = pack2(__uint_as_float(w[0]) * inv_l, __uint_as_float(w[1]) * inv_l, is_bf16);
            o.y = pack2(__uint_as_float(w[2]) * inv_l, __uint_as_float(w[3]) * inv_l, is_bf16);
            o.z = pack2(__uint_as_float(w[4]) * inv_l, __uint_as_float(w[5]) * inv_l, is_bf16);
            o.w = pack2(__uint_as_float(w[6]) * inv_l, __uint_as_float(w[7]) * inv_l, is_bf16);
            *reinterpret_cast<uint4*>(orow + half * 32 + 8 * ch) = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A2_TMEM_COLS);
    tmem_dealloc(tmem_P, A2_TMEM_COLS_P);
  }
}

}  // namespace dpt
