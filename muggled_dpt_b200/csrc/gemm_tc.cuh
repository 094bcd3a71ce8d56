// "Spatial GEMM": the one tensor-core kernel behind every Linear / 1x1 / 3x3 / ConvTranspose(k=s) layer of the DPT
// hot path (SURVEY.md §2.2 K1,K4,K6-K8,K10-K14,K16,K17,K19).
//
//   out[pixel(b,y,x), n] = epilogue( sum_{tap} sum_{c} A[b, y+dy(tap), x+dx(tap)+xoff, c] * Wt[n, tap*kpad + c] )
//
// * A is a channels-last activation tensor seen through a 4-D TMA tensor map (C, W, H, B). One M-tile is a
//   TH x TW patch of 128 pixels of one image; a 3x3 convolution is nine shifted TMA loads of the same patch
//   (out-of-bounds rows/columns are zero-filled by the TMA unit == zero padding). Token matrices [M, K] are the
//   degenerate case H = 1, TW = 128; "drop the cls token" is xoff = 1 on a (F, N, 1, B) map.
// * Wt is [N, taps*kpad] K-major (nn.Linear / packed conv weights), loaded by a 2-D TMA map.
// * tcgen05.mma (M=128, N=BLOCK_N, K=16, bf16/fp16 -> fp32; cta_group::2 CTA pairs on 256 x 256 / 256 x 128 tiles for
//   the large launches) accumulates in TMEM, double-buffered so the epilogue of tile i overlaps the main loop of tile
//   i+1. Persistent CTAs, static round-robin tile schedule.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue (two warpgroups, each
//   owning half of the accumulator columns; a warp can only read TMEM lanes 32*(warp%4)..+31).
// * Epilogue: TMEM -> regs -> (+bias, GELU/ReLU) -> swizzled smem staging -> coalesced 16-byte global IO with optional
//   residual / skip addends, optional second ReLU'd copy, fp32 or 16-bit output, output pixel remap
//   (y*so+oy, x*so+ox) for pixel-shuffle (ConvTranspose) stores - the sub-pixel either fixed per launch or picked by
//   the n-tile (shuffle_n: one launch per ConvTranspose); "head" mode reduces 32 channels to one depth value.
// * Folded LayerNorm (pre-norm transformer blocks): the consumer applies per-row (rstd, mean) and per-column weight
//   sums in its epilogue, the fp32-output producer writes the per-row partial statistics and a 16-bit copy of the
//   residual stream (GemmParams ln_* / stats_out / out16); RESPF: cp.async prefetch of the fp32 residual for short K.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

namespace dpt {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_STAGE_BYTES_PER_WARP = 4096;  // 32 rows x 128 B

// ACT_SWIGLU (16-bit outputs, BLOCK_N >= 128): the GEMM columns come in blocks of 64 = 32 gate columns followed by the 32
// linear columns of the same features (the packer interleaves the rows of the doubled Linear that way); the epilogue
// writes silu(gate) * linear, so the output has N / 2 columns and the [M, N] intermediate never exists.
enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_SIGMOID = 3, ACT_SWIGLU = 4 };
enum : int { OUT_HALF = 0, OUT_F32 = 1, OUT_HEAD = 2 };

struct __align__(64) GemmParams {
  CUtensorMap tmA;  // 4-D (C, W, H, B), box (64, TW, TH, 1), 128B swizzle
  CUtensorMap tmB;  // 2-D (taps*kpad, N), box (64, BLOCK_N) - (64, BLOCK_N/2) for the 2-CTA kernel -, 128B swizzle
  int W, H, B;      // input spatial extent used for M tiling (and output-row validity)
  int tw_log2;      // TW = 1 << tw_log2, TH = 128 >> tw_log2
  int tiles_x, tiles_y;
  int N, n_tiles;
  int num_taps;     // 1 (1x1 / linear) or 9 (3x3, pad 1)
  int kchunks;      // 64-wide K chunks per tap
  int a_xoff;       // added to the x coordinate of every A load (token mode: 1 skips the cls row)
  int is_bf16;      // 16-bit type of A / Wt / 16-bit outputs: 1 = bf16, 0 = fp16
  // epilogue
  const float* bias;  // [N] or null; bias_bstride != 0: one bias vector per image b at bias + b * bias_bstride
  long long bias_bstride;
  int act;
  int out_kind;
  void* out;
  long long ldo;      // elements between consecutive output pixels
  int OH, OW, so, oy, ox;
  int shuffle_n;      // > 0: ConvTranspose2d(k = s = so) in ONE launch - the N axis is [so*so][shuffle_n]: n-tile n_blk
                      // writes channels (n % shuffle_n) of sub-pixel (oy, ox) = divmod(n / shuffle_n, so); BLOCK_N | shuffle_n
  const void* add1;   // same dtype + pixel indexing as out (may alias out: in-place residual)
  long long ld_add1;
  const void* add2;
  long long ld_add2;
  void* out2_relu;    // optional second output relu(result), 16-bit, same pixel indexing
  long long ld_out2;
  // LayerNorm folded into the consumer GEMM (pre-norm transformer blocks): A holds the RAW 16-bit residual stream and
  // Wt = W (.) g, so  LN(x) W^T + b = rstd_m * acc - rstd_m * mean_m * colsum_n + b'_n  with the per-row statistics
  // produced by the epilogue of the GEMM that wrote the residual stream (stats_out below).
  const float* ln_stats;   // [rows, ln_parts, 2] partial (sum, sum of squares) of the fp32 row, or null
  const float* ln_colsum;  // [N] sum_k Wt[n, k] of the 16-bit weights
  int ln_parts;
  float ln_inv_f, ln_eps;  // 1 / row length, epsilon
  // OUT_F32 producer side: per-row partial statistics of the final fp32 values (one slot per 128-column half tile,
  // slot = n_blk * 2 + column half; deterministic, no atomics) and a 16-bit copy of the row
  float* stats_out;        // [rows, stats_parts, 2] or null
  int stats_parts;
  void* out16;             // 16-bit copy of the fp32 output (same pixel indexing), or null
  long long ld_out16;
  // SwinV2 cosine attention folded into the QKV GEMM (windowed_attention.py:99-111): with 32 features per head a
  // thread's 32-column epilogue unit is exactly one head of one token, so q <- normalize(q) * logit_scale[h] and
  // k <- normalize(k) (F.normalize, eps 1e-12) are applied to the fp32 accumulator before its single 16-bit rounding.
  const float* qk_logit;   // [heads] (already exp'd and clamped), or null
  int qk_features;         // F: columns [0, F) are q, [F, 2F) k, [2F, 3F) v
  // Plain 16-bit outputs (no addends, no pixel shuffle, BLOCK_N >= 128): each epilogue warp hands its staged 32-row x
  // 64-column chunk to the TMA (box = the warp's rows inside the tile, 128B swizzle = the staging layout) instead of
  // 8 shuffle + LDS + STG passes; rows / columns outside the tensor are clipped by the TMA unit.
  CUtensorMap tmOut;
  int tma_store;
  // 1 / n_tiles, 1 / tiles_x, 1 / tiles_y: the tile -> (n, x, y, image) decomposition runs once per tile on the critical
  // path of the producer and of every epilogue thread; fast_div replaces six integer divisions (~1000 clocks) there
  float inv_n_tiles, inv_tiles_x, inv_tiles_y;
  float head_w[32];   // OUT_HEAD: depth = act2(relu(acc + bias) . head_w + head_b)
  float head_b;
  int head_act;       // ACT_RELU or ACT_SIGMOID
};

// -DGEMM_TRACE (tools/gemm_trace.py, never in the shipped library): lane 0 of every epilogue warp and the MMA thread of
// one CTA stamp clock64() at the phase boundaries of their first eight tiles.
#ifdef GEMM_TRACE
__device__ long long g_gemm_trace[9][8][16];  // [epilogue warp 0..7 | 8 = MMA issuer][tile][stamp]
#define GEMM_STAMP(role, it_, k_)                                            \
  do {                                                                       \
    if (trace_cta && (it_) < 8) g_gemm_trace[role][it_][k_] = clock64();     \
  } while (0)
#else
#define GEMM_STAMP(role, it_, k_) do { } while (0)
#endif

// TWO_CTA: a CTA pair (cluster of 2, cta_group::2) computes a 256 x BLOCK_N tile; each CTA stages its own 128 rows of
// A and HALF of the B tile (the tensor core reads the other half from the peer's shared memory), which cuts the
// L2 -> SM operand traffic per FLOP by a third and buys two more pipeline stages.
// RESPF kernels (OUT_F32 with a short K loop, see gemm_tc_kernel) carry one more per-warp buffer - the prefetched fp32
// residual of the next 32-column unit - and pay for it with one pipeline stage.
template <int BLOCK_N, bool TWO_CTA = false, bool F32OUT = false>
struct GemmCfg {
  // BLOCK_N = 384 (CTA pairs, fp32 outputs only): 256 x 384 tiles for the small-M residual GEMMs - three n-tiles cover
  // N = 1024, so up to 24 m-pairs (M <= 6144) finish in ONE round on 72 of the 74 SM pairs where 128 x 128 tiles need
  // three. One accumulator stage (384 of the 512 TMEM columns), two MMAs per K step (N = 256 + N = 128), 4 stages.
  static constexpr bool WIDE = BLOCK_N == 384;
  static constexpr int ACC_STAGES = WIDE ? 1 : 2;
  static constexpr int STAGES_BASE = WIDE ? 4 : (TWO_CTA ? (BLOCK_N == 256 ? 6 : 8) : (BLOCK_N == 256 ? 4 : (BLOCK_N == 128 ? 6 : 8)));
  static constexpr int STAGES = (!F32OUT || WIDE) ? STAGES_BASE
                                : (BLOCK_N == 64 ? 6 : (BLOCK_N == 32 ? 8 : ((TWO_CTA && BLOCK_N == 128) ? STAGES_BASE - 2 : STAGES_BASE - 1)));
  static constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int B_BYTES = (TWO_CTA ? BLOCK_N / 2 : BLOCK_N) * GEMM_BLOCK_K * 2;
  static constexpr int STAGING_BYTES = GEMM_EPI_WARPS * GEMM_STAGE_BYTES_PER_WARP;
  static constexpr int RES_BYTES = F32OUT ? GEMM_EPI_WARPS * GEMM_STAGE_BYTES_PER_WARP : 0;
  // mbarriers + tmem ptr, bias tile + colsum tile [BLOCK_N] each (the wide fp32-output tiles have no folded LayerNorm)
  static constexpr int VEC_TILES = WIDE ? 1 : 2;
  static constexpr int BAR_BYTES = 256 + VEC_TILES * BLOCK_N * 4;
  static constexpr int SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + STAGING_BYTES + RES_BYTES + BAR_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
  static constexpr int TMEM_COLS = WIDE ? 512 : ((2 * BLOCK_N) < 32 ? 32 : (2 * BLOCK_N));
};

DPT_DEVICE float gelu_erf(float x) {
  // exact-erf GELU (nn.GELU default): 0.5 x (1 + erf(x / sqrt 2)). erf via Abramowitz-Stegun 7.1.28
  // (|err| < 3e-7): erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16, z >= 0.
  const float z = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(z, 0.0000430638f, 0.0002765672f);
  p = fmaf(z, p, 0.0001520143f);
  p = fmaf(z, p, 0.0092705272f);
  p = fmaf(z, p, 0.0422820123f);
  p = fmaf(z, p, 0.0705230784f);
  p = fmaf(z, p, 1.0f);
  p = p * p;
  p = p * p;
  p = p * p;
  p = p * p;
  const float e = 1.0f - rcp_approx(p);  // erf(|x|/sqrt2)
  const float half_x = 0.5f * x;
  return fmaf(copysignf(e, x), half_x, half_x);
}

// two GELUs at once with packed fp32x2 FMA (sm_100 FFMA2): same formula as gelu_erf
// two GELUs at once with packed fp32x2 FMA (sm_100 FFMA2). Same erf approximation as gelu_erf, arranged for fewer
// instructions: with a = |x| and r = 1 - erf(a / sqrt 2) = p(a)^-16 (the sqrt 2 folded into the coefficients),
// gelu(x) = relu(x) - 0.5 a r   [x >= 0: 0.5 x (2 - r); x < 0: 0.5 x r] - no copysign, no 1 - r.
DPT_DEVICE float2 gelu_erf2(float2 x) {
  const float2 a = make_float2(fabsf(x.x), fabsf(x.y));
  float2 p = __ffma2_rn(a, make_float2(5.3829750e-06f, 5.3829750e-06f), make_float2(4.8890636e-05f, 4.8890636e-05f));
  p = __ffma2_rn(a, p, make_float2(3.8003575e-05f, 3.8003575e-05f));
  p = __ffma2_rn(a, p, make_float2(3.2776263e-03f, 3.2776263e-03f));
  p = __ffma2_rn(a, p, make_float2(2.1141006e-02f, 2.1141006e-02f));
  p = __ffma2_rn(a, p, make_float2(4.9867347e-02f, 4.9867347e-02f));
  p = __ffma2_rn(a, p, make_float2(1.0f, 1.0f));
  p = __fmul2_rn(p, p);
  p = __fmul2_rn(p, p);
  p = __fmul2_rn(p, p);
  p = __fmul2_rn(p, p);
  const float2 rr = make_float2(rcp_approx(p.x), rcp_approx(p.y));
  const float2 nh = __fmul2_rn(a, make_float2(-0.5f, -0.5f));
  return __ffma2_rn(nh, rr, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

// floor(n / d) for 0 <= n < 2^23, d >= 1, inv = 1.0f / d: the float product is off by at most one either way
DPT_DEVICE int fast_div(int n, int d, float inv) {
  int q = (int)((float)n * inv);
  const int r = n - q * d;
  if (r < 0) --q;
  else if (r >= d) ++q;
  return q;
}

DPT_DEVICE uint32_t pack2(float a, float b, int is_bf16) {
  if (is_bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
DPT_DEVICE float2 unpack2(uint32_t u, int is_bf16) {
  if (is_bf16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// RESPF (OUT_F32 only): prefetch the fp32 residual with cp.async one unit ahead. Worth a pipeline stage when the K
// loop is short (proj: the epilogue's global-load latency is on the critical path), not when it is long (fc2).
template <int BLOCK_N, int OUT_KIND, int ACT, bool BF16, bool TWO_CTA = false, bool RESPF = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  static_assert(!RESPF || OUT_KIND == OUT_F32, "residual prefetch exists for fp32 outputs only");
  using Cfg = GemmCfg<BLOCK_N, TWO_CTA, RESPF>;
  static_assert(!TWO_CTA || BLOCK_N == 384 || BLOCK_N == 256 || BLOCK_N == 128, "the 2-CTA kernel is built for BLOCK_N = 384, 256 and 128");
  static_assert(BLOCK_N != 384 || (TWO_CTA && OUT_KIND == OUT_F32), "384-wide tiles: CTA pairs with fp32 output only");
  constexpr bool WIDE = Cfg::WIDE;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;
  constexpr int STAGES = Cfg::STAGES;

  // no static shared memory in this kernel: the dynamic segment starts at the CTA's (1024-aligned) window base;
  // checked below because the 128B-swizzle descriptors depend on it
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem_b + STAGES * Cfg::B_BYTES;
  uint8_t* resbuf = staging + Cfg::STAGING_BYTES;  // OUT_F32: per-warp prefetched residual unit
  float* bias_s = reinterpret_cast<float*>(resbuf + Cfg::RES_BYTES);  // bias [BLOCK_N] | LN colsum [BLOCK_N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + Cfg::VEC_TILES * BLOCK_N);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef GEMM_TRACE
  const bool trace_cta = blockIdx.x == 4 && lane == 0;
#endif

  const int m_tiles = p.B * p.tiles_y * p.tiles_x;
  const int num_kb = p.num_taps * p.kchunks;
  // Work items: tiles for the 1-CTA kernel; for the 2-CTA kernel a pair of vertically adjacent M-tiles
  // (m_tile = 2 * pair + cta_rank; an odd trailing M-tile computes against zero-filled rows and stores nothing)
  const uint32_t cta_rank = TWO_CTA ? cluster_ctarank() : 0u;
  const int total_tiles = (TWO_CTA ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles;
  const int work_first = TWO_CTA ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_stride = TWO_CTA ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp_idx == 0 && lane == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("dpt gemm: dynamic smem base not 1024-aligned\n");
      __trap();
    }
    prefetch_tmap(&p.tmA);
    prefetch_tmap(&p.tmB);
    if (p.tma_store) prefetch_tmap(&p.tmOut);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], TWO_CTA ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);  // leader's: both CTAs' epilogues
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    if constexpr (TWO_CTA) {
      tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (TWO_CTA) cluster_sync_all();  // barrier inits visible to the peer before any remote arrive / TMA
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();                // everything above overlapped the previous kernel's tail
  pdl_launch_dependents();   // and the next kernel may do the same with ours

  if (warp_idx == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const int TW = 1 << p.tw_log2, TH = GEMM_BLOCK_M >> p.tw_log2;
      for (int tile = work_first; tile < total_tiles; tile += work_stride) {
        const int m_idx = fast_div(tile, p.n_tiles, p.inv_n_tiles);
        const int n_blk = tile - m_idx * p.n_tiles;
        const int mt = m_idx * (TWO_CTA ? 2 : 1) + (int)cta_rank;
        const int mrow = fast_div(mt, p.tiles_x, p.inv_tiles_x);
        const int tx = mt - mrow * p.tiles_x;
        const int b = fast_div(mrow, p.tiles_y, p.inv_tiles_y);  // == p.B for the odd trailing M-tile of a pair: TMA zero-fills
        const int ty = mrow - b * p.tiles_y;
        const int x0 = tx * TW + p.a_xoff, y0 = ty * TH;
        for (int tap = 0; tap < p.num_taps; ++tap) {
          const int dy = p.num_taps == 9 ? tap / 3 - 1 : 0;
          const int dx = p.num_taps == 9 ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if constexpr (TWO_CTA) {
              // the leader's barrier collects the bytes of both CTAs; completion is signalled there
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
              tma_load_4d_2sm(smem_a + s * Cfg::A_BYTES, &p.tmA, &full_bar[s], kc * GEMM_BLOCK_K, x0 + dx, y0 + dy, b);
              if constexpr (WIDE) {
                // my half of the N = 256 MMA's B rows (two 64-row boxes), then my half of the N = 128 MMA's
                const int kx = (tap * p.kchunks + kc) * GEMM_BLOCK_K;
                uint8_t* sb = smem_b + s * Cfg::B_BYTES;
                tma_load_2d_2sm(sb, &p.tmB, &full_bar[s], kx, n_blk * BLOCK_N + (int)cta_rank * 128);
                tma_load_2d_2sm(sb + 8192, &p.tmB, &full_bar[s], kx, n_blk * BLOCK_N + (int)cta_rank * 128 + 64);
                tma_load_2d_2sm(sb + 16384, &p.tmB, &full_bar[s], kx, n_blk * BLOCK_N + 256 + (int)cta_rank * 64);
              } else
              tma_load_2d_2sm(smem_b + s * Cfg::B_BYTES, &p.tmB, &full_bar[s], (tap * p.kchunks + kc) * GEMM_BLOCK_K,
                              n_blk * BLOCK_N + (int)cta_rank * (BLOCK_N / 2));
            } else {
              mbar_arrive_expect_tx(&full_bar[s], Cfg::A_BYTES + Cfg::B_BYTES);
              tma_load_4d(smem_a + s * Cfg::A_BYTES, &p.tmA, &full_bar[s], kc * GEMM_BLOCK_K, x0 + dx, y0 + dy, b);
              tma_load_2d(smem_b + s * Cfg::B_BYTES, &p.tmB, &full_bar[s], (tap * p.kchunks + kc) * GEMM_BLOCK_K,
                          n_blk * BLOCK_N);
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp_idx == 1) {
    // ===================================== MMA issuer =====================================
    if (cta_rank == 0 && elect_one()) {  // 2-CTA: only the leader issues MMAs (for both CTAs)
      const uint32_t idesc = make_idesc_f16(TWO_CTA ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M, WIDE ? 256 : BLOCK_N, BF16, false, false);
      const uint32_t idesc_hi = make_idesc_f16(2 * GEMM_BLOCK_M, 128, BF16, false, false);  // WIDE: columns [256, 384)
      (void)idesc_hi;
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = work_first; tile < total_tiles; tile += work_stride, ++it) {
        const int as = ACC_STAGES == 2 ? (it & 1) : 0;
        const uint32_t aph = ACC_STAGES == 2 ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        GEMM_STAMP(8, it, 0);
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (kb == 0) GEMM_STAMP(8, it, 1);
          const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + s * Cfg::A_BYTES));
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + s * Cfg::B_BYTES));
          const uint64_t b_desc_hi = WIDE ? make_smem_desc_sw128(smem_u32(smem_b + s * Cfg::B_BYTES + 16384)) : 0;
          (void)b_desc_hi;
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in the (addr >> 4) field
            if constexpr (TWO_CTA) umma_f16_ss_2sm(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            if constexpr (WIDE)  // B rows 128.. of each CTA's stage
              umma_f16_ss_2sm(d_tmem + 256, a_desc + 2 * k, b_desc_hi + 2 * k, idesc_hi, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (TWO_CTA) umma_commit_2sm(&empty_bar[s]);  // frees the stage in both CTAs
          else umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if constexpr (TWO_CTA) umma_commit_2sm(&tmem_full[as]);
        else umma_commit(&tmem_full[as]);
        GEMM_STAMP(8, it, 2);
      }
    }
    __syncwarp();
  } else {
    // ===================================== epilogue =====================================
    const int ew = warp_idx - 2;      // 0..7
    const int et = threadIdx.x - 64;  // 0..255
    const int q = warp_idx & 3;       // TMEM lane quarter this warp may read
    const int wg = ew >> 2;           // column half
    constexpr int COLS_PER_WG = BLOCK_N >= 64 ? BLOCK_N / 2 : BLOCK_N;
    const bool wg_active = (BLOCK_N >= 64) || (wg == 0);
    uint8_t* stg = staging + ew * GEMM_STAGE_BYTES_PER_WARP;
    const int TW = 1 << p.tw_log2, TH = GEMM_BLOCK_M >> p.tw_log2;
    constexpr int is_bf16 = BF16 ? 1 : 0;
    const int r = q * 32 + lane;  // accumulator row (TMEM lane) owned by this thread
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    // Per-tile operands of the epilogue - output pixel of my row, bias / folded-LayerNorm column sums of the n-tile (one
    // or two columns per thread), the partial row statistics of my row - are fetched ONE TILE AHEAD: their global-load
    // latency (two dependent L2 round trips per tile otherwise) overlaps the previous tile's units.
    constexpr int NV = (BLOCK_N + GEMM_EPI_WARPS * 32 - 1) / (GEMM_EPI_WARPS * 32);
    struct EpiTile {
      int n_blk, b, n_base;
      int wx, wy;  // output (x, y) of this warp's first row (TMA store box origin)
      long long pix;
      bool row_ok;
      float bias_v[NV], cs_v[NV];
      float4 st4[4];
    };
    auto fetch_tile = [&](int tile, EpiTile& t) {
      const int m_idx = fast_div(tile, p.n_tiles, p.inv_n_tiles);
      t.n_blk = tile - m_idx * p.n_tiles;
      const int mt = m_idx * (TWO_CTA ? 2 : 1) + (int)cta_rank;
      const int mrow = fast_div(mt, p.tiles_x, p.inv_tiles_x);
      const int tx = mt - mrow * p.tiles_x;
      t.b = fast_div(mrow, p.tiles_y, p.inv_tiles_y);
      const int ty = mrow - t.b * p.tiles_y;
      // output pixel of my row
      const int x = tx * TW + (r & (TW - 1));
      const int y = ty * TH + (r >> p.tw_log2);
      t.row_ok = (x < p.W) && (y < p.H) && (t.b < p.B);
      t.wx = tx * TW + ((q * 32) & (TW - 1));
      t.wy = ty * TH + ((q * 32) >> p.tw_log2);
      // pixel-shuffle target: fixed per launch, or chosen by the n-tile (merged ConvTranspose)
      int sh_oy = p.oy, sh_ox = p.ox;
      t.n_base = t.n_blk * BLOCK_N;  // channel of column 0
      if (p.shuffle_n > 0) {
        const int sub_px = (t.n_blk * BLOCK_N) / p.shuffle_n;
        sh_oy = sub_px / p.so;
        sh_ox = sub_px % p.so;
        t.n_base -= sub_px * p.shuffle_n;
      }
      t.pix = ((long long)t.b * p.OH + (long long)y * p.so + sh_oy) * p.OW + (long long)x * p.so + sh_ox;
      // folded LayerNorm: the partial statistics of my row (ln_parts is even: two (sum, sum sq) pairs per 16-byte load)
      const bool has_ln = p.ln_stats != nullptr && t.row_ok;
      const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + t.pix * (2 * p.ln_parts));
      const int n4 = p.ln_parts >> 1;
#pragma unroll
      for (int k = 0; k < 4; ++k) t.st4[k] = (has_ln && k < n4) ? __ldg(sp + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      // bias (and column sums) of the n-tile, zero beyond N / when absent
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int e = et + v * GEMM_EPI_WARPS * 32;
        const int n = t.n_blk * BLOCK_N + e;
        const bool in = e < BLOCK_N && n < p.N && t.b < p.B;
        t.bias_v[v] = (p.bias != nullptr && in) ? __ldg(p.bias + (long long)t.b * p.bias_bstride + (t.n_base + e)) : 0.0f;
        t.cs_v[v] = (!WIDE && p.ln_stats != nullptr && in) ? __ldg(p.ln_colsum + n) : 0.0f;
      }
    };
    EpiTile cur;
    if (work_first < total_tiles) fetch_tile(work_first, cur);
    int it = 0;
    for (int tile = work_first; tile < total_tiles; tile += work_stride, ++it) {
      const int as = ACC_STAGES == 2 ? (it & 1) : 0;
      const uint32_t aph = ACC_STAGES == 2 ? ((it >> 1) & 1) : (it & 1);
      GEMM_STAMP(ew, it, 0);
      const int n_blk = cur.n_blk, b = cur.b, n_base = cur.n_base;
      const long long pix = cur.pix;
      const bool row_ok = cur.row_ok;
      const int wx = cur.wx, wy = cur.wy;
      (void)b; (void)wx; (void)wy;

      // bias / column-sum vectors -> smem. Single-buffered: the first barrier waits until every epilogue warp is done
      // with the previous tile's vectors.
      float* bs = bias_s;
      float* cs = bias_s + BLOCK_N;
      named_bar_sync(1, GEMM_EPI_WARPS * 32);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int e = et + v * GEMM_EPI_WARPS * 32;
        if (e < BLOCK_N) {
          bs[e] = cur.bias_v[v];
          if constexpr (!WIDE) cs[e] = cur.cs_v[v];
        }
      }
      named_bar_sync(1, GEMM_EPI_WARPS * 32);

      float ln_rstd = 1.0f, ln_rm = 0.0f;
      if (p.ln_stats != nullptr && row_ok) {  // partials summed in a fixed order
        float sum = 0.0f, sq = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sum += cur.st4[k].x + cur.st4[k].z;
          sq += cur.st4[k].y + cur.st4[k].w;
        }
        const float4* sp = reinterpret_cast<const float4*>(p.ln_stats + pix * (2 * p.ln_parts));
        const int n4 = p.ln_parts >> 1;
        for (int i = 4; i < n4; ++i) {
          const float4 t = __ldg(sp + i);
          sum += t.x + t.z;
          sq += t.y + t.w;
        }
        const float mean = sum * p.ln_inv_f;
        ln_rstd = rsqrtf(fmaxf(sq * p.ln_inv_f - mean * mean, 0.0f) + p.ln_eps);
        ln_rm = -ln_rstd * mean;
      }
      GEMM_STAMP(ew, it, 1);
      // The next tile's operands are fetched while this tile's units are processed: right after the first unit's math
      // (below), where the ~100 instructions of address arithmetic interleave with it instead of sitting, latency-bound,
      // between the barriers and the accumulator wait. Warps without units of their own fetch here.
      const bool fetch_in_units = wg_active && OUT_KIND != OUT_HEAD;
      if (!fetch_in_units && tile + work_stride < total_tiles) fetch_tile(tile + work_stride, cur);

      // OUT_F32: rows 4*i + (lane >> 3) of this warp are the ones this lane moves in the coalesced phase. The fp32
      // residual of a 32-column unit is fetched one unit ahead with cp.async into a per-warp buffer (unit 0: before the
      // accumulator is even complete), which takes the global-load latency off the critical path of the short-K
      // residual GEMMs (proj) without holding registers. Every lane reads back exactly the 16-byte slots it filled.
      constexpr bool F32O = OUT_KIND == OUT_F32;
      long long res_pix[F32O ? 8 : 1];
      bool res_ok[F32O ? 8 : 1];
      const uint32_t res_slot = smem_u32(resbuf + ew * GEMM_STAGE_BYTES_PER_WARP + lane * 16);  // + i * 512
      auto prefetch_residual = [&](int u) {
        if constexpr (F32O) {
          const int col = n_blk * BLOCK_N + wg * COLS_PER_WG + u * 32 + (lane & 7) * 4;
          if (RESPF && col < p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (res_ok[i])
                cp_async_16(res_slot + i * 512, reinterpret_cast<const float*>(p.add1) + res_pix[i] * p.ld_add1 + col);
          }
          cp_async_commit();
        }
      };
      if constexpr (F32O) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3);
          res_pix[i] = __shfl_sync(0xffffffffu, pix, rr);
          res_ok[i] = __shfl_sync(0xffffffffu, (int)row_ok, rr) != 0;
        }
        if (RESPF && wg_active && p.add1 != nullptr) prefetch_residual(0);
      }

      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      GEMM_STAMP(ew, it, 2);

      if (wg_active) {
        const int col_base = wg * COLS_PER_WG;  // within the tile
        const uint32_t t_acc = tmem_base + lane_addr + as * BLOCK_N + col_base;
        if constexpr (OUT_KIND == OUT_HEAD) {
          uint32_t v[32];
          tmem_ld32(t_acc, v);
          tmem_ld_wait();
          float acc = p.head_b;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc = fmaf(fmaxf(__uint_as_float(v[j]) + bs[j], 0.0f), p.head_w[j], acc);
          acc = p.head_act == ACT_SIGMOID ? rcp_approx(1.0f + __expf(-acc)) : fmaxf(acc, 0.0f);
          if (row_ok) {
            if (is_bf16) reinterpret_cast<__nv_bfloat16*>(p.out)[pix * p.ldo] = __float2bfloat16_rn(acc);
            else reinterpret_cast<__half*>(p.out)[pix * p.ldo] = __float2half_rn(acc);
          }
        } else {
          constexpr bool F32OUT = OUT_KIND == OUT_F32;
          constexpr int COLS_PER_STG = F32OUT ? 32 : 64;  // staging row = 128 B
          constexpr int ELEMS_PER_CHUNK = F32OUT ? 4 : 8;
          constexpr bool SWI = ACT == ACT_SWIGLU;             // two 32-column units -> 32 output columns
          static_assert(!SWI || (!F32OUT && COLS_PER_WG >= 64), "SwiGLU epilogue: 16-bit output, BLOCK_N >= 128");
          constexpr int OUT_COLS_PER_WG = SWI ? COLS_PER_WG / 2 : COLS_PER_WG;
          constexpr int NCOLS_HERE = COLS_PER_STG < OUT_COLS_PER_WG ? COLS_PER_STG : OUT_COLS_PER_WG;
          constexpr int UNITS = COLS_PER_WG / 32;             // 32-column TMEM loads per tile and warp
          constexpr int UNITS_PER_STG = (NCOLS_HERE / 32) * (SWI ? 2 : 1);  // 1 (fp32 out), 2 (16-bit out), 2 or 4 (SwiGLU)
          uint32_t gate[SWI ? 16 : 1];                        // silu(gate unit), packed 16-bit, until its linear unit arrives
          const int sub = lane & 7;  // 16-byte chunk within a 128-byte row segment (phase 2)
          // software pipeline: the TMEM load of unit u+1 is in flight while unit u is processed / stored
          // (OUT_F32 keeps a prefetched residual instead of a prefetched accumulator unit: registers)
          uint32_t vbuf[F32OUT ? 1 : 2][32];
          tmem_ld32(t_acc, vbuf[0]);
          // OUT_F32 + stats_out: partial row statistics of rows 4*i + (lane >> 3) over this warp's column half
          float st_sum[F32OUT ? 8 : 1], st_sq[F32OUT ? 8 : 1];
          if constexpr (F32OUT) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              st_sum[i] = 0.0f;
              st_sq[i] = 0.0f;
            }
          }
#pragma unroll
          for (int u = 0; u < UNITS; ++u) {
            uint32_t(&v)[32] = vbuf[F32OUT ? 0 : (u & 1)];
            if constexpr (F32OUT) {
              if (u > 0) tmem_ld32(t_acc + u * 32, vbuf[0]);
            }
            tmem_ld_wait_dep(v);
            if (u < 4) GEMM_STAMP(ew, it, 3 + 3 * u);
            if constexpr (!F32OUT) {
              if (u + 1 < UNITS) tmem_ld32(t_acc + (u + 1) * 32, vbuf[(u + 1) & 1]);
            }
            const int c0 = (u / UNITS_PER_STG) * COLS_PER_STG;   // first (output) column of this staging chunk
            const int cc = ((u % UNITS_PER_STG) >> (SWI ? 1 : 0)) * 32;  // column offset inside the staging chunk
            const int gc = col_base + u * 32;                    // first GEMM column of this unit within the tile
            // ---- phase 1: my row, 32 columns: +bias, activation -> swizzled staging
            {
              const float4* b4 = reinterpret_cast<const float4*>(bs + gc);
              float f[32];
              if (!WIDE && p.ln_stats != nullptr) {
                const float4* s4 = reinterpret_cast<const float4*>(cs + gc);
                const float2 rs2 = make_float2(ln_rstd, ln_rstd), rm2 = make_float2(ln_rm, ln_rm);
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // packed fp32x2 FMAs: acc * rstd + (rm * colsum + bias)
                  const float4 bb = b4[j];
                  const float4 ss = s4[j];
                  const float2 t0 = __ffma2_rn(rm2, make_float2(ss.x, ss.y), make_float2(bb.x, bb.y));
                  const float2 t1 = __ffma2_rn(rm2, make_float2(ss.z, ss.w), make_float2(bb.z, bb.w));
                  const float2 r0 = __ffma2_rn(make_float2(__uint_as_float(v[4 * j + 0]), __uint_as_float(v[4 * j + 1])), rs2, t0);
                  const float2 r1 = __ffma2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), rs2, t1);
                  f[4 * j + 0] = r0.x;
                  f[4 * j + 1] = r0.y;
                  f[4 * j + 2] = r1.x;
                  f[4 * j + 3] = r1.y;
                }
              } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = b4[j];
                f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x;
                f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
                f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z;
                f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
              }
              }
              if constexpr (ACT == ACT_NONE && !F32OUT) {
                if (p.qk_logit != nullptr) {
                  const int gcol = n_blk * BLOCK_N + gc;  // a multiple of 32 = one head
                  if (gcol < 2 * p.qk_features) {
                    float ss = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) ss = fmaf(f[j], f[j], ss);
                    float sc = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
                    if (gcol < p.qk_features) sc *= __ldg(p.qk_logit + (gcol >> 5));
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] *= sc;
                  }
                }
              }
              if constexpr (ACT == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float2 g = gelu_erf2(make_float2(f[j], f[j + 1]));
                  f[j] = g.x;
                  f[j + 1] = g.y;
                }
              } else if constexpr (ACT == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
              } else if constexpr (SWI) {
                if ((u & 1) == 0) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    gate[j] = pack2(f[2 * j] * rcp_approx(1.0f + __expf(-f[2 * j])),
                                    f[2 * j + 1] * rcp_approx(1.0f + __expf(-f[2 * j + 1])), is_bf16);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const float2 g = unpack2(gate[j], is_bf16);
                    f[2 * j] *= g.x;
                    f[2 * j + 1] *= g.y;
                  }
                }
              }
              if constexpr (SWI) {
                if ((u & 1) == 0) continue;  // the gate half: nothing to store yet
              }
              if constexpr (F32OUT) {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                  const int phys = ch ^ (lane & 7);
                  *reinterpret_cast<float4*>(stg + lane * 128 + phys * 16) =
                      make_float4(f[4 * ch], f[4 * ch + 1], f[4 * ch + 2], f[4 * ch + 3]);
                }
              } else {
                if (p.tma_store && (u % UNITS_PER_STG) == (SWI ? 1 : 0)) {
                  // first write into the staging chunk: the TMA store of the previous chunk must have read it
                  if (lane == 0) bulk_wait_read_all();
                  __syncwarp();
                }
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) {
                  const int phys = ((cc >> 3) + ch) ^ (lane & 7);
                  uint4 o;
                  o.x = pack2(f[8 * ch + 0], f[8 * ch + 1], is_bf16);
                  o.y = pack2(f[8 * ch + 2], f[8 * ch + 3], is_bf16);
                  o.z = pack2(f[8 * ch + 4], f[8 * ch + 5], is_bf16);
                  o.w = pack2(f[8 * ch + 6], f[8 * ch + 7], is_bf16);
                  *reinterpret_cast<uint4*>(stg + lane * 128 + phys * 16) = o;
                }
              }
            }
            if (u == (SWI ? 1 : 0) && tile + work_stride < total_tiles) fetch_tile(tile + work_stride, cur);  // see above
            if (u < 4) GEMM_STAMP(ew, it, 4 + 3 * u);
            if ((u % UNITS_PER_STG) != UNITS_PER_STG - 1) continue;  // staging chunk not complete yet
            __syncwarp();
            // ---- phase 2: coalesced global IO; 8 lanes cover one 128-byte row segment, 4 rows per pass, all
            //      eight passes' loads issued before the first store (the residual may alias the output)
            // first output column of this staging chunk (SwiGLU: half the GEMM column)
            const int ncol0 = SWI ? ((n_blk * BLOCK_N + col_base) >> 1) + c0 : n_blk * BLOCK_N + col_base + c0;
            const bool col_ok = (sub * ELEMS_PER_CHUNK < NCOLS_HERE) && (ncol0 + sub * ELEMS_PER_CHUNK) < (SWI ? p.N >> 1 : p.N);
            const long long coff = SWI ? (long long)ncol0 + sub * ELEMS_PER_CHUNK
                                       : (long long)n_base + col_base + c0 + sub * ELEMS_PER_CHUNK;  // output channel
            if constexpr (F32OUT) {
              const long long(&rpix)[8] = res_pix;
              bool ok[8];
              uint4 val[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + (lane >> 3);
                ok[i] = res_ok[i] && col_ok;
                val[i] = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((sub ^ (rr & 7)) * 16));
              }
              if (p.add1 != nullptr) {
                if constexpr (RESPF) {
                  cp_async_wait_all();  // this lane's own copies of this unit
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    if (ok[i]) {
                      const float4 a = lds_f4(res_slot + i * 512);
                      float4 o = *reinterpret_cast<float4*>(&val[i]);
                      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
                      val[i] = *reinterpret_cast<uint4*>(&o);
                    }
                  }
                  if (u + 1 < UNITS) prefetch_residual(u + 1);  // the slots were just read by this same lane
                } else {
                  float4 a[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    a[i] = ok[i] ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.add1) +
                                                                    rpix[i] * p.ld_add1 + coff)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    float4 o = *reinterpret_cast<float4*>(&val[i]);
                    o.x += a[i].x; o.y += a[i].y; o.z += a[i].z; o.w += a[i].w;
                    val[i] = *reinterpret_cast<uint4*>(&o);
                  }
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (ok[i]) *reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.out) + rpix[i] * p.ldo + coff) = val[i];
              if (p.stats_out != nullptr) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 o = *reinterpret_cast<float4*>(&val[i]);
                  if (ok[i]) {
                    const float2 a = make_float2(o.x, o.y), c = make_float2(o.z, o.w);
                    const float2 ps = __fadd2_rn(a, c);
                    const float2 pq = __ffma2_rn(a, a, __fmul2_rn(c, c));
                    st_sum[i] += ps.x + ps.y;
                    st_sq[i] += pq.x + pq.y;
                    if (p.out16 != nullptr) {
                      uint2 h;
                      h.x = pack2(o.x, o.y, is_bf16);
                      h.y = pack2(o.z, o.w, is_bf16);
                      *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out16) + rpix[i] * p.ld_out16 + coff) = h;
                    }
                  }
                }
              }
            } else if (NCOLS_HERE == 64 && p.tma_store) {
              fence_proxy_async_smem();  // the staging writes (generic proxy) -> visible to the TMA (async proxy)
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&p.tmOut, stg, ncol0, wx, wy, b);
                bulk_commit();
              }
            } else {
              const bool has_extra = p.add1 != nullptr || p.add2 != nullptr || p.out2_relu != nullptr;
#pragma unroll
              for (int hb = 0; hb < 2; ++hb) {  // two batches of four passes (register pressure)
                long long rpix[4];
                bool ok[4];
                uint4 val[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int rr = 16 * hb + 4 * i + (lane >> 3);
#if defined(GEMM_ABL) && GEMM_ABL == 2  // timing ablation only (wrong rows): no shuffles
                  rpix[i] = pix + rr;
                  ok[i] = row_ok && col_ok;
#else
                  rpix[i] = __shfl_sync(0xffffffffu, pix, rr);
                  ok[i] = __shfl_sync(0xffffffffu, (int)row_ok, rr) != 0 && col_ok;
#endif
                  val[i] = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((sub ^ (rr & 7)) * 16));
                }
                if (has_extra) {
                  const uint4 z = make_uint4(0, 0, 0, 0);
                  uint4 a1[4], a2[4];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    a1[i] = (p.add1 != nullptr && ok[i])
                                ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.add1) +
                                                                  rpix[i] * p.ld_add1 + coff)
                                : z;
                    a2[i] = (p.add2 != nullptr && ok[i])
                                ? *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.add2) +
                                                                  rpix[i] * p.ld_add2 + coff)
                                : z;
                  }
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    uint32_t* vv = reinterpret_cast<uint32_t*>(&val[i]);
                    const uint32_t* x1 = reinterpret_cast<const uint32_t*>(&a1[i]);
                    const uint32_t* x2 = reinterpret_cast<const uint32_t*>(&a2[i]);
                    uint4 rv;
                    uint32_t* rr32 = reinterpret_cast<uint32_t*>(&rv);
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                      float2 fv = unpack2(vv[w], is_bf16);
                      const float2 g1 = unpack2(x1[w], is_bf16), g2 = unpack2(x2[w], is_bf16);
                      fv.x += g1.x + g2.x;
                      fv.y += g1.y + g2.y;
                      vv[w] = pack2(fv.x, fv.y, is_bf16);
                      rr32[w] = pack2(fmaxf(fv.x, 0.f), fmaxf(fv.y, 0.f), is_bf16);
                    }
                    if (p.out2_relu != nullptr && ok[i])
                      *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out2_relu) + rpix[i] * p.ld_out2 + coff) = rv;
                  }
                }
#if defined(GEMM_ABL) && GEMM_ABL == 1  // timing ablation only: the stores happen for impossible values
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (ok[i] && val[i].x == 0x7fc07fc1u) *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + rpix[i] * p.ldo + coff) = val[i];
#else
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (ok[i]) *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + rpix[i] * p.ldo + coff) = val[i];
#endif
              }
            }
            __syncwarp();
            if (u < 4) GEMM_STAMP(ew, it, 5 + 3 * u);
          }
          if constexpr (F32OUT) {
            if (p.stats_out != nullptr) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                  st_sum[i] += __shfl_xor_sync(0xffffffffu, st_sum[i], o);
                  st_sq[i] += __shfl_xor_sync(0xffffffffu, st_sq[i], o);
                }
                if (sub == 0 && res_ok[i])
                  reinterpret_cast<float2*>(p.stats_out)[res_pix[i] * p.stats_parts + n_blk * 2 + wg] =
                      make_float2(st_sum[i], st_sq[i]);
              }
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld above)
      GEMM_STAMP(ew, it, 15);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (TWO_CTA) mbar_arrive_leader(&tmem_empty[as]);
        else mbar_arrive(&tmem_empty[as]);
      }
    }
  }

  if (p.tma_store && warp_idx >= 2 && lane == 0) bulk_wait_all();  // my TMA stores have left shared memory and landed
  tc_fence_before();
  if constexpr (TWO_CTA) cluster_sync_all();  // the peer's smem / TMEM stay valid until both CTAs are done
  else __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    if constexpr (TWO_CTA) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace dpt
