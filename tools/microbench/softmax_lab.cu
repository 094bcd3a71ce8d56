// softmax_lab - pipe-rate and softmax-sweep microbenchmarks for attn_tc.cuh (tools only, never linked into the library).
//
// Part 1: reciprocal throughput per SM sub-partition (clk per warp instruction) of the instructions the softmax of
//         attn_tc_kernel is made of: FFMA, FFMA2, FADD2, FMNMX, FMNMX3, MUFU.EX2, F2FP (bf16x2 / f16x2), IADD / SHL.
// Part 2: one "kv step" of the softmax warp (TMEM S row -> registers, row max, P = exp2(S*c - mu), row sum, 16-bit P ->
//         TMEM) in several instruction arrangements, with 1 / 2 / 4 warps per sub-partition, no MMA running: the floor
//         the softmax side can reach on its own.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I muggled_dpt_b200/csrc -I include \
//             -o tools/microbench/softmax_lab tools/microbench/softmax_lab.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

using namespace dpt;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------ part 1: pipe rates
enum { OP_FFMA, OP_FFMA2, OP_FADD2, OP_FMNMX, OP_FMNMX3, OP_EX2, OP_CVT_BF16, OP_CVT_F16, OP_IADD, OP_SHL, OP_FMUL2,
       OP_EX2_FFMA2, OP_EX2_FFMA2_MNMX, OP_FFMA2_MNMX3, OP_FFMA2_F2FP, OP_FFMA2_FADD2, OP_FFMA_FMNMX, OP_EX2_F16X2, OP_EX2_BF16X2, OP_EX2_F16, OP_COUNT };
static const char* kOpNames[OP_COUNT] = {"FFMA", "FFMA2", "FADD2", "FMNMX", "FMNMX3", "MUFU.EX2", "F2FP.bf16x2", "F2FP.f16x2",
                                         "IADD", "SHL", "FMUL2", "mix: 2 EX2 + 1 FFMA2", "mix: 2 EX2 + FFMA2 + FMNMX3", "mix: FFMA2 + FMNMX3", "mix: FFMA2 + F2FP",
                                         "mix: FFMA2 + FADD2", "mix: FFMA + FMNMX", "MUFU.EX2.f16x2", "MUFU.EX2.bf16x2", "MUFU.EX2.f16"};

template <int OP>
__global__ void rate_kernel(float* out, long long* clk, int iters, float seed) {
  constexpr int CH = 8;
  float a[CH], b[CH];
  float2 a2[CH];
  uint32_t u[CH];
  float m[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    m[i] = -1e30f;
    a[i] = seed + i * 0.001f + threadIdx.x * 1e-4f;
    b[i] = -a[i];
    a2[i] = make_float2(a[i], b[i]);
    u[i] = __float_as_uint(a[i]);
  }
  const float c = seed * 0.5f;
  const float2 c2 = make_float2(c, c), d2 = make_float2(0.001f, -0.001f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        if constexpr (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(b[i]));
        if constexpr (OP == OP_FFMA2) a2[i] = __ffma2_rn(a2[i], c2, d2);
        if constexpr (OP == OP_FMUL2) a2[i] = __fmul2_rn(a2[i], c2);
        if constexpr (OP == OP_FADD2) a2[i] = __fadd2_rn(a2[i], d2);
        if constexpr (OP == OP_FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        if constexpr (OP == OP_FMNMX3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c));
        if constexpr (OP == OP_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        if constexpr (OP == OP_CVT_BF16) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
        if constexpr (OP == OP_CVT_F16) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
        if constexpr (OP == OP_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
        if constexpr (OP == OP_SHL) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
        if constexpr (OP == OP_EX2_F16X2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
        if constexpr (OP == OP_EX2_BF16X2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
        if constexpr (OP == OP_EX2_F16) { unsigned short hs = (unsigned short)u[i]; asm volatile("ex2.approx.f16 %0, %0;" : "+h"(hs)); u[i] = hs; }
        if constexpr (OP == OP_FFMA2_MNMX3) {
          a2[i] = __ffma2_rn(a2[i], c2, d2);
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(b[i]), "f"(c));
        }
        if constexpr (OP == OP_FFMA2_F2FP) {
          a2[i] = __ffma2_rn(a2[i], c2, d2);
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i])));
        }
        if constexpr (OP == OP_FFMA2_FADD2) {
          a2[i] = __ffma2_rn(a2[i], c2, d2);
          float2 t = make_float2(a[i], b[i]);
          t = __fadd2_rn(t, d2);
          a[i] = t.x; b[i] = t.y;
        }
        if constexpr (OP == OP_FFMA_FMNMX) {
          asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(b[i]));
          asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(b[i]));
        }
        if constexpr (OP == OP_EX2_FFMA2 || OP == OP_EX2_FFMA2_MNMX) {
          a2[i] = __ffma2_rn(a2[i], c2, d2);
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
          if constexpr (OP == OP_EX2_FFMA2_MNMX) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(b[i]), "f"(c));
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += a[i] + b[i] + a2[i].x + a2[i].y + __uint_as_float(u[i]) + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

template <int OP>
void run_rate(float* d_out, long long* d_clk, int warps_per_smsp) {
  const int iters = 2000;
  rate_kernel<OP><<<148, 128 * warps_per_smsp>>>(d_out, d_clk, iters, 0.37f);
  CK(cudaDeviceSynchronize());
  long long c;
  CK(cudaMemcpy(&c, d_clk, 8, cudaMemcpyDeviceToHost));
  const double groups = (double)iters * 8 * 8;
  printf("  %-30s warps/SMSP=%d  %.2f clk per op-group per warp, %.2f clk per op-group per SMSP slot\n", kOpNames[OP],
         warps_per_smsp, c / groups, c / groups / warps_per_smsp);
}

// ------------------------------------------------------------------------------------------------ part 2: softmax step
// VARIANT 0: the shipped arrangement (max sweep with 2-input max; scale/shift sweep; ex2 sweep; sum + pack sweep)
// VARIANT 1: max sweep with 3-input max, then ONE fused sweep (ffma2 -> ex2 -> fadd2 + pack per pair, program order)
// VARIANT 2: like 1 but the fused sweep is software-pipelined by hand: the ex2 of pair-group g is issued between the
//            ffma2 of group g+1 and the sum / pack of group g-1 (groups of 8 pairs)
// VARIANT 3: no separate max sweep: the max of the raw scores accumulates inside the fused sweep (stabiliser from the
//            previous step; the rare "max outgrew mu" case would be handled after the fact)
template <int VARIANT, int POLY, bool BF16>
__global__ void __launch_bounds__(256, 1) sweep_kernel(float* out, long long* clk, int iters) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  const int q = warp & 3, slot = warp >> 2;  // TMEM lane quarter; which co-resident "CTA" this warp plays
  const uint32_t lane_addr = uint32_t(q * 32) << 16;
  // per slot: S at cols [slot*128, +128) (slots 0..1) ... with 4 slots the regions alias pairwise, timing only
  const uint32_t s_addr = tbase + lane_addr + (slot & 1) * 128 + (slot >> 1) * 256;
  const uint32_t p_addr = s_addr;  // P overwrites the first 64 columns of S (timing only)
  // fill S with plausible scores
  {
    uint32_t v[32];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-3.0f + 0.01f * ((lane * 7 + i * 13 + ci * 5) % 97));
      tmem_st32(s_addr + ci * 32, v);
    }
    tmem_st_wait();
  }
  __syncthreads();
  const float c = 0.18f;
  const float2 c2 = make_float2(c, c);
  float mu = 0.f;
  float2 l2[4] = {make_float2(0, 0), make_float2(0, 0), make_float2(0, 0), make_float2(0, 0)};
  float macc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t sv[4][32];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) tmem_ld32(s_addr + ci * 32, sv[ci]);
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) tmem_ld_wait_dep(sv[ci]);
    uint32_t pk[2][32];
    if constexpr (VARIANT == 0) {
      float m_t[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int i = 0; i < 32; i += 8)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            m_t[k] = fmaxf(m_t[k], fmaxf(__uint_as_float(sv[ci][i + 2 * k]), __uint_as_float(sv[ci][i + 2 * k + 1])));
      float m_tile = fmaxf(fmaxf(m_t[0], m_t[1]), fmaxf(m_t[2], m_t[3])) * c;
      if (m_tile > mu + 8.0f) mu = m_tile;
      const float2 neg_mu2 = make_float2(-mu, -mu);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 e = __ffma2_rn(make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1])), c2, neg_mu2);
          sv[ci][i] = __float_as_uint(e.x);
          sv[ci][i + 1] = __float_as_uint(e.y);
        }
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if (((i >> 1) & 7) >= 8 - POLY) {
            const float2 e = ex2_poly2<BF16 ? 3 : 4>(make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1])));
            sv[ci][i] = __float_as_uint(e.x);
            sv[ci][i + 1] = __float_as_uint(e.y);
          } else {
            sv[ci][i] = __float_as_uint(ex2_approx(__uint_as_float(sv[ci][i])));
            sv[ci][i + 1] = __float_as_uint(ex2_approx(__uint_as_float(sv[ci][i + 1])));
          }
        }
#pragma unroll
      for (int ci = 0; ci < 4; ++ci)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 pf = make_float2(__uint_as_float(sv[ci][2 * i]), __uint_as_float(sv[ci][2 * i + 1]));
          l2[i & 3] = __fadd2_rn(l2[i & 3], pf);
          pk[ci >> 1][(ci & 1) * 16 + i] = BF16 ? pack_bf16(pf.x, pf.y) : pack_f16(pf.x, pf.y);
        }
    } else {
      float m_tile = -INFINITY;
      if constexpr (VARIANT != 3) {
        float m_t[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
#pragma unroll
          for (int i = 0; i < 32; i += 8)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              m_t[k] = max3(m_t[k], __uint_as_float(sv[ci][i + 2 * k]), __uint_as_float(sv[ci][i + 2 * k + 1]));
        m_tile = fmaxf(fmaxf(m_t[0], m_t[1]), fmaxf(m_t[2], m_t[3])) * c;
        if (m_tile > mu + 8.0f) mu = m_tile;
      }
      const float2 neg_mu2 = make_float2(-mu, -mu);
      float m_in[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      auto scale = [&](int ci, int i) {
        const float2 e = __ffma2_rn(make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1])), c2, neg_mu2);
        if constexpr (VARIANT == 3) m_in[(i >> 1) & 3] = max3(m_in[(i >> 1) & 3], __uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1]));
        sv[ci][i] = __float_as_uint(e.x);
        sv[ci][i + 1] = __float_as_uint(e.y);
      };
      auto expo = [&](int ci, int i) {
        if (((i >> 1) & 7) >= 8 - POLY) {
          const float2 e = ex2_poly2<BF16 ? 3 : 4>(make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1])));
          sv[ci][i] = __float_as_uint(e.x);
          sv[ci][i + 1] = __float_as_uint(e.y);
        } else {
          sv[ci][i] = __float_as_uint(ex2_approx(__uint_as_float(sv[ci][i])));
          sv[ci][i + 1] = __float_as_uint(ex2_approx(__uint_as_float(sv[ci][i + 1])));
        }
      };
      auto sumpack = [&](int ci, int i) {
        const float2 pf = make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1]));
        l2[(i >> 1) & 3] = __fadd2_rn(l2[(i >> 1) & 3], pf);
        pk[ci >> 1][(ci & 1) * 16 + (i >> 1)] = BF16 ? pack_bf16(pf.x, pf.y) : pack_f16(pf.x, pf.y);
      };
      if constexpr (VARIANT == 1 || VARIANT == 3) {
#pragma unroll
        for (int ci = 0; ci < 4; ++ci)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            scale(ci, i);
            expo(ci, i);
            sumpack(ci, i);
          }
      } else {
        // groups of 8 pairs (16 columns), 16 groups over the row: scale(g) | ex2(g-1) | sum+pack(g-2)
        constexpr int NG = 16;
#pragma unroll
        for (int g = 0; g < NG + 2; ++g) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (g < NG) scale(g >> 1, (g & 1) * 16 + 2 * k);
            if (g >= 1 && g <= NG) expo((g - 1) >> 1, ((g - 1) & 1) * 16 + 2 * k);
            if (g >= 2) sumpack((g - 2) >> 1, ((g - 2) & 1) * 16 + 2 * k);
          }
        }
      }
      if constexpr (VARIANT == 3) {
        m_tile = fmaxf(fmaxf(m_in[0], m_in[1]), fmaxf(m_in[2], m_in[3])) * c;
        if (m_tile > mu + 8.0f) mu = m_tile;
      }
      macc += m_tile;
    }
    tmem_st32(p_addr, pk[0]);
    tmem_st32(p_addr + 32, pk[1]);
    tmem_st_wait();
  }
  const long long t1 = clock64();
  const float2 ls = __fadd2_rn(__fadd2_rn(l2[0], l2[1]), __fadd2_rn(l2[2], l2[3]));
  out[blockIdx.x * blockDim.x + threadIdx.x] = ls.x + ls.y + mu + macc;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

template <int VARIANT, int POLY, bool BF16>
void run_sweep(float* d_out, long long* d_clk) {
  const int iters = 400;
  for (int wps : {1, 2}) {
    sweep_kernel<VARIANT, POLY, BF16><<<148, 128 * wps>>>(d_out, d_clk, iters);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_clk, 8, cudaMemcpyDeviceToHost));
    printf("  variant %d poly %d/8 %s warps/SMSP=%d: %.0f clk per step per warp -> %.0f clk per 128x128 tile per SM\n", VARIANT,
           POLY, BF16 ? "bf16" : "f16", wps, (double)c / iters, (double)c / iters / wps);
  }
}

// ------------------------------------------------------------------------------------------------ part 3
// The 64-column step of attn64_tc.cuh (ld 2 x x32, max3, fused sweep, st x32) with 1..4 softmax warps per
// sub-partition, optionally with one extra warp issuing tcgen05 MMAs back to back (M128 N64 K16 SS + TS like the
// kernel's S and P@V) into other TMEM columns: does tensor-pipe activity slow the softmax warps down?
template <int POLY, bool WITH_MMA>
__global__ void __launch_bounds__(544, 1) sweep64_kernel(float* out, long long* clk, int iters, int nwarps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  __shared__ volatile int stop_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    stop_flag = 0;
  }
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  long long t0 = 0, t1 = 0;
  float res = 0.f;
  if (warp < nwarps) {
    const int q = warp & 3, slot = warp >> 2;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t s_addr = tbase + lane_addr + slot * 96;  // S 64 cols + P 32 cols per slot
    const uint32_t p_addr = s_addr + 64;
    {
      uint32_t v[32];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-3.0f + 0.01f * ((lane * 7 + i * 13 + ci * 5) % 97));
        tmem_st32(s_addr + ci * 32, v);
      }
      tmem_st_wait();
    }
    const float c = 0.18f;
    const float2 c2 = make_float2(c, c);
    float mu = 0.f;
    float2 l2[2] = {make_float2(0, 0), make_float2(0, 0)};
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t sv[2][32];
      tmem_ld32(s_addr, sv[0]);
      tmem_ld32(s_addr + 32, sv[1]);
      tmem_ld_wait_dep(sv[0]);
      tmem_ld_wait_dep(sv[1]);
      float m_t[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int i = 0; i < 32; i += 8)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            m_t[k] = max3(m_t[k], __uint_as_float(sv[ci][i + 2 * k]), __uint_as_float(sv[ci][i + 2 * k + 1]));
      const float m_tile = fmaxf(fmaxf(m_t[0], m_t[1]), fmaxf(m_t[2], m_t[3])) * c;
      if (m_tile > mu + 8.0f) mu = m_tile;
      const float2 neg_mu2 = make_float2(-mu, -mu);
      uint32_t pk[32];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 e = __ffma2_rn(make_float2(__uint_as_float(sv[ci][i]), __uint_as_float(sv[ci][i + 1])), c2, neg_mu2);
          float2 pf;
          if (((i >> 1) & 7) >= 8 - POLY) {
            pf = ex2_poly2<3>(e);
          } else {
            pf.x = ex2_approx(e.x);
            pf.y = ex2_approx(e.y);
          }
          l2[(i >> 1) & 1] = __fadd2_rn(l2[(i >> 1) & 1], pf);
          pk[ci * 16 + (i >> 1)] = pack_bf16(pf.x, pf.y);
        }
      tmem_st32(p_addr, pk);
      tmem_st_wait();
    }
    t1 = clock64();
    res = l2[0].x + l2[0].y + l2[1].x + l2[1].y + mu;
    if (threadIdx.x == 0) stop_flag = 1;
  } else if (WITH_MMA && warp == 16) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 64, true, false, false);
      const uint32_t idesc_o = make_idesc_f16(128, 64, true, false, true);
      const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem));
      const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + 16384));
      const uint32_t d0 = tbase + 384, d1 = tbase + 448, pa = tbase + 64;  // accumulators away from the softmax slots
      uint32_t ph = 0;
      while (!stop_flag) {
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ss(d0, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(d1, pa + 8 * k, b_desc + 128 * k, idesc_o, k != 0);
        }
        umma_commit(&bar);
        mbar_wait(&bar, ph);
        ph ^= 1;
      }
    }
    __syncwarp();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = res;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

template <int POLY, bool WITH_MMA>
void run_sweep64(float* d_out, long long* d_clk) {
  const int iters = 400;
  CK(cudaFuncSetAttribute(sweep64_kernel<POLY, WITH_MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792));
  for (int wps : {1, 2, 3, 4}) {
    sweep64_kernel<POLY, WITH_MMA><<<148, 544, 33792>>>(d_out, d_clk, iters, 4 * wps);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_clk, 8, cudaMemcpyDeviceToHost));
    printf("  64-col step poly %d/8 %s warps/SMSP=%d: %.0f clk per step per warp -> %.0f clk per 128x128 tile per SM\n", POLY,
           WITH_MMA ? "WITH concurrent MMAs" : "no MMA", wps, (double)c / iters, 2.0 * c / iters / wps);
  }
}

int main(int argc, char** argv) {
  const bool all = argc > 1;
  float* d_out;
  long long* d_clk;
  CK(cudaMalloc(&d_out, 148 * 512 * 4));
  CK(cudaMalloc(&d_clk, 64));
  printf("== part 1: pipe rates (8 independent chains per warp; 'op-group' = one op, or the listed mix)\n");
  for (int w : {2}) {
    run_rate<OP_FFMA>(d_out, d_clk, w);
    run_rate<OP_FFMA2>(d_out, d_clk, w);
    run_rate<OP_FMUL2>(d_out, d_clk, w);
    run_rate<OP_FADD2>(d_out, d_clk, w);
    run_rate<OP_FMNMX>(d_out, d_clk, w);
    run_rate<OP_FMNMX3>(d_out, d_clk, w);
    run_rate<OP_EX2>(d_out, d_clk, w);
    run_rate<OP_CVT_BF16>(d_out, d_clk, w);
    run_rate<OP_CVT_F16>(d_out, d_clk, w);
    run_rate<OP_IADD>(d_out, d_clk, w);
    run_rate<OP_SHL>(d_out, d_clk, w);
    run_rate<OP_EX2_FFMA2>(d_out, d_clk, w);
    run_rate<OP_EX2_FFMA2_MNMX>(d_out, d_clk, w);
    run_rate<OP_FFMA2_MNMX3>(d_out, d_clk, w);
    run_rate<OP_FFMA2_F2FP>(d_out, d_clk, w);
    run_rate<OP_FFMA2_FADD2>(d_out, d_clk, w);
    run_rate<OP_FFMA_FMNMX>(d_out, d_clk, w);
    run_rate<OP_EX2_F16X2>(d_out, d_clk, w);
    run_rate<OP_EX2_BF16X2>(d_out, d_clk, w);
    run_rate<OP_EX2_F16>(d_out, d_clk, w);
  }
  if (!all) return 0;
  printf("== part 2: softmax kv step (128 columns per thread), no MMA running\n");
  run_sweep<0, 2, true>(d_out, d_clk);
  run_sweep<0, 0, true>(d_out, d_clk);
  run_sweep<1, 0, true>(d_out, d_clk);
  run_sweep<1, 2, true>(d_out, d_clk);
  run_sweep<1, 3, true>(d_out, d_clk);
  run_sweep<1, 4, true>(d_out, d_clk);
  run_sweep<2, 2, true>(d_out, d_clk);
  run_sweep<2, 3, true>(d_out, d_clk);
  run_sweep<3, 0, true>(d_out, d_clk);
  run_sweep<3, 2, true>(d_out, d_clk);
  run_sweep<3, 3, true>(d_out, d_clk);
  run_sweep<1, 2, false>(d_out, d_clk);
  printf("== part 3: the 64-column step of attn64_tc.cuh, 1..4 warps per sub-partition, with / without concurrent MMAs\n");
  run_sweep64<2, false>(d_out, d_clk);
  run_sweep64<2, true>(d_out, d_clk);
  run_sweep64<0, false>(d_out, d_clk);
  run_sweep64<0, true>(d_out, d_clk);
  return 0;
}
