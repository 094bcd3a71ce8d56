// sync_lab - latencies of the synchronisation hops an attention kv step is made of (tools only).
//   1. mbarrier ping-pong between two warps: try_wait with suspend hint / try_wait without hint / test_wait spin
//   2. tcgen05.commit -> mbarrier -> waiting warp (no MMA pending, and behind one M128 N64 K16 MMA, and behind 4+4 MMAs)
//   3. tcgen05.ld x32 + wait::ld, tcgen05.st x32 + wait::st round trips
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I muggled_dpt_b200/csrc -I include \
//             -o tools/microbench/sync_lab tools/microbench/sync_lab.cu
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"

using namespace dpt;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ void wait_mode(uint64_t* bar, uint32_t parity) {
  if constexpr (MODE == 0) {
    mbar_wait(bar, parity);
  } else if constexpr (MODE == 1) {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  } else {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
  }
}

// warp 0 lane 0 and warp 1 lane 0 (different sub-partitions) bounce two barriers; other lanes idle at the end barrier
template <int MODE>
__global__ void pingpong_kernel(long long* clk, int iters) {
  __shared__ uint64_t bars[2];
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = 0, t1 = 0;
  if (lane == 0 && warp == 0) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mbar_arrive(&bars[0]);
      wait_mode<MODE>(&bars[1], i & 1);
    }
    t1 = clock64();
  } else if (lane == 0 && warp == 1) {
    for (int i = 0; i < iters; ++i) {
      wait_mode<MODE>(&bars[0], i & 1);
      mbar_arrive(&bars[1]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// 128 threads arrive (like p_ready), one thread waits then arrives on a count-1 barrier all 128 wait on (like s_full)
template <int MODE>
__global__ void fan_kernel(long long* clk, int iters) {
  __shared__ uint64_t bars[2];
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 128);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  if (warp < 4) {
    for (int i = 0; i < iters; ++i) {
      mbar_arrive(&bars[0]);
      wait_mode<MODE>(&bars[1], i & 1);
    }
  } else if (warp == 5 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      wait_mode<MODE>(&bars[0], i & 1);
      mbar_arrive(&bars[1]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

// MMA-side hop: NMMA x (M128 N64 K16 SS) then commit -> barrier; the same thread waits for it
template <int MODE>
__global__ void commit_kernel(long long* clk, int iters, int nmma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 128);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_f16(128, 64, true, false, false);
      const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem));
      const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem + 16384));
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        for (int k = 0; k < nmma; ++k) umma_f16_ss(tbase, a_desc + 2 * (k & 3), b_desc + 2 * (k & 3), idesc, k != 0);
        umma_commit(&bar);
        wait_mode<MODE>(&bar, i & 1);
      }
      t1 = clock64();
      clk[0] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 128);
  }
}

// TMEM round trips: NW warps per sub-partition each loop { ld x32 x2, wait, st x32, wait }
__global__ void tmem_rt_kernel(long long* clk, int iters, float* out) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  const uint32_t addr = tbase + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t v[2][32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[0][i] = v[1][i] = i;
  tmem_st32(addr, v[0]);
  tmem_st32(addr + 32, v[1]);
  tmem_st_wait();
  __syncthreads();
  long long t_ld = 0, t_st = 0;
  for (int it = 0; it < iters; ++it) {
    const long long a = clock64();
    tmem_ld32(addr, v[0]);
    tmem_ld32(addr + 32, v[1]);
    tmem_ld_wait_dep(v[0]);
    tmem_ld_wait_dep(v[1]);
    const long long b = clock64();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[0][i] += v[1][i];
    tmem_st32(addr, v[0]);
    tmem_st_wait();
    tc_fence_before();
    const long long c = clock64();
    t_ld += b - a;
    t_st += c - b;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(v[0][3]);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    clk[0] = t_ld;
    clk[1] = t_st;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

int main() {
  long long* d_clk;
  float* d_out;
  CK(cudaMalloc(&d_clk, 64));
  CK(cudaMalloc(&d_out, 148 * 512 * 4));
  const int iters = 2000;
  long long c[2];
  auto rd = [&]() { CK(cudaDeviceSynchronize()); CK(cudaMemcpy(c, d_clk, 16, cudaMemcpyDeviceToHost)); };
  pingpong_kernel<0><<<148, 128>>>(d_clk, iters); rd(); printf("pingpong try_wait+hint : %.0f clk per round trip (2 hops)\n", (double)c[0] / iters);
  pingpong_kernel<1><<<148, 128>>>(d_clk, iters); rd(); printf("pingpong try_wait      : %.0f clk per round trip\n", (double)c[0] / iters);
  pingpong_kernel<2><<<148, 128>>>(d_clk, iters); rd(); printf("pingpong test_wait spin: %.0f clk per round trip\n", (double)c[0] / iters);
  fan_kernel<0><<<148, 256>>>(d_clk, iters); rd(); printf("fan 128->1->128 try_wait+hint : %.0f clk per round trip\n", (double)c[0] / iters);
  fan_kernel<1><<<148, 256>>>(d_clk, iters); rd(); printf("fan 128->1->128 try_wait      : %.0f clk per round trip\n", (double)c[0] / iters);
  fan_kernel<2><<<148, 256>>>(d_clk, iters); rd(); printf("fan 128->1->128 test_wait spin: %.0f clk per round trip\n", (double)c[0] / iters);
  CK(cudaFuncSetAttribute(commit_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792));
  CK(cudaFuncSetAttribute(commit_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 33792));
  for (int nmma : {0, 1, 4, 8}) {
    commit_kernel<0><<<148, 128, 33792>>>(d_clk, iters, nmma); rd();
    printf("commit after %d MMA(M128 N64 K16) try_wait+hint: %.0f clk per issue->commit->wake\n", nmma, (double)c[0] / iters);
    commit_kernel<2><<<148, 128, 33792>>>(d_clk, iters, nmma); rd();
    printf("commit after %d MMA(M128 N64 K16) spin         : %.0f clk\n", nmma, (double)c[0] / iters);
  }
  for (int threads : {128, 256, 512}) {
    tmem_rt_kernel<<<148, threads>>>(d_clk, iters, d_out); rd();
    printf("TMEM %d warps/SMSP: ld 2 x x32 + wait %.0f clk, st x32 + wait + fence %.0f clk\n", threads / 128, (double)c[0] / iters, (double)c[1] / iters);
  }
  return 0;
}
