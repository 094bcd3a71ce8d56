// tcgen05.ld throughput per SM: how many bytes/clk can 1, 2, 4 (one per sub-partition) or 8 warps (two CTAs) read?
// build+run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld_bench tmem_ld_bench.cu && /tmp/tmem_ld_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(uint32_t* out, long long* cyc, int iters, int active_warps, int cols_log) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tptr + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(base + c * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= v[0] ^ v[31];
      }
    }
    t1 = clock64();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tptr));
}

int main() {
  uint32_t* out; long long* cyc;
  const int iters = 2000;
  cudaMalloc(&out, 296 * 128 * 4); cudaMalloc(&cyc, 296 * 8);
  for (int blocks_per_sm = 1; blocks_per_sm <= 2; ++blocks_per_sm)
    for (int aw = 1; aw <= 4; aw *= 2) {
      const int blocks = 148 * blocks_per_sm;
      k<<<blocks, 128>>>(out, cyc, 10, aw, 0);
      k<<<blocks, 128>>>(out, cyc, iters, aw, 0);
      cudaDeviceSynchronize();
      long long h[296]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
      const double bytes_per_sm = (double)blocks_per_sm * aw * iters * 4 * 32 * 32 * 4;
      printf("CTAs/SM %d, loading warps/CTA %d: %.1f B/clk/SM (clock64 units)  [%s]\n", blocks_per_sm, aw,
             bytes_per_sm / avg, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
