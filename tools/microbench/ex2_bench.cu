// Throughput of the exp2 variants a softmax inner loop can use on sm_100a (results per clock per SM).
// build+run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ex2_bench ex2_bench.cu && /tmp/ex2_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, long long* cyc, int iters) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = 0x3c003800u + threadIdx.x * 3 + i * 17;  // two small fp16s / one small float
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 3) asm volatile("{.reg .f32 a; mov.b32 a, %0; cvt.rn.f16x2.f32 %0, a, a;}" : "+r"(r[i]));
      if (MODE == 4) asm volatile("{.reg .f32 a; mov.b32 a, %0; cvt.rn.bf16x2.f32 %0, a, a;}" : "+r"(r[i]));
      if (MODE == 5) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(r[i]));
      if (MODE == 6) asm volatile("{.reg .f32 a; mov.b32 a, %0; fma.rn.f32 a, a, a, a; mov.b32 %0, a;}" : "+r"(r[i]));
      if (MODE == 7) asm volatile("{.reg .f16 lo, hi; .reg .f32 a; mov.b32 {lo, hi}, %0; cvt.f32.f16 a, lo; mov.b32 %0, a;}" : "+r"(r[i]));
      if (MODE == 8) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(r[i]));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_instr) {
  uint32_t* out; long long* cyc;
  const int blocks = 148 * 2, iters = 2000;
  cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&cyc, blocks * 8);
  k<MODE><<<blocks, 256>>>(out, cyc, 10);
  k<MODE><<<blocks, 256>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
  // 2 blocks x 256 threads per SM, each 16*iters instructions
  const double instr_per_sm = 2.0 * 256 * 16 * iters;
  printf("%-28s %8.1f thread-instr/clk/SM  -> %8.1f results/clk/SM   (%s)\n", name, instr_per_sm / avg,
         instr_per_sm / avg * per_instr, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("cvt.rn.f16x2.f32", 2);
  run<4>("cvt.rn.bf16x2.f32", 2);
  run<5>("fma.rn.f16x2", 2);
  run<6>("fma.rn.f32", 1);
  run<7>("cvt.f32.f16", 1);
  run<8>("fma.rn.bf16x2", 2);
  return 0;
}
