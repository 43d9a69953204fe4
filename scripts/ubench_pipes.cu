// Debug microbenchmark (not part of the library): per-SMSP issue cost of the instructions the attention softmax uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/_build/ubench_pipes scripts/ubench_pipes.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 256
template <int OP>
__global__ void k(float* out, long long* clk, float seed) {
  float a[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f; p[i] = (unsigned long long)__float_as_uint(a[i]) << 32 | __float_as_uint(a[i]); }
  const unsigned long long c2 = (unsigned long long)__float_as_uint(0.999f) << 32 | __float_as_uint(0.999f);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(0.999f));
      if (OP == 4) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(r) : "f"(a[i])); a[i] = __uint_as_float(r); }
      if (OP == 5) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 6) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
      if (OP == 7) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
      if (OP == 8) {   // mix: 1 MUFU + 1 FFMA2 + 1 FADD2 independent
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int OP>
void run(const char* name, int warps_per_smsp) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  const int threads = 128 * warps_per_smsp;
  k<OP><<<148, threads>>>(out, clk, 1.0f);
  k<OP><<<148, threads>>>(out, clk, 1.0f);
  long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %d warp/SMSP: %.2f clk per warp-instruction per SMSP (%.2f per instr seen by one warp)\n", name, warps_per_smsp,
         (double)c / (ITERS * 8.0 * warps_per_smsp), (double)c / (ITERS * 8.0));
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int w = 1; w <= 2; ++w) {
    run<0>("MUFU.EX2 f32", w); run<1>("FFMA2", w); run<2>("FADD2", w); run<3>("FFMA", w); run<4>("F2FP bf16x2", w);
    run<5>("FMNMX3", w); run<6>("MUFU.EX2 bf16x2", w); run<7>("MUFU.EX2 f16x2", w); run<8>("MUFU+FFMA2 pair", w);
  }
  return 0;
}
