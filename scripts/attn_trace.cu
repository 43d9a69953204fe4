// Debug harness (not part of the library): runs attention2_kernel at the bench shape with -DUTX_ATTN_TRACE and prints
// the SM-clock timeline of one CTA's MMA thread and two softmax warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DUTX_ATTN_TRACE -o scripts/_build/attn_trace \
//        scripts/attn_trace.cu unitex_b200/csrc/common.cu -lcuda
#include "../unitex_b200/csrc/attn2_sm100.cu"

#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
  const int S = argc > 1 ? atoi(argv[1]) : 9728, H = 24;
  const long ld = 3L * H * 128;
  std::vector<__nv_bfloat16> h(static_cast<size_t>(S) * ld);
  unsigned s = 12345;
  for (auto& v : h) { s = s * 1664525u + 1013904223u; v = __float2bfloat16(((s >> 8) & 0xffff) / 65536.0f * 2.f - 1.f); }
  __nv_bfloat16 *qkv, *out;
  cudaMalloc(&qkv, h.size() * 2);
  cudaMalloc(&out, static_cast<size_t>(S) * H * 128 * 2);
  cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  long long* tr;
  const size_t n = 3 * 128 * 8;
  cudaMalloc(&tr, n * 8);
  cudaMemset(tr, 0, n * 8);
  cudaMemcpyToSymbol(utx::g_attn_trace, &tr, sizeof(tr));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) utx::attention2_bf16(qkv, ld, out, H * 128, S, H, 0);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) utx::attention2_bf16(qkv, ld, out, H * 128, S, H, 0);
  cudaEventRecord(e1);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s / %s\n", cudaGetErrorString(cudaGetLastError()), utx::get_error()); return 1; }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("S=%d: %.3f ms per launch, %.1f TFLOP/s\n", S, ms / 5, 4.0 * S * S * 128 * H / (ms / 5 * 1e-3) / 1e12);
  std::vector<long long> t(n);
  cudaMemcpy(t.data(), tr, n * 8, cudaMemcpyDeviceToHost);
  auto T = [&](int role, int j, int p) { return t[(role * 128 + j) * 8 + p]; };
  const long long base = T(2, 20, 0);
  const char* names[3] = {"softmaxA", "softmaxB", "mma"};
  for (int j = 20; j < 26; ++j)
    for (int role = 0; role < 3; ++role) {
      printf("j=%d %-8s:", j, names[role]);
      for (int p = 0; p < 6; ++p) printf(" %7lld", T(role, j, p) - base);
      printf("\n");
    }
  printf("mma iteration period (clk): %lld\n", (T(2, 60, 0) - T(2, 20, 0)) / 40);
  printf("softmax points: 0 before S_FULL wait, 1 after, 2 max loads landed, 3 max/rescale done, 4 exp loop done, 5 P_FULL arrived\n");
  printf("mma points: 0 loop top, 1 V_FULL ok, 2 P_FULL(A) ok, 3 PV_A+S_A issued, 4 P_FULL(B) ok, 5 PV_B+S_B issued\n");
  return 0;
}
