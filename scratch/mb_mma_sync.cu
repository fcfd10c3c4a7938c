// Microbenchmark: peak rate of the legacy mma.sync.m16n8k16 bf16 path on this GPU (independent accumulators, no memory).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int NACC>
__global__ void __launch_bounds__(256) k(int iters, float* out) {
  float d[NACC][4];
  for (int i = 0; i < NACC; ++i) for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < NACC; ++i) for (int e = 0; e < 4; ++e) s += d[i][e];
  if (s == 12345.f) out[0] = s;
}
template <int NACC>
void run(int warps_per_sm) {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  float* out; cudaMalloc(&out, 4);
  const int iters = 20000;
  const int threads = 256, blocks = nsm * warps_per_sm / 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NACC><<<blocks, threads>>>(100, out);
  cudaEventRecord(e0);
  k<NACC><<<blocks, threads>>>(iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double mmas = double(blocks) * 8 * iters * NACC;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("NACC=%d warps/SM=%2d: %.1f TFLOP/s dense bf16, %.3f MMA/clk/SM (at %d MHz nominal)\n", NACC, warps_per_sm, mmas * 4096 / ms / 1e9,
         mmas / nsm / (ms * 1e-3 * clk * 1e3), clk / 1000);
}
int main() {
  run<1>(8); run<1>(32); run<2>(16); run<4>(8); run<4>(16); run<4>(32); run<8>(8); run<8>(16); run<8>(32);
  return 0;
}
