// tcgen05.mma issue/throughput micro-benchmark: straight-line issue of many MMAs by 1, 2 or 4 issuing warps.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo >> 4) & 0x3FFFu) << 16) | (uint64_t((sbo >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}
__host__ __device__ constexpr uint32_t idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24); }
template <int N>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc(N)), "r"(acc) : "memory");
}
template <int N>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc(N)), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}

// NI issuing warps (warps 4..4+NI-1, lane 0); each issues ITER x 8 MMAs (N, K=16), accumulate chain per 8; TS: A from TMEM
template <int N, int NI, bool TS, bool DEP, bool CE = false>
__global__ void __launch_bounds__(256, 1) k_issue(long long* out, int iters) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint64_t dummy[4];
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int e = tid; e < 98304 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x3C003C00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  long long t0 = clock64();
  if (warp >= 4 && warp < 4 + NI) {
    const int w = warp - 4;
    if ((tid & 31) == 0) {
      const uint64_t dA = umma_desc(smem_u32(smem) + w * 16384, 130 * 16, 128);
      const uint64_t dB = umma_desc(smem_u32(smem) + 65536, 256 * 16, 128);
      const uint32_t dbase = tb + (N <= 64 ? w * 128 : 0);
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t d = DEP ? dbase : dbase + (N <= 32 ? (i & 3) * 32 : 0);
          if (TS) mma_ts<N>(d, tb + 480 + (i & 1) * 8, dB, (DEP && i > 0) ? 1u : 0u);
          else mma_ss<N>(d, dA + (i % 3), dB, (DEP && i > 0) ? 1u : 0u);
        }
        if (CE) commit(smem_u32(&dummy[w]));  // one commit per batch of 8 MMAs, as the stage kernel does per layer
      }
      commit(smem_u32(&bar[w]));
    }
    __syncwarp();
  }
  if (warp < NI) mbar_wait(smem_u32(&bar[warp]), 0);
  __syncthreads();
  if (tid == 0) out[0] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

template <int N, int NI, bool TS, bool DEP, bool CE = false>
void run(long long* d, const char* name) {
  CK(cudaFuncSetAttribute(k_issue<N, NI, TS, DEP, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  long long c[2];
  for (int r = 0; r < 2; ++r) {
    const int iters = r ? 128 : 64;
    k_issue<N, NI, TS, DEP, CE><<<1, 256, 98304>>>(d, iters);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c[r], d, 8, cudaMemcpyDeviceToHost));
  }
  // slope between 64 and 128 iterations removes fixed costs
  const double per = double(c[1] - c[0]) / (64.0 * 8 * NI);
  printf("%-44s N=%3d issuers=%d: %7.1f cyc per MMA (aggregate), %7.1f cyc per MMA per issuer\n", name, N, NI, per, per * NI);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  long long* d; CK(cudaMalloc(&d, 64));
  run<32, 4, false, true, true>(d, "ss dependent + commit per 8");
  run<32, 2, false, true, true>(d, "ss dependent + commit per 8");
  run<32, 1, false, true, true>(d, "ss dependent + commit per 8");
  run<32, 1, false, true>(d, "ss dependent");
  run<32, 1, false, false>(d, "ss independent");
  run<32, 2, false, true>(d, "ss dependent");
  run<32, 4, false, true>(d, "ss dependent");
  run<32, 1, true, true>(d, "ts dependent");
  run<32, 4, true, true>(d, "ts dependent");
  run<64, 1, false, true>(d, "ss dependent");
  run<64, 4, false, true>(d, "ss dependent");
  run<96, 1, false, true>(d, "ss dependent");
  run<96, 4, false, true>(d, "ss dependent");
  run<128, 1, false, true>(d, "ss dependent");
  run<128, 4, false, true>(d, "ss dependent");
  run<256, 1, false, true>(d, "ss dependent");
  run<256, 2, false, true>(d, "ss dependent");
  run<64, 4, true, true>(d, "ts dependent");
  run<128, 4, true, true>(d, "ts dependent");
  return 0;
}
