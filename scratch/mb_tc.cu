// Micro-benchmarks that decide the design of the stage kernel (run once on a B200; results -> profiles/).
//  T1  tcgen05.shift semantics (direction, width, last row)
//  T2  tcgen05.mma with the A operand in TMEM: packing + numerics
//  T3  tensor-pipe cost of one conv layer for several operand arrangements
//  T4  tcgen05.ld / tcgen05.st throughput,  T5  SHFL throughput
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo >> 4) & 0x3FFFu) << 16) | (uint64_t((sbo >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}
__host__ __device__ constexpr uint32_t idesc(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void tshift(uint32_t a) { asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
#define LD8(r, t) asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t) : "memory")
#define ST8(t, r) asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(t), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory")
#define LD32(r, taddr)                                                                                               \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                    \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
      : "r"(taddr) : "memory")
#define ST32(taddr, r)                                                                                               \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, " \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),             \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),  \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),    \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),    \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory")

struct Out {
  uint32_t shift_before[128][16];
  uint32_t shift_after1[128][16];
  uint32_t shift_after2[128][16];
  float mma_ts[128][32];
  float mma_ts_shift[128][32];
  long long cyc[32];
};

// smem: A tile (K-major no swizzle, 130 rows x 4 planes) + B weights [k/8][n<=96][8]
__global__ void __launch_bounds__(512, 1) k_mb(Out* o) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char* sA = smem;                 // 4 planes x 130 rows x 16 B
  unsigned char* sB = smem + 16384;         // [k/8 (12)][n (96)][8] bf16 = 12*96*16 = 18432 B
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B: n<16: identity on k (b[n][k] = n==k), n==16: all ones, else 0.   N up to 96, K = 16 (2 core-matrix columns)
  for (int e = tid; e < 12 * 96 * 8; e += blockDim.x) {
    const int kk = e / (96 * 8), n = (e / 8) % 96, k8 = e % 8, k = kk * 8 + k8;
    float v = 0.f;
    if (kk < 2) v = (n < 16) ? (n == k ? 1.f : 0.f) : (n == 16 ? 1.f : 0.f);
    reinterpret_cast<__nv_bfloat16*>(sB)[e] = __float2bfloat16(v);
  }
  for (int e = tid; e < 16384 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(sA)[e] = 0x3F803F80u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
  const uint32_t barA = smem_u32(&bar);
  uint32_t ph = 0;
  const int row = (warp & 3) * 32 + (tid & 31);

  // ---------------- T1: shift semantics.  A region at columns 256..271 (16 columns), value = row*256 + col
  if (warp < 4) {
    uint32_t v[8];
    for (int h = 0; h < 2; ++h) {
      for (int c = 0; c < 8; ++c) v[c] = row * 256 + h * 8 + c;
      ST8(tb + lane_off + 256 + h * 8, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int h = 0; h < 2; ++h) {
      LD8(v, tb + lane_off + 256 + h * 8);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 8; ++c) o->shift_before[row][h * 8 + c] = v[c];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  for (int rep = 0; rep < 2; ++rep) {
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tshift(tb + 256);  // only the first 8 columns (32 bytes)?
      commit(barA);
    }
    mbar_wait(barA, ph); ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
      uint32_t v[8];
      for (int h = 0; h < 2; ++h) {
        LD8(v, tb + lane_off + 256 + h * 8);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int c = 0; c < 8; ++c) (rep ? o->shift_after2 : o->shift_after1)[row][h * 8 + c] = v[c];
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }

  // ---------------- T2: MMA with A in TMEM. a[row][k] = (row + k) % 7, packed 2 bf16 per column (k even in the low half)
  if (warp < 4) {
    uint32_t v[8];
    for (int c = 0; c < 8; ++c) {
      __nv_bfloat162 p = __floats2bfloat162_rn(float((row + 2 * c) % 7), float((row + 2 * c + 1) % 7));
      v[c] = *reinterpret_cast<uint32_t*>(&p);
    }
    ST8(tb + lane_off + 256, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  const uint64_t dB = umma_desc(smem_u32(sB), 96 * 16, 128);
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mma_ts(tb + 0, tb + 256, dB, idesc(32), 0);
    tshift(tb + 256);
    mma_ts(tb + 32, tb + 256, dB, idesc(32), 0);
    commit(barA);
  }
  mbar_wait(barA, ph); ph ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp < 4) {
    uint32_t v[32];
    LD32(v, tb + lane_off + 0);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) o->mma_ts[row][c] = __uint_as_float(v[c]);
    LD32(v, tb + lane_off + 32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int c = 0; c < 32; ++c) o->mma_ts_shift[row][c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();

  // ---------------- T3: tensor-pipe cost of one "layer" for several arrangements (REPS layers over 4 D regions)
  const uint64_t dA = umma_desc(smem_u32(sA), 130 * 16, 128);
  constexpr int REPS = 256;
  for (int variant = 0; variant < 10; ++variant) {
    long long t0 = 0;
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      t0 = clock64();
      for (int r = 0; r < REPS; ++r) {
        const uint32_t d = tb + (r & 3) * 32, d96 = tb + (r & 1) * 96, a = tb + 256 + (r & 3) * 32;
        switch (variant) {
          case 0:  // current kernel: 7 x (N=32, K=16), A from smem
            for (int i = 0; i < 7; ++i) mma_ss(d, dA + (i % 3), dB, idesc(32), i > 0);
            break;
          case 1:  // 7 x (N=32, K=16), A from TMEM, no shifts
            for (int i = 0; i < 7; ++i) mma_ts(d, a + (i & 1) * 8, dB, idesc(32), i > 0);
            break;
          case 2:  // A from TMEM with shifts: const, 2 mma, 2 shift, 2 mma, 2 shift, 2 mma
            mma_ts(d, a + 16, dB, idesc(32), 0);
            mma_ts(d, a, dB, idesc(32), 1); mma_ts(d, a + 8, dB, idesc(32), 1);
            tshift(a); tshift(a + 8);
            mma_ts(d, a, dB, idesc(32), 1); mma_ts(d, a + 8, dB, idesc(32), 1);
            tshift(a); tshift(a + 8);
            mma_ts(d, a, dB, idesc(32), 1); mma_ts(d, a + 8, dB, idesc(32), 1);
            break;
          case 3:  // N=96 trick: const(N=32) + 2 x (N=96, K=16), A from smem
            mma_ss(d96, dA, dB, idesc(96), 0); mma_ss(d96, dA + 1, dB, idesc(96), 1);
            break;
          case 4:  // shifts only: 4 per layer
            tshift(a); tshift(a + 8); tshift(a); tshift(a + 8);
            break;
          case 5:  // N=96 with A from TMEM
            mma_ts(d96, a, dB, idesc(96), 0); mma_ts(d96, a + 8, dB, idesc(96), 1);
            break;
          case 6:  // single MMA N=32 ss
            mma_ss(d, dA, dB, idesc(32), 0);
            break;
          case 7:  // single MMA N=32 ts
            mma_ts(d, a, dB, idesc(32), 0);
            break;
          case 8:  // single MMA N=256 ss
            mma_ss(tb, dA, dB, idesc(256), 0);
            break;
          case 9:  // single shift
            tshift(a);
            break;
        }
      }
      commit(barA);
    }
    mbar_wait(barA, ph); ph ^= 1;
    if (tid == 0) o->cyc[variant] = clock64() - t0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();
  }

  // ---------------- T4: tcgen05.ld / st throughput (x32), nw warps active
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int nw = cfg == 0 ? 4 : (cfg == 1 ? 8 : 16);
    __syncthreads();
    long long t0 = clock64();
    uint32_t acc = 0;
    if (warp < nw) {
      uint32_t v[32];
      for (int r = 0; r < 256; ++r) {
        LD32(v, tb + lane_off + ((r + warp) & 7) * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += v[0] ^ v[31];
      }
    }
    __syncthreads();
    if (tid == 0) o->cyc[10 + cfg] = clock64() - t0;
    if (acc == 0x12345) o->cyc[31] = acc;
    __syncthreads();
    t0 = clock64();
    if (warp < nw) {
      uint32_t v[32];
      for (int c = 0; c < 32; ++c) v[c] = c + tid;
      for (int r = 0; r < 256; ++r) {
        ST32(tb + lane_off + ((r + warp) & 7) * 32, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) o->cyc[13 + cfg] = clock64() - t0;
  }
  // ---------------- T5: SHFL throughput, 16 warps x 1024
  {
    __syncthreads();
    long long t0 = clock64();
    uint32_t x = tid;
#pragma unroll 16
    for (int r = 0; r < 1024; ++r) x = __shfl_up_sync(0xffffffffu, x, 1) + 1;
    __syncthreads();
    if (tid == 0) o->cyc[16] = clock64() - t0;
    if (x == 0x12345) o->cyc[31] = x;
    // independent shuffles (ILP 8)
    uint32_t y[8];
    for (int i = 0; i < 8; ++i) y[i] = tid + i;
    __syncthreads();
    t0 = clock64();
#pragma unroll 4
    for (int r = 0; r < 128; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = __shfl_up_sync(0xffffffffu, y[i], 1);
    __syncthreads();
    if (tid == 0) o->cyc[17] = clock64() - t0;
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s += y[i];
    if (s == 0x12345) o->cyc[31] = s;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}


// ---------------- clean tensor-pipe timing: one issuing lane, its warp-mates parked at __syncwarp, everyone else on the mbarrier
__global__ void __launch_bounds__(256, 1) k_t3(long long* cyc, int variant, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  unsigned char* sA = smem;
  unsigned char* sB = smem + 16384 * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tslot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int e = tid; e < 65536 / 4; e += blockDim.x) reinterpret_cast<uint32_t*>(smem)[e] = 0x3C003C00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tslot;
  const uint32_t barA = smem_u32(&bar);
  const uint64_t dA = umma_desc(smem_u32(sA), 130 * 16, 128);
  const uint64_t dA2 = umma_desc(smem_u32(sA) + 16384, 130 * 16, 128);
  const uint64_t dB = umma_desc(smem_u32(sB), 96 * 16, 128);
  long long t0 = 0;
  if (warp == 1) {
    if ((tid & 31) == 0) {
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tb + (r & 7) * 32;
        switch (variant) {
          case 0:  // 7 dependent N=32 (current kernel's batch), D rotates over 8 regions per batch
            for (int i = 0; i < 7; ++i) mma_ss(d, dA + (i % 3), dB, idesc(32), i > 0);
            break;
          case 1:  // 7 independent N=32 (7 different D)
            for (int i = 0; i < 7; ++i) mma_ss(tb + i * 32, dA + (i % 3), dB, idesc(32), 0);
            break;
          case 2:  // two interleaved dependent chains of 7 (2 batches)
            for (int i = 0; i < 7; ++i) { mma_ss(d, dA + (i % 3), dB, idesc(32), i > 0); mma_ss(d + 256, dA2 + (i % 3), dB, idesc(32), i > 0); }
            break;
          case 3:  // 1 x N=32
            mma_ss(d, dA, dB, idesc(32), 0);
            break;
          case 4:  // 2 dependent N=96
            mma_ss(tb + (r & 3) * 96, dA, dB, idesc(96), 0); mma_ss(tb + (r & 3) * 96, dA + 1, dB, idesc(96), 1);
            break;
          case 5:  // 7 dependent, A from TMEM
            for (int i = 0; i < 7; ++i) mma_ts(d, tb + 256 + (i & 1) * 8, dB, idesc(32), i > 0);
            break;
          case 6:  // 7 independent, A from TMEM
            for (int i = 0; i < 7; ++i) mma_ts(tb + i * 32, tb + 256 + (i & 1) * 8, dB, idesc(32), 0);
            break;
          case 7:  // 1 x N=256
            mma_ss(tb, dA, dB, idesc(256), 0);
            break;
          case 8:  // 1 x N=64
            mma_ss(d, dA, dB, idesc(64), 0);
            break;
          case 9:  // 1 x N=128
            mma_ss(tb + (r & 3) * 128, dA, dB, idesc(128), 0);
            break;
          case 10:  // 4 dependent N=64 (row-pair formulation: K=64 per 2 taps...) 
            for (int i = 0; i < 4; ++i) mma_ss(tb + (r & 3) * 64, dA + i, dB, idesc(64), i > 0);
            break;
          case 11:  // 4 shifts
            tshift(tb + 256); tshift(tb + 264); tshift(tb + 256); tshift(tb + 264);
            break;
          case 12:  // 7 dependent N=32 + commit per batch (as the real kernel does)
            for (int i = 0; i < 7; ++i) mma_ss(d, dA + (i % 3), dB, idesc(32), i > 0);
            if (r + 1 < reps) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar) + 0) : "memory");
            break;
        }
      }
      if (variant != 12) commit(barA);
      else commit(barA);
    }
    __syncwarp();
  }
  if (variant == 12) {
    // the barrier completes reps times; wait for each phase
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) { mbar_wait(barA, ph); ph ^= 1; }
  } else {
    mbar_wait(barA, 0);
  }
  if (tid == 32) cyc[variant] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  Out* d;
  CK(cudaMalloc(&d, sizeof(Out)));
  CK(cudaMemset(d, 0xEE, sizeof(Out)));
  CK(cudaFuncSetAttribute(k_mb, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  k_mb<<<1, 512, 65536>>>(d);
  CK(cudaDeviceSynchronize());
  std::vector<Out> hv(1);
  Out& h = hv[0];
  CK(cudaMemcpy(&h, d, sizeof(Out), cudaMemcpyDeviceToHost));
  auto show = [&](const char* name, uint32_t (*m)[16]) {
    printf("%s (value = row*256+col):\n", name);
    int rows[] = {0, 1, 2, 30, 31, 32, 33, 63, 64, 126, 127};
    for (int r : rows) {
      printf("  row %3d:", r);
      for (int c = 0; c < 16; ++c) printf(" %d.%d", m[r][c] >> 8, m[r][c] & 255);
      printf("\n");
    }
  };
  show("before", h.shift_before);
  show("after 1 shift of [col 0..7]", h.shift_after1);
  show("after 2 shifts", h.shift_after2);
  int bad = 0;
  for (int r = 0; r < 128; ++r) {
    for (int n = 0; n < 16; ++n) if (h.mma_ts[r][n] != float((r + n) % 7)) ++bad;
    float s = 0; for (int k = 0; k < 16; ++k) s += (r + k) % 7;
    if (h.mma_ts[r][16] != s) ++bad;
  }
  printf("T2 mma A-in-TMEM mismatches (assuming k=2c in low half): %d\n", bad);
  printf("  row0: "); for (int n = 0; n < 18; ++n) printf("%g ", h.mma_ts[0][n]); printf("\n");
  printf("  row5: "); for (int n = 0; n < 18; ++n) printf("%g ", h.mma_ts[5][n]); printf("\n");
  printf("  after shift row5: "); for (int n = 0; n < 18; ++n) printf("%g ", h.mma_ts_shift[5][n]); printf("\n");
  printf("  after shift row127: "); for (int n = 0; n < 18; ++n) printf("%g ", h.mma_ts_shift[127][n]); printf("\n");
  printf("  after shift row0: "); for (int n = 0; n < 18; ++n) printf("%g ", h.mma_ts_shift[0][n]); printf("\n");
  const char* names[] = {"7x mma_ss N32", "7x mma_ts N32", "7 mma_ts + 4 shift", "2x mma_ss N96", "4 shifts", "2x mma_ts N96", "1 mma_ss N32", "1 mma_ts N32", "1 mma_ss N256", "1 shift"};
  for (int v = 0; v < 10; ++v) printf("T3 %-20s %8.1f cyc/layer (256 layers, incl. ~issue+commit latency)\n", names[v], h.cyc[v] / 256.0);
  for (int c = 0; c < 3; ++c) printf("T4 ld x32 nw=%2d: %6.1f cyc per load-round; st: %6.1f\n", c == 0 ? 4 : (c == 1 ? 8 : 16), h.cyc[10 + c] / 256.0, h.cyc[13 + c] / 256.0);
  {
    long long* dc; CK(cudaMalloc(&dc, 64 * 8)); CK(cudaMemset(dc, 0, 64 * 8));
    CK(cudaFuncSetAttribute(k_t3, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const char* nm[] = {"7 dep N32 ss", "7 indep N32 ss", "2x7 interleaved dep N32", "1 N32", "2 dep N96", "7 dep N32 ts", "7 indep N32 ts", "1 N256", "1 N64", "1 N128", "4 dep N64", "4 shifts", "7 dep N32 + commit each"};
    for (int rep = 0; rep < 2; ++rep)
      for (int v = 0; v < 12; ++v) { k_t3<<<1, 256, 65536>>>(dc, v, rep ? 512 : 256); CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dc + v, 8, cudaMemcpyDeviceToHost));
        printf("T3v2 reps=%d %-28s total %8lld cyc  = %7.1f cyc/batch\n", rep ? 512 : 256, nm[v], c, double(c) / (rep ? 512 : 256)); }
  }
  printf("T5 shfl dependent chain, 16 warps: %.2f cyc per shfl round; independent: %.2f cyc per 16-warp round\n", h.cyc[16] / 1024.0, h.cyc[17] / 1024.0);
  return 0;
}
