// Local branch of Network2 on the tensor cores (MURAL_MODE_BF16): embedding gather + Linear(K1,H1)·ReLU·BN +
// Linear(H1,H2)·ReLU·BN + Linear(H2,n_class)   (MuRaL/model/model_snv.py:452-468, 492; eval BN folded forward).
//
// One CTA = 128 sites = the 128 rows of every MMA; two warpgroups share a tile (thread = site = TMEM lane, the warpgroups
// take alternate 8/16-column groups of the gather and of the accumulator read-back, which are the latency-bound parts of the
// per-tile chain); the three GEMMs are chained on chip:
//   A operand (activations, K-major SWIZZLE_NONE planes [k/8][128 rows][8] in shared memory)  x  B operand (weights
//   [k/8][N][8], resident in shared memory for the whole kernel)  ->  fp32 accumulator in TMEM  ->  tcgen05.ld, ReLU,
//   -> next layer's A operand.
// fp32-grade accuracy from bf16 MMAs: every operand is split x = hi + lo (two bf16) and a product is evaluated as
// hi*hi + hi*lo + lo*hi (the dropped lo*lo term is ~2^-18 relative), so the local logits agree with the fp32 kernel to
// ~1e-6.  The bias rides along as one extra K column (activation 1.0, weight row = bias), which splits it the same way.
// Tensor time per tile is ~3.2k cycles against ~2 ms per million sites of the CUDA-core kernel it replaces.
#include <cuda_bf16.h>
#include <string.h>

#include <vector>

#include "snv_model.cuh"

namespace mural {
namespace mlptc {

constexpr int TILE = 128;
constexpr int THREADS = 256;      // two warpgroups per tile
constexpr int PLANE = TILE * 16;  // bytes of one 8-column plane of an A operand
constexpr int NPF = 12;           // k-mer indices per thread the register prefetch can hold (n_cat <= 2 * NPF)

struct Dims {
  int n_cat, K1, H1, H2, NC, emb_rows;
  int K1p, P1, P2;  // K1+1, H1+1, H2+1 rounded up to 16 (the +1 is the bias column)
  int maxKp;        // max(K1p, P1, P2): hi planes at [0, maxKp/8), lo planes after them
  int w_off[3][2];  // byte offsets of W1/W2/W3 hi and lo inside the weight blob
  int w_bytes;
};

struct State {
  Dims d;
  uint8_t* d_w = nullptr;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ uint32_t idesc_n(int N) {  // kind::f16, D=F32, A=B=BF16, K-major, M=128
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t(N) >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
#define TMEM_LD16(r, taddr)                                                                                         \
  asm volatile(                                                                                                     \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                     \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"                            \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                    \
      : "r"(taddr)                                                                                                  \
      : "memory")

// x = hi + lo as two bf16 pairs
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float r0 = x0 - __uint_as_float(hi << 16), r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

__global__ void __launch_bounds__(THREADS, 1) k_local_mlp_tc(Dims d, const uint8_t* __restrict__ wblob, const float* __restrict__ emb,
                                                          const int32_t* __restrict__ cat32, const int64_t* __restrict__ cat64,
                                                          int64_t n, float* __restrict__ logits, int* __restrict__ err_flag) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sW = smem;                          // weight blob
  unsigned char* sA = smem + ((d.w_bytes + 127) & ~127);  // [2 (hi, lo)][maxKp/8][128][8] bf16
  int32_t* sCat = reinterpret_cast<int32_t*>(sA + 2 * (d.maxKp / 8) * PLANE);  // [128][n_cat] k-mer indices of the tile
  float* sEmb = reinterpret_cast<float*>(sCat + TILE * d.n_cat);             // [emb_rows][5]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid & (TILE - 1), wg = tid >> 7;   // site of the tile / warpgroup
  const int LO = (d.maxKp / 8) * PLANE;              // byte offset of the lo planes
  for (int e = tid; e < d.w_bytes / 16; e += THREADS) reinterpret_cast<uint4*>(sW)[e] = __ldg(reinterpret_cast<const uint4*>(wblob) + e);
  for (int e = tid; e < d.emb_rows * 5; e += THREADS) sEmb[e] = __ldg(emb + e);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;  // a warp reaches the TMEM lanes of its quarter of the warpgroup
  const uint32_t D1 = tmem, D2 = tmem + 256, D3 = tmem + 384;  // P1 <= 256, P2 <= 128 columns (checked on the host)
  const uint32_t barA = smem_u32(&bar);
  const uint32_t aBase = smem_u32(sA), wBase = smem_u32(sW);
  uint32_t phase = 0;

  // D[128 x N] = A[128 x Kp] * W[Kp x N] as three split products; issued by one thread, completion on the mbarrier
  auto gemm = [&](uint32_t dcol, int Kp, int N, int layer) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t id = idesc_n(N);
      uint32_t acc = 0;
      for (int term = 0; term < 3; ++term) {  // hi*hi, hi*lo, lo*hi
        const uint32_t a0 = aBase + (term == 2 ? LO : 0);
        const uint32_t w0 = wBase + d.w_off[layer][term == 1 ? 1 : 0];
        for (int j = 0; j < Kp / 16; ++j) {
          umma(dcol, umma_desc(a0 + 2 * j * PLANE, PLANE, 128), umma_desc(w0 + 2 * j * N * 16, N * 16, 128), id, acc);
          acc = 1;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barA) : "memory");
    }
    __syncwarp();  // the issuing lane's warp-mates must not spin in try_wait beside it (it would time-slice the issue)
    mbar_wait(barA, phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  };
  // ReLU(D[:, 0..P)) -> next A operand (hi/lo planes), column `one_col` forced to 1.0 (bias column of the next layer)
  auto relu_to_A = [&](uint32_t dcol, int P, int one_col) {
    // 16-column groups alternate between the warpgroups; two loads in flight per wait
    for (int c0 = wg; c0 < P / 16; c0 += 4) {
      uint32_t v[2][16];
      const bool two = c0 + 2 < P / 16;
      TMEM_LD16(v[0], dcol + lane_off + 16 * c0);
      if (two) TMEM_LD16(v[1], dcol + lane_off + 16 * (c0 + 2));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const int c = c0 + 2 * u;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x0 = fmaxf(__uint_as_float(v[u][2 * i]), 0.f), x1 = fmaxf(__uint_as_float(v[u][2 * i + 1]), 0.f);
          if (16 * c + 2 * i == one_col) x0 = 1.f;
          if (16 * c + 2 * i + 1 == one_col) x1 = 1.f;
          split2(x0, x1, hi[i], lo[i]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          *reinterpret_cast<uint4*>(sA + (2 * c + h) * PLANE + row * 16) = make_uint4(hi[4 * h], hi[4 * h + 1], hi[4 * h + 2], hi[4 * h + 3]);
          *reinterpret_cast<uint4*>(sA + LO + (2 * c + h) * PLANE + row * 16) = make_uint4(lo[4 * h], lo[4 * h + 1], lo[4 * h + 2], lo[4 * h + 3]);
        }
      }
    }
  };

  const int64_t n_tiles = (n + TILE - 1) / TILE;
  int pf[NPF];  // prefetched int32 k-mer indices of the next tile (n_cat <= NPF, checked on the host)
#pragma unroll
  for (int i = 0; i < NPF; ++i) {
    const int64_t gi = int64_t(blockIdx.x) * TILE * d.n_cat + tid + i * THREADS;
    pf[i] = (cat32 && tid + i * THREADS < TILE * d.n_cat && gi < n * d.n_cat) ? __ldg(cat32 + gi) : 0;
  }
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t site0 = tile * TILE;
    // ---- k-mer indices of the tile -> shared memory, range-checked like nn.Embedding.  int32 indices (the library's own
    // encoder) were prefetched into registers while the previous tile was computing.
    if (cat32) {
#pragma unroll
      for (int i = 0; i < NPF; ++i) {
        const int e = tid + i * THREADS;
        if (e < TILE * d.n_cat) {
          int idx = pf[i];
          if (idx < 0 || idx >= d.emb_rows) {  // nn.Embedding would raise IndexError
            if (err_flag && site0 * d.n_cat + e < n * d.n_cat) atomicOr(err_flag, 2);
            idx = 0;
          }
          sCat[e] = idx;
        }
      }
      const int64_t next0 = (tile + gridDim.x) * TILE * d.n_cat;
#pragma unroll
      for (int i = 0; i < NPF; ++i) {
        const int64_t gi = next0 + tid + i * THREADS;
        pf[i] = (tid + i * THREADS < TILE * d.n_cat && gi < n * d.n_cat) ? __ldg(cat32 + gi) : 0;
      }
    } else {
      for (int e = tid; e < TILE * d.n_cat; e += THREADS) {
        const int64_t gi = site0 * d.n_cat + e;
        int64_t idx = 0;
        if (gi < n * d.n_cat) {
          idx = cat64[gi];
          if (idx < 0 || idx >= d.emb_rows) {
            if (err_flag) atomicOr(err_flag, 2);
            idx = 0;
          }
        }
        sCat[e] = int32_t(idx);
      }
    }
    __syncthreads();
    // ---- embedding gather -> A1: thread = site; column k = 5*j + e is component e of the embedding of k-mer j,
    // column K1 = 1.0 (bias), one 16-byte store per 8-column plane
    {
      const int32_t* myCat = sCat + row * d.n_cat;
      for (int pl = wg; pl < d.K1p / 8; pl += 2) {  // 8-column planes alternate between the warpgroups
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float x[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int k = 8 * pl + 2 * i + h;
            float v = 0.f;
            if (k < d.K1) {
              const int j = k / 5;
              v = sEmb[myCat[j] * 5 + (k - 5 * j)];
            } else if (k == d.K1) {
              v = 1.f;
            }
            x[h] = v;
          }
          split2(x[0], x[1], hi[i], lo[i]);
        }
        *reinterpret_cast<uint4*>(sA + pl * PLANE + row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(sA + LO + pl * PLANE + row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    gemm(D1, d.K1p, d.P1, 0);
    relu_to_A(D1, d.P1, d.H1);
    gemm(D2, d.P1, d.P2, 1);
    relu_to_A(D2, d.P2, d.H2);
    gemm(D3, d.P2, 16, 2);
    if (wg == 0) {
      uint32_t v[16];
      TMEM_LD16(v, D3 + lane_off);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (site0 + row < n) {
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (o < d.NC) logits[(site0 + row) * d.NC + o] = __uint_as_float(v[o]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // every lane has read D3 and nobody still reads sA before the next gather overwrites it
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static size_t smem_bytes(const Dims& d) {
  return size_t((d.w_bytes + 127) & ~127) + size_t(2) * (d.maxKp / 8) * PLANE + size_t(TILE) * d.n_cat * 4 + size_t(d.emb_rows) * 5 * 4 + 64;
}

}  // namespace mlptc

void snv_mlp_tc_destroy(mural_snv_model* m) {
  if (!m->mlp_tc) return;
  mlptc::State* S = (mlptc::State*)m->mlp_tc;
  cudaFree(S->d_w);
  delete S;
  m->mlp_tc = nullptr;
}

// Builds the split-bf16 weight blob from the folded fp32 local-branch weights already resident in d_prep.
int snv_mlp_tc_prepare(mural_snv_model* m) {
  using namespace mlptc;
  snv_mlp_tc_destroy(m);
  Dims d{};
  d.n_cat = m->n_cat; d.K1 = m->k1; d.H1 = m->cfg.hidden1; d.H2 = m->cfg.hidden2; d.NC = m->cfg.n_class; d.emb_rows = m->emb_rows;
  auto ru16 = [](int x) { return (x + 15) & ~15; };
  d.K1p = ru16(d.K1 + 1); d.P1 = ru16(d.H1 + 1); d.P2 = ru16(d.H2 + 1);
  d.maxKp = d.K1p > d.P1 ? d.K1p : d.P1;
  if (d.P2 > d.maxKp) d.maxKp = d.P2;
  if (d.P1 > 256 || d.P2 > 128 || d.NC > 16 || d.n_cat > 2 * NPF) return 0;  // TMEM column plan / one N=16 head MMA
  const int Kp[3] = {d.K1p, d.P1, d.P2}, Np[3] = {d.P1, d.P2, 16};
  int off = 0;
  for (int l = 0; l < 3; ++l)
    for (int h = 0; h < 2; ++h) { d.w_off[l][h] = off; off += Kp[l] * Np[l] * 2; }
  d.w_bytes = off;
  if (smem_bytes(d) > 227 * 1024) return 0;  // fp32 kernel serves larger local branches
  // folded fp32 weights (LocalDev): W1t [K1][H1], b1, W2t [H1][H2] (BN folded), b2, W3t [H2][NC] (BN folded), b3
  const int K[3] = {d.K1, d.H1, d.H2}, N[3] = {d.H1, d.H2, d.NC};
  const float* dW[3] = {m->local.W1t, m->local.W2t, m->local.W3t};
  const float* dB[3] = {m->local.b1, m->local.b2, m->local.b3};
  std::vector<uint8_t> blob(size_t(d.w_bytes), 0);
  for (int l = 0; l < 3; ++l) {
    std::vector<float> W(size_t(K[l]) * N[l]), B(N[l]);
    CUDA_TRY(cudaMemcpy(W.data(), dW[l], W.size() * sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(B.data(), dB[l], B.size() * sizeof(float), cudaMemcpyDeviceToHost));
    __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(blob.data() + d.w_off[l][0]);
    __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(blob.data() + d.w_off[l][1]);
    for (int k = 0; k <= K[l]; ++k)
      for (int nn = 0; nn < N[l]; ++nn) {
        const float w = k < K[l] ? W[size_t(k) * N[l] + nn] : B[nn];  // row K = bias (activation column K is 1.0)
        const __nv_bfloat16 h = __float2bfloat16(w);
        const size_t idx = (size_t(k / 8) * Np[l] + nn) * 8 + (k % 8);   // [k/8][n][8]
        hi[idx] = h;
        lo[idx] = __float2bfloat16(w - __bfloat162float(h));
      }
  }
  State* S = new State();
  S->d = d;
  if (cudaMalloc((void**)&S->d_w, blob.size()) != cudaSuccess) {
    delete S;
    MURAL_FAIL("cudaMalloc of the local-branch tensor-core weight blob failed");
  }
  CUDA_TRY(cudaMemcpy(S->d_w, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  m->mlp_tc = S;
  return 0;
}

// returns -1 when the tensor-core local branch is not available for this model (caller uses the fp32 kernel)
int snv_local_launch_tc(mural_snv_model* m, const int32_t* cat32, const int64_t* cat64, int64_t ns, float* logits, int* err_flag,
                        cudaStream_t st) {
  using namespace mlptc;
  if (!m->mlp_tc) return -1;
  State* S = (State*)m->mlp_tc;
  const size_t smem = smem_bytes(S->d);
  static size_t configured = 0;
  if (smem > configured) {
    CUDA_TRY(cudaFuncSetAttribute(k_local_mlp_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int64_t tiles = cdiv(ns, TILE);
  const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
  LAUNCH(k_local_mlp_tc, grid, THREADS, smem, st, S->d, S->d_w, m->local.emb, cat32, cat64, ns, logits, err_flag);
  return 0;
}

}  // namespace mural
