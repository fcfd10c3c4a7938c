// One CNN branch of Network2 (eval) between the pooled stem output and the conv3 output as ONE kernel at fp32-equivalent
// precision: the 10 BN -> Conv1d(32,32,3) layers (two ResBlock pairs with their skips, conv2, conv3) and the two max-pools of
// MuRaL/model/model_snv.py:477-488 / 499-510, with every activation of a site resident in shared memory.
//
// The per-layer form of this path (snv_forward_fp32: k_conv_mma + k_pool, 12 launches per branch) moves each layer's fp32
// activations through HBM and sits at that formulation's bandwidth roofline; its main job in the product is the
// exception windows of MURAL_MODE_AUTO, a few thousand sites per call that run beside the bf16 pass — there the 24 small,
// HBM-bound launches are what the recompute costs.  Here a CTA takes one site through the whole chain: the arithmetic is the
// one of k_conv_mma<2> (snv_conv_mma.cu: operands split x = hi + lo in bf16, products lo*hi + hi*lo + hi*hi on
// mma.sync.m16n8k16 with fp32 accumulation, same order), so the result is bit-identical to the per-layer path
// (tests/test_gpu_snv_forward.py::test_site_chain_equals_per_layer_path); only x0 is read from and conv3's output written to
// global memory.  Zero padding per site is two zero rows around the staged (BN-applied) tile, as the reference pads after the
// BatchNorm.
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>

#include "snv_model.cuh"

namespace mural {
namespace chain {

constexpr int C = 32, THREADS = 256, WARPS = THREADS / 32;
constexpr int UW = 20;          // word stride of a staged row (16 packed bf16 pairs + 4): the 8 rows x 4 words of a fragment load hit 32 distinct banks
constexpr int WS = 40;          // uint32 row stride of the packed weights
constexpr int K2 = 3 * C / 2;   // 48 packed k-pairs
constexpr int NLAYER = 10;

constexpr int WP_WORDS = 2 * K2 * WS;   // packed weights of one layer: [hi | lo][k-pair][co (row stride WS)]
constexpr int WREG = (WP_WORDS / 4 + THREADS - 1) / THREADS;   // 16-byte words of a layer's weights per thread

struct Args {
  ConvLayerDev layer[NLAYER];  // rb1[0..3], conv2, rb2[0..3], conv3
  const uint32_t* wsplit;      // [NLAYER][WP_WORDS] split packed weights of these layers (k_chain_weights)
  const float* x0;             // [n][L1][C]  pool-1 output of the stem
  float* h;                    // [n][L3][C]  conv3 output after ReLU
  int64_t n;
  int L1, L2, L3;
  int pool2[3], pool3[3];      // (kernel, stride, pad)
  int LP;                      // rows per activation buffer (>= L1)
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xFFFF0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - hx, y - hy);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// Y[L][C] = conv(BN(act(X))) + bias (+ R1) (+ R2), optionally ReLU; X, R1, R2 in shared memory (row stride C), Y in shared
// memory (row stride C) or global memory.  The staged tile holds every (BN-applied) activation pair already split into packed
// bf16 hi / lo words — one split per element instead of one per tap and column tile — so an A fragment is 8 word loads; a warp's
// unit of work is (16-row tile, pair of 8-column tiles), which keeps all 8 warps busy on the short stage-2 / stage-3 rows too.
// Ends with a CTA barrier.
__device__ __forceinline__ void conv_layer(const float* X, float* Y, const float* R1, const float* R2, int L, const ConvLayerDev& P,
                                           const uint32_t* __restrict__ wnext, uint4 (&wreg)[WREG], int relu_out, uint32_t* us_hi,
                                           uint32_t* us_lo, uint32_t* wp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int MT = (L + 15) >> 4;
  // weights of this layer: fetched into registers while the previous layer's MMAs ran (split packed pairs prepared once per
  // model load, k_chain_weights)
#pragma unroll
  for (int i = 0; i < WREG; ++i)
    if (tid + i * THREADS < WP_WORDS / 4) reinterpret_cast<uint4*>(wp)[tid + i * THREADS] = wreg[i];
  // staged tile: row k holds site row k - 1 after act + BN affine; row 0 and rows > L are the zero padding
  for (int e = tid; e < (MT * 16 + 2) * (C / 4); e += THREADS) {
    const int k = e >> 3, c4 = (e & 7) * 4;
    const int r = k - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0 && r < L) {
      v = *reinterpret_cast<const float4*>(X + r * C + c4);
      if (P.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      const float4 a = __ldg(reinterpret_cast<const float4*>(P.a + c4)), b = __ldg(reinterpret_cast<const float4*>(P.b + c4));
      v.x = fmaf(v.x, a.x, b.x); v.y = fmaf(v.y, a.y, b.y); v.z = fmaf(v.z, a.z, b.z); v.w = fmaf(v.w, a.w, b.w);
    }
    uint2 hi, lo;
    split2(v.x, v.y, hi.x, lo.x);
    split2(v.z, v.w, hi.y, lo.y);
    *reinterpret_cast<uint2*>(us_hi + k * UW + (c4 >> 1)) = hi;
    *reinterpret_cast<uint2*>(us_lo + k * UW + (c4 >> 1)) = lo;
  }
  __syncthreads();
  // the next layer's weights travel from L2 during this layer's MMAs
#pragma unroll
  for (int i = 0; i < WREG; ++i)
    if (tid + i * THREADS < WP_WORDS / 4) wreg[i] = __ldg(reinterpret_cast<const uint4*>(wnext) + tid + i * THREADS);
  for (int u = warp; u < 2 * MT; u += WARPS) {
    const int mt = u >> 1, np = u & 1;
    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {  // k16 steps: tap = ks / 2, channels (ks & 1) * 16 ..
      const int tap = ks >> 1, kw = (ks & 1) * 8;
      uint32_t ah[4], al[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {   // a0: (g, 2t) a1: (g+8, 2t) a2: (g, 2t+8) a3: (g+8, 2t+8)
        const int idx = (mt * 16 + g + tap + 8 * (q & 1)) * UW + kw + t + 4 * (q >> 1);
        ah[q] = us_hi[idx];
        al[q] = us_lo[idx];
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int kb = ks * 8 + t, co = (np * 2 + j) * 8 + g;
        const uint32_t bh0 = wp[kb * WS + co], bh1 = wp[(kb + 4) * WS + co];
        const uint32_t bl0 = wp[K2 * WS + kb * WS + co], bl1 = wp[K2 * WS + (kb + 4) * WS + co];
        mma_bf16(acc[j], ah, bl0, bl1);   // smallest products first, the order of k_conv_mma<2>
        mma_bf16(acc[j], al, bh0, bh1);
        mma_bf16(acc[j], ah, bh0, bh1);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = mt * 16 + g + 8 * h;
      if (r >= L) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int co = (np * 2 + j) * 8 + 2 * t;
        const float2 b = __ldg(reinterpret_cast<const float2*>(P.bias + co));
        float2 y = make_float2(acc[j][2 * h] + b.x, acc[j][2 * h + 1] + b.y);
        if (R1) { const float2 q = *reinterpret_cast<const float2*>(R1 + r * C + co); y.x += q.x; y.y += q.y; }
        if (R2) { const float2 q = *reinterpret_cast<const float2*>(R2 + r * C + co); y.x += q.x; y.y += q.y; }
        if (relu_out) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); }
        *reinterpret_cast<float2*>(Y + r * C + co) = y;
      }
    }
  }
  __syncthreads();
}

// MaxPool1d with -inf padding (k_pool of snv_forward.cu)
__device__ __forceinline__ void pool_layer(const float* X, float* Y, int Lin, int Lout, const int* p) {
  for (int e = threadIdx.x; e < Lout * C; e += THREADS) {
    const int j = e >> 5, c = e & 31;
    int lo = j * p[1] - p[2], hi = lo + p[0];
    lo = lo < 0 ? 0 : lo;
    hi = hi > Lin ? Lin : hi;
    float mx = -FLT_MAX;
    for (int q = lo; q < hi; ++q) mx = fmaxf(mx, X[q * C + c]);
    Y[e] = mx;
  }
  __syncthreads();
}

// Wt[tap][ci][co] fp32 of one layer per block -> packed (k, k+1) pairs per column, split into hi | lo (the staging loop of k_conv_mma)
struct WeightPtrs { const float* Wt[2 * NLAYER]; };
__global__ void __launch_bounds__(256) k_chain_weights(WeightPtrs w, uint32_t* __restrict__ out) {
  const float* Wt = w.Wt[blockIdx.x];
  uint32_t* wp = out + size_t(blockIdx.x) * WP_WORDS;
  for (int e = threadIdx.x; e < K2 * WS; e += 256) {
    const int k2 = e / WS, co = e % WS;
    uint32_t hi = 0, lo = 0;
    if (co < C) split2(__ldg(Wt + (2 * k2) * C + co), __ldg(Wt + (2 * k2 + 1) * C + co), hi, lo);
    wp[k2 * WS + co] = hi;
    wp[K2 * WS + k2 * WS + co] = lo;
  }
}

__global__ void __launch_bounds__(THREADS) k_site_chain(Args a) {
  extern __shared__ __align__(16) float sm[];
  float* A = sm;                 // x0, later the pool-2 output / the stage-2 output
  float* B = A + a.LP * C;
  float* Cb = B + a.LP * C;
  float* D = Cb + a.LP * C;
  const int us_words = (((a.L1 + 15) >> 4) * 16 + 2) * UW;
  uint32_t* us_hi = reinterpret_cast<uint32_t*>(D + a.LP * C);
  uint32_t* us_lo = us_hi + us_words;
  uint32_t* wp = us_lo + us_words;
  uint4 wreg[WREG];   // weights of the next conv layer (layer 0 here; every conv_layer call fetches its successor's)
#pragma unroll
  for (int i = 0; i < WREG; ++i)
    if (threadIdx.x + i * THREADS < WP_WORDS / 4) wreg[i] = __ldg(reinterpret_cast<const uint4*>(a.wsplit) + threadIdx.x + i * THREADS);
  for (int64_t site = blockIdx.x; site < a.n; site += gridDim.x) {
    const float4* x0 = reinterpret_cast<const float4*>(a.x0 + site * a.L1 * C);
    for (int e = threadIdx.x; e < a.L1 * (C / 4); e += THREADS) reinterpret_cast<float4*>(A)[e] = __ldg(x0 + e);
    __syncthreads();
    // stage 1 at length L1: two ResBlocks + outer skip (model_snv.py:477-479 / 499-501)
    conv_layer(A, B, nullptr, nullptr, a.L1, a.layer[0], a.wsplit + 1 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);
    conv_layer(B, Cb, A, nullptr, a.L1, a.layer[1], a.wsplit + 2 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);   // y1 = x0 + f(x0)
    conv_layer(Cb, B, nullptr, nullptr, a.L1, a.layer[2], a.wsplit + 3 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);
    conv_layer(B, D, Cb, A, a.L1, a.layer[3], a.wsplit + 4 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);         // y1 + f(y1) + x0
    // pool 2, conv2, stage 2 (:480-485 / 502-507)
    pool_layer(D, A, a.L1, a.L2, a.pool2);
    conv_layer(A, B, nullptr, nullptr, a.L2, a.layer[4], a.wsplit + 5 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);   // jump2
    conv_layer(B, Cb, nullptr, nullptr, a.L2, a.layer[5], a.wsplit + 6 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);
    conv_layer(Cb, D, B, nullptr, a.L2, a.layer[6], a.wsplit + 7 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);
    conv_layer(D, Cb, nullptr, nullptr, a.L2, a.layer[7], a.wsplit + 8 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);
    conv_layer(Cb, A, D, B, a.L2, a.layer[8], a.wsplit + 9 * WP_WORDS, wreg, 0, us_hi, us_lo, wp);
    // pool 3, conv3 + ReLU (:486-488 / 508-510)
    pool_layer(A, Cb, a.L2, a.L3, a.pool3);
    conv_layer(Cb, a.h + site * a.L3 * C, nullptr, nullptr, a.L3, a.layer[9], a.wsplit + 0 * WP_WORDS, wreg, 1, us_hi, us_lo, wp);
  }
}

}  // namespace chain

// 0: launched; -1: shape not served by the fused kernel (caller keeps the per-layer path)
int snv_site_chain_launch(mural_snv_model* m, int br, const float* x0, float* h, int64_t ns, cudaStream_t st) {
  using namespace chain;
  if (getenv("MURAL_NO_SITE_CHAIN") != nullptr) return -1;   // parity switch (read per call: tests toggle it)
  if (m->cfg.channels != C || m->cfg.kernel_size != 3) return -1;
  const BranchDev& B = m->br[br];
  Args a{};
  for (int i = 0; i < 4; ++i) { a.layer[i] = B.rb1[i]; a.layer[5 + i] = B.rb2[i]; }
  a.layer[4] = B.conv2;
  a.layer[9] = B.conv3;
  for (int i = 0; i < NLAYER; ++i)
    if (a.layer[i].ks != 3 || a.layer[i].precise) return -1;
  if (!m->chain_ready) {   // split weights of both branches, once per weight load
    if (!m->d_chain) CUDA_TRY(cudaMalloc((void**)&m->d_chain, sizeof(uint32_t) * 2 * NLAYER * WP_WORDS));
    WeightPtrs w{};
    for (int b = 0; b < 2; ++b) {
      const BranchDev& Bb = m->br[b];
      for (int i = 0; i < 4; ++i) { w.Wt[b * NLAYER + i] = Bb.rb1[i].Wt; w.Wt[b * NLAYER + 5 + i] = Bb.rb2[i].Wt; }
      w.Wt[b * NLAYER + 4] = Bb.conv2.Wt;
      w.Wt[b * NLAYER + 9] = Bb.conv3.Wt;
    }
    LAUNCH(k_chain_weights, 2 * NLAYER, 256, 0, st, w, m->d_chain);
    m->chain_ready = true;
  }
  a.wsplit = m->d_chain + size_t(br) * NLAYER * WP_WORDS;
  a.x0 = x0; a.h = h; a.n = ns;
  a.L1 = B.L1; a.L2 = B.L2; a.L3 = B.L3;
  for (int i = 0; i < 3; ++i) { a.pool2[i] = B.pool[1][i]; a.pool3[i] = B.pool[2][i]; }
  a.LP = B.L1;
  const size_t smem = sizeof(float) * (size_t(4) * a.LP * C + size_t(((a.L1 + 15) / 16) * 16 + 2) * 2 * UW) + sizeof(uint32_t) * WP_WORDS;
  if (smem > 200 * 1024) return -1;   // windows beyond ~2 Kb radius: per-layer path
  static bool conf = false;
  if (!conf) {
    CUDA_TRY(cudaFuncSetAttribute(k_site_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // two CTAs of the 1 Kb window (107 KB each) per SM need the largest shared-memory carve-out
    CUDA_TRY(cudaFuncSetAttribute(k_site_chain, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    conf = true;
  }
  LAUNCH(k_site_chain, (unsigned)(ns < (int64_t(1) << 20) ? ns : (int64_t(1) << 20)), THREADS, smem, st, a);
  return 0;
}

}  // namespace mural
