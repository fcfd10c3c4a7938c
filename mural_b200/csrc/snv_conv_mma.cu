// Conv1d(32,32,3) with the ReLU/BatchNorm prologue on tensor cores at fp32-equivalent precision — the conv of the
// "fp32-equivalent" predict mode (MURAL_MODE_FP32, the MURAL_MODE_AUTO recompute) and of the training forward / dgrad
// (MuRaL/model/model_snv.py:794-812 ResBlock convs, :362-376 conv2/conv3; training loop MuRaL/training.py:424-427).
//
// Implicit GEMM per 128-row tile:  Y[128 x 32] = U'[128 x 96] * W[96 x 32],  k = tap*32 + ci, where U' row r holds the
// three input rows r-1, r, r+1 after u = a*act(x) + b, with the taps that fall outside the row's site zeroed (the reference
// pads AFTER the BatchNorm).  Both operands are split  x = hi + lo  into two bf16 values and multiplied as
// hi*hi + lo*hi + hi*lo  (mma.sync.m16n8k16, fp32 accumulate): 16 mantissa bits per operand, products exact in the fp32
// accumulator, i.e. ~1e-5 relative per product instead of bf16's 4e-3 — at three tensor-core MMAs instead of 96 FMAs per
// output.  Weights are split once per CTA into shared memory as packed B fragments; activations are staged as fp32 and
// split when the A fragments are built.  Same contract as k_conv<32> (snv_forward.cu): rows = n_sites*L flattened.
#include <cuda_bf16.h>

#include "snv_model.cuh"

namespace mural {
namespace cmma {

constexpr int C = 32, KS = 3, TILE = 128;
constexpr int US = 40;  // fp32 row stride of the activation tile: 8-byte fragment loads of 4 rows x 4 column pairs hit 32 distinct banks
constexpr int WS = 40;  // uint32 row stride of the packed weights: 4 k-pairs x 8 columns hit 32 distinct banks
constexpr int K2 = KS * C / 2;  // 48 packed k-pairs

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x, y) -> packed bf16 pair hi (x in the low half = lower k) and the packed remainder lo
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xFFFF0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - hx, y - hy);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// third level of the split (24 mantissa bits in all): x = hi + mid + lo
__device__ __forceinline__ void split3(float x, float y, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float rx = x - __uint_as_float(hi << 16), ry = y - __uint_as_float(hi & 0xFFFF0000u);
  split2(rx, ry, mid, lo);
}

// NS = 2: x = hi + lo, products hi*hi + lo*hi + hi*lo (~1e-5 per product; the predict modes).
// NS = 3: x = hi + mid + lo, the six products down to 2^-24 (fp32-level): the training forward / dgrad, whose BatchNorm
// backward and bias gradients are sums with heavy cancellation (a conv bias in front of a train-mode BatchNorm has a
// mathematically vanishing gradient), where 16 mantissa bits show up as 1e-2 relative errors.
#ifndef MURAL_CMMA_MINB
#define MURAL_CMMA_MINB 4
#endif
template <int NS>
__global__ void __launch_bounds__(128, NS == 2 ? MURAL_CMMA_MINB : 3) k_conv_mma(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ res1,
                                                                  const float* __restrict__ res2, int64_t rows, int L, ConvLayerDev P, int relu_out) {
  __shared__ __align__(16) float us[(TILE + 2) * US];
  __shared__ uint32_t wp[NS][K2 * WS];  // [hi|(mid)|lo][k-pair][co]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  // weights: Wt[tap][ci][co] fp32 -> packed (k, k+1) pairs per column, split
  for (int e = tid; e < K2 * C; e += 128) {
    const int k2 = e >> 5, co = e & 31;
    uint32_t sp[3];
    const float w0 = __ldg(P.Wt + (2 * k2) * C + co), w1 = __ldg(P.Wt + (2 * k2 + 1) * C + co);
    if (NS == 2) split2(w0, w1, sp[0], sp[1]);
    else split3(w0, w1, sp[0], sp[1], sp[2]);
#pragma unroll
    for (int s = 0; s < NS; ++s) wp[s][k2 * WS + co] = sp[s];
  }
  const int64_t n_tiles = (rows + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * TILE;
    __syncthreads();  // previous tile's fragment loads are done (also orders the weight staging on the first pass)
    // activation tile rows r0-1 .. r0+128 after act + BN affine; rows outside [0, rows) are zero
    for (int e = tid; e < (TILE + 2) * (C / 4); e += 128) {
      const int k = e >> 3, c4 = (e & 7) * 4;
      const int64_t r = r0 - 1 + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= 0 && r < rows) {
        v = __ldg(reinterpret_cast<const float4*>(in + r * C + c4));
        if (P.relu_in) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        const float4 a = __ldg(reinterpret_cast<const float4*>(P.a + c4)), b = __ldg(reinterpret_cast<const float4*>(P.b + c4));
        v.x = fmaf(v.x, a.x, b.x); v.y = fmaf(v.y, a.y, b.y); v.z = fmaf(v.z, a.z, b.z); v.w = fmaf(v.w, a.w, b.w);
      }
      *reinterpret_cast<float4*>(us + k * US + c4) = v;
    }
    __syncthreads();
    // this warp: rows 32*warp .. +31 as two m16 tiles; per row the taps that stay inside the row's site
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    bool v0[2][2], v2[2][2];  // [m-tile][row g / g+8]: tap 0 / tap 2 valid
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pin = int((r0 + warp * 32 + mt * 16 + g + 8 * h) % L);
        v0[mt][h] = pin >= 1;
        v2[mt][h] = pin <= L - 2;
      }
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {          // k16 steps: tap = ks / 2, channels (ks & 1) * 16 ..
      const int tap = ks >> 1, cb = (ks & 1) * 16;
      uint32_t as[NS][2][4];  // [split level][m-tile][fragment register]
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int rl = warp * 32 + mt * 16 + g + tap;  // smem row of (row g, this tap): tile row + tap (tile row 0 = r0 - 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) {              // a0: (g, 2t) a1: (g+8, 2t) a2: (g, 2t+8) a3: (g+8, 2t+8)
          const int h = q & 1;
          const float2 v = *reinterpret_cast<const float2*>(us + (rl + 8 * h) * US + cb + 2 * t + 8 * (q >> 1));
          const bool ok = tap == 1 || (tap == 0 ? v0[mt][h] : v2[mt][h]);
          if (NS == 2) split2(ok ? v.x : 0.f, ok ? v.y : 0.f, as[0][mt][q], as[1][mt][q]);
          else split3(ok ? v.x : 0.f, ok ? v.y : 0.f, as[0][mt][q], as[1][mt][q], as[NS - 1][mt][q]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int kb = ks * 8 + t, co = nt * 8 + g;
        uint32_t b0[NS], b1[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s) { b0[s] = wp[s][kb * WS + co]; b1[s] = wp[s][(kb + 4) * WS + co]; }
        // smallest products first: level pairs (i, j) with i + j descending
#pragma unroll
        for (int lev = NS - 1; lev >= 0; --lev)
#pragma unroll
          for (int i = 0; i <= lev; ++i)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_bf16(acc[mt][nt], as[i][mt], b0[lev - i], b1[lev - i]);
      }
    }
    // epilogue: c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g+8, ..) of each n8 tile
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t r = r0 + warp * 32 + mt * 16 + g + 8 * h;
        if (r >= rows) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int co = nt * 8 + 2 * t;
          const float2 b = __ldg(reinterpret_cast<const float2*>(P.bias + co));
          float2 y = make_float2(acc[mt][nt][2 * h] + b.x, acc[mt][nt][2 * h + 1] + b.y);
          if (res1) { const float2 q = __ldg(reinterpret_cast<const float2*>(res1 + r * C + co)); y.x += q.x; y.y += q.y; }
          if (res2) { const float2 q = __ldg(reinterpret_cast<const float2*>(res2 + r * C + co)); y.x += q.x; y.y += q.y; }
          if (relu_out) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); }
          *reinterpret_cast<float2*>(out + r * C + co) = y;
        }
      }
  }
}

}  // namespace cmma

static int m_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// C == 32, ks == 3 only (every shipped MuRaL-snv checkpoint); conv_any falls back to k_conv<C> otherwise
int conv32_mma(const float* in, float* out, const float* r1, const float* r2, int64_t n, int L, const ConvLayerDev& P, int relu_out,
               cudaStream_t st) {
  const int64_t rows = n * L;
  const int64_t tiles = cdiv(rows, cmma::TILE);
  const int64_t cap = int64_t(m_sm_count()) * (P.precise ? 3 : MURAL_CMMA_MINB);
  if (P.precise) LAUNCH_N("k_conv_mma<3>", cmma::k_conv_mma<3>, (unsigned)(tiles < cap ? tiles : cap), 128, 0, st, in, out, r1, r2, rows, L, P, relu_out);
  else LAUNCH_N("k_conv_mma<2>", cmma::k_conv_mma<2>, (unsigned)(tiles < cap ? tiles : cap), 128, 0, st, in, out, r1, r2, rows, L, P, relu_out);
  return 0;
}


// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co][ci][tap] += sum_r dy[r][co] * u[r + tap - 1][ci]  (u = a*act(x) + b inside the row's site, 0 outside),
// dbias[co] += sum_r dy[r][co]  — the weight gradient of one BN -> Conv1d(32,32,3) layer of the training backward
// (autograd of MuRaL/training.py:427 through model_snv.py:794-812) as a tensor-core GEMM
//   D[96 (tap,ci) x 32 co] += U'^T[96 x rows] * dY[rows x 32]
// reduced over the CTA's rows in fp32 accumulator registers (one atomicAdd per element and CTA at the end).  Operands are
// split hi + lo (bf16) once, when a 128-row tile is staged row-major into shared memory; ldmatrix.trans turns the row-major
// tiles into the K-paired fragments of mma.sync.m16n8k16, a tap is the same activation tile shifted by one row, and the
// per-site zero padding is a per-row mask on the dY fragments of taps 0 and 2.  Warp w: tap w % 3, rows 64*(w / 3) .. +63.
namespace wmma3 {
constexpr int C = 32, TILE = 128, RS = 40;  // RS: bf16 row stride (80 B: ldmatrix's 8 row addresses fall into 8 distinct 16-byte slots)
constexpr int THREADS = 192;

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}

__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, float4 v) {
  uint32_t h0, l0, h1, l1;
  cmma::split2(v.x, v.y, h0, l0);
  cmma::split2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(hi) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(lo) = make_uint2(l0, l1);
}

__global__ void __launch_bounds__(THREADS, 2) k_wgrad_mma(const float* __restrict__ x, const float* __restrict__ dy, int64_t rows, int L,
                                                          int relu, const float* __restrict__ a, const float* __restrict__ b,
                                                          float* __restrict__ G, int64_t w_off, int64_t b_off) {
  __shared__ __align__(16) __nv_bfloat16 uh[(TILE + 2) * RS], ul[(TILE + 2) * RS];  // rows r0-1 .. r0+128
  __shared__ __align__(16) __nv_bfloat16 dh[TILE * RS], dl[TILE * RS];
  __shared__ uint32_t mk[2][TILE / 2];  // [tap 0 | tap 2][row pair]: 0xFFFF per row whose tap stays inside its site
  __shared__ float sbias[C];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tap = warp % 3, kh = warp / 3;
  const int c4 = (tid & 7) * 4;  // THREADS % 8 == 0: a thread stages the same 4 channels every pass
  if (tid < C) sbias[tid] = 0.f;
  const float4 av = __ldg(reinterpret_cast<const float4*>(a + c4)), bv = __ldg(reinterpret_cast<const float4*>(b + c4));
  float acc[2][4][4];
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mi][nt][i] = 0.f;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t n_tiles = (rows + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * TILE;
    __syncthreads();
    for (int e = tid; e < (TILE + 2) * 8; e += THREADS) {
      const int k = e >> 3;
      const int64_t r = r0 - 1 + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= 0 && r < rows) {
        v = __ldg(reinterpret_cast<const float4*>(x + r * C + c4));
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        v.x = fmaf(v.x, av.x, bv.x); v.y = fmaf(v.y, av.y, bv.y); v.z = fmaf(v.z, av.z, bv.z); v.w = fmaf(v.w, av.w, bv.w);
      }
      store_split4(uh + k * RS + c4, ul + k * RS + c4, v);
    }
    for (int e = tid; e < TILE * 8; e += THREADS) {
      const int k = e >> 3;
      const int64_t r = r0 + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows) v = __ldg(reinterpret_cast<const float4*>(dy + r * C + c4));
      bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
      store_split4(dh + k * RS + c4, dl + k * RS + c4, v);
    }
    if (tid < TILE / 2) {
      uint32_t m0 = 0, m2 = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t r = r0 + 2 * tid + h;
        const int pin = int(r % L);
        if (r < rows && pin >= 1) m0 |= 0xFFFFu << (16 * h);
        if (r < rows && pin <= L - 2) m2 |= 0xFFFFu << (16 * h);
      }
      mk[0][tid] = m0;
      mk[1][tid] = m2;
    }
    __syncthreads();
    const int j = lane >> 3, i8 = lane & 7, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int k0 = kh * 64 + ks * 16;  // first row of this k16 step inside the tile
      // B (dY) fragments of the 4 n-tiles, masked for taps 0 / 2; matrices: j=0 rows k0.., n0 | j=1 rows k0+8.., n0 | j=2 rows k0.., n0+8 | j=3
      uint32_t bh[2][4], bl[2][4];
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        const int off = (k0 + (j & 1) * 8 + i8) * RS + np * 16 + (j >> 1) * 8;
        ldsm_x4_t(bh[np], dh + off);
        ldsm_x4_t(bl[np], dl + off);
      }
      if (tap != 1) {
        const uint32_t ma = mk[tap >> 1][(k0 >> 1) + t], mb = mk[tap >> 1][(k0 >> 1) + 4 + t];
#pragma unroll
        for (int np = 0; np < 2; ++np) {
          bh[np][0] &= ma; bh[np][1] &= mb; bh[np][2] &= ma; bh[np][3] &= mb;
          bl[np][0] &= ma; bl[np][1] &= mb; bl[np][2] &= ma; bl[np][3] &= mb;
        }
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        // A (U'^T) fragment of channels 16*mi ..: matrices j=0 rows k0.., ci0 | j=1 rows k0.., ci0+8 | j=2 rows k0+8.., ci0 | j=3
        uint32_t ah[4], al[4];
        const int off = (k0 + tap + (j >> 1) * 8 + i8) * RS + mi * 16 + (j & 1) * 8;   // smem row = tile row + tap (row 0 = r0 - 1)
        ldsm_x4_t(ah, uh + off);
        ldsm_x4_t(al, ul + off);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t h0 = bh[nt >> 1][(nt & 1) * 2], h1 = bh[nt >> 1][(nt & 1) * 2 + 1];
          const uint32_t l0 = bl[nt >> 1][(nt & 1) * 2], l1 = bl[nt >> 1][(nt & 1) * 2 + 1];
          cmma::mma_bf16(acc[mi][nt], al, h0, h1);
          cmma::mma_bf16(acc[mi][nt], ah, l0, l1);
          cmma::mma_bf16(acc[mi][nt], ah, h0, h1);
        }
      }
    }
  }
  // ---- flush: D element (m = ci, n = co) of tap `tap`
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ci = mi * 16 + g + 8 * (i >> 1), co = nt * 8 + 2 * t + (i & 1);
        atomicAdd(G + w_off + (int64_t(co) * C + ci) * 3 + tap, acc[mi][nt][i]);
      }
  atomicAdd(&sbias[c4], bsum.x); atomicAdd(&sbias[c4 + 1], bsum.y); atomicAdd(&sbias[c4 + 2], bsum.z); atomicAdd(&sbias[c4 + 3], bsum.w);
  __syncthreads();
  if (tid < C) atomicAdd(G + b_off + tid, sbias[tid]);
}
}  // namespace wmma3

int wgrad32_mma(const float* x, const float* dy, int64_t rows, int L, int relu, const float* a, const float* b, float* G, int64_t w_off,
                int64_t b_off, cudaStream_t st) {
  const int64_t tiles = cdiv(rows, wmma3::TILE);
  const int64_t cap = int64_t(m_sm_count()) * 2;
  LAUNCH_N("k_wgrad_mma", wmma3::k_wgrad_mma, (unsigned)(tiles < cap ? tiles : cap), wmma3::THREADS, 0, st, x, dy, rows, L, relu, a, b, G,
           w_off, b_off);
  return 0;
}

}  // namespace mural
