// MuRaL-snv Network2 training step on the GPU: train-mode forward (batch-statistic BatchNorm, dropout),
// full backward, gradient-norm clipping and Adam / AdamW(amsgrad) / SGD(nesterov) updates.
// Reference loop body: MuRaL/training.py:404-452 (forward, CE(sum), backward, clip_grad_norm_(10), optimizer.step).
//
// fp32 CUDA-core kernels, one launch per layer (this is the correctness-first path of round 1; the conv
// forward reuses k_conv of snv_forward.cu).  Parameters and gradients live in ONE flat fp32 buffer each, in the
// layout of mural_snv_model_tensor(): trainable tensors first, BatchNorm running statistics after, so the
// data-parallel exchange is a single all-reduce over grads[0 : n_trainable) and the optimizer is one kernel.
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "snv_model.cuh"

namespace mural {
namespace train {

constexpr float BN_EPS = 1e-5f;
constexpr float BN_MOM = 0.1f;

// ---------------------------------------------------------------------------------------------- rng (dropout)
__device__ __forceinline__ uint32_t hash32(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return uint32_t((x ^ (x >> 31)) >> 32);
}
// keep-scale of element idx of dropout layer `layer` at (seed, step): 0 or 1/(1-p)
__device__ __forceinline__ float drop_scale(float p, uint64_t seed, uint32_t step, uint32_t layer, uint64_t idx) {
  if (p <= 0.f) return 1.f;
  const uint32_t u = hash32(seed ^ (uint64_t(step) << 40) ^ (uint64_t(layer) << 34) ^ idx);
  return (u >= uint32_t(double(p) * 4294967296.0)) ? 1.f / (1.f - p) : 0.f;
}

// ---------------------------------------------------------------------------------------------- weight prep
struct ConvDesc {  // one BN -> Conv1d(Cin = C, Cout = C, ks) layer
  int64_t w, b, g, be, rm, rv;  // offsets into the flat blob
  int ks, relu_in;
};
constexpr int N_CONV = 20;  // per model: 2 branches x (4 + 1 + 4 + 1)

__global__ void k_prep_conv(const float* __restrict__ P, const ConvDesc* __restrict__ d, int C, float* __restrict__ Wt,
                            float* __restrict__ Wf, int64_t stride) {
  const ConvDesc L = d[blockIdx.x];
  float* wt = Wt + blockIdx.x * stride;
  float* wf = Wf + blockIdx.x * stride;
  for (int e = threadIdx.x; e < L.ks * C * C; e += blockDim.x) {
    const int t = e / (C * C), ci = (e / C) % C, co = e % C;
    const float v = P[L.w + (int64_t(co) * C + ci) * L.ks + t];
    wt[(t * C + ci) * C + co] = v;                   // forward: [tap][ci][co]
    wf[((L.ks - 1 - t) * C + co) * C + ci] = v;      // dgrad: conv over dy with flipped taps, [tap'][co][ci]
  }
}

// ---------------------------------------------------------------------------------------------- per-channel stats
// sums of act(x) and act(x)^2 over all rows (train-mode BatchNorm1d over (B, L)); double accumulation.
// A thread owns 4 channels (one 16-byte load per row) and walks rows two at a time: at batch 4096 the pass is HBM-bound and the
// scalar one-load-in-flight form reached 25 % of the bandwidth (8 KB in flight per SM against the ~28 KB Little's law asks for).
__global__ void k_stats(const float* __restrict__ x, int64_t rows, int C, int relu, double* __restrict__ out) {
  extern __shared__ double sh[];  // [2][4 * blockDim]
  const int C4 = C >> 2, c4 = threadIdx.x % C4, lane_r = threadIdx.x / C4, rpb = blockDim.x / C4;
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  const int64_t stride = int64_t(gridDim.x) * rpb;
  auto acc = [&](float4 v) {
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    s[0] += v.x; q[0] += double(v.x) * v.x;
    s[1] += v.y; q[1] += double(v.y) * v.y;
    s[2] += v.z; q[2] += double(v.z) * v.z;
    s[3] += v.w; q[3] += double(v.w) * v.w;
  };
  int64_t r = int64_t(blockIdx.x) * rpb + lane_r;
  if (lane_r < rpb) {
    for (; r + stride < rows; r += 2 * stride) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + r * C) + c4);
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (r + stride) * C) + c4);
      acc(v0);
      acc(v1);
    }
    if (r < rows) acc(__ldg(reinterpret_cast<const float4*>(x + r * C) + c4));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { sh[threadIdx.x * 4 + k] = s[k]; sh[4 * blockDim.x + threadIdx.x * 4 + k] = q[k]; }
  __syncthreads();
  if (threadIdx.x < C) {   // channel c = 4 * c4' + k lives at slot (lane_r * C4 + c4') * 4 + k = lane_r * C + c
    double ts = 0, tq = 0;
    for (int k = 0; k < rpb; ++k) { ts += sh[k * C + threadIdx.x]; tq += sh[4 * blockDim.x + k * C + threadIdx.x]; }
    atomicAdd(out + threadIdx.x, ts);
    atomicAdd(out + C + threadIdx.x, tq);
  }
}

// stats -> affine (a = gamma*invstd, b = beta - mean*a), saved mean/invstd, running-stat update (momentum 0.1,
// unbiased variance), like nn.BatchNorm1d in training mode
__global__ void k_bn_finalize(const double* __restrict__ st, double N, int C, float* __restrict__ P, int64_t g, int64_t be,
                              int64_t rm, int64_t rv, float* __restrict__ a, float* __restrict__ b, float* __restrict__ mu,
                              float* __restrict__ invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = st[c] / N;
  double var = st[C + c] / N - mean * mean;
  if (var < 0) var = 0;
  const float is = float(1.0 / sqrt(var + double(BN_EPS)));
  mu[c] = float(mean);
  invstd[c] = is;
  a[c] = P[g + c] * is;
  b[c] = P[be + c] - float(mean) * P[g + c] * is;
  const double unb = N > 1 ? var * N / (N - 1) : var;
  P[rm + c] = (1.f - BN_MOM) * P[rm + c] + BN_MOM * float(mean);
  P[rv + c] = (1.f - BN_MOM) * P[rv + c] + BN_MOM * float(unb);
}

// ---------------------------------------------------------------------------------------------- conv backward
// s1 = sum du, s2 = sum du * zhat with zhat = (act(x) - mu) * invstd   (4 channels per thread, two rows in flight: see k_stats)
__global__ void k_bn_bwd_reduce(const float* __restrict__ du, const float* __restrict__ x, int64_t rows, int C, int relu,
                                const float* __restrict__ mu, const float* __restrict__ invstd, double* __restrict__ out) {
  extern __shared__ double sh[];
  const int C4 = C >> 2, c4 = threadIdx.x % C4, lane_r = threadIdx.x / C4, rpb = blockDim.x / C4;
  const float4 m = *reinterpret_cast<const float4*>(mu + 4 * c4), is = *reinterpret_cast<const float4*>(invstd + 4 * c4);
  double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  const int64_t stride = int64_t(gridDim.x) * rpb;
  auto acc = [&](float4 v, const float4 d) {
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    s[0] += d.x; q[0] += double(d.x) * ((v.x - m.x) * is.x);
    s[1] += d.y; q[1] += double(d.y) * ((v.y - m.y) * is.y);
    s[2] += d.z; q[2] += double(d.z) * ((v.z - m.z) * is.z);
    s[3] += d.w; q[3] += double(d.w) * ((v.w - m.w) * is.w);
  };
  int64_t r = int64_t(blockIdx.x) * rpb + lane_r;
  if (lane_r < rpb) {
    for (; r + stride < rows; r += 2 * stride) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + r * C) + c4), d0 = __ldg(reinterpret_cast<const float4*>(du + r * C) + c4);
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (r + stride) * C) + c4);
      const float4 d1 = __ldg(reinterpret_cast<const float4*>(du + (r + stride) * C) + c4);
      acc(v0, d0);
      acc(v1, d1);
    }
    if (r < rows) acc(__ldg(reinterpret_cast<const float4*>(x + r * C) + c4), __ldg(reinterpret_cast<const float4*>(du + r * C) + c4));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { sh[threadIdx.x * 4 + k] = s[k]; sh[4 * blockDim.x + threadIdx.x * 4 + k] = q[k]; }
  __syncthreads();
  if (threadIdx.x < C) {
    double ts = 0, tq = 0;
    for (int k = 0; k < rpb; ++k) { ts += sh[k * C + threadIdx.x]; tq += sh[4 * blockDim.x + k * C + threadIdx.x]; }
    atomicAdd(out + threadIdx.x, ts);
    atomicAdd(out + C + threadIdx.x, tq);
  }
}

// dx = relu'(x) * a * (du - s1/N - zhat*s2/N); out = dx (+ add1) (+ add2); block 0 also emits dgamma, dbeta.  One thread = 4 channels
// of a row (16-byte loads / stores).
__global__ void k_bn_bwd_apply(const float* __restrict__ du, const float* __restrict__ x, int64_t rows, int C, int relu,
                               const float* __restrict__ mu, const float* __restrict__ invstd, const float* __restrict__ a,
                               const double* __restrict__ st, double N, const float* __restrict__ add1,
                               const float* __restrict__ add2, float* __restrict__ out, float* __restrict__ G, int64_t g_off,
                               int64_t be_off) {
  const int64_t e4 = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < C) {
    atomicAdd(G + be_off + threadIdx.x, float(st[threadIdx.x]));
    atomicAdd(G + g_off + threadIdx.x, float(st[C + threadIdx.x]));
  }
  const int C4 = C >> 2;
  if (e4 >= rows * C4) return;
  const int c = int(e4 % C4) * 4;
  const float4 xv = __ldg(reinterpret_cast<const float4*>(x) + e4), d = __ldg(reinterpret_cast<const float4*>(du) + e4);
  const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {d.x, d.y, d.z, d.w};
  float o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float z = relu ? fmaxf(xs[k], 0.f) : xs[k];
    const float zh = (z - mu[c + k]) * invstd[c + k];
    float dz = a[c + k] * (ds[k] - float(st[c + k] / N) - zh * float(st[C + c + k] / N));
    if (relu && xs[k] <= 0.f) dz = 0.f;
    o[k] = dz;
  }
  if (add1) { const float4 t = __ldg(reinterpret_cast<const float4*>(add1) + e4); o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w; }
  if (add2) { const float4 t = __ldg(reinterpret_cast<const float4*>(add2) + e4); o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w; }
  reinterpret_cast<float4*>(out)[e4] = make_float4(o[0], o[1], o[2], o[3]);
}

// dW[co][ci][t] += sum_r dy[r][co] * u[r+t-pad][ci] (u = a*act(x)+b inside the site, 0 outside), dbias[co] += sum_r dy
template <int C>
__global__ void __launch_bounds__(256) k_wgrad(const float* __restrict__ x, const float* __restrict__ dy, int64_t rows, int L,
                                               int ks, int relu, const float* __restrict__ a, const float* __restrict__ b,
                                               float* __restrict__ G, int64_t w_off, int64_t b_off) {
  // dW[co][ci][t] = sum_r dy[r][co] * u[r + t - half][ci] over rows whose tap stays inside the site (u = BN(act(x))).
  // A thread owns a 4(ci) x 4(co) block of one tap: per row two LDS.128 feed 16 FMAs.  Tiles of TR rows are staged in
  // shared memory; the per-(row, tap) validity is folded into a masked copy of dy per tap.
  constexpr int TR = 64;
  constexpr int CS = C + 4;  // row stride (floats), keeps float4 alignment and spreads banks
  extern __shared__ __align__(16) float shf[];
  float* us = shf;                           // [(TR + ks - 1)][CS]
  float* ds = us + (TR + ks - 1) * CS;       // [TR][CS]
  int* ps = reinterpret_cast<int*>(ds + TR * CS);  // [TR] position of the row inside its site (very negative: beyond rows)
  const int half = ks / 2, tid = threadIdx.x;
  constexpr int NB = (C / 4) * (C / 4);      // 4x4 blocks per tap
  const int n_blocks = ks * NB;
  constexpr int MAXB = (7 * NB + 255) / 256;  // blocks per thread bound (ks <= 7)
  float acc[MAXB][16];
#pragma unroll
  for (int k = 0; k < MAXB; ++k)
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[k][e] = 0.f;
  float accb = 0.f;
  const int64_t n_tiles = (rows + TR - 1) / TR;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * TR;
    __syncthreads();
    for (int e = tid; e < (TR + ks - 1) * C; e += 256) {
      const int k = e / C, ci = e - k * C;
      const int64_t r = r0 - half + k;
      float v = 0.f;
      if (r >= 0 && r < rows) {
        v = x[r * C + ci];
        if (relu) v = fmaxf(v, 0.f);
        v = fmaf(v, a[ci], b[ci]);
      }
      us[k * CS + ci] = v;
    }
    for (int e = tid; e < TR * C; e += 256) {
      const int k = e / C, co = e - k * C;
      const int64_t r = r0 + k;
      ds[k * CS + co] = r < rows ? dy[r * C + co] : 0.f;
    }
    if (tid < TR) ps[tid] = (r0 + tid < rows) ? int((r0 + tid) % L) : -1000000;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < MAXB; ++k) {
      const int blk = tid + k * 256;
      if (blk < n_blocks) {
        const int t = blk / NB, ci4 = ((blk % NB) / (C / 4)) * 4, co4 = (blk % (C / 4)) * 4;
#pragma unroll 4
        for (int r = 0; r < TR; ++r) {
          const int q = ps[r] + t - half;
          if (q < 0 || q >= L) continue;  // tap falls on the zero padding of this row's site (warp-uniform: same t per warp mostly)
          const float4 d4 = *reinterpret_cast<const float4*>(ds + r * CS + co4);
          const float4 u4 = *reinterpret_cast<const float4*>(us + (r + t) * CS + ci4);
          const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, uv[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 0; o < 4; ++o) acc[k][i * 4 + o] = fmaf(dv[o], uv[i], acc[k][i * 4 + o]);
        }
      }
    }
    if (tid < C) {
      float s = 0.f;
      for (int r = 0; r < TR; ++r) s += ds[r * CS + tid];
      accb += s;
    }
  }
#pragma unroll
  for (int k = 0; k < MAXB; ++k) {
    const int blk = tid + k * 256;
    if (blk < n_blocks) {
      const int t = blk / NB, ci4 = ((blk % NB) / (C / 4)) * 4, co4 = (blk % (C / 4)) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int o = 0; o < 4; ++o) atomicAdd(G + w_off + (int64_t(co4 + o) * C + ci4 + i) * ks + t, acc[k][i * 4 + o]);
    }
  }
  if (tid < C) atomicAdd(G + b_off + tid, accb);
}

// ---------------------------------------------------------------------------------------------- pooling
__global__ void k_pool_fwd_idx(const float* __restrict__ in, float* __restrict__ out, int32_t* __restrict__ idx, int64_t n,
                               int Lin, int Lout, int C, int pk, int ps, int pp) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n * Lout * C) return;
  const int c = int(e % C);
  const int64_t sj = e / C;
  const int j = int(sj % Lout);
  const int64_t site = sj / Lout;
  int lo = j * ps - pp, hi = lo + pk;
  lo = lo < 0 ? 0 : lo;
  hi = hi > Lin ? Lin : hi;
  float mx = -FLT_MAX;
  int am = lo;
  for (int p = lo; p < hi; ++p) {
    const float v = in[(site * Lin + p) * C + c];
    if (v > mx) { mx = v; am = p; }
  }
  out[e] = mx;
  idx[e] = am;
}
// all MuRaL pools have stride == kernel: every input position belongs to exactly one window
__global__ void k_pool_bwd(const float* __restrict__ gy, const int32_t* __restrict__ idx, float* __restrict__ gx, int64_t n, int Lin,
                           int Lout, int C, int ps, int pp) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n * Lin * C) return;
  const int c = int(e % C);
  const int64_t sp = e / C;
  const int p = int(sp % Lin);
  const int64_t site = sp / Lin;
  const int j = (p + pp) / ps;
  float g = 0.f;
  if (j < Lout) {
    const int64_t o = (site * Lout + j) * C + c;
    if (idx[o] == p) g = gy[o];
  }
  gx[e] = g;
}
// global max over the L3 rows (torch.max(dim=2)), input already ReLU'd
__global__ void k_gmax_fwd(const float* __restrict__ h, float* __restrict__ out, int32_t* __restrict__ idx, int64_t n, int L3, int C) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n * C) return;
  const int c = int(e % C);
  const int64_t site = e / C;
  float mx = -FLT_MAX;
  int am = 0;
  for (int p = 0; p < L3; ++p) {
    const float v = h[(site * L3 + p) * C + c];
    if (v > mx) { mx = v; am = p; }
  }
  out[e] = mx;
  idx[e] = am;
}
// d(conv3 pre-ReLU output) from d(gmax): routed to the arg-max row, masked by the ReLU
__global__ void k_gmax_bwd(const float* __restrict__ gy, const int32_t* __restrict__ idx, const float* __restrict__ h,
                           float* __restrict__ gx, int64_t n, int L3, int C) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n * L3 * C) return;
  const int c = int(e % C);
  const int64_t sp = e / C;
  const int p = int(sp % L3);
  const int64_t site = sp / L3;
  const int64_t o = site * C + c;
  gx[e] = (idx[o] == p && h[e] > 0.f) ? gy[o] : 0.f;
}

// ---------------------------------------------------------------------------------------------- dense [n][F] layers
// y = x W^T + b (optionally ReLU); W is [N][K] row-major (nn.Linear layout inside the flat blob)
__global__ void k_linear_fwd(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b, int64_t n,
                             int K, int N, int relu, float* __restrict__ y) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n * N) return;
  const int o = int(e % N);
  const int64_t i = e / N;
  float acc = b[o];
  for (int k = 0; k < K; ++k) acc = fmaf(x[i * K + k], W[o * K + k], acc);
  y[e] = relu ? fmaxf(acc, 0.f) : acc;
}
// dy <- dy * (y > 0) when relu; dW[o][k] += sum_i dy[i][o] x[i][k]; db[o] += sum_i dy[i][o]
// gW[o][k] += sum_i dy'[i][o] * x[i][k],  gb[o] += sum_i dy'[i][o]   (dy' = dy masked by the ReLU of the layer's output).
// One CTA per chunk of LBW_TS samples (staged in shared memory), threads own 4(o) x 4(k) blocks, atomicAdd at the end.
constexpr int LBW_TS = 64;
__global__ void __launch_bounds__(256) k_linear_bwd_w(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                                      int64_t n, int K, int N, int relu, float* __restrict__ gW, float* __restrict__ gb) {
  extern __shared__ __align__(16) float lsm[];
  const int Np = (N + 3) & ~3, Kp = (K + 1 + 3) & ~3;  // column K of x is the constant 1 (bias gradient)
  float* sd = lsm;                 // [TS][Np]
  float* sx = sd + LBW_TS * Np;    // [TS][Kp]
  const int tid = threadIdx.x;
  const int64_t i0 = int64_t(blockIdx.x) * LBW_TS;
  for (int e = tid; e < LBW_TS * Np; e += 256) {
    const int i = e / Np, o = e - i * Np;
    float d = 0.f;
    if (o < N && i0 + i < n) {
      d = dy[(i0 + i) * N + o];
      if (relu && y[(i0 + i) * N + o] <= 0.f) d = 0.f;
    }
    sd[e] = d;
  }
  for (int e = tid; e < LBW_TS * Kp; e += 256) {
    const int i = e / Kp, k = e - i * Kp;
    float v = 0.f;
    if (i0 + i < n) v = k < K ? x[(i0 + i) * K + k] : (k == K ? 1.f : 0.f);
    sx[e] = v;
  }
  __syncthreads();
  const int nbo = Np / 4, nbk = Kp / 4;
  for (int blk = tid; blk < nbo * nbk; blk += 256) {
    const int o4 = (blk / nbk) * 4, k4 = (blk % nbk) * 4;
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int i = 0; i < LBW_TS; ++i) {
      const float4 d4 = *reinterpret_cast<const float4*>(sd + i * Np + o4);
      const float4 x4 = *reinterpret_cast<const float4*>(sx + i * Kp + k4);
      const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a * 4 + c] = fmaf(dv[a], xv[c], acc[a * 4 + c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int o = o4 + a, k = k4 + c;
        if (o < N && k < K) atomicAdd(gW + o * K + k, acc[a * 4 + c]);
        else if (o < N && k == K) atomicAdd(gb + o, acc[a * 4 + c]);
      }
  }
}
static inline unsigned lbw_grid(int64_t n) { return (unsigned)((n + LBW_TS - 1) / LBW_TS); }
static inline size_t lbw_smem(int K, int N) {
  const size_t bytes = sizeof(float) * LBW_TS * size_t(((N + 3) & ~3) + ((K + 1 + 3) & ~3));
  static size_t conf = 48 * 1024;
  if (bytes > conf) {
    cudaFuncSetAttribute(k_linear_bwd_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    conf = bytes;
  }
  return bytes;
}
__global__ void k_linear_bwd_x(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ W, int64_t n,
                               int K, int N, int relu, float* __restrict__ dx) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n * K) return;
  const int k = int(e % K);
  const int64_t i = e / K;
  float acc = 0.f;
  for (int o = 0; o < N; ++o) {
    float d = dy[i * N + o];
    if (relu && y[i * N + o] <= 0.f) d = 0.f;
    acc = fmaf(d, W[o * K + k], acc);
  }
  dx[e] = acc;
}
// BatchNorm1d over the batch of an [n][F] tensor + dropout; one block per feature
// ldy: row stride of y (0 = F): the continuous features are normalised straight into the tail columns of the first Linear's input
__global__ void k_bn1d_fwd(const float* __restrict__ x, int64_t n, int F, float* __restrict__ P, int64_t g, int64_t be, int64_t rm,
                           int64_t rv, float* __restrict__ mu, float* __restrict__ invstd, float p_drop, uint64_t seed,
                           const uint32_t* __restrict__ step_p, uint32_t layer, float* __restrict__ y, int ldy = 0) {
  if (ldy == 0) ldy = F;
  __shared__ double sh[2][256];
  const uint32_t step = __ldg(step_p);
  const int f = blockIdx.x;
  double s = 0, q = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[i * F + f];
    s += v;
    q += double(v) * v;
  }
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  const double mean = sh[0][0] / double(n);
  double var = sh[1][0] / double(n) - mean * mean;
  if (var < 0) var = 0;
  const float is = float(1.0 / sqrt(var + double(BN_EPS)));
  if (threadIdx.x == 0) {
    mu[f] = float(mean);
    invstd[f] = is;
    const double unb = n > 1 ? var * double(n) / double(n - 1) : var;
    P[rm + f] = (1.f - BN_MOM) * P[rm + f] + BN_MOM * float(mean);
    P[rv + f] = (1.f - BN_MOM) * P[rv + f] + BN_MOM * float(unb);
  }
  const float ga = P[g + f], bb = P[be + f];
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = (x[i * F + f] - float(mean)) * is * ga + bb;
    y[i * ldy + f] = v * drop_scale(p_drop, seed, step, layer, uint64_t(i) * F + f);
  }
}
// backward of dropout(BN1d(x)): dx, dgamma, dbeta
__global__ void k_bn1d_bwd(const float* __restrict__ dy, const float* __restrict__ x, int64_t n, int F, const float* __restrict__ P,
                           int64_t g, const float* __restrict__ mu, const float* __restrict__ invstd, float p_drop, uint64_t seed,
                           const uint32_t* __restrict__ step_p, uint32_t layer, float* __restrict__ dx, float* __restrict__ G, int64_t g_off,
                           int64_t be_off, int lddy = 0) {   // lddy: row stride of dy (0 = F); dx == NULL: parameter gradients only
  __shared__ double sh[2][256];
  if (lddy == 0) lddy = F;
  const uint32_t step = __ldg(step_p);
  const int f = blockIdx.x;
  const float m = mu[f], is = invstd[f], ga = P[g + f];
  double s = 0, q = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = dy[i * lddy + f] * drop_scale(p_drop, seed, step, layer, uint64_t(i) * F + f);
    s += d;
    q += double(d) * ((x[i * F + f] - m) * is);
  }
  sh[0][threadIdx.x] = s;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  const float s1 = float(sh[0][0] / double(n)), s2 = float(sh[1][0] / double(n));
  if (threadIdx.x == 0) {
    G[be_off + f] += float(sh[0][0]);
    G[g_off + f] += float(sh[1][0]);
  }
  if (!dx) return;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = dy[i * lddy + f] * drop_scale(p_drop, seed, step, layer, uint64_t(i) * F + f);
    const float xh = (x[i * F + f] - m) * is;
    dx[i * F + f] = ga * is * (d - s1 - xh * s2);
  }
}
// ld: row stride of y / dy (K1 + n_cont: the normalised continuous features follow the embeddings, model_snv.py:457-463)
__global__ void k_emb_fwd(const float* __restrict__ E, const int32_t* __restrict__ cat, int64_t n, int n_cat, float p_drop,
                          uint64_t seed, const uint32_t* __restrict__ step_p, float* __restrict__ y, int ld) {
  const uint32_t step = __ldg(step_p);
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int K1 = n_cat * 5;
  if (e >= n * K1) return;
  const int k = int(e % K1);
  const int64_t i = e / K1;
  y[i * ld + k] = E[cat[i * n_cat + k / 5] * 5 + k % 5] * drop_scale(p_drop, seed, step, 100, uint64_t(e));
}
__global__ void k_emb_bwd(const float* __restrict__ dy, const int32_t* __restrict__ cat, int64_t n, int n_cat, float p_drop,
                          uint64_t seed, const uint32_t* __restrict__ step_p, float* __restrict__ gE, int ld) {
  const uint32_t step = __ldg(step_p);
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int K1 = n_cat * 5;
  if (e >= n * K1) return;
  const int k = int(e % K1);
  const int64_t i = e / K1;
  atomicAdd(gE + cat[i * n_cat + k / 5] * 5 + k % 5, dy[i * ld + k] * drop_scale(p_drop, seed, step, 100, uint64_t(e)));
}
// dropout applied in place on an [n][F] tensor (distal_fc: BatchNorm -> Dropout -> Linear; BN handled by k_bn1d_fwd)

// ---------------------------------------------------------------------------------------------- combine + loss
// log(clamp((sm(local) + (sm(mid)+sm(large))/2)/2, 1e-9)), softmaxes saved for the backward (model_snv.py:515-523)
__global__ void k_combine_fwd(const float* __restrict__ ll, const float* __restrict__ lm, const float* __restrict__ lg, int64_t n,
                              int NC, float* __restrict__ sm, float* __restrict__ logp) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float* src[3] = {ll, lm, lg};
  float s[3][16];
  for (int k = 0; k < 3; ++k) {
    float mx = -FLT_MAX, sum = 0.f;
    for (int o = 0; o < NC; ++o) mx = fmaxf(mx, src[k][i * NC + o]);
    for (int o = 0; o < NC; ++o) { s[k][o] = expf(src[k][i * NC + o] - mx); sum += s[k][o]; }
    for (int o = 0; o < NC; ++o) { s[k][o] /= sum; sm[(int64_t(k) * n + i) * NC + o] = s[k][o]; }
  }
  for (int o = 0; o < NC; ++o) logp[i * NC + o] = logf(fmaxf((s[0][o] + (s[1][o] + s[2][o]) / 2.f) / 2.f, 1e-9f));
}
__global__ void k_combine_bwd(const float* __restrict__ dlogp, const float* __restrict__ sm, int64_t n, int NC,
                              float* __restrict__ dl, float* __restrict__ dm, float* __restrict__ dg) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float* dst[3] = {dl, dm, dg};
  float dpm[16];
  for (int o = 0; o < NC; ++o) {
    const float pm = (sm[(0 * n + i) * NC + o] + (sm[(1 * n + i) * NC + o] + sm[(2 * n + i) * NC + o]) / 2.f) / 2.f;
    dpm[o] = pm > 1e-9f ? dlogp[i * NC + o] / pm : 0.f;
  }
  for (int k = 0; k < 3; ++k) {
    const float w = k == 0 ? 0.5f : 0.25f;
    float dot = 0.f;
    for (int o = 0; o < NC; ++o) dot += w * dpm[o] * sm[(int64_t(k) * n + i) * NC + o];
    for (int o = 0; o < NC; ++o) {
      const float s = sm[(int64_t(k) * n + i) * NC + o];
      dst[k][i * NC + o] = s * (w * dpm[o] - dot);
    }
  }
}
// CrossEntropyLoss(reduction='sum') on the log-probs and its gradient (training.py:425)
__global__ void k_ce_grad(const float* __restrict__ logp, const int32_t* __restrict__ meta, int64_t n, int NC, double* __restrict__ loss,
                          float* __restrict__ dlogp) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  float v = 0.f;
  if (i < n) {
    const float* p = logp + i * NC;
    float mx = -FLT_MAX, s = 0.f;
    for (int o = 0; o < NC; ++o) mx = fmaxf(mx, p[o]);
    for (int o = 0; o < NC; ++o) s += expf(p[o] - mx);
    int y = (meta[i] >> 1) & 0x7f;
    if (y >= NC) y = 0;
    for (int o = 0; o < NC; ++o) dlogp[i * NC + o] = expf(p[o] - mx) / s - (o == y ? 1.f : 0.f);
    v = -(p[y] - mx - logf(s));
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0 && loss) atomicAdd(loss, double(v));
}

// ---------------------------------------------------------------------------------------------- stem (train)
__global__ void k_sym_gather(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta, int R, int L,
                             int local_R, int order, int n_cat, uint8_t* __restrict__ sym_out, int32_t* __restrict__ cat_out) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  uint8_t* sym = sm_raw;
  const int64_t site = blockIdx.x;
  const int m = meta[site];
  load_window(G, int(uint32_t(m) >> 8), int64_t(pos[site]) - R, L, m & 1, sym);
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += blockDim.x) sym_out[site * L + i] = sym[i];
  for (int j = threadIdx.x; j < n_cat; j += blockDim.x) {
    int idx = 0;
    bool bad = false;
    for (int d = 0; d < order; ++d) {
      const int s = sym[R - local_R + j + d];
      bad |= s > 3;
      idx = idx * 4 + (s & 3);
    }
    cat_out[site * n_cat + j] = bad ? (1 << (2 * order)) : idx;
  }
}
__global__ void k_sym_hist(const uint8_t* __restrict__ sym, int64_t n, int L, int off0, int L0, unsigned long long* __restrict__ cnt) {
  __shared__ unsigned int h[16];
  if (threadIdx.x < 16) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t tot = n * L0;
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < tot; e += int64_t(gridDim.x) * blockDim.x) {
    const int64_t site = e / L0;
    const int p = int(e - site * L0);
    atomicAdd(&h[sym[site * L + off0 + p] & 15], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 16 && h[threadIdx.x]) atomicAdd(cnt + threadIdx.x, (unsigned long long)h[threadIdx.x]);
}
__constant__ float c_e[16][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}, {.5f, 0, .5f, 0}, {0, .5f, 0, .5f},
                                 {.5f, .5f, 0, 0}, {0, .5f, .5f, 0}, {.5f, 0, 0, .5f}, {0, 0, .5f, .5f},
                                 {0, (float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3)}, {(float)(1.0 / 3), 0, (float)(1.0 / 3), (float)(1.0 / 3)},
                                 {(float)(1.0 / 3), (float)(1.0 / 3), 0, (float)(1.0 / 3)}, {(float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3), 0},
                                 {.25f, .25f, .25f, .25f}, {0, 0, 0, 0}};
// BN(4) batch statistics from the symbol histogram, running-stat update, and the per-tap table of the stem
__global__ void k_stem_bn_table(const unsigned long long* __restrict__ cnt, float* __restrict__ P, int64_t g, int64_t be, int64_t rm,
                                int64_t rv, int64_t w, int C, int ks, float* __restrict__ ab /*[4] a,[4] b,[4] mu,[4] invstd*/,
                                float* __restrict__ T) {
  __shared__ float a[4], b[4];
  if (threadIdx.x < 4) {
    const int c = threadIdx.x;
    double N = 0, s = 0, q = 0;
    for (int k = 0; k < 15; ++k) {
      const double ck = double(cnt[k]);
      N += ck;
      s += ck * c_e[k][c];
      q += ck * double(c_e[k][c]) * c_e[k][c];
    }
    const double mean = s / N;
    double var = q / N - mean * mean;
    if (var < 0) var = 0;
    const float is = float(1.0 / sqrt(var + double(BN_EPS)));
    a[c] = P[g + c] * is;
    b[c] = P[be + c] - float(mean) * a[c];
    ab[c] = a[c]; ab[4 + c] = b[c]; ab[8 + c] = float(mean); ab[12 + c] = is;
    const double unb = N > 1 ? var * N / (N - 1) : var;
    P[rm + c] = (1.f - BN_MOM) * P[rm + c] + BN_MOM * float(mean);
    P[rv + c] = (1.f - BN_MOM) * P[rv + c] + BN_MOM * float(unb);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ks * 16 * C; e += blockDim.x) {
    const int t = e / (16 * C), sy = (e / C) % 16, co = e % C;
    float acc = 0.f;
    if (sy < 15)
      for (int c = 0; c < 4; ++c) acc += P[w + (int64_t(co) * 4 + c) * ks + t] * (a[c] * c_e[sy][c] + b[c]);
    T[e] = acc;
  }
}
template <int C>
__global__ void __launch_bounds__(128) k_stem_train(const uint8_t* __restrict__ sym_g, int L, int ks, const float* __restrict__ T,
                                                    const float* __restrict__ bias, int L0, int off0, int L1, int pk, int ps, int pp,
                                                    float* __restrict__ out, int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  float* sT = reinterpret_cast<float*>(sm_raw);
  float* sB = sT + ks * 16 * C;
  uint8_t* sym = reinterpret_cast<uint8_t*>(sB + C);
  const int64_t site = blockIdx.x;
  for (int e = threadIdx.x; e < ks * 16 * C; e += blockDim.x) sT[e] = T[e];
  for (int e = threadIdx.x; e < C; e += blockDim.x) sB[e] = bias[e];
  for (int i = threadIdx.x; i < L0; i += blockDim.x) sym[i] = sym_g[site * L + off0 + i];
  __syncthreads();
  const int half = ks / 2;
  // grid.y slices of the site's bins: at small batches one CTA per site leaves SMs idle and its serial walk over all bins is on
  // the critical path of the step
  const int jb = (L1 + gridDim.y - 1) / gridDim.y, j0 = blockIdx.y * jb, j1 = (j0 + jb < L1) ? j0 + jb : L1;
  for (int e = j0 * C + threadIdx.x; e < j1 * C; e += blockDim.x) {
    const int j = e / C, c = e - j * C;
    int lo = j * ps - pp, hi = lo + pk;
    lo = lo < 0 ? 0 : lo;
    hi = hi > L0 ? L0 : hi;
    float mx = -FLT_MAX;
    int am = lo;
    if (ks == 3) {   // shipped kernel size: the three symbols of a position slide along the bin (one symbol load per position)
      const float* t0 = sT + c;
      const float b = sB[c];
      int s0 = lo >= 1 ? sym[lo - 1] : SYM_PAD, s1 = sym[lo];
      for (int p = lo; p < hi; ++p) {
        const int s2 = p + 1 < L0 ? sym[p + 1] : SYM_PAD;
        const float v = ((b + t0[s0 * C]) + t0[(16 + s1) * C]) + t0[(32 + s2) * C];   // same summation order as the generic loop
        if (v > mx) { mx = v; am = p; }
        s0 = s1; s1 = s2;
      }
    } else {
      for (int p = lo; p < hi; ++p) {
        float v = sB[c];
        for (int t = 0; t < ks; ++t) {
          const int q = p + t - half;
          const int s = (q >= 0 && q < L0) ? sym[q] : SYM_PAD;
          v += sT[(t * 16 + s) * C + c];
        }
        if (v > mx) { mx = v; am = p; }
      }
    }
    out[(site * L1) * int64_t(C) + e] = mx;
    idx[(site * L1) * int64_t(C) + e] = am;
  }
}
// gradient w.r.t. the stem table: Gt[t][sym][co] += g at every (arg-max position, in-range tap); dbias[co] += g
template <int C>
__global__ void __launch_bounds__(256) k_stem_bwd(const float* __restrict__ g, const int32_t* __restrict__ idx,
                                                  const uint8_t* __restrict__ sym_g, int64_t n, int L, int ks, int L0, int off0,
                                                  int L1, float* __restrict__ Gt, float* __restrict__ gbias) {
  extern __shared__ float sG[];  // [ks][16][C] + [C]
  const int nG = ks * 16 * C;
  for (int e = threadIdx.x; e < nG + C; e += blockDim.x) sG[e] = 0.f;
  __syncthreads();
  const int half = ks / 2;
  const int64_t tot = n * L1 * C;
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < tot; e += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(e % C);
    const int64_t site = e / (int64_t(L1) * C);
    const float gv = g[e];
    if (gv == 0.f) continue;
    const int p = idx[e];
    atomicAdd(&sG[nG + c], gv);
    for (int t = 0; t < ks; ++t) {
      const int q = p + t - half;
      if (q >= 0 && q < L0) atomicAdd(&sG[(t * 16 + sym_g[site * L + off0 + q]) * C + c], gv);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nG; e += blockDim.x)
    if (sG[e] != 0.f) atomicAdd(Gt + e, sG[e]);
  for (int e = threadIdx.x; e < C; e += blockDim.x) atomicAdd(gbias + e, sG[nG + e]);
}
// table gradient -> conv1 weight / BN(4) gamma, beta gradients (single CTA)
__global__ void k_stem_param_grad(const float* __restrict__ Gt, const float* __restrict__ P, int64_t w, int C, int ks,
                                  const float* __restrict__ ab, float* __restrict__ G, int64_t gw, int64_t gg, int64_t gbe) {
  __shared__ float dxbn[15][4];
  const float *a = ab, *b = ab + 4, *mu = ab + 8, *is = ab + 12;
  for (int e = threadIdx.x; e < C * 4 * ks; e += blockDim.x) {  // dW[co][ci][t] = sum_s Gt[t][s][co] * xbn_s[ci]
    const int co = e / (4 * ks), ci = (e / ks) % 4, t = e % ks;
    float acc = 0.f;
    for (int s = 0; s < 15; ++s) acc += Gt[(t * 16 + s) * C + co] * (a[ci] * c_e[s][ci] + b[ci]);
    G[gw + e] += acc;
  }
  for (int e = threadIdx.x; e < 15 * 4; e += blockDim.x) {  // d(BN output of symbol s)[ci] = sum_t sum_co W * Gt
    const int s = e / 4, ci = e % 4;
    float acc = 0.f;
    for (int t = 0; t < ks; ++t)
      for (int co = 0; co < C; ++co) acc += P[w + (int64_t(co) * 4 + ci) * ks + t] * Gt[(t * 16 + s) * C + co];
    dxbn[s][ci] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int ci = threadIdx.x;
    float db = 0.f, dg = 0.f;
    for (int s = 0; s < 15; ++s) {
      db += dxbn[s][ci];
      dg += dxbn[s][ci] * ((c_e[s][ci] - mu[ci]) * is[ci]);
    }
    G[gbe + ci] += db;
    G[gg + ci] += dg;
  }
}

// ---------------------------------------------------------------------------------------------- optimizer
__global__ void k_sumsq(const float* __restrict__ g, int64_t n, float scale, double* __restrict__ out) {
  double s = 0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const double v = double(g[i]) * scale;
    s += v * v;
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, sh[0]);
}
__global__ void k_bump_u32(uint32_t* c) { *c += 1u; }
__global__ void k_bump_i64(int64_t* c) { *c += 1; }
// kind 0: Adam (coupled L2), 1: AdamW + amsgrad, 2: SGD momentum 0.98 nesterov  (training.py:346-357);
// gradient first scaled by grad_scale, then by the clip_grad_norm_ coefficient min(1, max_norm/(norm+1e-6))
__global__ void k_opt_step(int kind, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                           float* __restrict__ vmax, int64_t n, float lr, float wd, int64_t step, float max_norm, float grad_scale,
                           const double* __restrict__ sumsq, const float* __restrict__ lr_p, const int64_t* __restrict__ step_p) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  if (lr_p) lr = __ldg(lr_p);          // graph-captured steps: hyper-parameters that change per step live on the device
  if (step_p) step = __ldg(step_p);
  float clip = 1.f;
  if (max_norm > 0.f) {
    const float norm = float(sqrt(*sumsq));
    clip = fminf(1.f, max_norm / (norm + 1e-6f));
  }
  float gr = g[i] * grad_scale * clip;
  float w = p[i];
  if (kind == 2) {
    const float mom = 0.98f;
    gr += wd * w;
    const float buf = step == 1 ? gr : mom * m[i] + gr;
    m[i] = buf;
    w -= lr * (gr + mom * buf);
  } else {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    if (kind == 0) gr += wd * w;
    else w *= (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gr;
    const float vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi;
    v[i] = vi;
    const float bc1 = 1.f - powf(b1, float(step)), bc2 = 1.f - powf(b2, float(step));
    float vv = vi;
    if (kind == 1) { vv = fmaxf(vmax[i], vi); vmax[i] = vv; }
    w -= (lr / bc1) * mi / (sqrtf(vv) / sqrtf(bc2) + eps);
  }
  p[i] = w;
}

}  // namespace train
}  // namespace mural

using namespace mural;
using namespace mural::train;

// ================================================================================================ host
struct mural_snv_train {
  mural_snv_model* m;
  int64_t cap = 0;  // sites the tape is sized for
  ConvDesc h_conv[N_CONV];
  ConvDesc* d_conv = nullptr;
  float *d_Wt = nullptr, *d_Wf = nullptr;  // [N_CONV][ks_max*C*C]
  float* d_bn = nullptr;                   // per conv layer: a, b, mu, invstd  [N_CONV][4][C]
  float* d_const = nullptr;                // ones[C], zeros[C]
  double* d_stat = nullptr;                // scratch [2*C] doubles (+ sumsq)
  float* d_stem = nullptr;                 // per branch: ab[16], T[ks*16*C], Gt[ks*16*C]
  unsigned long long* d_cnt = nullptr;     // [2][16]
  void* d_tape = nullptr;
  int64_t tape_bytes = 0;
  // dropout
  float p_emb = 0.f, p_local = 0.f, p_fc = 0.f;
  uint64_t seed = 0;
  uint32_t* d_step = nullptr;  // forward counter on the device (dropout stream position): bumped by a kernel, so that a
                               // captured CUDA graph of the step draws new masks at every replay
  // tape pointers (valid after forward)
  int64_t n = 0;
  uint8_t* sym = nullptr;
  int32_t* cat = nullptr;
  struct Br { float *x0, *t1, *y1, *t2, *z1, *x2, *j2, *t1b, *y1b, *t2b, *z2, *x3, *h, *gm, *gmn, *logit;
              int32_t *i1, *i2, *i3, *ig; float *mu, *is; } br[2];
  float *e0, *r1, *d1, *r2, *d2, *ll, *mu1, *is1, *mu2, *is2, *smx, *mu0, *is0;
  const float* cont = nullptr;   // cont_x of the last forward (kept for the backward of first_bn_layer)
  float* gbuf[2][5];   // gradient scratch of each CNN branch
  float* gsm[9];       // [0..2] d(logits) of the three branches; then (ga, gb) of the local chain and of each CNN branch
  // the mid branch, the large branch and the local MLP are independent chains of small kernels between the window gather
  // and the combine: they run on three streams (fork / join with events, also inside a captured graph)
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  // weight gradients leave the backward chains (nothing reads them before the optimizer): one more stream per conv branch;
  // ev_w[br] = completion of the branch's latest weight-gradient kernel, awaited before a buffer it reads is overwritten
  cudaStream_t wside[2] = {nullptr, nullptr};
  cudaEvent_t ev_wfork[2] = {nullptr, nullptr}, ev_w[2] = {nullptr, nullptr};
  bool w_pending[2] = {false, false};
};

static int64_t off_of(const mural_snv_model* m, const std::string& n) { return m->layout[m->index.at(n)].offset; }

extern "C" int mural_snv_train_create(mural_snv_model_t* m, mural_snv_train_t** out) {
  MURAL_CHECK(m && out, "NULL argument");
  mural_snv_train* T = new mural_snv_train();
  T->m = m;
  const int C = m->cfg.channels, ks = m->cfg.kernel_size;
  int k = 0;
  for (int br = 0; br < 2; ++br) {
    const std::string s = br ? "_2" : "";
    auto add = [&](const std::string& bn, const std::string& cv, int kss, int relu) {
      T->h_conv[k++] = ConvDesc{off_of(m, cv + ".weight"), off_of(m, cv + ".bias"), off_of(m, bn + ".weight"), off_of(m, bn + ".bias"),
                                off_of(m, bn + ".running_mean"), off_of(m, bn + ".running_var"), kss, relu};
    };
    for (int g = 1; g <= 2; ++g) {
      if (g == 2) add("conv2" + s + ".0", "conv2" + s + ".1", ks, 0);
      for (int i = 0; i < 2; ++i) {
        const std::string p = "RBs" + std::to_string(g) + s + "." + std::to_string(i);
        add(p + ".bn1", p + ".conv1", 3, 1);
        add(p + ".bn2", p + ".conv2", 3, 1);
      }
    }
    add("conv3" + s + ".0", "conv3" + s + ".1", ks, 0);
  }
  // per branch order: rb1[0..3] (0-3), conv2 (4), rb2[0..3] (5-8), conv3 (9)
  CUDA_TRY(cudaSetDevice(m->device));
  const int64_t wsz = int64_t(7) * C * C;
  CUDA_TRY(cudaMalloc((void**)&T->d_conv, sizeof(ConvDesc) * N_CONV));
  CUDA_TRY(cudaMemcpy(T->d_conv, T->h_conv, sizeof(ConvDesc) * N_CONV, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc((void**)&T->d_Wt, sizeof(float) * N_CONV * wsz));
  CUDA_TRY(cudaMalloc((void**)&T->d_Wf, sizeof(float) * N_CONV * wsz));
  CUDA_TRY(cudaMalloc((void**)&T->d_bn, sizeof(float) * N_CONV * 4 * C));
  CUDA_TRY(cudaMalloc((void**)&T->d_const, sizeof(float) * 2 * C));
  std::vector<float> cst(2 * C, 0.f);
  for (int i = 0; i < C; ++i) cst[i] = 1.f;
  CUDA_TRY(cudaMemcpy(T->d_const, cst.data(), sizeof(float) * 2 * C, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc((void**)&T->d_stat, sizeof(double) * (2 * 1024 + 8)));
  CUDA_TRY(cudaMalloc((void**)&T->d_stem, sizeof(float) * 2 * (16 + 2 * int64_t(ks) * 16 * C)));
  CUDA_TRY(cudaMalloc((void**)&T->d_cnt, sizeof(unsigned long long) * 32));
  CUDA_TRY(cudaMalloc((void**)&T->d_step, 4));
  CUDA_TRY(cudaMemset(T->d_step, 0, 4));
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaStreamCreateWithFlags(&T->side[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&T->ev_join[i], cudaEventDisableTiming));
    CUDA_TRY(cudaStreamCreateWithFlags(&T->wside[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&T->ev_wfork[i], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&T->ev_w[i], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&T->ev_fork, cudaEventDisableTiming));
  *out = T;
  return 0;
}

extern "C" void mural_snv_train_destroy(mural_snv_train_t* T) {
  if (!T) return;
  cudaFree(T->d_conv); cudaFree(T->d_Wt); cudaFree(T->d_Wf); cudaFree(T->d_bn); cudaFree(T->d_const);
  cudaFree(T->d_stat); cudaFree(T->d_stem); cudaFree(T->d_cnt); cudaFree(T->d_tape); cudaFree(T->d_step);
  for (int i = 0; i < 2; ++i) {
    if (T->side[i]) cudaStreamDestroy(T->side[i]);
    if (T->ev_join[i]) cudaEventDestroy(T->ev_join[i]);
    if (T->wside[i]) cudaStreamDestroy(T->wside[i]);
    if (T->ev_wfork[i]) cudaEventDestroy(T->ev_wfork[i]);
    if (T->ev_w[i]) cudaEventDestroy(T->ev_w[i]);
  }
  if (T->ev_fork) cudaEventDestroy(T->ev_fork);
  delete T;
}

extern "C" int mural_snv_train_set_dropout(mural_snv_train_t* T, float p_emb, float p_local, float p_fc, uint64_t seed) {
  MURAL_CHECK(T, "NULL argument");
  MURAL_CHECK(p_emb >= 0 && p_emb < 1 && p_local >= 0 && p_local < 1 && p_fc >= 0 && p_fc < 1, "dropout must be in [0,1)");
  T->p_emb = p_emb; T->p_local = p_local; T->p_fc = p_fc; T->seed = seed;
  return 0;
}

namespace {
struct Carver {
  char* p;
  template <typename X> X* take(int64_t n) {
    X* r = reinterpret_cast<X*>(p);
    p += (n * sizeof(X) + 255) & ~int64_t(255);
    return r;
  }
};

int ensure_tape(mural_snv_train* T, int64_t n) {
  if (T->cap >= n && T->d_tape) return 0;
  const mural_snv_model* m = T->m;
  const int C = m->cfg.channels, NC = m->cfg.n_class, H1 = m->cfg.hidden1, H2 = m->cfg.hidden2, K1 = m->k1;
  for (int pass = 0; pass < 2; ++pass) {
    Carver cv{pass ? (char*)T->d_tape : (char*)nullptr};
    T->sym = cv.take<uint8_t>(n * m->L);
    T->cat = cv.take<int32_t>(n * m->n_cat);
    for (int br = 0; br < 2; ++br) {
      const BranchDev& B = m->br[br];
      auto& b = T->br[br];
      const int64_t s1 = n * B.L1 * C, s2 = n * B.L2 * C, s3 = n * B.L3 * C;
      b.x0 = cv.take<float>(s1); b.t1 = cv.take<float>(s1); b.y1 = cv.take<float>(s1); b.t2 = cv.take<float>(s1); b.z1 = cv.take<float>(s1);
      b.x2 = cv.take<float>(s2); b.j2 = cv.take<float>(s2); b.t1b = cv.take<float>(s2); b.y1b = cv.take<float>(s2);
      b.t2b = cv.take<float>(s2); b.z2 = cv.take<float>(s2);
      b.x3 = cv.take<float>(s3); b.h = cv.take<float>(s3);
      b.gm = cv.take<float>(n * C); b.gmn = cv.take<float>(n * C); b.logit = cv.take<float>(n * NC);
      b.i1 = cv.take<int32_t>(s1); b.i2 = cv.take<int32_t>(s2); b.i3 = cv.take<int32_t>(s3); b.ig = cv.take<int32_t>(n * C);
      b.mu = cv.take<float>(C); b.is = cv.take<float>(C);
    }
    const int K1c = m->k1c > 0 ? m->k1c : K1;
    T->e0 = cv.take<float>(n * K1c); T->r1 = cv.take<float>(n * H1); T->d1 = cv.take<float>(n * H1);
    T->mu0 = cv.take<float>(K1c - K1 + 1); T->is0 = cv.take<float>(K1c - K1 + 1);
    T->r2 = cv.take<float>(n * H2); T->d2 = cv.take<float>(n * H2); T->ll = cv.take<float>(n * NC);
    T->mu1 = cv.take<float>(H1); T->is1 = cv.take<float>(H1); T->mu2 = cv.take<float>(H2); T->is2 = cv.take<float>(H2);
    T->smx = cv.take<float>(3 * n * NC);
    for (int br = 0; br < 2; ++br)
      for (int i = 0; i < 5; ++i) T->gbuf[br][i] = cv.take<float>(n * m->br[br].L1 * C);
    const int64_t Fmax = (H1 > K1c ? H1 : K1c) > C ? (H1 > K1c ? H1 : K1c) : C;
    for (int i = 0; i < 9; ++i) T->gsm[i] = cv.take<float>(n * Fmax);
    if (!pass) {
      cudaFree(T->d_tape);
      T->d_tape = nullptr;
      T->tape_bytes = int64_t(cv.p - (char*)nullptr);
      CUDA_TRY(cudaMalloc(&T->d_tape, T->tape_bytes));
    }
  }
  T->cap = n;
  return 0;
}

inline unsigned gridn(int64_t n, int b = 256) { return (unsigned)cdiv(n, b); }
}  // namespace

// ------------------------------------------------------------------------------------------------ forward
static int conv_train_fwd(mural_snv_train* T, float* P, int li, const float* x, float* y, const float* r1, const float* r2,
                          int64_t n, int L, int relu_out, cudaStream_t st) {
  const int C = T->m->cfg.channels;
  double* stat = T->d_stat + (li / 10) * 256;  // per-branch scratch: the branches run concurrently
  const ConvDesc& d = T->h_conv[li];
  float* bn = T->d_bn + int64_t(li) * 4 * C;
  const int64_t rows = n * L;
  CUDA_TRY(cudaMemsetAsync(stat, 0, sizeof(double) * 2 * C, st));
  const int thr = 256, rpb = thr / (C / 4);   // a thread owns 4 channels of a row
  int grid = (int)cdiv(rows, rpb * 4);
  if (grid > 1184) grid = 1184;
  if (grid < 1) grid = 1;
  LAUNCH(k_stats, grid, thr, sizeof(double) * 8 * thr, st, x, rows, C, d.relu_in, stat);
  LAUNCH(k_bn_finalize, 1, 64, 0, st, stat, double(rows), C, P, d.g, d.be, d.rm, d.rv, bn, bn + C, bn + 2 * C, bn + 3 * C);
  ConvLayerDev cl{T->d_Wt + int64_t(li) * 7 * C * C, P + d.b, bn, bn + C, d.ks, d.relu_in, 1};
  return conv_any(C, x, y, r1, r2, n, L, cl, relu_out, st);
}

extern "C" int mural_snv_train_forward(mural_snv_train_t* T, const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta,
                                       int64_t n, float* d_blob, float* d_logp, void* stream) {
  MURAL_CHECK(T && g && d_pos && d_meta && d_blob && d_logp, "NULL argument");
  MURAL_CHECK(n >= 2, "BatchNorm in training mode needs more than one site per batch");  // training.py:415 skips them
  mural_snv_model* m = T->m;
  cudaStream_t st = (cudaStream_t)stream;
  const int C = m->cfg.channels, ks = m->cfg.kernel_size, NC = m->cfg.n_class, H1 = m->cfg.hidden1, H2 = m->cfg.hidden2, K1 = m->k1;
  const int NCT = m->cfg.n_cont, K1c = K1 + NCT;
  MURAL_CHECK(NCT == 0 || m->d_cont != nullptr, "this model has continuous features: call mural_snv_set_cont before the training forward");
  MURAL_CHECK(ks * C * C <= 48 * 256, "unsupported conv shape for training (CNN_kernel_size * C^2 must be <= 12288)");
  if (int rc = ensure_tape(T, n)) return rc;
  T->n = n;
  LAUNCH(k_bump_u32, 1, 1, 0, st, T->d_step);
  float* P = d_blob;
  const int64_t wsz = int64_t(7) * C * C;
  LAUNCH(k_prep_conv, N_CONV, 256, 0, st, P, T->d_conv, C, T->d_Wt, T->d_Wf, wsz);
  LAUNCH(k_sym_gather, (unsigned)n, 128, (size_t(m->L) + 15) & ~size_t(15), st, g->view, d_pos, d_meta, m->cfg.distal_radius, m->L,
         m->cfg.local_radius, m->cfg.local_order, m->n_cat, T->sym, T->cat);
  CUDA_TRY(cudaMemsetAsync(T->d_cnt, 0, sizeof(unsigned long long) * 32, st));
  // fork: mid branch -> side[0], local MLP -> side[1]; the large branch (the longest chain) stays on the caller's stream
  cudaStream_t const st_main = st;
  CUDA_TRY(cudaEventRecord(T->ev_fork, st_main));
  for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamWaitEvent(T->side[i], T->ev_fork, 0));
  for (int br = 0; br < 2; ++br) {
    st = br ? st_main : T->side[0];
    const BranchDev& B = m->br[br];
    auto& b = T->br[br];
    const std::string s = br ? "_2" : "";
    const int off0 = br ? 0 : m->L / 2 - 100;
    float* stem = T->d_stem + br * (16 + 2 * int64_t(ks) * 16 * C);
    LAUNCH(k_sym_hist, 296, 256, 0, st, T->sym, n, m->L, off0, B.L0, T->d_cnt + br * 16);
    LAUNCH(k_stem_bn_table, 1, 256, 0, st, T->d_cnt + br * 16, P, off_of(m, "conv1" + s + ".0.weight"), off_of(m, "conv1" + s + ".0.bias"),
           off_of(m, "conv1" + s + ".0.running_mean"), off_of(m, "conv1" + s + ".0.running_var"), off_of(m, "conv1" + s + ".1.weight"), C,
           ks, stem, stem + 16);
    const size_t smem = sizeof(float) * (size_t(ks) * 16 * C + C) + ((size_t(B.L0) + 15) & ~size_t(15));
    unsigned stem_parts = (unsigned)((8 * 148 + n - 1) / n);   // ~8 CTAs of 128 threads per SM in all
    if (stem_parts > 8) stem_parts = 8;
    if (stem_parts > (unsigned)B.L1) stem_parts = (unsigned)B.L1;
    if (stem_parts < 1) stem_parts = 1;
#define STEMT(CC)                                                                                                         \
  case CC: {                                                                                                             \
    static size_t conf = 0;                                                                                              \
    if (smem > 48 * 1024 && smem > conf) {                                                                               \
      CUDA_TRY(cudaFuncSetAttribute(k_stem_train<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
      conf = smem;                                                                                                       \
    }                                                                                                                    \
    LAUNCH(k_stem_train<CC>, dim3((unsigned)n, stem_parts), 128, smem, st, T->sym, m->L, ks, stem + 16, P + off_of(m, "conv1" + s + ".1.bias"), B.L0, \
           off0, B.L1, B.pool[0][0], B.pool[0][1], B.pool[0][2], b.x0, b.i1);                                            \
  } break;
    switch (C) { STEMT(16) STEMT(32) STEMT(64) default: MURAL_FAIL("unsupported channel count"); }
#undef STEMT
    const int base = br * 10;
    if (int rc = conv_train_fwd(T, P, base + 0, b.x0, b.t1, nullptr, nullptr, n, B.L1, 0, st)) return rc;
    if (int rc = conv_train_fwd(T, P, base + 1, b.t1, b.y1, b.x0, nullptr, n, B.L1, 0, st)) return rc;
    if (int rc = conv_train_fwd(T, P, base + 2, b.y1, b.t2, nullptr, nullptr, n, B.L1, 0, st)) return rc;
    if (int rc = conv_train_fwd(T, P, base + 3, b.t2, b.z1, b.y1, b.x0, n, B.L1, 0, st)) return rc;
    LAUNCH(k_pool_fwd_idx, gridn(n * B.L2 * C), 256, 0, st, b.z1, b.x2, b.i2, n, B.L1, B.L2, C, B.pool[1][0], B.pool[1][1], B.pool[1][2]);
    if (int rc = conv_train_fwd(T, P, base + 4, b.x2, b.j2, nullptr, nullptr, n, B.L2, 0, st)) return rc;
    if (int rc = conv_train_fwd(T, P, base + 5, b.j2, b.t1b, nullptr, nullptr, n, B.L2, 0, st)) return rc;
    if (int rc = conv_train_fwd(T, P, base + 6, b.t1b, b.y1b, b.j2, nullptr, n, B.L2, 0, st)) return rc;
    if (int rc = conv_train_fwd(T, P, base + 7, b.y1b, b.t2b, nullptr, nullptr, n, B.L2, 0, st)) return rc;
    if (int rc = conv_train_fwd(T, P, base + 8, b.t2b, b.z2, b.y1b, b.j2, n, B.L2, 0, st)) return rc;
    LAUNCH(k_pool_fwd_idx, gridn(n * B.L3 * C), 256, 0, st, b.z2, b.x3, b.i3, n, B.L2, B.L3, C, B.pool[2][0], B.pool[2][1], B.pool[2][2]);
    if (int rc = conv_train_fwd(T, P, base + 9, b.x3, b.h, nullptr, nullptr, n, B.L3, 1, st)) return rc;
    LAUNCH(k_gmax_fwd, gridn(n * C), 256, 0, st, b.h, b.gm, b.ig, n, B.L3, C);
    const std::string fc = br ? "distal_fc2" : "distal_fc1";
    LAUNCH(k_bn1d_fwd, C, 256, 0, st, b.gm, n, C, P, off_of(m, fc + ".0.weight"), off_of(m, fc + ".0.bias"),
           off_of(m, fc + ".0.running_mean"), off_of(m, fc + ".0.running_var"), b.mu, b.is, T->p_fc, T->seed, T->d_step, 10u + br, b.gmn);
    LAUNCH(k_linear_fwd, gridn(n * NC), 256, 0, st, b.gmn, P + off_of(m, fc + ".2.weight"), P + off_of(m, fc + ".2.bias"), n, C, NC, 0,
           b.logit);
  }
  // local branch (model_snv.py:452-468, 492)
  st = T->side[1];
  LAUNCH(k_emb_fwd, gridn(n * K1), 256, 0, st, P + off_of(m, "emb_layer.weight"), T->cat, n, m->n_cat, T->p_emb, T->seed, T->d_step, T->e0, K1c);
  T->cont = m->d_cont;
  m->d_cont = m->d_cont_cur = nullptr;   // consumed, as in the eval forward
  if (NCT > 0)   // first_bn_layer(cont_data), batch statistics, no dropout, concatenated behind the embeddings (model_snv.py:457-463)
    LAUNCH(k_bn1d_fwd, NCT, 256, 0, st, T->cont, n, NCT, P, off_of(m, "first_bn_layer.weight"), off_of(m, "first_bn_layer.bias"),
           off_of(m, "first_bn_layer.running_mean"), off_of(m, "first_bn_layer.running_var"), T->mu0, T->is0, 0.f, T->seed, T->d_step, 0u,
           T->e0 + K1, K1c);
  LAUNCH(k_linear_fwd, gridn(n * H1), 256, 0, st, T->e0, P + off_of(m, "lin_layers.0.weight"), P + off_of(m, "lin_layers.0.bias"), n, K1c, H1,
         1, T->r1);
  LAUNCH(k_bn1d_fwd, H1, 256, 0, st, T->r1, n, H1, P, off_of(m, "bn_layers.0.weight"), off_of(m, "bn_layers.0.bias"),
         off_of(m, "bn_layers.0.running_mean"), off_of(m, "bn_layers.0.running_var"), T->mu1, T->is1, T->p_local, T->seed, T->d_step, 1u, T->d1);
  LAUNCH(k_linear_fwd, gridn(n * H2), 256, 0, st, T->d1, P + off_of(m, "lin_layers.1.weight"), P + off_of(m, "lin_layers.1.bias"), n, H1, H2,
         1, T->r2);
  LAUNCH(k_bn1d_fwd, H2, 256, 0, st, T->r2, n, H2, P, off_of(m, "bn_layers.1.weight"), off_of(m, "bn_layers.1.bias"),
         off_of(m, "bn_layers.1.running_mean"), off_of(m, "bn_layers.1.running_var"), T->mu2, T->is2, T->p_local, T->seed, T->d_step, 2u, T->d2);
  LAUNCH(k_linear_fwd, gridn(n * NC), 256, 0, st, T->d2, P + off_of(m, "local_fc.0.weight"), P + off_of(m, "local_fc.0.bias"), n, H2, NC, 0,
         T->ll);
  // join
  st = st_main;
  for (int i = 0; i < 2; ++i) {
    CUDA_TRY(cudaEventRecord(T->ev_join[i], T->side[i]));
    CUDA_TRY(cudaStreamWaitEvent(st, T->ev_join[i], 0));
  }
  LAUNCH(k_combine_fwd, gridn(n, 128), 128, 0, st, T->ll, T->br[0].logit, T->br[1].logit, n, NC, T->smx, d_logp);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ backward
// gradient of one BN->conv layer: parameter grads (+=) and dx written to `out` (= dx + add1 + add2)
static int conv_train_bwd(mural_snv_train* T, const float* P, float* G, int li, const float* x, const float* dy, float* du,
                          const float* add1, const float* add2, float* out, int64_t n, int L, cudaStream_t st) {
  const int C = T->m->cfg.channels;
  const ConvDesc& d = T->h_conv[li];
  float* bn = T->d_bn + int64_t(li) * 4 * C;
  double* stat = T->d_stat + (li / 10) * 256;
  const int64_t rows = n * L;
  const size_t smem = sizeof(float) * (size_t(64 + d.ks - 1) * (C + 4) + size_t(64) * (C + 4)) + sizeof(int) * 64;
  static const bool use_mma = getenv("MURAL_NO_CONV_MMA") == nullptr && getenv("MURAL_NO_WGRAD_MMA") == nullptr;
  int wg = (int)cdiv(rows, 64);
  if (wg > 296) wg = 296;
  // the weight gradient reads x and dy only and feeds nothing downstream: fork it onto the branch's weight-gradient stream
  // (captured as a parallel graph branch); the chain joins it before a buffer it reads can be overwritten (wait_wgrad)
  const int wbr = li / 10;
  cudaStream_t st_chain = st;
  const bool prev_pending = T->w_pending[wbr];
  CUDA_TRY(cudaEventRecord(T->ev_wfork[wbr], st_chain));
  CUDA_TRY(cudaStreamWaitEvent(T->wside[wbr], T->ev_wfork[wbr], 0));
  st = T->wside[wbr];
  if (C == 32 && d.ks == 3 && use_mma) {
    if (int rc = wgrad32_mma(x, dy, rows, L, d.relu_in, bn, bn + C, G, d.w, d.b, st)) return rc;
  } else
#define WG(CC)                                                                                                  \
  case CC: {                                                                                                   \
    static size_t conf = 0;                                                                                    \
    if (smem > 48 * 1024 && smem > conf) {                                                                     \
      CUDA_TRY(cudaFuncSetAttribute(k_wgrad<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
      conf = smem;                                                                                             \
    }                                                                                                          \
    LAUNCH(k_wgrad<CC>, wg, 256, smem, st, x, dy, rows, L, d.ks, d.relu_in, bn, bn + C, G, d.w, d.b);          \
  } break;
  switch (C) { WG(16) WG(32) WG(64) default: MURAL_FAIL("unsupported channel count"); }
#undef WG
  st = st_chain;
  ConvLayerDev cl{T->d_Wf + int64_t(li) * 7 * C * C, T->d_const + C, T->d_const, T->d_const + C, d.ks, 0, 0};  // dgrad: two-level split MMA
  if (int rc = conv_any(C, dy, du, nullptr, nullptr, n, L, cl, 0, st)) return rc;
  CUDA_TRY(cudaMemsetAsync(stat, 0, sizeof(double) * 2 * C, st));
  const int thr = 256, rpb = thr / (C / 4);
  int grid = (int)cdiv(rows, rpb * 4);
  if (grid > 1184) grid = 1184;
  if (grid < 1) grid = 1;
  LAUNCH(k_bn_bwd_reduce, grid, thr, sizeof(double) * 8 * thr, st, du, x, rows, C, d.relu_in, bn + 2 * C, bn + 3 * C, stat);
  // `out` may be the dy the PREVIOUS layer's weight gradient is still reading (the gradient buffers rotate): join that one first
  if (prev_pending) CUDA_TRY(cudaStreamWaitEvent(st, T->ev_w[wbr], 0));
  LAUNCH(k_bn_bwd_apply, gridn(rows * (C / 4)), 256, 0, st, du, x, rows, C, d.relu_in, bn + 2 * C, bn + 3 * C, bn, stat, double(rows), add1,
         add2, out, G, d.g, d.be);
  CUDA_TRY(cudaEventRecord(T->ev_w[wbr], T->wside[wbr]));   // after the wait above: ev_w now stands for THIS layer's weight gradient
  T->w_pending[wbr] = true;
  return 0;
}
// the chain is about to overwrite a buffer the branch's latest weight gradient may still read (pool backward, stem, end of chain)
static int wait_wgrad(mural_snv_train* T, int br, cudaStream_t st) {
  if (T->w_pending[br]) {
    CUDA_TRY(cudaStreamWaitEvent(st, T->ev_w[br], 0));
    T->w_pending[br] = false;
  }
  return 0;
}

extern "C" int mural_snv_train_backward(mural_snv_train_t* T, const float* d_blob, const float* d_dlogp, float* d_grads, void* stream) {
  MURAL_CHECK(T && d_blob && d_dlogp && d_grads, "NULL argument");
  MURAL_CHECK(T->n > 0, "backward without a preceding forward");
  mural_snv_model* m = T->m;
  cudaStream_t st = (cudaStream_t)stream;
  T->w_pending[0] = T->w_pending[1] = false;
  const int64_t n = T->n;
  const int C = m->cfg.channels, ks = m->cfg.kernel_size, NC = m->cfg.n_class, H1 = m->cfg.hidden1, H2 = m->cfg.hidden2, K1 = m->k1;
  const float* P = d_blob;
  float* G = d_grads;
  CUDA_TRY(cudaMemsetAsync(G, 0, sizeof(float) * m->n_trainable, st));
  float *dl = T->gsm[0], *dm = T->gsm[1], *dg = T->gsm[2];
  LAUNCH(k_combine_bwd, gridn(n, 128), 128, 0, st, d_dlogp, T->smx, n, NC, dl, dm, dg);
  // fork as in the forward: three independent chains, each with its own scratch (gsm pairs, gbuf sets, BN-sum scratch)
  cudaStream_t const st_main = st;
  CUDA_TRY(cudaEventRecord(T->ev_fork, st_main));
  for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamWaitEvent(T->side[i], T->ev_fork, 0));
  // ---- local branch
  {
    st = T->side[1];
    float *ga = T->gsm[3], *gb = T->gsm[4];
    const int64_t W3 = off_of(m, "local_fc.0.weight"), W2 = off_of(m, "lin_layers.1.weight"), W1 = off_of(m, "lin_layers.0.weight");
    LAUNCH(k_linear_bwd_w, lbw_grid(n), 256, lbw_smem(H2, NC), st, dl, nullptr, T->d2, n, H2, NC, 0, G + W3, G + off_of(m, "local_fc.0.bias"));
    LAUNCH(k_linear_bwd_x, gridn(n * H2), 256, 0, st, dl, nullptr, P + W3, n, H2, NC, 0, ga);
    LAUNCH(k_bn1d_bwd, H2, 256, 0, st, ga, T->r2, n, H2, P, off_of(m, "bn_layers.1.weight"), T->mu2, T->is2, T->p_local, T->seed, T->d_step, 2u, gb,
           G, off_of(m, "bn_layers.1.weight"), off_of(m, "bn_layers.1.bias"));
    LAUNCH(k_linear_bwd_w, lbw_grid(n), 256, lbw_smem(H1, H2), st, gb, T->r2, T->d1, n, H1, H2, 1, G + W2, G + off_of(m, "lin_layers.1.bias"));
    LAUNCH(k_linear_bwd_x, gridn(n * H1), 256, 0, st, gb, T->r2, P + W2, n, H1, H2, 1, ga);
    LAUNCH(k_bn1d_bwd, H1, 256, 0, st, ga, T->r1, n, H1, P, off_of(m, "bn_layers.0.weight"), T->mu1, T->is1, T->p_local, T->seed, T->d_step, 1u, gb,
           G, off_of(m, "bn_layers.0.weight"), off_of(m, "bn_layers.0.bias"));
    const int NCT = m->cfg.n_cont, K1c = K1 + NCT;
    LAUNCH(k_linear_bwd_w, lbw_grid(n), 256, lbw_smem(K1c, H1), st, gb, T->r1, T->e0, n, K1c, H1, 1, G + W1, G + off_of(m, "lin_layers.0.bias"));
    LAUNCH(k_linear_bwd_x, gridn(n * K1c), 256, 0, st, gb, T->r1, P + W1, n, K1c, H1, 1, ga);
    LAUNCH(k_emb_bwd, gridn(n * K1), 256, 0, st, ga, T->cat, n, m->n_cat, T->p_emb, T->seed, T->d_step, G + off_of(m, "emb_layer.weight"), K1c);
    if (NCT > 0)   // first_bn_layer: parameter gradients only (cont_x is an input)
      LAUNCH(k_bn1d_bwd, NCT, 256, 0, st, ga + K1, T->cont, n, NCT, P, off_of(m, "first_bn_layer.weight"), T->mu0, T->is0, 0.f, T->seed, T->d_step,
             0u, (float*)nullptr, G, off_of(m, "first_bn_layer.weight"), off_of(m, "first_bn_layer.bias"), K1c);
  }
  // ---- CNN branches
  for (int br = 0; br < 2; ++br) {
    const BranchDev& B = m->br[br];
    auto& b = T->br[br];
    const std::string s = br ? "_2" : "";
    const std::string fc = br ? "distal_fc2" : "distal_fc1";
    const float* dlogit = br ? dg : dm;
    st = br ? st_main : T->side[0];
    float *ga = T->gsm[5 + 2 * br], *gb = T->gsm[6 + 2 * br];
    float *G0 = T->gbuf[br][0], *G1 = T->gbuf[br][1], *G2 = T->gbuf[br][2], *G3 = T->gbuf[br][3], *D = T->gbuf[br][4];
    LAUNCH(k_linear_bwd_w, lbw_grid(n), 256, lbw_smem(C, NC), st, dlogit, nullptr, b.gmn, n, C, NC, 0, G + off_of(m, fc + ".2.weight"),
           G + off_of(m, fc + ".2.bias"));
    LAUNCH(k_linear_bwd_x, gridn(n * C), 256, 0, st, dlogit, nullptr, P + off_of(m, fc + ".2.weight"), n, C, NC, 0, ga);
    LAUNCH(k_bn1d_bwd, C, 256, 0, st, ga, b.gm, n, C, P, off_of(m, fc + ".0.weight"), b.mu, b.is, T->p_fc, T->seed, T->d_step, 10u + br, gb, G,
           off_of(m, fc + ".0.weight"), off_of(m, fc + ".0.bias"));
    LAUNCH(k_gmax_bwd, gridn(n * B.L3 * C), 256, 0, st, gb, b.ig, b.h, G0, n, B.L3, C);  // G0 = d(conv3 out)
    const int base = br * 10;
    if (int rc = conv_train_bwd(T, P, G, base + 9, b.x3, G0, D, nullptr, nullptr, G1, n, B.L3, st)) return rc;   // G1 = d x3
    if (int rc = wait_wgrad(T, br, st)) return rc;
    LAUNCH(k_pool_bwd, gridn(n * B.L2 * C), 256, 0, st, G1, b.i3, G0, n, B.L2, B.L3, C, B.pool[2][1], B.pool[2][2]);  // G0 = d z2
    // stage 2: z2 = y1b + C8(t2b) + j2 ; t2b = C7(y1b) ; y1b = j2 + C6(t1b) ; t1b = C5(j2) ; j2 = C4(x2)
    if (int rc = conv_train_bwd(T, P, G, base + 8, b.t2b, G0, D, nullptr, nullptr, G1, n, B.L2, st)) return rc;  // G1 = d t2b
    if (int rc = conv_train_bwd(T, P, G, base + 7, b.y1b, G1, D, G0, nullptr, G2, n, B.L2, st)) return rc;        // G2 = d y1b
    if (int rc = conv_train_bwd(T, P, G, base + 6, b.t1b, G2, D, nullptr, nullptr, G1, n, B.L2, st)) return rc;  // G1 = d t1b
    if (int rc = conv_train_bwd(T, P, G, base + 5, b.j2, G1, D, G0, G2, G3, n, B.L2, st)) return rc;             // G3 = d j2
    if (int rc = conv_train_bwd(T, P, G, base + 4, b.x2, G3, D, nullptr, nullptr, G1, n, B.L2, st)) return rc;   // G1 = d x2
    if (int rc = wait_wgrad(T, br, st)) return rc;
    LAUNCH(k_pool_bwd, gridn(n * B.L1 * C), 256, 0, st, G1, b.i2, G0, n, B.L1, B.L2, C, B.pool[1][1], B.pool[1][2]);  // G0 = d z1
    // stage 1: z1 = y1 + C3(t2) + x0 ; t2 = C2(y1) ; y1 = x0 + C1(t1) ; t1 = C0(x0)
    if (int rc = conv_train_bwd(T, P, G, base + 3, b.t2, G0, D, nullptr, nullptr, G1, n, B.L1, st)) return rc;   // G1 = d t2
    if (int rc = conv_train_bwd(T, P, G, base + 2, b.y1, G1, D, G0, nullptr, G2, n, B.L1, st)) return rc;         // G2 = d y1
    if (int rc = conv_train_bwd(T, P, G, base + 1, b.t1, G2, D, nullptr, nullptr, G1, n, B.L1, st)) return rc;   // G1 = d t1
    if (int rc = conv_train_bwd(T, P, G, base + 0, b.x0, G1, D, G0, G2, G3, n, B.L1, st)) return rc;             // G3 = d x0
    // stem
    if (int rc = wait_wgrad(T, br, st)) return rc;   // joins the branch's weight-gradient stream into the chain (and so into ev_join)
    float* stem = T->d_stem + br * (16 + 2 * int64_t(ks) * 16 * C);
    float* Gt = stem + 16 + int64_t(ks) * 16 * C;
    CUDA_TRY(cudaMemsetAsync(Gt, 0, sizeof(float) * ks * 16 * C, st));
    const int off0 = br ? 0 : m->L / 2 - 100;
    const size_t smem = sizeof(float) * (size_t(ks) * 16 * C + C);
#define SB(CC)                                                                                                       \
  case CC:                                                                                                          \
    LAUNCH(k_stem_bwd<CC>, 148 * 8, 256, smem, st, G3, b.i1, T->sym, n, m->L, ks, B.L0, off0, B.L1, Gt,                  \
           G + off_of(m, "conv1" + s + ".1.bias"));                                                                 \
    break;
    switch (C) { SB(16) SB(32) SB(64) default: MURAL_FAIL("unsupported channel count"); }
#undef SB
    LAUNCH(k_stem_param_grad, 1, 256, 0, st, Gt, P, off_of(m, "conv1" + s + ".1.weight"), C, ks, stem, G, off_of(m, "conv1" + s + ".1.weight"),
           off_of(m, "conv1" + s + ".0.weight"), off_of(m, "conv1" + s + ".0.bias"));
  }
  st = st_main;
  for (int i = 0; i < 2; ++i) {  // join: the caller's stream sees every gradient
    CUDA_TRY(cudaEventRecord(T->ev_join[i], T->side[i]));
    CUDA_TRY(cudaStreamWaitEvent(st, T->ev_join[i], 0));
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_conv32_wgrad(const float* d_x, const float* d_dy, int64_t n, int32_t L, int32_t relu_in, const float* d_a,
                                  const float* d_b, float* d_dW, float* d_dbias, int32_t impl, void* stream) {
  MURAL_CHECK(d_x && d_dy && d_a && d_b && d_dW && d_dbias && n > 0 && L > 0, "bad argument");
  MURAL_CHECK(d_dbias >= d_dW, "dbias must not precede dW (both are addressed as offsets from dW)");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = n * L, b_off = d_dbias - d_dW;
  if (impl) {
    if (int rc = wgrad32_mma(d_x, d_dy, rows, L, relu_in, d_a, d_b, d_dW, 0, b_off, st)) return rc;
  } else {
    const size_t smem = sizeof(float) * (size_t(64 + 3 - 1) * (32 + 4) + size_t(64) * (32 + 4)) + sizeof(int) * 64;
    int wg = (int)cdiv(rows, 64);
    if (wg > 296) wg = 296;
    LAUNCH(k_wgrad<32>, wg, 256, smem, st, d_x, d_dy, rows, L, 3, relu_in, d_a, d_b, d_dW, int64_t(0), b_off);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_ce_sum_grad(const float* d_logp, const int32_t* d_meta, int64_t n, int32_t n_class, double* d_loss,
                                 float* d_dlogp, void* stream) {
  MURAL_CHECK(d_logp && d_meta && d_dlogp && n_class >= 2 && n_class <= 16, "bad argument");
  if (n == 0) return 0;
  LAUNCH(k_ce_grad, gridn(n), 256, 0, (cudaStream_t)stream, d_logp, d_meta, n, n_class, d_loss, d_dlogp);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_optimizer_step(int32_t kind, float* d_params, const float* d_grads, float* d_m, float* d_v, float* d_vmax,
                                    int64_t n, float lr, float weight_decay, int64_t step, float max_norm, float grad_scale,
                                    double* d_scratch, void* stream) {
  MURAL_CHECK(d_params && d_grads && d_scratch && n > 0, "bad argument");
  MURAL_CHECK(kind >= 0 && kind <= 2, "optimizer kind must be 0 (Adam), 1 (AdamW/amsgrad) or 2 (SGD nesterov)");
  MURAL_CHECK(d_m && (kind == 2 || d_v) && (kind != 1 || d_vmax), "optimizer state buffer missing");
  MURAL_CHECK(step >= 1, "step counts from 1");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(d_scratch, 0, sizeof(double), st));
  if (max_norm > 0.f) LAUNCH(k_sumsq, 296, 256, 0, st, d_grads, n, grad_scale, d_scratch);
  LAUNCH(k_opt_step, gridn(n), 256, 0, st, kind, d_params, d_grads, d_m, d_v, d_vmax, n, lr, weight_decay, step, max_norm, grad_scale,
         d_scratch, (const float*)nullptr, (const int64_t*)nullptr);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_optimizer_step_dev(int32_t kind, float* d_params, const float* d_grads, float* d_m, float* d_v, float* d_vmax,
                                        int64_t n, const float* d_lr, float weight_decay, int64_t* d_step, float max_norm,
                                        float grad_scale, double* d_scratch, void* stream) {
  MURAL_CHECK(d_params && d_grads && d_scratch && d_lr && d_step && n > 0, "bad argument");
  MURAL_CHECK(kind >= 0 && kind <= 2, "optimizer kind must be 0 (Adam), 1 (AdamW/amsgrad) or 2 (SGD nesterov)");
  MURAL_CHECK(d_m && (kind == 2 || d_v) && (kind != 1 || d_vmax), "optimizer state buffer missing");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(d_scratch, 0, sizeof(double), st));
  LAUNCH(k_bump_i64, 1, 1, 0, st, d_step);
  if (max_norm > 0.f) LAUNCH(k_sumsq, 296, 256, 0, st, d_grads, n, grad_scale, d_scratch);
  LAUNCH(k_opt_step, gridn(n), 256, 0, st, kind, d_params, d_grads, d_m, d_v, d_vmax, n, 0.f, weight_decay, int64_t(1), max_norm,
         grad_scale, d_scratch, d_lr, (const int64_t*)d_step);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
