// Training step of MuRaL-indel `UNet_Small` (MuRaL/model/model_indel.py:6-176; loop body MuRaL/training.py:404-452) as a
// tape of generic ops: this header holds the per-work-item arithmetic of every op.
//
// One source, two builds: with nvcc every op is a grid-stride kernel over work items; with g++ (-DINDEL_EMU,
// tests/emu/indel_train_emu.cpp) the same functors run in a serial loop on host memory — test infrastructure with which the
// arithmetic is checked against fp64 autograd of the oracle and the reference's own gradients without a GPU
// (tests/test_indel_train_emu.py).
// First version: correctness and structure (unit = conv -> train-mode BatchNorm -> activation (+ residuals), tape,
// gradient accumulation); the kernels are plain one-item-per-thread loops, to be tiled once profiled.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define HD __host__ __device__ __forceinline__
#ifndef INDEL_TRAIN_LAUNCH  // the library routes launches through its accounting macro (common.cuh)
#define INDEL_TRAIN_LAUNCH(name, kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif
#else
#define HD inline
#endif

namespace indel_train {

// ------------------------------------------------------------------------------------------------ execution layer
struct ConvFwd; struct ConvBwdX; struct ConvBwdW; struct MaxL;
#ifdef __CUDACC__
template <class F>
__global__ void __launch_bounds__(256) k_run(int64_t n, F f) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) f(i);
}
struct Exec {
  cudaStream_t st = nullptr;
  int64_t launches = 0;
  // Weight gradients leave the backward chain: nothing downstream reads them before the optimizer, so they run on a side stream
  // while the main stream goes on with the input gradient (at the reference's batch of 32 no kernel of the chain fills the GPU).
  // dz (the gradient in front of a unit's conv) alternates between two buffers; slot `par` may be overwritten again only after
  // the weight-gradient kernel that read it has finished (wait_w).
  cudaStream_t main_st = nullptr, side_st = nullptr;
  cudaEvent_t ev_dz = nullptr, ev_w[2] = {nullptr, nullptr};
  bool w_pending[2] = {false, false};
  void side_init() {
    if (!side_st) {
      cudaStreamCreateWithFlags(&side_st, cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&ev_dz, cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_w[0], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&ev_w[1], cudaEventDisableTiming);
    }
  }
  void wait_w(int par) { if (w_pending[par]) { cudaStreamWaitEvent(st, ev_w[par], 0); w_pending[par] = false; } }
  void begin_w() { side_init(); main_st = st; cudaEventRecord(ev_dz, st); cudaStreamWaitEvent(side_st, ev_dz, 0); st = side_st; }
  void end_w(int par) { cudaEventRecord(ev_w[par], side_st); w_pending[par] = true; st = main_st; }
  template <class F> void run(int64_t n, const F& f) {
    if (n <= 0) return;
    int64_t g = (n + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;  // grid-stride: a multiple of the SM count
    INDEL_TRAIN_LAUNCH(F::kName, (k_run<F>), (unsigned)g, 256, st, n, f);
    ++launches;
  }
  void* alloc(size_t b) { void* p = nullptr; cudaMalloc(&p, b); return p; }
  void free_(void* p) { cudaFree(p); }
  void zero(void* p, size_t b) { cudaMemsetAsync(p, 0, b, st); }
  // shared-memory-tiled fast paths (indel_train_tiled.cuh); false = not applicable, the caller runs the functor
  bool conv_fwd(const ConvFwd& f);
  bool conv_bwd_x(const ConvBwdX& f, float* scratch);
  bool conv_bwd_w(const ConvBwdW& f);
  bool max_rows(const MaxL& f, int64_t rows);
  template <class F> bool row_reduce(const F& f, int64_t rows, double* stat);   // BnStats / UnitBwdReduce: one CTA (or warp) per row
};
template <class T> HD void atomic_add(T* p, T v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;  // host instantiation of the functors is never executed in the CUDA build
#endif
}
#else
struct Exec {
  int64_t launches = 0;
  void wait_w(int) {}
  void begin_w() {}
  void end_w(int) {}
  template <class F> void run(int64_t n, const F& f) { for (int64_t i = 0; i < n; ++i) f(i); ++launches; }
  void* alloc(size_t b) { return calloc(b ? b : 1, 1); }
  void free_(void* p) { free(p); }
  void zero(void* p, size_t b) { memset(p, 0, b); }
  bool conv_fwd(const ConvFwd&) { return false; }
  bool conv_bwd_x(const ConvBwdX&, float*) { return false; }
  bool conv_bwd_w(const ConvBwdW&) { return false; }
  bool max_rows(const MaxL&, int64_t) { return false; }
  template <class F> bool row_reduce(const F&, int64_t, double*) { return false; }
};
template <class T> inline void atomic_add(T* p, T v) { *p += v; }
#endif

enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_SOFTPLUS = 3 };

HD float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }
HD float act_fwd(float z, int act) {
  if (act == ACT_SILU) return z * sigmoidf_(z);
  if (act == ACT_RELU) return z > 0.f ? z : 0.f;
  if (act == ACT_SOFTPLUS) return z > 20.f ? z : log1pf(expf(z));  // nn.Softplus(beta=1, threshold=20)
  return z;
}
HD float act_bwd(float z, int act) {
  if (act == ACT_SILU) { const float s = sigmoidf_(z); return s * (1.f + z * (1.f - s)); }
  if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (act == ACT_SOFTPLUS) return z > 20.f ? 1.f : sigmoidf_(z);
  return 1.f;
}
HD uint32_t hash32(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return uint32_t(x);
}
// keep-scale of element idx of a dropout layer at (seed, step): 0 or 1/(1-p)
HD float drop_scale(float p, uint64_t seed, uint32_t step, uint64_t idx) {
  if (p <= 0.f) return 1.f;
  const uint32_t u = hash32(seed ^ (uint64_t(step) << 40) ^ idx);
  return (u * (1.0f / 4294967296.0f)) < p ? 0.f : 1.f / (1.f - p);
}

// ------------------------------------------------------------------------------------------------ ops (work-item functors)
struct BumpStep {   // one work item: advances the dropout stream position at the start of a training forward
  static constexpr const char* kName = "k_indel_train<BumpStep>";
  uint32_t* p;
  HD void operator()(int64_t) const { *p += 1u; }
};
struct ConvDims { int B, Cin, Lin, Cout, Lout, k, stride, pad, up; };
constexpr int CONV_W_SPLIT = 32;  // work items per (co, ci, t, b) in the weight gradient (one warp walks a row pair coalesced)  // input is read through a nearest upsample by `up`

struct ConvFwd {
  static constexpr const char* kName = "k_indel_train<ConvFwd>";  // item: output element (b, co, lo)
  const float* x; const float* W; const float* bias; float* y; ConvDims d;
  HD void operator()(int64_t i) const {
    const int lo = int(i % d.Lout), co = int((i / d.Lout) % d.Cout), b = int(i / (int64_t(d.Lout) * d.Cout));
    float acc = bias ? bias[co] : 0.f;
    const int Lv = d.Lin * d.up;
    for (int ci = 0; ci < d.Cin; ++ci) {
      const float* xr = x + (int64_t(b) * d.Cin + ci) * d.Lin;
      const float* wr = W + (int64_t(co) * d.Cin + ci) * d.k;
      for (int t = 0; t < d.k; ++t) {
        const int j = lo * d.stride + t - d.pad;
        if (j >= 0 && j < Lv) acc += wr[t] * xr[j / d.up];
      }
    }
    y[i] = acc;
  }
};
struct ConvBwdX {
  static constexpr const char* kName = "k_indel_train<ConvBwdX>";  // item: input element (b, ci, jin); dx += W^T dy
  const float* dy; const float* W; float* dx; ConvDims d;
  HD void operator()(int64_t i) const {
    const int jin = int(i % d.Lin), ci = int((i / d.Lin) % d.Cin), b = int(i / (int64_t(d.Lin) * d.Cin));
    float acc = 0.f;
    for (int u = 0; u < d.up; ++u) {
      const int j = jin * d.up + u;
      for (int t = 0; t < d.k; ++t) {
        const int num = j + d.pad - t;
        if (num < 0 || num % d.stride) continue;
        const int lo = num / d.stride;
        if (lo >= d.Lout) continue;
        for (int co = 0; co < d.Cout; ++co)
          acc += W[(int64_t(co) * d.Cin + ci) * d.k + t] * dy[(int64_t(b) * d.Cout + co) * d.Lout + lo];
      }
    }
    dx[i] += acc;
  }
};
struct ConvBwdW {
  static constexpr const char* kName = "k_indel_train<ConvBwdW>";
  // item: (co, ci, t, b, s) with s the fastest index: item s takes output positions s, s + ROW_SPLIT, ... so that the threads
  // of a warp read neighbouring positions of the same rows; dW += sum_lo dy * x, db += sum_lo dy
  const float* x; const float* dy; float* dW; float* db; ConvDims d;
  HD void operator()(int64_t i) const {
    const int s0 = int(i % CONV_W_SPLIT);
    const int64_t r = i / CONV_W_SPLIT;
    const int b = int(r % d.B), t = int((r / d.B) % d.k), ci = int((r / (int64_t(d.B) * d.k)) % d.Cin),
              co = int(r / (int64_t(d.B) * d.k * d.Cin));
    if (s0 >= d.Lout) return;
    const float* dr = dy + (int64_t(b) * d.Cout + co) * d.Lout;
    const float* xr = x + (int64_t(b) * d.Cin + ci) * d.Lin;
    const int Lv = d.Lin * d.up;
    float acc = 0.f, sb = 0.f;
    for (int lo = s0; lo < d.Lout; lo += CONV_W_SPLIT) {
      const int j = lo * d.stride + t - d.pad;
      if (j >= 0 && j < Lv) acc += dr[lo] * xr[j / d.up];
      sb += dr[lo];
    }
    atomic_add(dW + (int64_t(co) * d.Cin + ci) * d.k + t, acc);
    if (db && ci == 0 && t == 0) atomic_add(db + co, sb);
  }
};
constexpr int ROW_SPLIT = 64;  // work items per (b, c) row in the row reductions: item s takes positions s, s + 64, ... so that
                               // neighbouring threads read neighbouring addresses and a batch of 8-channel rows still fills the GPU
struct BnStats {
  static constexpr const char* kName = "k_indel_train<BnStats>";
  // item: (row (b, c), s); stat[c] += sum, stat[C + c] += sum of squares (double)
  const float* x; double* stat; int C, L;
  HD int channels() const { return C; }
  HD int length() const { return L; }
  HD bool reduces() const { return true; }
  HD void at(int64_t row, int l, double& s, double& q) const { const float v = x[row * L + l]; s += v; q += double(v) * v; }
  HD void operator()(int64_t i) const {
    const int64_t row = i / ROW_SPLIT;
    const int s0 = int(i % ROW_SPLIT), c = int(row % C);
    if (s0 >= L) return;
    double s = 0, q = 0;
    for (int l = s0; l < L; l += ROW_SPLIT) at(row, l, s, q);
    atomic_add(stat + c, s);
    atomic_add(stat + C + c, q);
  }
};
struct BnFinalize {
  static constexpr const char* kName = "k_indel_train<BnFinalize>";  // item: channel; nn.BatchNorm1d training mode (biased var for the batch, unbiased for running_var, momentum 0.1)
  const double* stat; double N; int C; float* rm; float* rv; float* mean; float* invstd;
  HD void operator()(int64_t c) const {
    const double m = stat[c] / N;
    double var = stat[C + c] / N - m * m;
    if (var < 0) var = 0;
    mean[c] = float(m);
    invstd[c] = float(1.0 / sqrt(var + 1e-5));
    const double unb = N > 1 ? var * N / (N - 1) : var;
    rm[c] = 0.9f * rm[c] + 0.1f * float(m);
    rv[c] = 0.9f * rv[c] + 0.1f * float(unb);
  }
};
struct UnitOut {
  static constexpr const char* kName = "k_indel_train<UnitOut>";  // item: element; y = drop * act(bn(t)) + res1 + res2
  const float* t; const float* mean; const float* invstd; const float* gamma; const float* beta;  // gamma == nullptr: no BN
  const float* res1; const float* res2; float* y; int C, L, act; float p; uint64_t seed;
  const uint32_t* step_p;   // dropout stream position, read from memory so that a captured step graph advances it (BumpStep)
  HD float z_of(int64_t i) const {
    if (!gamma) return t[i];
    const int c = int((i / L) % C);
    return (t[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
  }
  HD void operator()(int64_t i) const {
    float v = act_fwd(z_of(i), act) * drop_scale(p, seed, *step_p, uint64_t(i));
    if (res1) v += res1[i];
    if (res2) v += res2[i];
    y[i] = v;
  }
};
struct UnitBwdReduce {
  static constexpr const char* kName = "k_indel_train<UnitBwdReduce>";
  // item: (row (b, c), s); dz = dy * drop * act'(z); s1 = sum dz, s2 = sum dz * xhat; also stores dz
  UnitOut u; const float* dy; float* dz; double* stat;
  HD int channels() const { return u.C; }
  HD int length() const { return u.L; }
  HD bool reduces() const { return u.gamma != nullptr; }
  HD void at(int64_t row, int l, double& s1, double& s2) const {
    const int c = int(row % u.C);
    const int64_t i = row * u.L + l;
    const float g = dy[i] * drop_scale(u.p, u.seed, *u.step_p, uint64_t(i)) * act_bwd(u.z_of(i), u.act);
    dz[i] = g;
    if (u.gamma) { s1 += g; s2 += double(g) * ((u.t[i] - u.mean[c]) * u.invstd[c]); }
  }
  HD void operator()(int64_t it) const {
    const int64_t row = it / ROW_SPLIT;
    const int s0 = int(it % ROW_SPLIT), c = int(row % u.C);
    if (s0 >= u.L) return;
    double s1 = 0, s2 = 0;
    for (int l = s0; l < u.L; l += ROW_SPLIT) at(row, l, s1, s2);
    if (u.gamma) { atomic_add(stat + c, s1); atomic_add(stat + u.C + c, s2); }
  }
};
struct UnitBwdApply {
  static constexpr const char* kName = "k_indel_train<UnitBwdApply>";  // item: element; dt = gamma * invstd * (dz - s1/N - xhat * s2/N)  (in place on dz)
  UnitOut u; float* dz; const double* stat; double N;
  HD void operator()(int64_t i) const {
    const int c = int((i / u.L) % u.C);
    const float xh = (u.t[i] - u.mean[c]) * u.invstd[c];
    dz[i] = u.gamma[c] * u.invstd[c] * (dz[i] - float(stat[c] / N) - xh * float(stat[u.C + c] / N));
  }
};
struct BnParamGrad {
  static constexpr const char* kName = "k_indel_train<BnParamGrad>";  // item: channel
  const double* stat; int C; float* dgamma; float* dbeta;
  HD void operator()(int64_t c) const { dgamma[c] += float(stat[C + c]); dbeta[c] += float(stat[c]); }
};
struct AddTo {
  static constexpr const char* kName = "k_indel_train<AddTo>"; const float* src; float* dst; HD void operator()(int64_t i) const { dst[i] += src[i]; } };
struct FlipCL {
  static constexpr const char* kName = "k_indel_train<FlipCL>"; const float* x; float* y; int C, L;  // y[b, c, l] = x[b, C-1-c, L-1-l]  (torch.flip(x, [1, 2]))
  HD void operator()(int64_t i) const {
    const int l = int(i % L), c = int((i / L) % C); const int64_t b = i / (int64_t(L) * C);
    y[i] = x[(b * C + (C - 1 - c)) * L + (L - 1 - l)];
  } };
struct AddFlipL {
  static constexpr const char* kName = "k_indel_train<AddFlipL>"; const float* a; const float* bb; float* o; int L;  // o = a + flip(bb, [2])
  HD void operator()(int64_t i) const { const int l = int(i % L); o[i] = a[i] + bb[i - l + (L - 1 - l)]; } };
struct AddFlipLBwd {
  static constexpr const char* kName = "k_indel_train<AddFlipLBwd>"; const float* d_o; float* da; float* db; int L;
  HD void operator()(int64_t i) const { const int l = int(i % L); da[i] += d_o[i]; db[i - l + (L - 1 - l)] += d_o[i]; } };
struct UpReduce {
  static constexpr const char* kName = "k_indel_train<UpReduce>";  // item: input element (b, ci, jin); dx += sum of the `up` virtual positions it was repeated to
  const float* dxv; float* dx; int up;
  HD void operator()(int64_t i) const {
    float s = 0.f;
    for (int u = 0; u < up; ++u) s += dxv[i * up + u];
    dx[i] += s;
  } };
struct MaxL {
  static constexpr const char* kName = "k_indel_train<MaxL>"; const float* x; float* y; int32_t* arg; int L;  // item: row (b, c)
  HD void operator()(int64_t r) const {
    const float* p = x + r * L; int a = 0; float m = p[0];
    for (int l = 1; l < L; ++l) if (p[l] > m) { m = p[l]; a = l; }
    y[r] = m; arg[r] = a;
  } };
struct MaxLBwd {
  static constexpr const char* kName = "k_indel_train<MaxLBwd>"; const float* dy; const int32_t* arg; float* dx; int L;
  HD void operator()(int64_t r) const { dx[r * L + arg[r]] += dy[r]; } };
struct CeGrad {
  static constexpr const char* kName = "k_indel_train<CeGrad>";  // CrossEntropyLoss(reduction='sum') on the network output (training.py:327,425); item: site
  const float* out; const int32_t* label; int NC; float* dout; double* loss;
  HD void operator()(int64_t i) const {
    const float* p = out + i * NC;
    float mx = p[0]; for (int o = 1; o < NC; ++o) mx = p[o] > mx ? p[o] : mx;
    float s = 0.f; for (int o = 0; o < NC; ++o) s += expf(p[o] - mx);
    const int y = label[i];
    for (int o = 0; o < NC; ++o) dout[i * NC + o] = expf(p[o] - mx) / s - (o == y ? 1.f : 0.f);
    atomic_add(loss, double(-(p[y] - mx - logf(s))));
  } };

}  // namespace indel_train

#ifdef __CUDACC__
#include "indel_train_tiled.cuh"
namespace indel_train {
#ifndef INDEL_TRAIN_LAUNCH_SMEM
#define INDEL_TRAIN_LAUNCH_SMEM(name, kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
namespace tiled {
inline bool tileable(const ConvDims& d) {
  return d.stride == 1 && d.Lout >= 256 && (d.k == 1 || d.k == 5 || d.k == 7) && d.pad == (d.k - 1) / 2 && d.Lout == d.Lin * d.up;
}
// At the reference's batch of 32 a (tile, sample) grid of the deeper levels is a few dozen CTAs that each walk all output-channel
// chunks in turn: split the chunks over grid.z until the grid has ~4 CTAs per SM (the input tile is re-staged per z slice).
inline int z_split(int base_ctas, int Cout) {
  const int chunks = (Cout + 7) / 8;
  int z = (4 * 148 + base_ctas - 1) / (base_ctas > 0 ? base_ctas : 1);
  if (z > chunks) z = chunks;
  return z < 1 ? 1 : z;
}
template <int K> inline bool launch_conv(const ConvT& a, int B, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t(a.Cin) * XS + size_t(a.Cin) * K * 8);
  if (smem > 200 * 1024) return false;
  static size_t conf = 0;
  if (smem > 48 * 1024 && smem > conf) {
    if (cudaFuncSetAttribute(k_conv_tiled<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    conf = smem;
  }
  dim3 grid((unsigned)((a.Lout + TL - 1) / TL), (unsigned)B, (unsigned)z_split(int((a.Lout + TL - 1) / TL) * B, a.Cout));
  INDEL_TRAIN_LAUNCH_SMEM("k_indel_conv_tiled", (k_conv_tiled<K>), grid, 128, smem, st, a);
  return true;
}
template <int K> inline bool launch_conv_strided(const ConvT& a, int stride, int B, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t(a.Cin) * ((TLS - 1) * stride + K) + size_t(a.Cin) * K * 8);
  if (smem > 200 * 1024) return false;
  static size_t conf = 0;
  if (smem > 48 * 1024 && smem > conf) {
    if (cudaFuncSetAttribute(k_conv_strided<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    conf = smem;
  }
  dim3 grid((unsigned)((a.Lout + TLS - 1) / TLS), (unsigned)B, (unsigned)z_split(int((a.Lout + TLS - 1) / TLS) * B, a.Cout));
  INDEL_TRAIN_LAUNCH_SMEM("k_indel_conv_strided", (k_conv_strided<K>), grid, 128, smem, st, a, stride);
  return true;
}
inline bool run_conv(const ConvT& a, int K, int B, cudaStream_t st) {
  return K == 7 ? launch_conv<7>(a, B, st) : (K == 5 ? launch_conv<5>(a, B, st) : launch_conv<1>(a, B, st));
}
template <int K> inline bool launch_wgrad(WgradT a, int B, cudaStream_t st) {
  a.TLe = a.Lout < TL ? a.Lout : TL;
  a.n_tiles = (a.Lout + a.TLe - 1) / a.TLe;
  a.RSx = ((a.TLe - 1) * a.stride + K) | 1;
  a.RSd = a.TLe | 1;
  const size_t smem = sizeof(float) * (size_t(a.Cin) * a.RSx + size_t(a.Cout) * a.RSd);
  if (smem > 200 * 1024 || (a.Cin * a.Cout > 256 * 8 && a.n_tiles > 1)) return false;   // pair batches re-stage the tile: short rows only
  static size_t conf = 0;
  if (smem > 48 * 1024 && smem > conf) {
    if (cudaFuncSetAttribute(k_wgrad_tiled<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    conf = smem;
  }
  int G = 296 / (B > 0 ? B : 1);
  if (G < 1) G = 1;
  if (G > a.n_tiles) G = a.n_tiles;
  dim3 grid((unsigned)G, (unsigned)B);
  INDEL_TRAIN_LAUNCH_SMEM("k_indel_wgrad_tiled", (k_wgrad_tiled<K>), grid, 256, smem, st, a);
  return true;
}
}  // namespace tiled

inline bool Exec::conv_fwd(const ConvFwd& f) {
  const ConvDims& d = f.d;
  tiled::ConvT a{f.x, f.W, f.bias, f.y, d.Cin, d.Lin, d.Cout, d.Lout, d.pad, d.up, d.Cin * d.k, d.k, 0, 0};
  bool ok = false;
  if (tiled::tileable(d)) ok = tiled::run_conv(a, d.k, d.B, st);
  else if ((d.k == 1 || d.k == 5 || d.k == 7) && d.pad == (d.k - 1) / 2)       // strided / short rows
    ok = d.k == 7 ? tiled::launch_conv_strided<7>(a, d.stride, d.B, st)
                  : (d.k == 5 ? tiled::launch_conv_strided<5>(a, d.stride, d.B, st) : tiled::launch_conv_strided<1>(a, d.stride, d.B, st));
  if (ok) ++launches;
  return ok;
}
inline bool Exec::conv_bwd_x(const ConvBwdX& f, float* scratch) {
  const ConvDims& d = f.d;
  if (d.stride == 1 && d.Lout < 256 && d.Lout == d.Lin * d.up && (d.k == 1 || d.k == 5 || d.k == 7) && d.pad == (d.k - 1) / 2 &&
      (d.up == 1 || scratch)) {   // short rows (the deep levels): the generic tile kernel on the transposed / flipped weights
    tiled::ConvT a{f.dy, f.W, nullptr, d.up == 1 ? f.dx : scratch, d.Cout, d.Lout, d.Cin, d.Lout, d.pad, 1, d.k, d.Cin * d.k, 1, d.up == 1 ? 1 : 0};
    const bool ok = d.k == 7 ? tiled::launch_conv_strided<7>(a, 1, d.B, st)
                             : (d.k == 5 ? tiled::launch_conv_strided<5>(a, 1, d.B, st) : tiled::launch_conv_strided<1>(a, 1, d.B, st));
    if (!ok) return false;
    ++launches;
    if (d.up > 1) run(int64_t(d.B) * d.Cin * d.Lin, UpReduce{scratch, f.dx, d.up});
    return true;
  }
  if (!tiled::tileable(d) || (d.up > 1 && !scratch)) return false;
  // correlation over dy with the channel axes swapped and the taps flipped; with an upsample in front the result is the
  // gradient of the virtual upsampled input (scratch), folded back by UpReduce
  tiled::ConvT a{f.dy, f.W, nullptr, d.up == 1 ? f.dx : scratch, d.Cout, d.Lout, d.Cin, d.Lout, d.pad, 1, d.k, d.Cin * d.k, 1, d.up == 1 ? 1 : 0};
  if (!tiled::run_conv(a, d.k, d.B, st)) return false;
  ++launches;
  if (d.up > 1) run(int64_t(d.B) * d.Cin * d.Lin, UpReduce{scratch, f.dx, d.up});
  return true;
}
inline bool Exec::conv_bwd_w(const ConvBwdW& f) {
  const ConvDims& d = f.d;
  if (!(d.k == 1 || d.k == 5 || d.k == 7) || d.pad != (d.k - 1) / 2 || d.Lout < 1) return false;   // any stride, any length
  tiled::WgradT a{f.x, f.dy, f.dW, f.db, d.Cin, d.Lin, d.Cout, d.Lout, d.pad, d.up, d.stride, 0, 0, 0, 0};
  const bool ok = d.k == 7 ? tiled::launch_wgrad<7>(a, d.B, st) : (d.k == 5 ? tiled::launch_wgrad<5>(a, d.B, st) : tiled::launch_wgrad<1>(a, d.B, st));
  if (ok) ++launches;
  return ok;
}
// per-row sums of a BatchNorm pass: threads of a row walk it coalesced, the row's two sums are reduced by shuffles (+ shared
// memory across warps) and leave as ONE pair of double atomics per row — the work-item form issues 64 pairs per row onto
// 2*C addresses, which serialises (2.5 ms per step at batch 32)
// segs > 1 (long rows, few of them — the top levels at batch 32 are 256-512 rows of 8000): a row is cut into `segs` pieces that
// go to different CTAs, each leaving its own pair of atomics.
template <class F, int TPR>
__global__ void __launch_bounds__(256) k_row_reduce(F f, int64_t rows_in, double* stat, int segs) {
  constexpr int RPB = 256 / TPR;   // rows per block
  __shared__ double sh[2][8];
  const int sub = threadIdx.x / TPR, lt = threadIdx.x % TPR;
  const int L = f.length(), C = f.channels();
  const int Ls = (L + segs - 1) / segs;
  const int64_t rows = rows_in * segs;   // virtual rows = (row, segment)
  for (int64_t row0 = int64_t(blockIdx.x) * RPB; row0 < rows; row0 += int64_t(gridDim.x) * RPB) {
    const int64_t vrow = row0 + sub;
    const int64_t row = vrow / segs;
    const int seg = int(vrow - row * segs);
    const int l_end = (seg + 1) * Ls < L ? (seg + 1) * Ls : L;
    double s = 0, q = 0;
    if (vrow < rows)
      for (int l = seg * Ls + lt; l < l_end; l += TPR) f.at(row, l, s, q);
    if (!f.reduces()) continue;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (TPR == 32) {
      if (lt == 0 && vrow < rows) { atomicAdd(stat + int(row % C), s); atomicAdd(stat + C + int(row % C), q); }
    } else {
      __syncthreads();
      if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
      __syncthreads();
      if (threadIdx.x == 0 && vrow < rows) {
        double ts = 0, tq = 0;
        for (int w = 0; w < 8; ++w) { ts += sh[0][w]; tq += sh[1][w]; }
        atomicAdd(stat + int(row % C), ts);
        atomicAdd(stat + C + int(row % C), tq);
      }
    }
  }
}
template <class F> inline bool Exec::row_reduce(const F& f, int64_t rows, double* stat) {
  if (rows <= 0) return true;
  if (f.length() > 256) {
    int segs = int((148 * 8) / rows);                          // aim at ~8 CTAs per SM
    const int max_segs = f.length() / 512 > 0 ? f.length() / 512 : 1;   // at least 512 positions (2 per thread) per piece
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
    const int64_t g = rows * segs < 148 * 8 ? rows * segs : 148 * 8;
    INDEL_TRAIN_LAUNCH_SMEM(F::kName, (k_row_reduce<F, 256>), (unsigned)g, 256, 0, st, f, rows, stat, segs);
  } else {
    int64_t g = (rows + 7) / 8;
    if (g > 148 * 8) g = 148 * 8;
    INDEL_TRAIN_LAUNCH_SMEM(F::kName, (k_row_reduce<F, 32>), (unsigned)g, 256, 0, st, f, rows, stat, 1);
  }
  ++launches;
  return true;
}
inline bool Exec::max_rows(const MaxL& f, int64_t rows) {
  INDEL_TRAIN_LAUNCH_SMEM("k_indel_max_rows", tiled::k_max_rows, (unsigned)((rows + 7) / 8), 256, 0, st, f.x, f.y, f.arg, rows, f.L);
  ++launches;
  return true;
}
}  // namespace indel_train
#endif
