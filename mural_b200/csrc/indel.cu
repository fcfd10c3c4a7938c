// MuRaL-indel UNet_Small eval forward (MuRaL/model/model_indel.py:21-176), fp32 CUDA-core path.
//
// Every BatchNorm of the U-Net sits directly AFTER a convolution, so in eval mode it folds exactly into that
// convolution's weights and bias (no padding subtlety as in Network2).  The network then is a sequence of
// generic Conv1d ops with fused prologue (nearest-neighbour upsampling of the input by an integer factor) and
// epilogue (bias, SiLU / ReLU / Softplus, up to two residual adds), on channels-last activations [site][L][C].
// The one-hot input never exists: the stem reads symbols from the packed genome; with use_reverse the
// strand-symmetric stem  conv(x) + flip(conv(flip(x,[1,2])),[2])  (model_indel.py:154-155) is evaluated as
// two table lookups per tap (a channel-flipped one-hot column is the one-hot column of the complement symbol).
#include <float.h>
#include <math.h>
#include <algorithm>
#include <string.h>

#include "indel_tc.cuh"
#include "snv_model.cuh"

namespace mural {
namespace indel {

enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2, ACT_SOFTPLUS = 3 };

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_SILU) return __fdividef(x, 1.f + __expf(-x));  // fast intrinsics: ~1e-6 relative, far inside the 1e-3 gate; the
                                                                  // precise expf + division cost as much as the layer's FMAs
  if (act == ACT_RELU) return fmaxf(x, 0.f);
  if (act == ACT_SOFTPLUS) return x > 20.f ? x : log1pf(expf(x));  // nn.Softplus(beta=1, threshold=20)
  return x;
}

constexpr int TP_MIN = 64;  // output positions per CTA for wide layers; narrow layers take more so that all 128 threads have an item

struct ConvOp {
  const float* W;     // [ks][Cin][Cout], BatchNorm folded in
  const float* bias;  // [Cout]
  int Cin, Cout, ks, stride, up, act;
};

// out[s][p][co] = act(bias[co] + sum_t sum_ci W[t][ci][co] * in[s][(p*stride + t - ks/2) / up][ci]) (+ res1)(+ res2)
// positions are taken in the (virtually) upsampled input of length Lin*up; outside it the input is zero.
__global__ void __launch_bounds__(128) k_conv_gen(const float* __restrict__ in, float* __restrict__ out,
                                                  const float* __restrict__ res1, const float* __restrict__ res2, int Lin, int Lout,
                                                  ConvOp P, int TP) {
  extern __shared__ __align__(16) float sm[];
  const int Cin = P.Cin, Cout = P.Cout, ks = P.ks, half = ks / 2;
  float* ws = sm;                                 // [ks*Cin*Cout]
  const int xrows = (TP - 1) * P.stride + ks;
  const int xs_stride = Cin + 4;                  // Cin % 4 == 0: rows stay 16-byte aligned for float4 loads
  float* xs = ws + ((ks * Cin * Cout + 3) & ~3);  // [xrows][Cin+4]
  const int tid = threadIdx.x;
  const int64_t site = blockIdx.y;
  const int p0 = blockIdx.x * TP;
  for (int e = tid; e < ks * Cin * Cout; e += 128) ws[e] = P.W[e];
  const int Lv = Lin * P.up;  // virtual (upsampled) input length
  const float* ins = in + site * int64_t(Lin) * Cin;
  {  // (k, ci) advance by 128 elements per step without a division per element
    const int dk = 128 / Cin, dci = 128 - dk * Cin;
    int k = tid / Cin, ci = tid - k * Cin;
    for (int e = tid; e < xrows * Cin; e += 128) {
      const int q = p0 * P.stride - half + k;
      float v = 0.f;
      if (q >= 0 && q < Lv) v = ins[int64_t(P.up == 1 ? q : q / P.up) * Cin + ci];
      xs[k * xs_stride + ci] = v;
      k += dk; ci += dci;
      if (ci >= Cin) { ci -= Cin; ++k; }
    }
  }
  __syncthreads();
  const int CG = Cout >> 2;
  for (int item = tid; item < (TP / 4) * CG; item += 128) {
    const int pg = item / CG, cg = item - pg * CG;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int t = 0; t < ks; ++t) {
      const float* wt = ws + t * Cin * Cout + 4 * cg;
      const float* xt = xs + (pg * 4 * P.stride + t) * xs_stride;
      for (int ci = 0; ci < Cin; ci += 4) {  // 4 input channels per step: 8 LDS.128 feed 64 FMAs (same summation order as ci = 0, 1, 2, ...)
        const float4 w0 = *reinterpret_cast<const float4*>(wt + ci * Cout);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + (ci + 1) * Cout);
        const float4 w2 = *reinterpret_cast<const float4*>(wt + (ci + 2) * Cout);
        const float4 w3 = *reinterpret_cast<const float4*>(wt + (ci + 3) * Cout);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(xt + i * P.stride * xs_stride + ci);
          acc[i][0] = fmaf(x.x, w0.x, acc[i][0]); acc[i][1] = fmaf(x.x, w0.y, acc[i][1]);
          acc[i][2] = fmaf(x.x, w0.z, acc[i][2]); acc[i][3] = fmaf(x.x, w0.w, acc[i][3]);
          acc[i][0] = fmaf(x.y, w1.x, acc[i][0]); acc[i][1] = fmaf(x.y, w1.y, acc[i][1]);
          acc[i][2] = fmaf(x.y, w1.z, acc[i][2]); acc[i][3] = fmaf(x.y, w1.w, acc[i][3]);
          acc[i][0] = fmaf(x.z, w2.x, acc[i][0]); acc[i][1] = fmaf(x.z, w2.y, acc[i][1]);
          acc[i][2] = fmaf(x.z, w2.z, acc[i][2]); acc[i][3] = fmaf(x.z, w2.w, acc[i][3]);
          acc[i][0] = fmaf(x.w, w3.x, acc[i][0]); acc[i][1] = fmaf(x.w, w3.y, acc[i][1]);
          acc[i][2] = fmaf(x.w, w3.z, acc[i][2]); acc[i][3] = fmaf(x.w, w3.w, acc[i][3]);
        }
      }
    }
    const float4 b = *reinterpret_cast<const float4*>(P.bias + 4 * cg);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p = p0 + pg * 4 + i;
      if (p >= Lout) continue;
      const int64_t o = (site * Lout + p) * int64_t(Cout) + 4 * cg;
      float4 v = make_float4(apply_act(acc[i][0] + b.x, P.act), apply_act(acc[i][1] + b.y, P.act), apply_act(acc[i][2] + b.z, P.act),
                             apply_act(acc[i][3] + b.w, P.act));
      if (res1) { const float4 r = *reinterpret_cast<const float4*>(res1 + o); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
      if (res2) { const float4 r = *reinterpret_cast<const float4*>(res2 + o); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
      *reinterpret_cast<float4*>(out + o) = v;
    }
  }
}

__constant__ float c_oh[16][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}, {.5f, 0, .5f, 0}, {0, .5f, 0, .5f},
                                  {.5f, .5f, 0, 0}, {0, .5f, .5f, 0}, {.5f, 0, 0, .5f}, {0, 0, .5f, .5f},
                                  {0, (float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3)}, {(float)(1.0 / 3), 0, (float)(1.0 / 3), (float)(1.0 / 3)},
                                  {(float)(1.0 / 3), (float)(1.0 / 3), 0, (float)(1.0 / 3)}, {(float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3), 0},
                                  {.25f, .25f, .25f, .25f}, {0, 0, 0, 0}};

// stem: symbols -> network input [site][L][4].  T == NULL: plain one-hot.  Otherwise the use_reverse stem:
//   x'[p][co] = 2*bias[co] + sum_t T[t][sym[p+t-h]][co] + T[t][comp(sym[p-t+h])][co]     (out-of-range taps: 0)
__global__ void __launch_bounds__(256) k_indel_stem(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta,
                                                    const uint8_t* __restrict__ sym_in, int R, int L, int ks,
                                                    const float* __restrict__ T, const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smr[];
  float* sT = reinterpret_cast<float*>(smr);  // [ks][16][4]
  uint8_t* sym = smr + sizeof(float) * ks * 16 * 4;
  const int64_t site = blockIdx.x;
  if (T) for (int e = threadIdx.x; e < ks * 64; e += blockDim.x) sT[e] = T[e];
  if (sym_in) {
    for (int i = threadIdx.x; i < L; i += blockDim.x) sym[i] = sym_in[site * L + i];
  } else {
    const int m = meta[site];
    load_window(G, int(uint32_t(m) >> 8), int64_t(pos[site]) - R + 1, L, m & 1, sym);  // indel window starts at start-R+1
  }
  __syncthreads();
  const int half = ks / 2;
  float4* o4 = reinterpret_cast<float4*>(out) + site * int64_t(L);
  for (int p = threadIdx.x; p < L; p += blockDim.x) {
    float4 v;
    if (!T) {
      const int s = sym[p];
      v = make_float4(c_oh[s][0], c_oh[s][1], c_oh[s][2], c_oh[s][3]);
    } else {
      v = make_float4(2.f * bias[0], 2.f * bias[1], 2.f * bias[2], 2.f * bias[3]);
      for (int t = 0; t < ks; ++t) {
        const int q1 = p + t - half, q2 = p - t + half;
        if (q1 >= 0 && q1 < L) {
          const float4 w = *reinterpret_cast<const float4*>(sT + (t * 16 + sym[q1]) * 4);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        if (q2 >= 0 && q2 < L) {
          const float4 w = *reinterpret_cast<const float4*>(sT + (t * 16 + comp_sym(sym[q2])) * 4);
          v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
      }
    }
    o4[p] = v;
  }
}

// torch.max over L of softplus(out_conv) -> BN(out_fc.0) folded into Linear(out_fc.2) -> Softplus
__global__ void __launch_bounds__(128) k_indel_head(const float* __restrict__ x, int64_t n, int L, int C, const float* __restrict__ Wfc,
                                                    const float* __restrict__ bfc, int NC, float* __restrict__ out) {
  __shared__ float red[128];
  __shared__ float gm[64];
  const int64_t site = blockIdx.x;
  const int c = threadIdx.x % C, part = threadIdx.x / C, parts = 128 / C;
  float mx = -FLT_MAX;
  if (part < parts)
    for (int p = part; p < L; p += parts) mx = fmaxf(mx, x[(site * L + p) * C + c]);
  red[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x < C) {
    for (int k = 1; k < parts; ++k) mx = fmaxf(mx, red[k * C + threadIdx.x]);
    gm[threadIdx.x] = mx;
  }
  __syncthreads();
  if (threadIdx.x < NC) {
    float acc = bfc[threadIdx.x];
    for (int k = 0; k < C; ++k) acc = fmaf(gm[k], Wfc[k * NC + threadIdx.x], acc);
    out[site * NC + threadIdx.x] = apply_act(acc, ACT_SOFTPLUS);
  }
}

}  // namespace indel
}  // namespace mural

using namespace mural;
using namespace mural::indel;

// ================================================================================================ host
struct mural_indel_model {
  mural_indel_config_t cfg;
  int device;
  int L;
  int ch[6], len[6];
  std::vector<TensorEntry> layout;
  std::map<std::string, int> index;
  int64_t n_blob = 0, n_trainable = 0;
  float* d_prep = nullptr;
  struct Op { int64_t W, b; int Cin, Cout, ks, stride, up, act; };
  std::vector<Op> ops;  // stem conv (optional) is separate; order documented in mural_indel_forward
  int64_t stemT = -1, stemB = -1, Wfc = -1, bfc = -1;
  bool loaded = false;
  void* d_ws = nullptr;
  int64_t ws_bytes = 0;
  // tensor-core path (indel_tc.cuh): one fused kernel per U-Net level; pre-split bf16 B fragments + biases in d_tc
  struct TcLevel { int64_t Wl, W5, W1, bias; int KCl, KC5, Cin, CinP, stride, up, F, HA, dmin, ND, NC8, MT, NW, SG, TP, RA, RS, n_tiles, rows_in, smem, Lin, Lout, occ; };
  std::vector<TcLevel> tcl;  // encoder levels 0..5, decoder steps 0..4 (levels 4..0)
  int64_t tcWo0 = -1, tcWo1 = -1;
  int tcKCo = 0;
  uint32_t* d_tc = nullptr;
  bool tc_ok = false;
  int mode = 0;  // 0: tensor-core path when the shapes allow it, 1: fp32 CUDA-core kernels
};

extern "C" int mural_indel_model_create(const mural_indel_config_t* cfg, int device, mural_indel_model_t** out) {
  MURAL_CHECK(cfg && out, "NULL argument");
  *out = nullptr;
  const int C = cfg->channels, ks = cfg->kernel_size;
  MURAL_CHECK(C >= 4 && C <= 16 && C % 4 == 0, "CNN_out_channels must be 4, 8, 12 or 16 for the indel model in this build");
  MURAL_CHECK(ks >= 1 && ks <= 15 && (ks & 1), "CNN_kernel_size must be odd and <= 15");
  MURAL_CHECK(cfg->n_class >= 2 && cfg->n_class <= 16, "n_class must be in [2,16]");
  MURAL_CHECK(cfg->distal_radius >= 8 && cfg->distal_radius <= 40000, "distal_radius out of range");
  mural_indel_model* m = new mural_indel_model();
  m->cfg = *cfg;
  m->device = device;
  m->L = 2 * cfg->distal_radius;
  int L = m->L;
  for (int i = 0; i < 6; ++i) {
    m->ch[i] = C * (i + 1);
    const int s = cfg->downsize[i];
    if (s < 1 || s > 16) { delete m; MURAL_FAIL("down_list entries must be in [1,16]"); }
    L = (L - 1) / s + 1;  // Conv1d(k, stride s, padding (k-1)/2)
    m->len[i] = L;
  }
  // decoder shape check: nn.Upsample(scale_factor=down[lvl+1]) must reproduce the encoder length (model_indel.py:165-170)
  for (int lvl = 4; lvl >= 0; --lvl)
    if (m->len[lvl + 1] * cfg->downsize[lvl + 1] != m->len[lvl]) {
      delete m;
      MURAL_FAIL("window length is not compatible with down_list (skip connections would not line up)");
    }
  std::vector<TensorEntry> P, B;
  auto p = [&](const std::string& n, int64_t k) { P.push_back({n, 0, k, 0}); };
  auto bn = [&](const std::string& n, int64_t c) {
    p(n + ".weight", c); p(n + ".bias", c);
    B.push_back({n + ".running_mean", 0, c, 1});
    B.push_back({n + ".running_var", 0, c, 1});
  };
  auto cblock = [&](const std::string& n, int c) {
    p(n + ".conv.0.weight", int64_t(2 * c) * c * 5); bn(n + ".conv.1", 2 * c);
    p(n + ".conv.3.weight", int64_t(c) * 2 * c); bn(n + ".conv.4", c);
  };
  if (cfg->use_reverse) { p("conv.0.weight", 4 * 4 * ks); p("conv.0.bias", 4); bn("conv.1", 4); }
  for (int i = 0; i < 6; ++i) {
    const std::string s = "uplblocks." + std::to_string(i);
    const int cin = i ? m->ch[i - 1] : 4;
    p(s + ".0.weight", int64_t(m->ch[i]) * cin * ks); p(s + ".0.bias", m->ch[i]); bn(s + ".1", m->ch[i]);
  }
  for (int i = 0; i < 6; ++i) cblock("upblocks." + std::to_string(i) + ".0", m->ch[i]);
  for (int i = 0; i < 5; ++i) {
    const std::string s = "downlblocks." + std::to_string(i);
    p(s + ".1.weight", int64_t(m->ch[4 - i]) * m->ch[5 - i] * ks); p(s + ".1.bias", m->ch[4 - i]); bn(s + ".2", m->ch[4 - i]);
  }
  for (int i = 0; i < 5; ++i) cblock("downblocks." + std::to_string(i) + ".0", m->ch[4 - i]);
  p("out_conv.0.weight", int64_t(C) * C); p("out_conv.0.bias", C); bn("out_conv.1", C);
  p("out_conv.3.weight", int64_t(C) * C); p("out_conv.3.bias", C);
  bn("out_fc.0", C);
  p("out_fc.2.weight", int64_t(cfg->n_class) * C); p("out_fc.2.bias", cfg->n_class);
  int64_t off = 0;
  for (auto& e : P) { e.offset = off; off += e.numel; m->layout.push_back(e); }
  m->n_trainable = off;
  for (auto& e : B) { e.offset = off; off += e.numel; m->layout.push_back(e); }
  m->n_blob = off;
  for (size_t i = 0; i < m->layout.size(); ++i) m->index[m->layout[i].name] = (int)i;
  *out = m;
  return 0;
}

extern "C" void mural_indel_model_destroy(mural_indel_model_t* m) {
  if (!m) return;
  cudaFree(m->d_prep);
  cudaFree(m->d_ws);
  cudaFree(m->d_tc);
  delete m;
}
extern "C" int32_t mural_indel_model_n_tensors(const mural_indel_model_t* m) { return m ? (int32_t)m->layout.size() : 0; }
extern "C" int64_t mural_indel_model_n_params(const mural_indel_model_t* m) { return m ? m->n_blob : 0; }
extern "C" int mural_indel_model_tensor(const mural_indel_model_t* m, int32_t i, const char** name, int64_t* offset, int64_t* numel,
                                        int32_t* is_buffer) {
  MURAL_CHECK(m && i >= 0 && i < (int32_t)m->layout.size(), "tensor index out of range");
  const TensorEntry& e = m->layout[i];
  if (name) *name = e.name.c_str();
  if (offset) *offset = e.offset;
  if (numel) *numel = e.numel;
  if (is_buffer) *is_buffer = e.is_buffer;
  return 0;
}


// ---- tensor-core path: host preparation ---------------------------------------------------------------------------------
static inline uint32_t bf16_rn_bits(float x) {  // round to nearest even; inputs are finite
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return u >> 16;
}
static inline float bf16_bits_to_float(uint32_t b) {
  uint32_t u = b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// Folded conv weights W[t][ci][co] -> B fragments of mma.m16n8k16 over k = t*CinP + ci (zero for ci >= Cin and k >= ks*CinP):
// [k chunk][column tile][lane] x {b0 hi, b1 hi, b0 lo, b1 lo}, a packed pair holding k (low half) and k+1.
static int64_t make_frags(std::vector<uint32_t>& buf, const float* W, int ks, int Cin, int CinP, int Cout) {
  const int K = ks * CinP, KC = (K + 15) / 16, NT = Cout / 8;
  const int64_t off = (int64_t)buf.size();
  buf.resize(off + int64_t(KC) * NT * 32 * 4, 0u);
  auto w = [&](int k, int n) -> float {
    if (k >= K) return 0.f;
    const int t = k / CinP, ci = k - t * CinP;
    return ci < Cin ? W[(int64_t(t) * Cin + ci) * Cout + n] : 0.f;
  };
  for (int kc = 0; kc < KC; ++kc)
    for (int nt = 0; nt < NT; ++nt)
      for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, q = lane & 3, n = nt * 8 + g;
        uint32_t* d = buf.data() + off + ((int64_t(kc) * NT + nt) * 32 + lane) * 4;
        for (int h = 0; h < 2; ++h) {
          const int k = kc * 16 + 2 * q + 8 * h;
          const float w0 = w(k, n), w1 = w(k + 1, n);
          const uint32_t h0 = bf16_rn_bits(w0), h1 = bf16_rn_bits(w1);
          const uint32_t l0 = bf16_rn_bits(w0 - bf16_bits_to_float(h0)), l1 = bf16_rn_bits(w1 - bf16_bits_to_float(h1));
          d[h] = h0 | (h1 << 16);
          d[2 + h] = l0 | (l1 << 16);
        }
      }
  return off;
}

typedef void (*LevelKernel)(const indel_tc::LevelParams);
// Kernel instantiations: (column tiles NC8 = C/8, row tiles per warp MT, warps per CTA NW).  Narrow levels are long: two row
// tiles per warp reuse each weight fragment twice; wide levels hold many column tiles per row tile in registers (MT = 1), and
// when their weights leave room for one CTA per SM only, that CTA brings 16 warps.
static LevelKernel level_kernel(int NC8, int MT, int NW, bool tail) {
  using namespace indel_tc;
  if (tail) {
    if (NW != 8) return nullptr;
    if (NC8 == 1) return MT == 2 ? (LevelKernel)k_unet_level<1, 2, 8, true> : MT == 1 ? (LevelKernel)k_unet_level<1, 1, 8, true> : nullptr;
    if (NC8 == 2) return MT == 2 ? (LevelKernel)k_unet_level<2, 2, 8, true> : MT == 1 ? (LevelKernel)k_unet_level<2, 1, 8, true> : nullptr;
    return nullptr;
  }
  if (MT == 2 && NW == 8) return NC8 == 1 ? (LevelKernel)k_unet_level<1, 2, 8, false> : NC8 == 2 ? (LevelKernel)k_unet_level<2, 2, 8, false> : nullptr;
  if (MT != 1) return nullptr;
  if (NW == 8) switch (NC8) {
    case 1: return k_unet_level<1, 1, 8, false>;
    case 2: return k_unet_level<2, 1, 8, false>;
    case 3: return k_unet_level<3, 1, 8, false>;
    case 4: return k_unet_level<4, 1, 8, false>;
    case 5: return k_unet_level<5, 1, 8, false>;
    case 6: return k_unet_level<6, 1, 8, false>;
  }
  if (NW == 16) switch (NC8) {
    case 3: return k_unet_level<3, 1, 16, false>;
    case 4: return k_unet_level<4, 1, 16, false>;
    case 5: return k_unet_level<5, 1, 16, false>;
    case 6: return k_unet_level<6, 1, 16, false>;
  }
  return nullptr;
}
// resident CTAs per SM allowed by the register file (ptxas -v of the instantiations above)
static int level_reg_ctas(int NC8, int NW) { return NW == 16 ? 1 : NC8 == 1 ? 4 : NC8 <= 4 ? 2 + (NC8 >= 3) : 2; }

// Builds the fragment buffers and the tile geometry of the 11 level kernels; leaves tc_ok false (fp32 kernels are used) when
// a level does not fit: channels not a multiple of 8, a level wider than 48 channels, or > 227 KB of shared memory.
static int indel_tc_prepare(mural_indel_model* m, const std::vector<float>& prep) {
  using namespace indel_tc;
  m->tc_ok = false;
  m->tcl.clear();
  const int C = m->cfg.channels, ks = m->cfg.kernel_size;
  if (C % 8 != 0 || m->ch[5] > 48) return 0;
  m->tcKCo = (C / 8 + 1) / 2;
  std::vector<uint32_t> buf;
  auto fbits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
  auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
  for (int step = 0; step < 11; ++step) {
    const bool dec = step >= 6, tail = step == 10;
    const int lvl = dec ? 4 - (step - 6) : step;
    const int o0 = dec ? 18 + 3 * (step - 6) : 3 * step;  // lconv, conv5, conv1 of this level in m->ops
    const mural_indel_model::Op &ol = m->ops[o0], &o5 = m->ops[o0 + 1], &o1 = m->ops[o0 + 2];
    if (o5.ks != 5 || o1.ks != 1 || ol.ks != ks) return 0;
    mural_indel_model::TcLevel T{};
    T.NC8 = m->ch[lvl] / 8;
    T.Cin = ol.Cin;
    T.CinP = (ol.Cin + 7) & ~7;
    T.stride = ol.stride;
    T.Lin = dec ? m->len[lvl + 1] : (lvl ? m->len[lvl - 1] : m->L);
    T.Lout = m->len[lvl];
    T.KC5 = (5 * m->ch[lvl] + 15) / 16;
    // lconv weights as written (unfolded) and, for a decoder level, with the nearest upsampling folded in: output position
    // F*m + ph, tap t reads low-resolution row m + floor((ph + t - half) / F); per phase ph the taps with the same row offset d
    // are summed (fp64) into Wf[ph][d - dmin] (ND = dmax - dmin + 1 taps instead of ks).
    const int64_t Wl_plain = make_frags(buf, prep.data() + ol.W, ks, ol.Cin, T.CinP, ol.Cout);
    int64_t Wl_fold = -1;
    int fF = 1, f_dmin = 0, fND = ks;
    if (ol.up > 1 && ol.stride == 1) {
      fF = ol.up;
      const int half = ks / 2;
      f_dmin = fdiv(-half, fF);
      fND = fdiv(fF - 1 + ks - 1 - half, fF) - f_dmin + 1;
      std::vector<float> Wf(size_t(fND) * ol.Cin * ol.Cout);
      for (int ph = 0; ph < fF; ++ph) {
        std::vector<double> acc(Wf.size(), 0.0);
        for (int t = 0; t < ks; ++t) {
          const int d = fdiv(ph + t - half, fF) - f_dmin;
          for (int e = 0; e < ol.Cin * ol.Cout; ++e) acc[size_t(d) * ol.Cin * ol.Cout + e] += double(prep[ol.W + size_t(t) * ol.Cin * ol.Cout + e]);
        }
        for (size_t e = 0; e < Wf.size(); ++e) Wf[e] = float(acc[e]);
        const int64_t off = make_frags(buf, Wf.data(), fND, ol.Cin, T.CinP, ol.Cout);
        if (ph == 0) Wl_fold = off;
      }
    }
    T.W5 = make_frags(buf, prep.data() + o5.W, 5, o5.Cin, o5.Cin, o5.Cout);
    T.W1 = make_frags(buf, prep.data() + o1.W, 1, o1.Cin, o1.Cin, o1.Cout);
    const int Cl = m->ch[lvl];
    T.bias = (int64_t)buf.size();
    for (int c = 0; c < Cl; ++c) buf.push_back(fbits(prep[ol.b + c]));
    for (int c = 0; c < 2 * Cl; ++c) buf.push_back(fbits(prep[o5.b + c]));
    for (int c = 0; c < Cl; ++c) buf.push_back(fbits(prep[o1.b + c]));
    for (int c = 0; c < Cl; ++c) buf.push_back(fbits(tail ? prep[m->ops[33].b + c] : 0.f));
    for (int c = 0; c < Cl; ++c) buf.push_back(fbits(tail ? prep[m->ops[34].b + c] : 0.f));
    while (buf.size() & 3) buf.push_back(0u);

    // ---- shape of the level kernel: (folded?, MT, NW, tile rows, site group), all candidates that fit the 227 KB of an SM;
    // cost = tensor work per site / (share of warps that get a row tile x occupancy factor).  Few resident warps cannot hide
    // the phase barriers and the global-load latency (measured on the shipped shapes: 8 warps ~0.55, 16 ~0.8 of the 24+ rate).
    auto shape = [&](bool fold, int MT, int NW, int ra_max, int sg) {
      T.MT = MT; T.NW = NW;
      T.F = fold ? fF : 1; T.HA = fold ? fF : 2; T.dmin = fold ? f_dmin : 0; T.ND = fold ? fND : ks;
      T.up = fold ? 1 : ol.up;                                // folded: the kernel stages low-resolution rows
      T.Wl = fold ? Wl_fold : Wl_plain;
      T.KCl = (T.ND * T.CinP + 15) / 16;
      const int unit = 16 * MT * T.F;                         // a row tile holds one phase: RA is a multiple of F row-tile groups
      const int cap = ((ra_max + unit - 1) / unit) * unit;
      const int usable = ((cap - T.HA - 2) / T.F) * T.F;      // outputs per tile (a multiple of F: tiles start on a phase-0 position);
                                                              // the rest is the halo of Conv5 (+ alignment when folded)
      if (usable < T.F) return false;
      T.n_tiles = (T.Lout + usable - 1) / usable;
      T.TP = (((T.Lout + T.n_tiles - 1) / T.n_tiles + T.F - 1) / T.F) * T.F;
      T.RA = ((T.TP + T.HA + 2 + unit - 1) / unit) * unit;
      T.rows_in = T.F > 1 ? T.RA / T.F + T.ND : (T.RA - 1) * T.stride + ks + 1;
      T.RS = (T.rows_in + T.stride - 1) / T.stride;
      T.SG = T.n_tiles == 1 ? sg : 1;
      LevelParams P{};
      P.KCl = T.KCl; P.KC5 = T.KC5; P.KCo = m->tcKCo; P.CinP = T.CinP; P.RA = T.RA; P.rows_in = T.rows_in; P.RS = T.RS;
      P.stride = T.stride; P.SG = T.SG; P.F = T.F; P.HA = T.HA;
      T.smem = smem_layout(T.NC8, tail, P).total;
      return T.SG == sg && T.smem <= 227 * 1024;
    };
    double best = -1.0;
    int bF = 0, bMT = 0, bNW = 0, bRA = 0, bSG = 0;
    for (int fold = 0; fold <= (Wl_fold >= 0 ? 1 : 0); ++fold)
      for (int MT : {2, 1})
        for (int NW : {8, 16}) {
          if (!level_kernel(T.NC8, MT, NW, tail)) continue;
          for (int ra_max = RA_MAX; ra_max >= 16; ra_max /= 2)
            for (int sg = 1; sg <= 16; sg *= 2) {
              if (!shape(fold != 0, MT, NW, ra_max, sg)) break;
              const int tiles = T.SG * (T.RA / 16), per_round = NW * MT;
              const double util = double(tiles) / double(((tiles + per_round - 1) / per_round) * per_round);
              const int ctas = std::min((227 * 1024) / T.smem, level_reg_ctas(T.NC8, NW));
              const double occ = sqrt(std::min(1.0, double(ctas * NW) / 24.0));
              const double work = double(T.n_tiles) * T.RA * (double(T.KCl) * T.NC8 + 2.0 * T.KC5 * T.NC8 + double(T.NC8) * T.NC8);
              const double score = util * occ / work;
              if (score > best * (1.0 + 1e-9)) { best = score; bF = fold; bMT = MT; bNW = NW; bRA = ra_max; bSG = sg; }
            }
        }
    if (best < 0) return 0;
    shape(bF != 0, bMT, bNW, bRA, bSG);
    if (getenv("MURAL_INDEL_DEBUG"))
      fprintf(stderr, "indel level %2d: NC8=%d fold=%d(F=%d ND=%d) MT=%d NW=%d Lout=%d n_tiles=%d TP=%d RA=%d SG=%d KCl=%d smem=%d\n", step, T.NC8,
              bF, T.F, T.ND, T.MT, T.NW, T.Lout, T.n_tiles, T.TP, T.RA, T.SG, T.KCl, T.smem);
    m->tcl.push_back(T);
  }
  m->tcWo0 = make_frags(buf, prep.data() + m->ops[33].W, 1, C, C, C);
  m->tcWo1 = make_frags(buf, prep.data() + m->ops[34].W, 1, C, C, C);
  for (int step = 0; step < 11; ++step) {  // levels of equal shape share a kernel: opt in to the largest request
    const mural_indel_model::TcLevel T = m->tcl[step];
    int need = 0;
    for (int o = 0; o < 11; ++o)
      if (level_kernel(m->tcl[o].NC8, m->tcl[o].MT, m->tcl[o].NW, o == 10) == level_kernel(T.NC8, T.MT, T.NW, step == 10))
        need = std::max(need, m->tcl[o].smem);
    static std::map<LevelKernel, int> configured;  // process-wide and monotone: several models (radii) share the kernels
    LevelKernel k = level_kernel(T.NC8, T.MT, T.NW, step == 10);
    if (configured[k] < need) {
      CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
      configured[k] = need;
    }
    int occ = 1;   // resident CTAs per SM of this level's launch: the persistent grid is n_sm * occ
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, T.NW * 32, T.smem));
    m->tcl[step].occ = occ < 1 ? 1 : occ;
  }
  cudaFree(m->d_tc);
  m->d_tc = nullptr;
  CUDA_TRY(cudaMalloc((void**)&m->d_tc, buf.size() * 4));
  CUDA_TRY(cudaMemcpy(m->d_tc, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice));
  m->tc_ok = true;
  return 0;
}

extern "C" int mural_indel_model_load(mural_indel_model_t* m, const float* h_blob, int64_t n) {
  MURAL_CHECK(m && h_blob, "NULL argument");
  MURAL_CHECK(n == m->n_blob, "parameter blob has the wrong length");
  for (int64_t i = 0; i < n; ++i) MURAL_CHECK(isfinite(h_blob[i]), "non-finite value in the parameter blob");
  const int C = m->cfg.channels, ks = m->cfg.kernel_size, NC = m->cfg.n_class;
  std::vector<float> prep;
  auto T = [&](const std::string& nme) { return h_blob + m->layout[m->index.at(nme)].offset; };
  auto alloc = [&](int64_t k) { int64_t o = (int64_t)prep.size(); prep.resize(o + ((k + 3) & ~int64_t(3)), 0.f); return o; };
  // conv (weight [co][ci][t], optional bias) followed by BatchNorm `bnn` (may be empty) -> [t][ci][co] + bias
  auto fold = [&](const std::string& w, const std::string& b, const std::string& bnn, int cin, int cout, int kk, int stride, int up,
                  int act) {
    mural_indel_model::Op op{alloc(int64_t(kk) * cin * cout), alloc(cout), cin, cout, kk, stride, up, act};
    const float* W = T(w);
    const float* bi = b.empty() ? nullptr : T(b);
    for (int co = 0; co < cout; ++co) {
      double a = 1.0, sh = 0.0;
      if (!bnn.empty()) {
        a = double(T(bnn + ".weight")[co]) / sqrt(double(T(bnn + ".running_var")[co]) + 1e-5);
        sh = double(T(bnn + ".bias")[co]) - double(T(bnn + ".running_mean")[co]) * a;
      }
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < kk; ++t) prep[op.W + (int64_t(t) * cin + ci) * cout + co] = float(a * double(W[(int64_t(co) * cin + ci) * kk + t]));
      prep[op.b + co] = float(a * (bi ? double(bi[co]) : 0.0) + sh);
    }
    m->ops.push_back(op);
    return op;
  };
  m->ops.clear();
  m->stemT = m->stemB = -1;
  if (m->cfg.use_reverse) {  // BN(conv(one-hot)) as a per-tap table over the 15 symbols
    const float *W = T("conv.0.weight"), *bi = T("conv.0.bias");
    m->stemT = alloc(int64_t(ks) * 16 * 4);
    m->stemB = alloc(4);
    const double oh[15][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}, {.5, 0, .5, 0}, {0, .5, 0, .5}, {.5, .5, 0, 0},
                              {0, .5, .5, 0}, {.5, 0, 0, .5}, {0, 0, .5, .5}, {0, 1, 1, 1}, {1, 0, 1, 1}, {1, 1, 0, 1}, {1, 1, 1, 0},
                              {.25, .25, .25, .25}};
    for (int co = 0; co < 4; ++co) {
      const double a = double(T("conv.1.weight")[co]) / sqrt(double(T("conv.1.running_var")[co]) + 1e-5);
      const double sh = double(T("conv.1.bias")[co]) - double(T("conv.1.running_mean")[co]) * a;
      prep[m->stemB + co] = float(a * double(bi[co]) + sh);
      for (int t = 0; t < ks; ++t)
        for (int s = 0; s < 15; ++s) {
          double acc = 0;
          for (int c = 0; c < 4; ++c) {
            double e = oh[s][c];
            if (s >= 10 && s <= 13) e = e ? double(float(1.0 / 3.0)) : 0.0;
            acc += double(W[(co * 4 + c) * ks + t]) * e;
          }
          prep[m->stemT + (int64_t(t) * 16 + s) * 4 + co] = float(a * acc);
        }
    }
  }
  for (int i = 0; i < 6; ++i) {  // encoder: ops 3i, 3i+1, 3i+2
    const std::string l = "uplblocks." + std::to_string(i), u = "upblocks." + std::to_string(i) + ".0.conv";
    const int cin = i ? m->ch[i - 1] : 4, c = m->ch[i];
    fold(l + ".0.weight", l + ".0.bias", l + ".1", cin, c, ks, m->cfg.downsize[i], 1, ACT_NONE);
    fold(u + ".0.weight", "", u + ".1", c, 2 * c, 5, 1, 1, ACT_SILU);
    fold(u + ".3.weight", "", u + ".4", 2 * c, c, 1, 1, 1, ACT_NONE);
  }
  for (int i = 0; i < 5; ++i) {  // decoder: ops 18 + 3i ...
    const std::string l = "downlblocks." + std::to_string(i), u = "downblocks." + std::to_string(i) + ".0.conv";
    const int cin = m->ch[5 - i], c = m->ch[4 - i];
    fold(l + ".1.weight", l + ".1.bias", l + ".2", cin, c, ks, 1, m->cfg.downsize[5 - i], ACT_NONE);
    fold(u + ".0.weight", "", u + ".1", c, 2 * c, 5, 1, 1, ACT_SILU);
    fold(u + ".3.weight", "", u + ".4", 2 * c, c, 1, 1, 1, ACT_NONE);
  }
  fold("out_conv.0.weight", "out_conv.0.bias", "out_conv.1", C, C, 1, 1, 1, ACT_RELU);     // op 33
  fold("out_conv.3.weight", "out_conv.3.bias", "", C, C, 1, 1, 1, ACT_SOFTPLUS);           // op 34
  {  // BN(out_fc.0) folded into Linear(out_fc.2): Wfc[c][o]
    m->Wfc = alloc(int64_t(C) * NC);
    m->bfc = alloc(NC);
    const float *W = T("out_fc.2.weight"), *bi = T("out_fc.2.bias");
    for (int o = 0; o < NC; ++o) {
      double acc = bi[o];
      for (int c = 0; c < C; ++c) {
        const double a = double(T("out_fc.0.weight")[c]) / sqrt(double(T("out_fc.0.running_var")[c]) + 1e-5);
        const double sh = double(T("out_fc.0.bias")[c]) - double(T("out_fc.0.running_mean")[c]) * a;
        prep[m->Wfc + int64_t(c) * NC + o] = float(double(W[o * C + c]) * a);
        acc += double(W[o * C + c]) * sh;
      }
      prep[m->bfc + o] = float(acc);
    }
  }
  CUDA_TRY(cudaSetDevice(m->device));
  cudaFree(m->d_prep);
  m->d_prep = nullptr;
  CUDA_TRY(cudaMalloc((void**)&m->d_prep, prep.size() * sizeof(float)));
  CUDA_TRY(cudaMemcpy(m->d_prep, prep.data(), prep.size() * sizeof(float), cudaMemcpyHostToDevice));
  if (int rc = indel_tc_prepare(m, prep)) return rc;
  m->loaded = true;
  return 0;
}

static int run_conv(const mural_indel_model* m, int oi, const float* in, float* out, const float* r1, const float* r2, int64_t ns,
                    int Lin, int Lout, cudaStream_t st) {
  const mural_indel_model::Op& o = m->ops[oi];
  ConvOp P{m->d_prep + o.W, m->d_prep + o.b, o.Cin, o.Cout, o.ks, o.stride, o.up, o.act};
  // items per CTA = (TP/4) position groups x (Cout/4) channel groups: keep >= 128 of them (the U-Net's widest levels are
  // also its narrowest in channels: Cout = 8 / 16 at L = 8000 / 2000)
  int TP = TP_MIN;
  while ((TP / 4) * (o.Cout / 4) < 128 && TP < 512) TP *= 2;
  const int xrows = (TP - 1) * o.stride + o.ks;
  MURAL_CHECK(o.Cin % 4 == 0 && o.Cout % 4 == 0, "indel conv: channel counts must be multiples of 4");
  const size_t smem = sizeof(float) * (((size_t(o.ks) * o.Cin * o.Cout + 3) & ~size_t(3)) + size_t(xrows) * (o.Cin + 4));
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CUDA_TRY(cudaFuncSetAttribute(k_conv_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((unsigned)cdiv(Lout, TP), (unsigned)ns);
  LAUNCH(k_conv_gen, grid, 128, smem, st, in, out, r1, r2, Lin, Lout, P, TP);
  return 0;
}

// Tensor-core forward: stem, 6 encoder levels, 5 decoder levels (the last with out_conv + position max fused), head.
static int indel_forward_tc(mural_indel_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta, const uint8_t* d_sym,
                            int64_t n, float* d_out, cudaStream_t st) {
  using namespace indel_tc;
  const int C = m->cfg.channels, ks = m->cfg.kernel_size, NC = m->cfg.n_class, L = m->L;
  // workspace per site (floats): X [L*4], encoder outputs E[0..5], two decoder buffers, position max [C]
  int64_t dmax = 0;
  for (int lvl = 1; lvl <= 4; ++lvl) dmax = std::max<int64_t>(dmax, int64_t(m->len[lvl]) * m->ch[lvl]);
  int64_t per = int64_t(L) * 4 + 2 * dmax + C;
  for (int i = 0; i < 6; ++i) per += int64_t(m->len[i]) * m->ch[i];
  per = (per + 3) & ~int64_t(3);
  int64_t chunk = (int64_t(2) << 30) / (per * 4);
  if (chunk < 1) chunk = 1;
  if (chunk > 4096) chunk = 4096;
  if (chunk > n) chunk = n;
  if (m->ws_bytes < chunk * per * 4) {
    cudaFree(m->d_ws);
    m->d_ws = nullptr;
    m->ws_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_ws, chunk * per * 4));
    m->ws_bytes = chunk * per * 4;
  }
  auto al4 = [](int64_t v) { return (v + 3) & ~int64_t(3); };
  float* w = (float*)m->d_ws;
  float* X = w; w += al4(chunk * int64_t(L) * 4);
  float* E[6];
  for (int i = 0; i < 6; ++i) { E[i] = w; w += al4(chunk * int64_t(m->len[i]) * m->ch[i]); }
  float* D[2];
  D[0] = w; w += al4(chunk * dmax);
  D[1] = w; w += al4(chunk * dmax);
  float* gmax = w;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  GenomeView gv = G ? *G : GenomeView{};
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int64_t ns = (n - s0 < chunk) ? (n - s0) : chunk;
    const size_t ssm = sizeof(float) * ks * 64 + ((size_t(L) + 15) & ~size_t(15));
    static size_t conf = 0;
    if (ssm > 48 * 1024 && ssm > conf) {
      CUDA_TRY(cudaFuncSetAttribute(k_indel_stem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
      conf = ssm;
    }
    // (building the network input inside the first level kernel while staging was measured: +0.5 ms on that kernel for the
    // 0.2 ms of this one — a per-tile symbol fetch and two extra barriers in a latency-bound loop.  Kept separate.)
    LAUNCH(k_indel_stem, (unsigned)ns, 256, ssm, st, gv, d_pos ? d_pos + s0 : nullptr, d_meta ? d_meta + s0 : nullptr,
           d_sym ? d_sym + s0 * L : nullptr, m->cfg.distal_radius, L, ks, m->stemT >= 0 ? m->d_prep + m->stemT : nullptr,
           m->stemB >= 0 ? m->d_prep + m->stemB : nullptr, X);
    LAUNCH(k_fill_f32, 64, 256, 0, st, gmax, ns * C, -INFINITY);
    const float* x = X;
    for (int step = 0; step < 11; ++step) {
      const mural_indel_model::TcLevel& T = m->tcl[step];
      const bool dec = step >= 6, tail = step == 10;
      const int lvl = dec ? 4 - (step - 6) : step;
      LevelParams P{};
      P.in = x;
      P.skip = dec ? E[lvl] : nullptr;
      P.out = dec ? D[step & 1] : E[lvl];
      P.gmax = gmax;
      const uint4* tc = reinterpret_cast<const uint4*>(m->d_tc);
      P.Wl = tc + T.Wl / 4; P.W5 = tc + T.W5 / 4; P.W1 = tc + T.W1 / 4;
      P.Wo0 = tc + m->tcWo0 / 4; P.Wo1 = tc + m->tcWo1 / 4;
      P.bias = reinterpret_cast<const float*>(m->d_tc) + T.bias;
      P.Cin = T.Cin; P.CinP = T.CinP; P.ks = ks; P.stride = T.stride; P.up = T.up;
      P.Lin = T.Lin; P.Lout = T.Lout;
      P.KCl = T.KCl; P.KC5 = T.KC5; P.KCo = m->tcKCo;
      P.TP = T.TP; P.RA = T.RA; P.n_tiles = T.n_tiles; P.rows_in = T.rows_in; P.RS = T.RS; P.SG = T.SG;
      P.F = T.F; P.HA = T.HA; P.dmin = T.dmin;
      P.magic_stride = T.stride > 1 ? uint32_t(((1ull << 32) + T.stride - 1) / T.stride) : 0u;
      P.magic_up = T.up > 1 ? uint32_t(((1ull << 32) + T.up - 1) / T.up) : 0u;
      P.n_sites = ns;
      P.n_items = cdiv(ns, T.SG) * T.n_tiles;
      LevelKernel k = level_kernel(T.NC8, T.MT, T.NW, tail);
      const int THREADS = T.NW * 32;
      const int64_t grid = std::min<int64_t>(P.n_items, int64_t(n_sm) * T.occ);
      static const char* names[11] = {"k_unet_level/enc0", "k_unet_level/enc1", "k_unet_level/enc2", "k_unet_level/enc3", "k_unet_level/enc4",
                                      "k_unet_level/enc5", "k_unet_level/dec4", "k_unet_level/dec3", "k_unet_level/dec2", "k_unet_level/dec1",
                                      "k_unet_level/dec0+out"};
      LAUNCH_N(names[step], k, (unsigned)grid, THREADS, T.smem, st, P);
      x = P.out;
    }
    LAUNCH(k_indel_head_tc, (unsigned)cdiv(ns * NC, 128), 128, 0, st, gmax, ns, C, m->d_prep + m->Wfc, m->d_prep + m->bfc, NC, d_out + s0 * NC);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

static int indel_forward(mural_indel_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta, const uint8_t* d_sym,
                         int64_t n, float* d_out, cudaStream_t st) {
  if (m->tc_ok && m->mode == 0) return indel_forward_tc(m, G, d_pos, d_meta, d_sym, n, d_out, st);
  const int C = m->cfg.channels, ks = m->cfg.kernel_size, NC = m->cfg.n_class, L = m->L;
  // workspace per site (floats): X[L*4], per level A,E [len*ch], H [len0*2C max], decoder x / scratch
  int64_t per = int64_t(L) * 4;
  for (int i = 0; i < 6; ++i) per += 2 * int64_t(m->len[i]) * m->ch[i];
  per += int64_t(m->len[0]) * 2 * C;          // H (largest at level 0: len0 * 2C; every level has len*2ch <= that? checked below)
  int64_t hmax = 0, dmax = 0;                 // widest hidden tensor; widest decoder output (level 0 only when every stride is > 1:
                                              // a down_list entry of 1 keeps the length while the channels grow — found by memcheck)
  for (int i = 0; i < 6; ++i) hmax = std::max<int64_t>(hmax, int64_t(m->len[i]) * 2 * m->ch[i]);
  for (int i = 0; i < 5; ++i) dmax = std::max<int64_t>(dmax, int64_t(m->len[i]) * m->ch[i]);
  per += 2 * dmax;                            // decoder ping-pong
  per += hmax;
  int64_t chunk = (int64_t(768) << 20) / (per * 4);
  if (chunk < 1) chunk = 1;
  if (chunk > 1024) chunk = 1024;
  if (chunk > n) chunk = n;
  if (m->ws_bytes < chunk * per * 4) {
    cudaFree(m->d_ws);
    m->d_ws = nullptr;
    m->ws_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_ws, chunk * per * 4));
    m->ws_bytes = chunk * per * 4;
  }
  float* w = (float*)m->d_ws;
  float* X = w; w += chunk * int64_t(L) * 4;
  float *A[6], *E[6];
  for (int i = 0; i < 6; ++i) { A[i] = w; w += chunk * int64_t(m->len[i]) * m->ch[i]; E[i] = w; w += chunk * int64_t(m->len[i]) * m->ch[i]; }
  float* H = w; w += chunk * (int64_t(m->len[0]) * 2 * C + hmax);
  float* D0 = w; w += chunk * dmax;
  float* D1 = w;
  GenomeView gv = G ? *G : GenomeView{};
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int64_t ns = (n - s0 < chunk) ? (n - s0) : chunk;
    const size_t ssm = sizeof(float) * ks * 64 + ((size_t(L) + 15) & ~size_t(15));
    static size_t conf = 0;
    if (ssm > 48 * 1024 && ssm > conf) {
      CUDA_TRY(cudaFuncSetAttribute(k_indel_stem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
      conf = ssm;
    }
    LAUNCH(k_indel_stem, (unsigned)ns, 256, ssm, st, gv, d_pos ? d_pos + s0 : nullptr, d_meta ? d_meta + s0 : nullptr,
           d_sym ? d_sym + s0 * L : nullptr, m->cfg.distal_radius, L, ks, m->stemT >= 0 ? m->d_prep + m->stemT : nullptr,
           m->stemB >= 0 ? m->d_prep + m->stemB : nullptr, X);
    // encoder (model_indel.py:158-163)
    const float* x = X;
    int Lx = L;
    for (int i = 0; i < 6; ++i) {
      if (int rc = run_conv(m, 3 * i, x, A[i], nullptr, nullptr, ns, Lx, m->len[i], st)) return rc;
      if (int rc = run_conv(m, 3 * i + 1, A[i], H, nullptr, nullptr, ns, m->len[i], m->len[i], st)) return rc;
      if (int rc = run_conv(m, 3 * i + 2, H, E[i], A[i], nullptr, ns, m->len[i], m->len[i], st)) return rc;   // x + conv(x)
      x = E[i];
      Lx = m->len[i];
    }
    // decoder (model_indel.py:165-170): upsample -> conv+BN -> ConvBlock -> + encoder skip
    for (int i = 0; i < 5; ++i) {
      const int lvl = 4 - i;
      float* a = A[lvl];  // the encoder's A buffer of this level is free again
      if (int rc = run_conv(m, 18 + 3 * i, x, a, nullptr, nullptr, ns, Lx, m->len[lvl], st)) return rc;
      if (int rc = run_conv(m, 18 + 3 * i + 1, a, H, nullptr, nullptr, ns, m->len[lvl], m->len[lvl], st)) return rc;
      float* o = (i & 1) ? D1 : D0;
      if (int rc = run_conv(m, 18 + 3 * i + 2, H, o, a, E[lvl], ns, m->len[lvl], m->len[lvl], st)) return rc;
      x = o;
      Lx = m->len[lvl];
    }
    // out_conv + global max + out_fc (model_indel.py:172-174)
    if (int rc = run_conv(m, 33, x, A[0], nullptr, nullptr, ns, Lx, Lx, st)) return rc;
    if (int rc = run_conv(m, 34, A[0], E[0], nullptr, nullptr, ns, Lx, Lx, st)) return rc;
    LAUNCH(k_indel_head, (unsigned)ns, 128, 0, st, E[0], ns, Lx, C, m->d_prep + m->Wfc, m->d_prep + m->bfc, NC, d_out + s0 * NC);
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_indel_forward(mural_indel_model_t* m, const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta,
                                   int64_t n, float* d_out, void* stream) {
  MURAL_CHECK(m && g && (n == 0 || (d_pos && d_meta && d_out)), "NULL argument");
  MURAL_CHECK(m->loaded, "model weights not loaded (call mural_indel_model_load first)");
  MURAL_CHECK(g->device == m->device, "genome and model live on different devices");
  if (n == 0) return 0;
  return indel_forward(m, &g->view, d_pos, d_meta, nullptr, n, d_out, (cudaStream_t)stream);
}

extern "C" int mural_indel_forward_tensors(mural_indel_model_t* m, const float* d_distal, int64_t n, int32_t L, float* d_out,
                                           void* stream) {
  MURAL_CHECK(m && (n == 0 || (d_distal && d_out)), "NULL argument");
  MURAL_CHECK(m->loaded, "model weights not loaded (call mural_indel_model_load first)");
  MURAL_CHECK(L == m->L, "distal_x length does not match the model's distal_radius");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* d_sym = nullptr;
  const int64_t sym_bytes = (n * int64_t(L) + 255) & ~int64_t(255);
  CUDA_TRY(cudaMalloc((void**)&d_sym, sym_bytes + 256));
  int* d_bad = (int*)(d_sym + sym_bytes);
  cudaMemsetAsync(d_bad, 0, 4, st);
  int rc = onehot_to_symbols_checked(d_distal, n, L, d_sym, d_bad, st);
  if (rc == 0) {
    int bad = 0;
    cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (bad) rc = fail(__FILE__, __LINE__, "distal_x holds columns that are not reference one-hot vectors");
  }
  if (rc == 0) rc = indel_forward(m, nullptr, nullptr, nullptr, d_sym, n, d_out, st);
  cudaStreamSynchronize(st);
  cudaFree(d_sym);
  return rc;
}

extern "C" int mural_indel_model_config(const mural_indel_model_t* m, mural_indel_config_t* out) {
  MURAL_CHECK(m && out, "NULL argument");
  *out = m->cfg;
  return 0;
}

extern "C" int mural_indel_set_mode(mural_indel_model_t* m, int32_t mode) {
  MURAL_CHECK(m, "NULL argument");
  MURAL_CHECK(mode == 0 || mode == 1, "ValueError: indel mode must be 0 (tensor-core level kernels when the shapes allow) or 1 (fp32 kernels)");
  m->mode = mode;
  return 0;
}
extern "C" int mural_indel_tc_available(const mural_indel_model_t* m) { return m && m->tc_ok ? 1 : 0; }
