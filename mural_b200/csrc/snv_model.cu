// Network2 parameter layout and eval-mode folding (host side, double precision), upload of the
// prepared weights.  Reference: MuRaL/model/model_snv.py:290-437 (constructor), :439-525 (forward).
#include <math.h>
#include <string.h>

#include "snv_model.cuh"

using namespace mural;

namespace {

const double BN_EPS = 1e-5;  // nn.BatchNorm1d default, used everywhere in model_snv.py

struct LayoutBuilder {
  std::vector<TensorEntry> params, buffers;
  void p(const std::string& n, int64_t numel) { params.push_back({n, 0, numel, 0}); }
  void b(const std::string& n, int64_t numel) { buffers.push_back({n, 0, numel, 1}); }
  void bn(const std::string& n, int64_t c) {
    p(n + ".weight", c);
    p(n + ".bias", c);
    b(n + ".running_mean", c);
    b(n + ".running_var", c);
  }
  void conv(const std::string& n, int64_t co, int64_t ci, int64_t ks) {
    p(n + ".weight", co * ci * ks);
    p(n + ".bias", co);
  }
};

int pool_out(int L, int k, int s, int p) { return (L + 2 * p - k) / s + 1; }  // MaxPool1d, ceil_mode=False

}  // namespace

extern "C" int mural_snv_model_create(const mural_snv_config_t* cfg, int device, mural_snv_model_t** out) {
  MURAL_CHECK(cfg && out, "NULL argument");
  *out = nullptr;
  const int C = cfg->channels, ks = cfg->kernel_size, k = cfg->local_order;
  MURAL_CHECK(C == 16 || C == 32 || C == 64, "CNN_out_channels must be 16, 32 or 64 in this build");
  MURAL_CHECK(ks >= 1 && ks <= 7 && (ks & 1), "CNN_kernel_size must be odd and <= 7");
  MURAL_CHECK(k >= 1 && k <= 6, "local_order must be in [1,6]");
  MURAL_CHECK(cfg->n_class >= 2 && cfg->n_class <= 16, "n_class must be in [2,16]");
  MURAL_CHECK(cfg->hidden1 >= 1 && cfg->hidden1 <= 1024 && cfg->hidden2 >= 1 && cfg->hidden2 <= 1024,
              "hidden sizes must be in [1,1024]");
  const int L = 2 * cfg->distal_radius + 1;
  MURAL_CHECK(L > 200, "Error: distal seq len must be >200");  // model_snv.py:470
  MURAL_CHECK(cfg->distal_radius <= 40000, "distal_radius too large for this build (max 40000)");
  const int n_cat = 2 * cfg->local_radius + 1 - (k - 1);
  MURAL_CHECK(n_cat >= 1 && cfg->local_radius <= cfg->distal_radius, "local window must be inside the expanded window");
  mural_snv_model* m = new mural_snv_model();
  m->cfg = *cfg;
  m->device = device;
  m->n_cat = n_cat;
  m->emb_rows = (1 << (2 * k)) + 1;  // emb_padding_idx+1, nn_utils.py:196, model_snv.py:322
  m->k1 = 5 * n_cat;                 // embedding width is hard-coded to 5 (model_snv.py:322,325)
  MURAL_CHECK(cfg->n_cont >= 0 && cfg->n_cont <= 64, "n_cont must be in [0,64]");
  m->k1c = m->k1 + cfg->n_cont;
  m->L = L;
  LayoutBuilder lb;
  lb.p("emb_layer.weight", int64_t(m->emb_rows) * 5);
  lb.p("lin_layers.0.weight", int64_t(cfg->hidden1) * m->k1c);
  lb.p("lin_layers.0.bias", cfg->hidden1);
  lb.p("lin_layers.1.weight", int64_t(cfg->hidden2) * cfg->hidden1);
  lb.p("lin_layers.1.bias", cfg->hidden2);
  if (cfg->n_cont > 0) lb.bn("first_bn_layer", cfg->n_cont);
  lb.bn("bn_layers.0", cfg->hidden1);
  lb.bn("bn_layers.1", cfg->hidden2);
  for (int br = 0; br < 2; ++br) {
    const std::string s = br ? "_2" : "";
    lb.bn("conv1" + s + ".0", 4);
    lb.conv("conv1" + s + ".1", C, 4, ks);
    for (int g = 1; g <= 2; ++g) {
      for (int i = 0; i < 2; ++i) {
        const std::string rb = "RBs" + std::to_string(g) + s + "." + std::to_string(i);
        lb.bn(rb + ".bn1", C);
        lb.conv(rb + ".conv1", C, C, 3);
        lb.bn(rb + ".bn2", C);
        lb.conv(rb + ".conv2", C, C, 3);
      }
      const std::string cv = "conv" + std::to_string(g + 1) + s;
      lb.bn(cv + ".0", C);
      lb.conv(cv + ".1", C, C, ks);
    }
    const std::string fc = br ? "distal_fc2" : "distal_fc1";
    lb.bn(fc + ".0", C);
    lb.p(fc + ".2.weight", int64_t(cfg->n_class) * C);
    lb.p(fc + ".2.bias", cfg->n_class);
  }
  lb.p("local_fc.0.weight", int64_t(cfg->n_class) * cfg->hidden2);
  lb.p("local_fc.0.bias", cfg->n_class);
  int64_t off = 0;
  for (auto& e : lb.params) { e.offset = off; off += e.numel; m->layout.push_back(e); }
  m->n_trainable = off;
  for (auto& e : lb.buffers) { e.offset = off; off += e.numel; m->layout.push_back(e); }
  m->n_blob = off;
  for (size_t i = 0; i < m->layout.size(); ++i) m->index[m->layout[i].name] = (int)i;
  // geometry of the two branches (model_snv.py:356-371, 399-414)
  const int pools[2][3][3] = {{{3, 3, 1}, {3, 3, 1}, {3, 3, 1}}, {{15, 15, 7}, {7, 7, 3}, {3, 3, 1}}};
  for (int br = 0; br < 2; ++br) {
    BranchDev& B = m->br[br];
    memcpy(B.pool, pools[br], sizeof(B.pool));
    B.L0 = br ? L : 201;
    B.L1 = pool_out(B.L0, B.pool[0][0], B.pool[0][1], B.pool[0][2]);
    B.L2 = pool_out(B.L1, B.pool[1][0], B.pool[1][1], B.pool[1][2]);
    B.L3 = pool_out(B.L2, B.pool[2][0], B.pool[2][1], B.pool[2][2]);
    if (B.L3 < 1) { delete m; MURAL_FAIL("window too short for the pooling pyramid"); }
  }
  *out = m;
  return 0;
}

extern "C" void mural_snv_model_destroy(mural_snv_model_t* m) {
  if (!m) return;
  cudaFree(m->d_prep);
  cudaFree(m->d_chain);
  cudaFree(m->d_ws);
  cudaFree(m->d_io);
  cudaFree(m->d_auto);
  if (m->h_auto) cudaFreeHost(m->h_auto);
  if (m->auto_ev) cudaEventDestroy((cudaEvent_t)m->auto_ev);
  if (m->aux_ev) cudaEventDestroy((cudaEvent_t)m->aux_ev);
  if (m->aux_stream) cudaStreamDestroy((cudaStream_t)m->aux_stream);
  cudaFree(m->d_ws2);
  snv_tc_destroy(m);
  delete m;
}
extern "C" int32_t mural_snv_model_n_tensors(const mural_snv_model_t* m) { return m ? (int32_t)m->layout.size() : 0; }
extern "C" int64_t mural_snv_model_n_params(const mural_snv_model_t* m) { return m ? m->n_blob : 0; }
extern "C" int64_t mural_snv_model_n_trainable(const mural_snv_model_t* m) { return m ? m->n_trainable : 0; }
extern "C" int mural_snv_model_tensor(const mural_snv_model_t* m, int32_t i, const char** name, int64_t* offset,
                                      int64_t* numel, int32_t* is_buffer) {
  MURAL_CHECK(m && i >= 0 && i < (int32_t)m->layout.size(), "tensor index out of range");
  const TensorEntry& e = m->layout[i];
  if (name) *name = e.name.c_str();
  if (offset) *offset = e.offset;
  if (numel) *numel = e.numel;
  if (is_buffer) *is_buffer = e.is_buffer;
  return 0;
}

namespace {

struct Folder {
  const mural_snv_model* m;
  const float* blob;
  std::vector<float> prep;  // host image of d_prep
  const float* T(const std::string& n) const { return blob + m->layout[m->index.at(n)].offset; }
  int64_t alloc(int64_t n) {
    int64_t o = (int64_t)prep.size();
    prep.resize(o + ((n + 3) & ~int64_t(3)), 0.f);  // 16-byte granularity so float4 loads stay aligned
    return o;
  }
  // BN eval affine: a = gamma/sqrt(var+eps), b = beta - mean*a
  void bn_affine(const std::string& n, int c, std::vector<double>& a, std::vector<double>& b) const {
    const float *w = T(n + ".weight"), *be = T(n + ".bias"), *mu = T(n + ".running_mean"), *var = T(n + ".running_var");
    a.resize(c);
    b.resize(c);
    for (int i = 0; i < c; ++i) {
      a[i] = double(w[i]) / sqrt(double(var[i]) + BN_EPS);
      b[i] = double(be[i]) - double(mu[i]) * a[i];
    }
  }
};

// offsets (into prep) of one conv layer
struct ConvOff { int64_t Wt, bias, a, b; int ks, relu_in; };

ConvOff fold_conv(Folder& F, const std::string& bn, const std::string& conv, int C, int ks, int relu_in) {
  ConvOff o;
  o.ks = ks;
  o.relu_in = relu_in;
  o.Wt = F.alloc(int64_t(ks) * C * C);
  o.bias = F.alloc(C);
  o.a = F.alloc(C);
  o.b = F.alloc(C);
  const float* W = F.T(conv + ".weight");  // [co][ci][tap]
  const float* bi = F.T(conv + ".bias");
  for (int co = 0; co < C; ++co)
    for (int ci = 0; ci < C; ++ci)
      for (int t = 0; t < ks; ++t) F.prep[o.Wt + (int64_t(t) * C + ci) * C + co] = W[(int64_t(co) * C + ci) * ks + t];
  std::vector<double> a, b;
  F.bn_affine(bn, C, a, b);
  for (int c = 0; c < C; ++c) {
    F.prep[o.bias + c] = bi[c];
    F.prep[o.a + c] = (float)a[c];
    F.prep[o.b + c] = (float)b[c];
  }
  return o;
}

const double ONEHOT[16][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}, {.5, 0, .5, 0}, {0, .5, 0, .5},
                              {.5, .5, 0, 0}, {0, .5, .5, 0}, {.5, 0, 0, .5}, {0, 0, .5, .5},
                              {0, 1, 1, 1}, {1, 0, 1, 1}, {1, 1, 0, 1}, {1, 1, 1, 0}, {.25, .25, .25, .25}, {0, 0, 0, 0}};

}  // namespace

extern "C" int mural_snv_model_load(mural_snv_model_t* m, const float* h_blob, int64_t n) {
  MURAL_CHECK(m && h_blob, "NULL argument");
  MURAL_CHECK(n == m->n_blob, "parameter blob has the wrong length");
  for (int64_t i = 0; i < n; ++i) MURAL_CHECK(isfinite(h_blob[i]), "non-finite value in the parameter blob");
  const mural_snv_config_t& cfg = m->cfg;
  const int C = cfg.channels, ks = cfg.kernel_size, H1 = cfg.hidden1, H2 = cfg.hidden2, NC = cfg.n_class, K1 = m->k1;
  Folder F{m, h_blob, {}};
  // ---- local branch (model_snv.py:452-468, 492).  BN sits AFTER the ReLU, so it folds forward into
  // the next Linear: W2' = W2*diag(a1), b2' = b2 + W2*b1.
  int64_t o_emb = F.alloc(int64_t(m->emb_rows) * 5);
  memcpy(&F.prep[o_emb], F.T("emb_layer.weight"), sizeof(float) * m->emb_rows * 5);
  const int NCONT = cfg.n_cont, K1C = m->k1c;
  int64_t o_W1t = F.alloc(int64_t(K1C) * H1), o_b1 = F.alloc(H1);
  {
    const float* W = F.T("lin_layers.0.weight");  // [H1][K1 + n_cont]
    const float* bi = F.T("lin_layers.0.bias");
    // continuous features: first_bn_layer (eval affine a, b) sits in FRONT of the Linear (model_snv.py:457-463), so it folds
    // into the extra input columns: W' = W*diag(a), bias' = bias + W*b
    std::vector<double> a, b;
    if (NCONT > 0) F.bn_affine("first_bn_layer", NCONT, a, b);
    for (int o = 0; o < H1; ++o) {
      double acc = bi[o];
      for (int k = 0; k < K1C; ++k) {
        double w = W[int64_t(o) * K1C + k];
        if (k >= K1) { acc += w * b[k - K1]; w *= a[k - K1]; }
        F.prep[o_W1t + int64_t(k) * H1 + o] = float(w);
      }
      F.prep[o_b1 + o] = float(acc);
    }
  }
  int64_t o_W2t = F.alloc(int64_t(H1) * H2), o_b2 = F.alloc(H2);
  {
    std::vector<double> a, b;
    F.bn_affine("bn_layers.0", H1, a, b);
    const float *W = F.T("lin_layers.1.weight"), *bi = F.T("lin_layers.1.bias");  // [H2][H1]
    for (int o = 0; o < H2; ++o) {
      double acc = bi[o];
      for (int k = 0; k < H1; ++k) {
        F.prep[o_W2t + int64_t(k) * H2 + o] = (float)(double(W[int64_t(o) * H1 + k]) * a[k]);
        acc += double(W[int64_t(o) * H1 + k]) * b[k];
      }
      F.prep[o_b2 + o] = (float)acc;
    }
  }
  int64_t o_W3t = F.alloc(int64_t(H2) * NC), o_b3 = F.alloc(NC);
  {
    std::vector<double> a, b;
    F.bn_affine("bn_layers.1", H2, a, b);
    const float *W = F.T("local_fc.0.weight"), *bi = F.T("local_fc.0.bias");  // [NC][H2]
    for (int o = 0; o < NC; ++o) {
      double acc = bi[o];
      for (int k = 0; k < H2; ++k) {
        F.prep[o_W3t + int64_t(k) * NC + o] = (float)(double(W[int64_t(o) * H2 + k]) * a[k]);
        acc += double(W[int64_t(o) * H2 + k]) * b[k];
      }
      F.prep[o_b3 + o] = (float)acc;
    }
  }
  // ---- CNN branches
  struct BrOff { int64_t T, bias1, T4, Wfc, bfc; ConvOff rb1[4], conv2, rb2[4], conv3; } bo[2];
  for (int br = 0; br < 2; ++br) {
    const std::string s = br ? "_2" : "";
    // stem table: BN(4) then Conv1d(4->C) applied to a one-hot column == table lookup per tap.
    // Padding is applied AFTER the BN (nn.Sequential(BatchNorm1d, Conv1d(padding)), model_snv.py:350-353),
    // so an out-of-range tap contributes exactly 0: symbol 15 (PAD) has an all-zero row.
    std::vector<double> a, b;
    F.bn_affine("conv1" + s + ".0", 4, a, b);
    const float* W = F.T("conv1" + s + ".1.weight");  // [C][4][ks]
    bo[br].T = F.alloc(int64_t(ks) * 16 * C);
    bo[br].bias1 = F.alloc(C);
    memcpy(&F.prep[bo[br].bias1], F.T("conv1" + s + ".1.bias"), sizeof(float) * C);
    for (int t = 0; t < ks; ++t)
      for (int sy = 0; sy < 15; ++sy)
        for (int co = 0; co < C; ++co) {
          double acc = 0;
          for (int c = 0; c < 4; ++c) {
            // the encoder feeds float32 one-hot values (1/3 is float32(1/3)); BN runs in fp32 on them
            double e = ONEHOT[sy][c];
            if (sy >= 10 && sy <= 13) e = e ? double(float(1.0 / 3.0)) : 0.0;
            acc += double(W[(int64_t(co) * 4 + c) * ks + t]) * (a[c] * e + b[c]);
          }
          F.prep[bo[br].T + (int64_t(t) * 16 + sy) * C + co] = (float)acc;
        }
    // 4-mer table of the fast stem (ks == 3): both positions of a pair share one lookup.  Built in float with
    // the same summation order as the per-tap path (bias + tap0 + tap1 + tap2) so both paths agree bitwise.
    bo[br].T4 = -1;
    if (ks == 3) {
      bo[br].T4 = F.alloc(int64_t(256) * C);
      for (int k4 = 0; k4 < 256; ++k4) {
        const int b0 = k4 & 3, b1 = (k4 >> 2) & 3, b2 = (k4 >> 4) & 3, b3 = (k4 >> 6) & 3;
        for (int co = 0; co < C; ++co) {
          const float* Tt = &F.prep[bo[br].T];
          float v1 = F.prep[bo[br].bias1 + co], v2 = v1;
          v1 += Tt[(0 * 16 + b0) * C + co]; v1 += Tt[(1 * 16 + b1) * C + co]; v1 += Tt[(2 * 16 + b2) * C + co];
          v2 += Tt[(0 * 16 + b1) * C + co]; v2 += Tt[(1 * 16 + b2) * C + co]; v2 += Tt[(2 * 16 + b3) * C + co];
          F.prep[bo[br].T4 + int64_t(k4) * C + co] = v1 > v2 ? v1 : v2;
        }
      }
    }
    for (int g = 1; g <= 2; ++g) {
      ConvOff* rb = g == 1 ? bo[br].rb1 : bo[br].rb2;
      for (int i = 0; i < 2; ++i) {
        const std::string p = "RBs" + std::to_string(g) + s + "." + std::to_string(i);
        rb[2 * i] = fold_conv(F, p + ".bn1", p + ".conv1", C, 3, 1);
        rb[2 * i + 1] = fold_conv(F, p + ".bn2", p + ".conv2", C, 3, 1);
      }
      const std::string cv = "conv" + std::to_string(g + 1) + s;
      (g == 1 ? bo[br].conv2 : bo[br].conv3) = fold_conv(F, cv + ".0", cv + ".1", C, ks, 0);
    }
    const std::string fc = br ? "distal_fc2" : "distal_fc1";
    F.bn_affine(fc + ".0", C, a, b);
    bo[br].Wfc = F.alloc(int64_t(C) * NC);
    bo[br].bfc = F.alloc(NC);
    const float *Wf = F.T(fc + ".2.weight"), *bf = F.T(fc + ".2.bias");  // [NC][C]
    for (int o = 0; o < NC; ++o) {
      double acc = bf[o];
      for (int c = 0; c < C; ++c) {
        F.prep[bo[br].Wfc + int64_t(c) * NC + o] = (float)(double(Wf[int64_t(o) * C + c]) * a[c]);
        acc += double(Wf[int64_t(o) * C + c]) * b[c];
      }
      F.prep[bo[br].bfc + o] = (float)acc;
    }
  }
  // ---- upload
  CUDA_TRY(cudaSetDevice(m->device));
  if (m->prep_floats != (int64_t)F.prep.size()) {
    cudaFree(m->d_prep);
    m->d_prep = nullptr;
    CUDA_TRY(cudaMalloc((void**)&m->d_prep, F.prep.size() * sizeof(float)));
    m->prep_floats = (int64_t)F.prep.size();
  }
  CUDA_TRY(cudaMemcpy(m->d_prep, F.prep.data(), F.prep.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float* D = m->d_prep;
  m->local = LocalDev{D + o_emb, D + o_W1t, D + o_b1, D + o_W2t, D + o_b2, D + o_W3t, D + o_b3};
  auto mk = [&](const ConvOff& o) { return ConvLayerDev{D + o.Wt, D + o.bias, D + o.a, D + o.b, o.ks, o.relu_in}; };
  for (int br = 0; br < 2; ++br) {
    BranchDev& B = m->br[br];
    B.T = D + bo[br].T;
    B.bias1 = D + bo[br].bias1;
    B.T4 = bo[br].T4 >= 0 ? D + bo[br].T4 : nullptr;
    B.Wfc = D + bo[br].Wfc;
    B.bfc = D + bo[br].bfc;
    for (int i = 0; i < 4; ++i) { B.rb1[i] = mk(bo[br].rb1[i]); B.rb2[i] = mk(bo[br].rb2[i]); }
    B.conv2 = mk(bo[br].conv2);
    B.conv3 = mk(bo[br].conv3);
  }
  if (int rc = snv_tc_prepare(m, h_blob)) return rc;
  m->chain_ready = false;
  m->loaded = true;
  return 0;
}
