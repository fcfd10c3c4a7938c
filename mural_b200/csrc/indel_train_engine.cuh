// Tape engine of the MuRaL-indel training step (units, program of UNet_Small, forward / backward drivers) on top of
// indel_train_core.cuh.  Built by nvcc into libmural_b200.so (indel_train.cu) and by g++ into the host emulation.
#pragma once
#include "indel_train_core.cuh"

namespace indel_train {

struct Config { int radius, channels, kernel_size, n_class, down[6], use_reverse; };
struct Tensor { int C, L; int64_t off; };  // [B, C, L]; values at vals + off * B, gradients at grads + off * B
struct Unit {
  int in = -1, out = -1, t = -1;           // tensor ids: input, output, conv output before BN (== in when there is no conv)
  bool has_conv = false, has_bn = false;
  int64_t W = -1, b = -1, gamma = -1, beta = -1, rm = -1, rv = -1;  // offsets into the parameter blob
  int Cin = 0, Cout = 0, k = 1, stride = 1, up = 1, act = ACT_NONE, res1 = -1, res2 = -1;
  float p_drop = 0.f;
  int stat_slot = 0;                        // mean / invstd storage: stats + stat_slot
};
enum OpKind { OP_UNIT, OP_FLIP_CL, OP_ADD_FLIPL, OP_MAXL };
struct Op { OpKind kind; int a, b, o, unit; };

struct Engine {
  Config cfg;
  std::map<std::string, int64_t> off;       // parameter name -> offset in the blob (layout of mural_indel_model_tensor)
  std::vector<Tensor> tensors;
  std::vector<Unit> units;
  std::vector<Op> ops;
  int t_in = -1, t_out = -1;
  int64_t per_site = 0, max_per_site = 0;   // floats per site over all tensors / of the largest tensor
  int n_stat = 0;
  // buffers (device or host, see Exec)
  Exec ex;
  int64_t cap = 0;
  float *vals = nullptr, *grads = nullptr, *dz = nullptr, *dz_b = nullptr, *dxv = nullptr, *stats = nullptr;
  int32_t* arg = nullptr;
  double* dstat = nullptr;
  uint64_t seed = 0;
  uint32_t* step_mem = nullptr;   // dropout stream position (device memory in the CUDA build), bumped by every forward

  int tensor(int C, int L) { tensors.push_back({C, L, per_site}); per_site += int64_t(C) * L;
    if (int64_t(C) * L > max_per_site) max_per_site = int64_t(C) * L; return int(tensors.size()) - 1; }

  int unit(int in, const std::string& conv, bool bias, int Cout, int k, int stride, int up, const std::string& bn, int act,
           int res1 = -1, int res2 = -1, float p_drop = 0.f) {
    Unit u;
    u.in = in; u.act = act; u.res1 = res1; u.res2 = res2; u.p_drop = p_drop;
    const Tensor ti = tensors[in];
    int Lout = ti.L, C = ti.C;
    if (!conv.empty()) {
      u.has_conv = true; u.W = off.at(conv + ".weight"); u.b = bias ? off.at(conv + ".bias") : -1;
      u.Cin = ti.C; u.Cout = Cout; u.k = k; u.stride = stride; u.up = up;
      Lout = (ti.L * up + 2 * ((k - 1) / 2) - k) / stride + 1;
      C = Cout;
      u.t = tensor(C, Lout);
    } else {
      u.t = in;
    }
    if (!bn.empty()) {
      u.has_bn = true; u.gamma = off.at(bn + ".weight"); u.beta = off.at(bn + ".bias");
      u.rm = off.at(bn + ".running_mean"); u.rv = off.at(bn + ".running_var");
      u.stat_slot = n_stat; n_stat += 2 * C;
    }
    u.out = tensor(C, Lout);
    units.push_back(u);
    ops.push_back({OP_UNIT, -1, -1, -1, int(units.size()) - 1});
    return u.out;
  }

  // UNet_Small.forward (model_indel.py:151-176)
  void build() {
    const int C = cfg.channels, ks = cfg.kernel_size, L = 2 * cfg.radius;
    int ch[6]; for (int i = 0; i < 6; ++i) ch[i] = C * (i + 1);
    t_in = tensor(4, L);
    int cur = t_in;
    if (cfg.use_reverse) {                                                       // :154-155
      const int xf = tensor(4, L);
      ops.push_back({OP_FLIP_CL, t_in, -1, xf, -1});
      const int A = unit(t_in, "conv.0", true, 4, ks, 1, 1, "conv.1", ACT_NONE);
      const int Bf = unit(xf, "conv.0", true, 4, ks, 1, 1, "conv.1", ACT_NONE);
      const int O = tensor(4, L);
      ops.push_back({OP_ADD_FLIPL, A, Bf, O, -1});
      cur = O;
    }
    int enc[6];
    for (int i = 0; i < 6; ++i) {                                                // :158-163
      const std::string s = std::to_string(i);
      const int lo = unit(cur, "uplblocks." + s + ".0", true, ch[i], ks, cfg.down[i], 1, "uplblocks." + s + ".1", ACT_NONE);
      const int h = unit(lo, "upblocks." + s + ".0.conv.0", false, 2 * ch[i], 5, 1, 1, "upblocks." + s + ".0.conv.1", ACT_SILU);
      cur = enc[i] = unit(h, "upblocks." + s + ".0.conv.3", false, ch[i], 1, 1, 1, "upblocks." + s + ".0.conv.4", ACT_NONE, lo);
    }
    for (int i = 0; i < 5; ++i) {                                                // :165-170
      const std::string s = std::to_string(i);
      const int c = ch[4 - i];
      const int lo = unit(cur, "downlblocks." + s + ".1", true, c, ks, 1, cfg.down[5 - i], "downlblocks." + s + ".2", ACT_NONE);
      const int h = unit(lo, "downblocks." + s + ".0.conv.0", false, 2 * c, 5, 1, 1, "downblocks." + s + ".0.conv.1", ACT_SILU);
      cur = unit(h, "downblocks." + s + ".0.conv.3", false, c, 1, 1, 1, "downblocks." + s + ".0.conv.4", ACT_NONE, lo, enc[4 - i]);
    }
    cur = unit(cur, "out_conv.0", true, C, 1, 1, 1, "out_conv.1", ACT_RELU);     // :172
    cur = unit(cur, "out_conv.3", true, C, 1, 1, 1, "", ACT_SOFTPLUS);
    const int mx = tensor(C, 1);
    ops.push_back({OP_MAXL, cur, -1, mx, -1});                                   // :173
    cur = unit(mx, "", false, 0, 1, 1, 1, "out_fc.0", ACT_NONE, -1, -1, 0.1f);   // :174 BatchNorm1d -> Dropout(0.1)
    t_out = unit(cur, "out_fc.2", true, cfg.n_class, 1, 1, 1, "", ACT_SOFTPLUS); //      Linear -> Softplus
  }

  void ensure(int64_t B) {
    if (B <= cap) return;
    ex.free_(vals); ex.free_(grads); ex.free_(dz); ex.free_(dz_b); ex.free_(dxv); ex.free_(stats); ex.free_(arg); ex.free_(dstat);
    vals = (float*)ex.alloc(sizeof(float) * per_site * B);
    grads = (float*)ex.alloc(sizeof(float) * per_site * B);
    dz = (float*)ex.alloc(sizeof(float) * max_per_site * B);
    dz_b = (float*)ex.alloc(sizeof(float) * max_per_site * B);  // second slot: see Exec::wait_w
    dxv = (float*)ex.alloc(sizeof(float) * max_per_site * B);   // gradient of an upsampled input before it is folded back (tiled path)
    stats = (float*)ex.alloc(sizeof(float) * (n_stat + 1));
    arg = (int32_t*)ex.alloc(sizeof(int32_t) * cfg.channels * 6 * B);
    dstat = (double*)ex.alloc(sizeof(double) * 2 * 1024);
    if (!step_mem) { step_mem = (uint32_t*)ex.alloc(sizeof(uint32_t)); ex.zero(step_mem, sizeof(uint32_t)); }
    cap = B;
  }
  float* V(int t, int64_t B) { return vals + tensors[t].off * B; }
  float* G(int t, int64_t B) { return grads + tensors[t].off * B; }

  UnitOut unit_out(const Unit& u, float* P, int64_t B) {
    const Tensor to = tensors[u.out];
    UnitOut o;
    o.t = V(u.t, B); o.mean = stats + u.stat_slot; o.invstd = stats + u.stat_slot + to.C;
    o.gamma = u.has_bn ? P + u.gamma : nullptr; o.beta = u.has_bn ? P + u.beta : nullptr;
    o.res1 = u.res1 >= 0 ? V(u.res1, B) : nullptr; o.res2 = u.res2 >= 0 ? V(u.res2, B) : nullptr;
    o.y = V(u.out, B); o.C = to.C; o.L = to.L; o.act = u.act; o.p = u.p_drop; o.seed = seed ^ (uint64_t(u.out) << 48); o.step_p = step_mem;
    return o;
  }
  ConvDims dims(const Unit& u, int64_t B) {
    const Tensor ti = tensors[u.in], tt = tensors[u.t];
    return ConvDims{int(B), u.Cin, ti.L, u.Cout, tt.L, u.k, u.stride, (u.k - 1) / 2, u.up};
  }

  // x: [B, 4, L] one-hot windows already in vals of t_in (caller copies / gathers them there); P: parameter blob (running
  // statistics are updated in place); out: [B, n_class]
  void forward(float* P, int64_t B) {
    ex.run(1, BumpStep{step_mem});
    for (const Op& op : ops) {
      if (op.kind == OP_FLIP_CL) {
        const Tensor t = tensors[op.a];
        ex.run(B * t.C * t.L, FlipCL{V(op.a, B), V(op.o, B), t.C, t.L});
      } else if (op.kind == OP_ADD_FLIPL) {
        const Tensor t = tensors[op.a];
        ex.run(B * t.C * t.L, AddFlipL{V(op.a, B), V(op.b, B), V(op.o, B), t.L});
      } else if (op.kind == OP_MAXL) {
        const Tensor t = tensors[op.a];
        const MaxL f{V(op.a, B), V(op.o, B), arg, t.L};
        if (!ex.max_rows(f, B * t.C)) ex.run(B * t.C, f);
      } else {
        const Unit& u = units[op.unit];
        const Tensor tt = tensors[u.t];
        if (u.has_conv) {
          const ConvDims d = dims(u, B);
          const ConvFwd f{V(u.in, B), P + u.W, u.b >= 0 ? P + u.b : nullptr, V(u.t, B), d};
          if (!ex.conv_fwd(f)) ex.run(B * d.Cout * d.Lout, f);
        }
        if (u.has_bn) {
          ex.zero(dstat, sizeof(double) * 2 * tt.C);
          const BnStats fs{V(u.t, B), dstat, tt.C, tt.L};
          if (!ex.row_reduce(fs, B * tt.C, dstat)) ex.run(B * tt.C * ROW_SPLIT, fs);
          ex.run(tt.C, BnFinalize{dstat, double(B) * tt.L, tt.C, P + u.rm, P + u.rv, stats + u.stat_slot, stats + u.stat_slot + tt.C});
        }
        ex.run(B * tt.C * tt.L, unit_out(u, P, B));
      }
    }
  }

  // dOut (gradient of the loss w.r.t. the network output) must be in grads of t_out; every other gradient buffer is zeroed
  // here first.  Gp: parameter gradients (+=), same layout as the blob.
  void backward(float* P, float* Gp, int64_t B, const float* d_out) {
    ex.zero(grads, sizeof(float) * per_site * B);
    ex.run(B * cfg.n_class, AddTo{d_out, G(t_out, B)});
    float* const dz_slot[2] = {this->dz, dz_b};
    int par = 0;
    for (int oi = int(ops.size()) - 1; oi >= 0; --oi) {
      const Op& op = ops[oi];
      if (op.kind == OP_FLIP_CL) continue;  // input of the network: no gradient needed
      if (op.kind == OP_ADD_FLIPL) {
        const Tensor t = tensors[op.a];
        ex.run(B * t.C * t.L, AddFlipLBwd{G(op.o, B), G(op.a, B), G(op.b, B), t.L});
      } else if (op.kind == OP_MAXL) {
        const Tensor t = tensors[op.a];
        ex.run(B * t.C, MaxLBwd{G(op.o, B), arg, G(op.a, B), t.L});
      } else {
        const Unit& u = units[op.unit];
        const Tensor tt = tensors[u.t];
        const int64_t n = B * tt.C * tt.L;
        par ^= 1;
        float* const dz = dz_slot[par];
        ex.wait_w(par);                      // the weight-gradient kernel that read this slot two units ago
        if (u.res1 >= 0) ex.run(n, AddTo{G(u.out, B), G(u.res1, B)});
        if (u.res2 >= 0) ex.run(n, AddTo{G(u.out, B), G(u.res2, B)});
        const UnitOut uo = unit_out(u, P, B);
        ex.zero(dstat, sizeof(double) * 2 * tt.C);
        const UnitBwdReduce fr{uo, G(u.out, B), dz, dstat};
        if (!ex.row_reduce(fr, B * tt.C, dstat)) ex.run(B * tt.C * ROW_SPLIT, fr);
        if (u.has_bn) {
          ex.run(tt.C, BnParamGrad{dstat, tt.C, Gp + u.gamma, Gp + u.beta});
          ex.run(n, UnitBwdApply{uo, dz, dstat, double(B) * tt.L});
        }
        if (u.has_conv) {
          const ConvDims d = dims(u, B);
          const ConvBwdW fw{V(u.in, B), dz, Gp + u.W, u.b >= 0 ? Gp + u.b : nullptr, d};
          ex.begin_w();                      // side stream (CUDA build): parameter gradients are read by the optimizer only
          if (!ex.conv_bwd_w(fw)) ex.run(int64_t(d.Cout) * d.Cin * d.k * B * CONV_W_SPLIT, fw);
          ex.end_w(par);
          if (u.in != t_in) {
            const ConvBwdX fx{dz, P + u.W, G(u.in, B), d};
            if (!ex.conv_bwd_x(fx, dxv)) ex.run(B * d.Cin * d.Lin, fx);
          }
        } else {
          ex.run(n, AddTo{dz, G(u.in, B)});
        }
      }
    }
    ex.wait_w(0);                            // the caller's stream continues (all-reduce, optimizer) after every weight gradient
    ex.wait_w(1);
  }
};

}  // namespace indel_train
