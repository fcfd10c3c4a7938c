// TEST INFRASTRUCTURE: host build (g++ -DINDEL_EMU) of the indel training tape.  The op arithmetic and the tape of
// mural_b200/csrc/indel_train_{core,engine}.cuh are compiled as plain C++ — every "kernel" is a serial loop over its work
// items on host memory — so that tests/test_indel_train_emu.py can check forward, loss and every gradient against fp64
// autograd of the oracle without a GPU.  Never linked into libmural_b200.so.
#include "../../mural_b200/csrc/indel_train_engine.cuh"
using namespace indel_train;

extern "C" int indel_train_emu_step(int radius, int channels, int ks, int n_class, const int* down, int use_reverse, int n_names,
                                    const char** names, const int64_t* offsets, float* blob, int64_t n_blob, const float* x_onehot,
                                    const int32_t* labels, int64_t B, float dropout_off, float* out, float* grads, double* loss) {
  Engine E;
  E.cfg = Config{radius, channels, ks, n_class, {down[0], down[1], down[2], down[3], down[4], down[5]}, use_reverse};
  for (int i = 0; i < n_names; ++i) E.off[names[i]] = offsets[i];
  E.build();
  if (dropout_off != 0.f) for (auto& u : E.units) u.p_drop = 0.f;
  E.ensure(B);
  memcpy(E.V(E.t_in, B), x_onehot, sizeof(float) * 4 * 2 * radius * B);
  E.forward(blob, B);
  memcpy(out, E.V(E.t_out, B), sizeof(float) * n_class * B);
  std::vector<float> d_out(n_class * B);
  *loss = 0;
  E.ex.run(B, CeGrad{E.V(E.t_out, B), labels, n_class, d_out.data(), loss});
  memset(grads, 0, sizeof(float) * n_blob);
  E.backward(blob, grads, B, d_out.data());
  return int(E.ex.launches);
}
