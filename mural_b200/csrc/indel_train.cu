// Training step of MuRaL-indel (UNet_Small): train-mode forward (batch-statistic BatchNorm with running-stat update, dropout
// in out_fc) and full backward in fp32.  Reference: MuRaL/model/model_indel.py:6-176 (network), MuRaL/training.py:404-452 (loop
// body; the loss is CrossEntropyLoss(sum) applied to the Softplus outputs, mural_ce_sum_grad).  SURVEY 8(d) config 4.
//
// The network is run as a tape of generic ops (indel_train_engine.cuh): unit = conv -> BatchNorm -> activation (+ residuals),
// flips of the reverse-strand stem, max over positions.  The arithmetic of every op lives in indel_train_core.cuh and was
// checked against fp64 autograd of the oracle through a host build of the same source (tests/emu/indel_train_emu.cpp, tests/test_indel_train_emu.py).
#include "common.cuh"

#define INDEL_TRAIN_LAUNCH(name, kernel, grid, block, stream, ...) LAUNCH_N(name, kernel, grid, block, 0, stream, __VA_ARGS__)
#define INDEL_TRAIN_LAUNCH_SMEM(name, kernel, grid, block, smem, stream, ...) LAUNCH_N(name, kernel, grid, block, smem, stream, __VA_ARGS__)
#include "indel_train_engine.cuh"

using namespace mural;

struct mural_indel_train {
  mural_indel_model_t* m = nullptr;
  indel_train::Engine E;
  int64_t n = 0;  // sites of the last forward
  int64_t n_blob = 0;
};

extern "C" int mural_indel_train_create(mural_indel_model_t* m, mural_indel_train_t** out) {
  MURAL_CHECK(m && out, "NULL argument");
  *out = nullptr;
  mural_indel_config_t cfg;
  if (int rc = mural_indel_model_config(m, &cfg)) return rc;
  mural_indel_train* T = new mural_indel_train();
  T->m = m;
  T->E.cfg = indel_train::Config{cfg.distal_radius, cfg.channels, cfg.kernel_size, cfg.n_class,
                                 {cfg.downsize[0], cfg.downsize[1], cfg.downsize[2], cfg.downsize[3], cfg.downsize[4], cfg.downsize[5]},
                                 cfg.use_reverse};
  for (int32_t i = 0; i < mural_indel_model_n_tensors(m); ++i) {
    const char* name = nullptr;
    int64_t off = 0, num = 0;
    int32_t buf = 0;
    mural_indel_model_tensor(m, i, &name, &off, &num, &buf);
    T->E.off[name] = off;
  }
  T->n_blob = mural_indel_model_n_params(m);
  T->E.build();
  // the decoder adds encoder outputs: lengths must line up (checked by mural_indel_model_create as well)
  *out = T;
  return 0;
}

extern "C" void mural_indel_train_destroy(mural_indel_train_t* T) {
  if (!T) return;
  indel_train::Engine& E = T->E;
  cudaFree(E.vals); cudaFree(E.grads); cudaFree(E.dz); cudaFree(E.dz_b); cudaFree(E.dxv); cudaFree(E.stats); cudaFree(E.arg); cudaFree(E.dstat); cudaFree(E.step_mem);
  if (E.ex.side_st) {
    cudaStreamDestroy(E.ex.side_st);
    cudaEventDestroy(E.ex.ev_dz);
    cudaEventDestroy(E.ex.ev_w[0]);
    cudaEventDestroy(E.ex.ev_w[1]);
  }
  delete T;
}

extern "C" int mural_indel_train_set_dropout(mural_indel_train_t* T, float p_fc, uint64_t seed) {
  MURAL_CHECK(T, "NULL argument");
  MURAL_CHECK(p_fc >= 0 && p_fc < 1, "dropout must be in [0,1)");
  for (auto& u : T->E.units)
    if (!u.has_conv && u.has_bn) u.p_drop = p_fc;  // out_fc: BatchNorm1d -> Dropout -> Linear (model_indel.py:146-149)
  T->E.seed = seed;
  return 0;
}

static int train_forward(mural_indel_train* T, int64_t n, float* d_blob, float* d_out, cudaStream_t st) {
  indel_train::Engine& E = T->E;
  E.ex.st = st;
  E.forward(d_blob, n);
  CUDA_TRY(cudaMemcpyAsync(d_out, E.V(E.t_out, n), sizeof(float) * n * E.cfg.n_class, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaGetLastError());
  T->n = n;
  return 0;
}

extern "C" int mural_indel_train_forward(mural_indel_train_t* T, const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta,
                                         int64_t n, float* d_blob, float* d_out, void* stream) {
  MURAL_CHECK(T && g && d_pos && d_meta && d_blob && d_out, "NULL argument");
  MURAL_CHECK(n >= 2, "BatchNorm in training mode needs more than one site per batch");  // training.py:415 skips them
  indel_train::Engine& E = T->E;
  E.ensure(n);
  MURAL_CHECK(E.vals && E.grads && E.dz && E.dz_b, "cudaMalloc of the training workspace failed");
  // one-hot windows of the batch, [n, 4, 2R] (seq_ohe_encoder, preprocessing.py:756-816), straight into the input tensor
  if (int rc = mural_encode_onehot(g, d_pos, d_meta, n, E.cfg.radius, MURAL_MODEL_INDEL, E.V(E.t_in, n), stream)) return rc;
  return train_forward(T, n, d_blob, d_out, (cudaStream_t)stream);
}

extern "C" int mural_indel_train_forward_tensors(mural_indel_train_t* T, const float* d_distal, int64_t n, int32_t L, float* d_blob,
                                                 float* d_out, void* stream) {
  MURAL_CHECK(T && d_distal && d_blob && d_out, "NULL argument");
  MURAL_CHECK(n >= 2, "BatchNorm in training mode needs more than one site per batch");
  indel_train::Engine& E = T->E;
  MURAL_CHECK(L == 2 * E.cfg.radius, "distal_x length does not match the model's distal_radius");
  E.ensure(n);
  MURAL_CHECK(E.vals && E.grads && E.dz && E.dz_b, "cudaMalloc of the training workspace failed");
  CUDA_TRY(cudaMemcpyAsync(E.V(E.t_in, n), d_distal, sizeof(float) * n * 4 * L, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return train_forward(T, n, d_blob, d_out, (cudaStream_t)stream);
}

extern "C" int mural_indel_train_backward(mural_indel_train_t* T, float* d_blob, const float* d_dout, float* d_grads, void* stream) {
  MURAL_CHECK(T && d_blob && d_dout && d_grads, "NULL argument");
  MURAL_CHECK(T->n > 0, "backward without a preceding forward");
  indel_train::Engine& E = T->E;
  cudaStream_t st = (cudaStream_t)stream;
  E.ex.st = st;
  CUDA_TRY(cudaMemsetAsync(d_grads, 0, sizeof(float) * T->n_blob, st));
  E.backward(d_blob, d_grads, T->n, d_dout);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
