// Output epilogue of run_predict: softmax of the returned log-probs, FullDirichlet calibrator apply,
// optional Poisson calibration.  Reference: MuRaL/scripts/run_predict.py:214-225,
// dirichlet_python/dirichletcal/calib/fulldirichlet.py:78-80, calib/multinomial.py:60-64,235-244,
// dirichletcal/utils.py:5-7, MuRaL/model/calibration.py:10-23.
#include <float.h>
#include <string.h>

#include "common.cuh"

namespace mural {

constexpr int CAL_MAXK = 16;
struct CalWeights { double w[CAL_MAXK][CAL_MAXK + 1]; };

__global__ void k_calibrate(const float* __restrict__ logp, int64_t n, int K, CalWeights W, int use_w, int poisson,
                            double* __restrict__ out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float* p = logp + i * K;
  // F.softmax(pred_y, dim=1) in fp32 (run_predict.py:214)
  float mx = -FLT_MAX;
  for (int o = 0; o < K; ++o) mx = fmaxf(mx, p[o]);
  float pf[CAL_MAXK];
  float sum = 0.f;
  for (int o = 0; o < K; ++o) { pf[o] = expf(p[o] - mx); sum += pf[o]; }
  double pr[CAL_MAXK];
  for (int o = 0; o < K; ++o) { pf[o] = pf[o] / sum; pr[o] = double(pf[o]); }
  if (use_w) {
    // clip_for_log in the input dtype (float32: tiny = FLT_MIN, 1-tiny rounds to 1), log in float32
    double s[CAL_MAXK];
    for (int o = 0; o < K; ++o) s[o] = double(logf(fminf(fmaxf(pf[o], FLT_MIN), 1.0f)));
    double z[CAL_MAXK], zm = -DBL_MAX;
    for (int r = 0; r < K; ++r) {
      double acc = W.w[r][K];  // intercept column (the appended ones, multinomial.py:62)
      for (int c = 0; c < K; ++c) acc += s[c] * W.w[r][c];
      z[r] = acc;
      zm = fmax(zm, acc);
    }
    double zs = 0;
    for (int r = 0; r < K; ++r) { z[r] = exp(z[r] - zm); zs += z[r]; }
    for (int r = 0; r < K; ++r) pr[r] = z[r] / zs;
  }
  if (poisson) {  // calibration.py:10-23
    const double p0 = fmin(fmax(pr[0], 1e-10), 1.0);
    const double lam = -log(p0);
    for (int o = 1; o < K; ++o) pr[o] = lam * pr[o] / (1.0 - p0);
    pr[0] = 1.0 - lam;
  }
  for (int o = 0; o < K; ++o) out[i * K + o] = pr[o];
}

}  // namespace mural
using namespace mural;

extern "C" int mural_calibrate(const float* d_logp, int64_t n, int32_t n_class, const double* h_weights, int32_t poisson,
                               double* d_prob, void* stream) {
  MURAL_CHECK(d_logp && d_prob, "NULL argument");
  MURAL_CHECK(n_class >= 2 && n_class <= CAL_MAXK, "n_class out of range");
  if (n == 0) return 0;
  CalWeights W;
  memset(&W, 0, sizeof(W));
  if (h_weights)
    for (int r = 0; r < n_class; ++r)
      for (int c = 0; c <= n_class; ++c) W.w[r][c] = h_weights[r * (n_class + 1) + c];
  LAUNCH(k_calibrate, (unsigned)cdiv(n, 128), 128, 0, (cudaStream_t)stream, d_logp, n, n_class, W, h_weights != nullptr, poisson,
         d_prob);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
