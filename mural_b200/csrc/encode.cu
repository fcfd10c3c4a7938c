// Bit-exact encoders: k-mer index matrix and one-hot expanded window straight from the packed genome.
// Reference: MuRaL/data/preprocessing.py seq_digit_encoder :636-723, seq_ohe_encoder :756-816.
#include "snv_model.cuh"

namespace mural {

// float bit patterns of the reference's one-hot columns ('+' table, preprocessing.py:758-772)
__constant__ float c_onehot[16][4] = {
    {1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1},
    {.5f, 0, .5f, 0}, {0, .5f, 0, .5f}, {.5f, .5f, 0, 0}, {0, .5f, .5f, 0}, {.5f, 0, 0, .5f}, {0, 0, .5f, .5f},
    {0, (float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3)}, {(float)(1.0 / 3), 0, (float)(1.0 / 3), (float)(1.0 / 3)},
    {(float)(1.0 / 3), (float)(1.0 / 3), 0, (float)(1.0 / 3)}, {(float)(1.0 / 3), (float)(1.0 / 3), (float)(1.0 / 3), 0},
    {.25f, .25f, .25f, .25f}, {0, 0, 0, 0}};

// oriented symbol at oriented window index i of a site
__device__ __forceinline__ int site_symbol(const GenomeView& G, int chrom, int64_t wstart, int W, int strand, int i) {
  const int64_t q = wstart + (strand ? (W - 1 - i) : i);
  int s = SYM_N;
  if (q >= 0 && q < G.chrom_len[chrom]) s = genome_symbol(G, G.chrom_off[chrom] + q);
  return strand ? comp_sym(s) : s;
}

__global__ void k_encode_local(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta,
                               int64_t n, int radius, int order, int W, int n_k, int w_shift,
                               int64_t* __restrict__ out) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t >= n * n_k) return;
  const int64_t site = t / n_k;
  const int j = int(t - site * n_k);
  const int m = meta[site];
  const int strand = m & 1, chrom = int(uint32_t(m) >> 8);
  const int64_t wstart = int64_t(pos[site]) - radius + w_shift;
  int idx = 0;
  bool bad = false;
  for (int d = 0; d < order; ++d) {  // most significant base first (preprocessing.py:710)
    const int s = site_symbol(G, chrom, wstart, W, strand, j + d);
    bad |= (s > 3);
    idx = idx * 4 + (s & 3);
  }
  out[t] = bad ? (int64_t(1) << (2 * order)) : idx;  // 4**order (preprocessing.py:722)
}

// same k-mer indices as k_encode_local but int32, feeding the local-branch MLP pre-pass of the network
__global__ void k_local_idx32(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta, int64_t n,
                              int radius, int order, int W, int n_k, int32_t* __restrict__ out) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t >= n * n_k) return;
  const int64_t site = t / n_k;
  const int j = int(t - site * n_k);
  const int m = meta[site];
  const int strand = m & 1, chrom = int(uint32_t(m) >> 8);
  const int64_t wstart = int64_t(pos[site]) - radius;
  int idx = 0;
  bool bad = false;
  for (int d = 0; d < order; ++d) {
    const int s = site_symbol(G, chrom, wstart, W, strand, j + d);
    bad |= (s > 3);
    idx = idx * 4 + (s & 3);
  }
  out[t] = bad ? (1 << (2 * order)) : idx;
}

__global__ void k_encode_onehot(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta,
                                int64_t n, int radius, int W, int w_shift, float* __restrict__ out) {
  const int64_t site = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= W) return;
  const int m = meta[site];
  const int strand = m & 1, chrom = int(uint32_t(m) >> 8);
  const int64_t wstart = int64_t(pos[site]) - radius + w_shift;
  const int s = site_symbol(G, chrom, wstart, W, strand, i);
  float* o = out + site * 4 * int64_t(W) + i;
#pragma unroll
  for (int c = 0; c < 4; ++c) o[int64_t(c) * W] = c_onehot[s][c];
}

__global__ void k_onehot_to_symbols(const float* __restrict__ x, int64_t n, int W, uint8_t* __restrict__ sym,
                                    int* __restrict__ flag) {
  const int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (t >= n * W) return;
  const int64_t site = t / W;
  const int i = int(t - site * W);
  const float* p = x + site * 4 * int64_t(W) + i;
  const float v0 = p[0], v1 = p[W], v2 = p[2 * int64_t(W)], v3 = p[3 * int64_t(W)];
  int s = 255;
#pragma unroll
  for (int k = 0; k < N_SYM; ++k)
    if (v0 == c_onehot[k][0] && v1 == c_onehot[k][1] && v2 == c_onehot[k][2] && v3 == c_onehot[k][3]) s = k;
  sym[t] = uint8_t(s);
  if (s == 255 && flag) atomicOr(flag, 1);
}

static inline int window_len(int radius, int model_type) { return 2 * radius + (model_type == MURAL_MODEL_SNV ? 1 : 0); }

int snv_local_idx_launch(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta, int64_t n,
                         int32_t* cat32, cudaStream_t st) {
  const int W = 2 * m->cfg.local_radius + 1;
  const int64_t total = n * m->n_cat;
  LAUNCH(k_local_idx32, (unsigned)cdiv(total, 256), 256, 0, st, *G, d_pos, d_meta, n, m->cfg.local_radius, m->cfg.local_order, W,
         m->n_cat, cat32);
  return 0;
}

int onehot_to_symbols_checked(const float* d_onehot, int64_t n, int32_t W, uint8_t* d_sym, int* d_flag, cudaStream_t st) {
  LAUNCH(k_onehot_to_symbols, (unsigned)cdiv(n * W, 256), 256, 0, st, d_onehot, n, W, d_sym, d_flag);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace mural
using namespace mural;

extern "C" int mural_encode_local(const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta, int64_t n,
                                  int32_t radius, int32_t order, int32_t model_type, int64_t* d_out, void* stream) {
  MURAL_CHECK(g && (n == 0 || (d_pos && d_meta && d_out)), "NULL argument");
  MURAL_CHECK(order >= 1 && order <= 12, "local_order must be in [1,12]");
  MURAL_CHECK(model_type == MURAL_MODEL_SNV || model_type == MURAL_MODEL_INDEL, "model_type must be snv or indel");
  const int W = window_len(radius, model_type);
  const int n_k = W - (order - 1);
  MURAL_CHECK(radius >= 0 && n_k > 0, "local window shorter than the k-mer");
  if (n == 0) return 0;
  const int64_t total = n * n_k;
  LAUNCH(k_encode_local, (unsigned)cdiv(total, 256), 256, 0, (cudaStream_t)stream, g->view, d_pos, d_meta, n, radius, order, W,
         n_k, model_type == MURAL_MODEL_SNV ? 0 : 1, d_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_encode_onehot(const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta, int64_t n,
                                   int32_t radius, int32_t model_type, float* d_out, void* stream) {
  MURAL_CHECK(g && (n == 0 || (d_pos && d_meta && d_out)), "NULL argument");
  MURAL_CHECK(model_type == MURAL_MODEL_SNV || model_type == MURAL_MODEL_INDEL, "model_type must be snv or indel");
  const int W = window_len(radius, model_type);
  MURAL_CHECK(radius >= 0 && W > 0, "empty window");
  if (n == 0) return 0;
  for (int64_t s0 = 0; s0 < n; s0 += 65535) {  // gridDim.y limit
    const int64_t ns = (n - s0 < 65535) ? (n - s0) : 65535;
    dim3 grid((unsigned)cdiv(W, 256), (unsigned)ns);
    LAUNCH(k_encode_onehot, grid, 256, 0, (cudaStream_t)stream, g->view, d_pos + s0, d_meta + s0, ns, radius, W,
           model_type == MURAL_MODEL_SNV ? 0 : 1, d_out + s0 * 4 * int64_t(W));
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_onehot_to_symbols(const float* d_onehot, int64_t n, int32_t W, uint8_t* d_sym, void* stream) {
  MURAL_CHECK(d_onehot && d_sym && W > 0, "bad argument");
  if (n == 0) return 0;
  return onehot_to_symbols_checked(d_onehot, n, W, d_sym, nullptr, (cudaStream_t)stream);
}

namespace {
template <typename OutT, typename F>
int host_roundtrip(const mural_genome_t* g, const int32_t* h_pos, const int32_t* h_meta, int64_t n, int64_t out_per_site,
                   OutT* h_out, F&& run) {
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(g->device));
  int32_t *d_pos = nullptr, *d_meta = nullptr;
  OutT* d_out = nullptr;
  // bounded chunks so a genome-wide call never needs more than ~256 MB of staging
  int64_t chunk = (int64_t(256) << 20) / (out_per_site * (int64_t)sizeof(OutT));
  if (chunk < 1) chunk = 1;
  if (chunk > n) chunk = n;
  CUDA_TRY(cudaMalloc(&d_pos, chunk * 4));
  CUDA_TRY(cudaMalloc(&d_meta, chunk * 4));
  CUDA_TRY(cudaMalloc(&d_out, chunk * out_per_site * sizeof(OutT)));
  int rc = 0;
  for (int64_t s0 = 0; s0 < n && rc == 0; s0 += chunk) {
    const int64_t ns = (n - s0 < chunk) ? (n - s0) : chunk;
    cudaMemcpy(d_pos, h_pos + s0, ns * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_meta, h_meta + s0, ns * 4, cudaMemcpyHostToDevice);
    rc = run(d_pos, d_meta, ns, d_out);
    if (rc == 0 && cudaMemcpy(h_out + s0 * out_per_site, d_out, ns * out_per_site * sizeof(OutT), cudaMemcpyDeviceToHost) !=
                       cudaSuccess)
      rc = fail(__FILE__, __LINE__, std::string("D2H: ") + cudaGetErrorString(cudaGetLastError()));
  }
  cudaFree(d_pos);
  cudaFree(d_meta);
  cudaFree(d_out);
  return rc;
}
}  // namespace

extern "C" int mural_encode_local_host(const mural_genome_t* g, const int32_t* h_pos, const int32_t* h_meta, int64_t n,
                                       int32_t radius, int32_t order, int32_t model_type, int64_t* h_out) {
  MURAL_CHECK(g && (n == 0 || (h_pos && h_meta && h_out)), "NULL argument");
  const int n_k = window_len(radius, model_type) - (order - 1);
  MURAL_CHECK(n_k > 0, "local window shorter than the k-mer");
  return host_roundtrip<int64_t>(g, h_pos, h_meta, n, n_k, h_out, [&](int32_t* p, int32_t* m, int64_t ns, int64_t* o) {
    return mural_encode_local(g, p, m, ns, radius, order, model_type, o, nullptr);
  });
}

extern "C" int mural_encode_onehot_host(const mural_genome_t* g, const int32_t* h_pos, const int32_t* h_meta, int64_t n,
                                        int32_t radius, int32_t model_type, float* h_out) {
  MURAL_CHECK(g && (n == 0 || (h_pos && h_meta && h_out)), "NULL argument");
  const int W = window_len(radius, model_type);
  MURAL_CHECK(W > 0, "empty window");
  return host_roundtrip<float>(g, h_pos, h_meta, n, 4 * int64_t(W), h_out, [&](int32_t* p, int32_t* m, int64_t ns, float* o) {
    return mural_encode_onehot(g, p, m, ns, radius, model_type, o, nullptr);
  });
}
