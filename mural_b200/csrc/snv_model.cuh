// MuRaL-snv Network2 (MuRaL/model/model_snv.py:290-525): parameter layout + device-resident prepared weights.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace mural {

struct TensorEntry {
  std::string name;  // reference state_dict key
  int64_t offset;    // element offset in the flat fp32 blob
  int64_t numel;
  int32_t is_buffer;  // 0 trainable parameter, 1 BatchNorm running statistic
};

// One Conv1d(C,C,ks) with its preceding BatchNorm (and optional ReLU before the BN), eval-folded:
//   y = bias + sum_tap sum_ci Wt[tap][ci][co] * pad0( a[ci]*act(x[.,ci]) + b[ci] )
struct ConvLayerDev {
  const float* Wt;    // [ks][C][C]  (tap, ci, co)
  const float* bias;  // [C]
  const float* a;     // [C] BN scale  gamma/sqrt(var+eps)
  const float* b;     // [C] BN shift  beta - mean*a
  int ks;
  int relu_in;  // ReLU applied to the input before the BN affine (ResBlock convs)
  int precise = 0;  // 1: fp32-level products required (training forward) -> fp32 FMA kernel; 0: two-level bf16 split MMA (~1e-5 per product)
};

struct BranchDev {
  const float* T;      // stem table [ks][16 symbols][C]: conv1 applied to BN(4)(one-hot column)
  const float* bias1;  // [C]
  const float* T4;     // ks==3 only: [256 4-mers][C] = max(conv1 at p, conv1 at p+1) incl. bias; index b(p-1) | b(p)<<2 | b(p+1)<<4 | b(p+2)<<6
  ConvLayerDev rb1[4];  // RBs1.0.conv1, RBs1.0.conv2, RBs1.1.conv1, RBs1.1.conv2
  ConvLayerDev conv2;
  ConvLayerDev rb2[4];
  ConvLayerDev conv3;
  const float* Wfc;  // [C][n_class]   BN(distal_fc.0) folded into Linear(distal_fc.2)
  const float* bfc;  // [n_class]
  int pool[3][3];    // (kernel, stride, pad) of the three max-pools
  int L0, L1, L2, L3;  // lengths: input, after pool1, after pool2, after pool3
};

struct LocalDev {
  const float* emb;  // [4^k+1][5]
  const float* W1t;  // [5*n_cat][h1]
  const float* b1;   // [h1]
  const float* W2t;  // [h1][h2]   bn_layers.0 folded in
  const float* b2;
  const float* W3t;  // [h2][n_class]  bn_layers.1 folded in (local_fc)
  const float* b3;
};

}  // namespace mural

struct mural_snv_model {
  mural_snv_config_t cfg;
  int device;
  int n_cat, emb_rows, k1;  // k1 = 5*n_cat
  int k1c = 0;              // k1 + n_cont: inputs of the first Linear (continuous features follow the embeddings, model_snv.py:460)
  const float* d_cont = nullptr;  // cont_x of the next forward (mural_snv_set_cont); consumed by it
  const float* d_cont_cur = nullptr;  // ... advanced to the chunk being processed
  int L;                    // 2R+1
  std::vector<mural::TensorEntry> layout;
  std::map<std::string, int> index;
  int64_t n_blob, n_trainable;
  // prepared eval weights
  float* d_prep = nullptr;
  int64_t prep_floats = 0;
  mural::LocalDev local;
  mural::BranchDev br[2];  // 0 = middle-scale (201 bp), 1 = large-scale (full window)
  bool loaded = false;
  // bf16 tcgen05 path (snv_tc.cu) and the tensor-core local branch (snv_mlp_tc.cu)
  void* tc = nullptr;
  void* mlp_tc = nullptr;
  void* tail = nullptr;  // warp-level tail kernel (snv_tail.cu)
  uint32_t* d_chain = nullptr;  // split (hi | lo) packed conv weights of both branches for the fused per-site chain (snv_site_chain.cu)
  bool chain_ready = false;     // cleared whenever the weights are (re)loaded
  // forward workspace (grown on demand)
  void* d_ws = nullptr;
  int64_t ws_bytes = 0;
  int64_t chunk_sites = 0;
  // persistent H2D/D2H staging of mural_snv_predict_host
  void* d_io = nullptr;
  int64_t io_bytes = 0;
  // MURAL_MODE_AUTO scratch: device site list, pinned count, event
  void* d_auto = nullptr;
  int64_t auto_bytes = 0;
  void* h_auto = nullptr;
  void* auto_ev = nullptr;
  int64_t last_auto_sites = 0;
  // ... its fp32-equivalent recompute runs on a side stream, concurrently with the bf16 pass, in its own workspace
  void* aux_stream = nullptr;
  void* aux_ev = nullptr;
  void* d_ws2 = nullptr;
  int64_t ws2_bytes = 0;
  // debug taps (parity tests): host copies of intermediate activations of the last chunk
  bool debug = false;
  bool slow_stem = false;  // parity switch: force the generic per-tap stem kernel
  std::map<std::string, std::vector<float>> tap_store;
};

namespace mural {
// bf16 tcgen05 path (snv_tc.cu): prepared bf16 weights live behind m->tc
int snv_tc_prepare(mural_snv_model* m, const float* h_blob);
void snv_tc_destroy(mural_snv_model* m);
int snv_forward_tc(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                   const uint8_t* d_sym, const int64_t* d_cat, int64_t n, float* d_logp, cudaStream_t st);
// tensor-core local branch (snv_mlp_tc.cu); launch returns -1 when unavailable for this model
int snv_mlp_tc_prepare(mural_snv_model* m);
void snv_mlp_tc_destroy(mural_snv_model* m);
int snv_local_launch_tc(mural_snv_model* m, const int32_t* cat32, const int64_t* cat64, int64_t ns, float* logits, int* err_flag,
                        cudaStream_t st);
// warp-level tail: pool 3 + conv3 + global max + heads + combine (snv_tail.cu); launch returns -1 when unavailable
int snv_tail_prepare(mural_snv_model* m, const float* h_blob);
void snv_tail_destroy(mural_snv_model* m);
int snv_tail_launch(mural_snv_model* m, const void* z2_mid, int64_t ra_mid, const void* z2_large, int64_t ra_large,
                    const float* local_logits, int64_t ns, float* logp, float* tg0, float* tg1, float* tl0, float* tl1, cudaStream_t st);
// shared launch helpers (snv_forward.cu)
int snv_stem_launch(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                    const uint8_t* d_sym, int64_t ns, float* mid_out, float* large_out, int32_t* cat_out, cudaStream_t st);
int snv_stem_launch_planes(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                           const uint8_t* d_sym, int64_t ns, float* mid_out, int64_t mid_rows_alloc, float* large_out,
                           int64_t large_rows_alloc, int32_t* cat_out, cudaStream_t st, bool out_bf16 = false,
                           const int* skip_flag = nullptr);
// dense-site stem + stage-1 lattice (snv_dense_stem.cu)
// Device-side description of a chunk (written by k_chunk_span, read by every kernel of the dense path: no host sync).
struct ChunkInfo {
  long long g_lo;  // chromosome coordinate of table index 0 (may be negative: overhang is imputed with N)
  int n_pos;       // table rows in use
  int dense;       // 1: tables + gather (+ lattice), 0: per-site kernels
  int chrom;
  int has[2];      // strands present
  // stage-1 lattice geometry per branch: 2*ps1 pseudo-sites (strand x phase) of M lattice steps each
  int M[2];
  int lat_rows[2];   // 2*ps1*(M+1)+1
  int lat_tiles[2];  // RB4 tiles covering lat_rows
};
constexpr int LAT_EO = 5;   // stage-1 output rows per window end that depend on the site (4 convs + the clipped first/last bin)
constexpr int LAT_EI = 9;   // stage-1 input rows per window end those outputs read
constexpr int LAT_EL = 2 * LAT_EI;  // length of a site's edge pseudo-site: [rows 0..8 | rows L1-9..L1-1]
struct LatticeBufs {        // per branch; all bf16 planes [4][ra][8]
  void* lat_in; void* lat_out; int64_t lat_ra;
  void* edge_in; void* edge_out; int64_t edge_ra;
};
int64_t snv_dense_cap(int64_t chunk);
bool snv_lattice_supported(const mural_snv_model* m);
size_t snv_dense_bytes(const mural_snv_model* m, int64_t chunk);
int snv_dense_stem_launch(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta, int64_t ns,
                          int64_t chunk, void* mid_out, int64_t mid_ra, void* large_out, int64_t large_ra, void* d_scratch,
                          const int** d_flag, cudaStream_t st, const LatticeBufs* lattice = nullptr,
                          const ChunkInfo** d_info = nullptr);
// fused per-site chain of one CNN branch at fp32-equivalent precision (snv_site_chain.cu); -1: shape not served
int snv_site_chain_launch(mural_snv_model* m, int br, const float* x0, float* h, int64_t ns, cudaStream_t st);
int snv_local_idx_launch(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta, int64_t n,
                         int32_t* cat32, cudaStream_t st);
int snv_local_launch(mural_snv_model* m, const int32_t* cat32, const int64_t* cat64, int64_t ns, float* logits, int* err_flag,
                     cudaStream_t st);
int snv_head_launch(mural_snv_model* m, const float* h_mid, const float* h_large, const float* local_logits, int64_t ns,
                    float* logp, float* tg0, float* tg1, float* tl0, float* tl1, cudaStream_t st);
int conv_any(int C, const float* in, float* out, const float* r1, const float* r2, int64_t n, int L, const ConvLayerDev& P,
             int relu_out, cudaStream_t st);
int conv32_mma(const float* in, float* out, const float* r1, const float* r2, int64_t n, int L, const ConvLayerDev& P, int relu_out,
               cudaStream_t st);  // snv_conv_mma.cu: C == 32, ks == 3, split-bf16 mma.sync
int wgrad32_mma(const float* x, const float* dy, int64_t rows, int L, int relu, const float* a, const float* b, float* G, int64_t w_off,
                int64_t b_off, cudaStream_t st);  // snv_conv_mma.cu: weight gradient of a C == 32, ks == 3 layer
int snv_ensure_workspace(mural_snv_model* m, int64_t bytes);
int onehot_to_symbols_checked(const float* d_onehot, int64_t n, int32_t W, uint8_t* d_sym, int* d_flag, cudaStream_t st);
// aux_ws: use the side workspace (d_ws2) so that the call may run concurrently with a forward that owns d_ws
int snv_forward_fp32(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                     const uint8_t* d_sym, const int64_t* d_cat, int64_t n, float* d_logp, cudaStream_t st, bool aux_ws = false);
}
