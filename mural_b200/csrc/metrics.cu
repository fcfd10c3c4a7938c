// Per-epoch validation metrics of the reference's Evaluator as device-side grouped reductions.
// Reference: MuRaL/evaluation/evaluation.py:48-67 (freq_kmer_comp_multi), :124-193 (corr_calc_sub), :196-203 (calc_avg_prob),
// :545-566 (evaluate_regional_score); called every epoch from MuRaL/training.py:488-520.
//
// Both reductions are HBM-bound integer work (one pass over n sites: codes, label, n_class probabilities).  The sums are kept
// as 64-bit integers — counts exactly, probabilities in fixed point (2^-36) — so the result does not depend on the order in
// which the atomics land: the tables are bit-reproducible from run to run and for any grid size.  The handful of Pearson
// correlations over the (tiny) tables are formed by the caller in fp64.
#include "common.cuh"

namespace mural {

constexpr int MET_MAXK = 16;      // n_class
constexpr int MET_THREADS = 256;
constexpr int MET_PER_THREAD = 8;  // consecutive sites per thread in the window-run kernels
constexpr int MET_SMEM_WORDS = 12288;  // 96 KB of 64-bit table entries per CTA

__device__ __forceinline__ int64_t cdiv_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }
__device__ __forceinline__ long long to_fixed(double p) { return __double2ll_rn(p * MURAL_METRIC_SCALE); }

// ---- k-mer context groups ---------------------------------------------------------------------------------------------
// table[region][group][0] = sites, [1 + c] = sites with label c, [1 + K + c] = fixed-point sum of prob c
// SMEM: the CTA accumulates in shared memory, in `copies` private replicas of the table (warp w uses replica w % copies) so that
// the few hot groups of a short k-mer do not serialise every warp on the same addresses; replicas are summed at the flush.
template <bool SMEM>
__global__ void __launch_bounds__(MET_THREADS) k_kmer_groups(const int64_t* __restrict__ flank, int n_cols, int d, const int32_t* __restrict__ meta,
                                                           const double* __restrict__ prob, int K, int64_t region_size, int G, int copies,
                                                           unsigned long long* __restrict__ table, int* __restrict__ err) {
  extern __shared__ unsigned long long s_tab[];
  const int W = 1 + 2 * K;
  const int GW = G * W;
  const int64_t r = blockIdx.y;
  unsigned long long* out = table + r * int64_t(GW);
  if (SMEM) {
    for (int e = threadIdx.x; e < GW * copies; e += blockDim.x) s_tab[e] = 0ull;
    __syncthreads();
  }
  unsigned long long* acc = SMEM ? s_tab + ((threadIdx.x >> 5) % copies) * GW : out;
  const int mid = n_cols / 2;
  const int64_t lo = r * region_size, per = cdiv_dev(region_size, gridDim.x);
  const int64_t b = lo + blockIdx.x * per, e_ = (b + per < lo + region_size) ? b + per : lo + region_size;
  bool bad = false;
  for (int64_t i = b + threadIdx.x; i < e_; i += blockDim.x) {
    const int64_t* f = flank + i * n_cols;
    int g = 0;
    for (int j = d; j >= 1; --j) {  // us_d .. us1, then ds1 .. ds_d: pandas groupby key order, most significant first
      const int64_t c = __ldg(f + mid - j);
      bad |= (c < 0 || c > 4);
      g = g * 5 + int(c);
    }
    for (int j = 1; j <= d; ++j) {
      const int64_t c = __ldg(f + mid + j);
      bad |= (c < 0 || c > 4);
      g = g * 5 + int(c);
    }
    if (bad) break;
    const int y = (__ldg(meta + i) >> 1) & 0x7f;
    unsigned long long* row = acc + int64_t(g) * W;
    atomicAdd(row, 1ull);
    if (y < K) atomicAdd(row + 1 + y, 1ull);
    for (int c = 0; c < K; ++c) atomicAdd(row + 1 + K + c, (unsigned long long)to_fixed(__ldg(prob + i * K + c)));
  }
  if (bad) atomicOr(err, 1);
  if (SMEM) {
    __syncthreads();
    for (int e = threadIdx.x; e < GW; e += blockDim.x) {
      unsigned long long v = 0ull;
      for (int c = 0; c < copies; ++c) v += s_tab[c * GW + e];
      if (v) atomicAdd(out + e, v);
    }
  }
}

// ---- runs of consecutive sites that share (chrom, start // window) ---------------------------------------------------------
__device__ __forceinline__ int64_t run_key(const int32_t* __restrict__ meta, const int32_t* __restrict__ start, const int64_t* __restrict__ order,
                                           int64_t i, int window) {
  const int64_t s = order ? __ldg(order + i) : i;
  return (int64_t(__ldg(meta + s) >> 8) << 32) | int64_t(__ldg(start + s) / window);
}

// run starts inside each CTA's slice of MET_THREADS * MET_PER_THREAD sites
__global__ void __launch_bounds__(MET_THREADS) k_run_count(const int32_t* __restrict__ meta, const int32_t* __restrict__ start,
                                                         const int64_t* __restrict__ order, int64_t n, int window,
                                                         unsigned long long* __restrict__ block_runs) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int64_t base = (int64_t(blockIdx.x) * MET_THREADS + threadIdx.x) * MET_PER_THREAD;
  int cnt = 0;
  if (base < n) {
    int64_t prev = base ? run_key(meta, start, order, base - 1, window) : -1;
    for (int u = 0; u < MET_PER_THREAD && base + u < n; ++u) {
      const int64_t k = run_key(meta, start, order, base + u, window);
      cnt += (k != prev);
      prev = k;
    }
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0) block_runs[blockIdx.x] = (unsigned long long)s_cnt;
}

// exclusive scan of the per-CTA run counts (one CTA; a few thousand entries for millions of sites); total -> block_runs[n_blocks]
__global__ void __launch_bounds__(1024) k_run_scan(unsigned long long* __restrict__ block_runs, int64_t n_blocks) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0ull;
  __syncthreads();
  for (int64_t b0 = 0; b0 < n_blocks; b0 += blockDim.x) {
    const int64_t i = b0 + threadIdx.x;
    const unsigned long long v = i < n_blocks ? block_runs[i] : 0ull;
    unsigned long long x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long w = s_warp[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      s_warp[threadIdx.x] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long before = s_carry + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0ull) + x - v;
    if (i < n_blocks) block_runs[i] = before;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) block_runs[n_blocks] = s_carry;
}

// rows[run][0] = sites, [1 + c] = sites with label c, [1 + K + c] = fixed-point sum of prob c
__global__ void __launch_bounds__(MET_THREADS) k_run_sums(const int32_t* __restrict__ meta, const int32_t* __restrict__ start,
                                                        const int64_t* __restrict__ order, const double* __restrict__ prob, int K, int64_t n,
                                                        int window, const unsigned long long* __restrict__ block_runs,
                                                        unsigned long long* __restrict__ rows) {
  __shared__ int s_warp[MET_THREADS / 32];
  const int W = 1 + 2 * K;
  const int64_t base = (int64_t(blockIdx.x) * MET_THREADS + threadIdx.x) * MET_PER_THREAD;
  int64_t keys[MET_PER_THREAD];
  int cnt = 0;
  int64_t prev = -1;
  if (base < n) {
    prev = base ? run_key(meta, start, order, base - 1, window) : -1;
    int64_t p = prev;
    for (int u = 0; u < MET_PER_THREAD && base + u < n; ++u) {
      keys[u] = run_key(meta, start, order, base + u, window);
      cnt += (keys[u] != p);
      p = keys[u];
    }
  }
  // exclusive scan of cnt over the CTA
  int x = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  int before = x - cnt;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) before += s_warp[w];
  if (base >= n) return;
  // run index of the site before this thread's first one is (starts so far) - 1; thread-local accumulation, one flush per run
  int64_t run = int64_t(block_runs[blockIdx.x]) + before - 1;
  unsigned long long acc[1 + 2 * MET_MAXK];
  for (int c = 0; c < W; ++c) acc[c] = 0ull;
  auto flush = [&]() {
    if (acc[0]) {
      unsigned long long* row = rows + run * W;
      for (int c = 0; c < W; ++c)
        if (acc[c]) atomicAdd(row + c, acc[c]);
    }
    for (int c = 0; c < W; ++c) acc[c] = 0ull;
  };
  for (int u = 0; u < MET_PER_THREAD && base + u < n; ++u) {
    if (keys[u] != prev) {
      flush();
      ++run;
      prev = keys[u];
    }
    const int64_t s = order ? __ldg(order + base + u) : base + u;
    const int y = (__ldg(meta + s) >> 1) & 0x7f;
    acc[0] += 1;
    for (int c = 0; c < K; ++c) {
      acc[1 + c] += (y == c);
      acc[1 + K + c] += (unsigned long long)to_fixed(__ldg(prob + s * K + c));
    }
  }
  flush();
}

}  // namespace mural
using namespace mural;

extern "C" int mural_kmer_group_stats(const int64_t* d_flank, int64_t n, int32_t n_cols, int32_t k, const int32_t* d_meta,
                                      const double* d_prob, int32_t n_class, int64_t region_size, int64_t* d_table, void* stream) {
  MURAL_CHECK(d_flank && d_meta && d_prob && d_table, "NULL argument");
  MURAL_CHECK(n_class >= 1 && n_class <= MET_MAXK, "n_class out of range");
  const int d = k / 2;
  MURAL_CHECK(d >= 1 && d <= 4 && 2 * d + 1 <= n_cols, "ValueError: k-mer length does not fit the local columns");
  MURAL_CHECK(n >= 0 && region_size >= 0, "negative size");
  MURAL_CHECK(n <= (int64_t(1) << 27), "ValueError: more than 2^27 sites per call (fixed-point sums)");
  cudaStream_t st = (cudaStream_t)stream;
  if (region_size == 0) region_size = n;
  const int64_t n_regions = region_size ? n / region_size : 0;
  if (n_regions == 0) return 0;
  MURAL_CHECK(n_regions <= 65535, "ValueError: more than 65535 regions");
  int G = 1;
  for (int j = 0; j < 2 * d; ++j) G *= 5;
  const int W = 1 + 2 * n_class;
  CUDA_TRY(cudaMemsetAsync(d_table, 0, size_t(n_regions) * G * W * 8, st));
  int* d_err = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&d_err, 4, st));
  CUDA_TRY(cudaMemsetAsync(d_err, 0, 4, st));
  // enough CTAs per region to fill the GPU, each with at least a few thousand sites
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  // the pass is latency-bound (one dependent load chain per site): many small CTAs keep enough loads in flight
  int64_t per_region = cdiv(int64_t(8) * sms, n_regions);
  const int64_t cap = cdiv(region_size, 1024);
  if (per_region > cap) per_region = cap;
  // every CTA zeroes and flushes a whole table: keep that below the work of accumulating its sites
  const int64_t cap_flush = region_size * 4 / (int64_t(G) * (1 + 2 * n_class));
  if (G * (1 + 2 * n_class) <= MET_SMEM_WORDS && per_region > cap_flush) per_region = cap_flush;
  if (per_region < 1) per_region = 1;
  const dim3 grid((unsigned)per_region, (unsigned)n_regions);
  auto* tab = reinterpret_cast<unsigned long long*>(d_table);
  if (G * W <= MET_SMEM_WORDS) {
    int copies = (MET_SMEM_WORDS / 2) / (G * W);  // replicas within 48 KB so that four CTAs fit an SM
    if (copies > MET_THREADS / 32) copies = MET_THREADS / 32;
    if (copies < 1) copies = 1;
    const size_t smem = size_t(G) * W * 8 * copies;
    static bool configured = false;
    if (!configured) {
      CUDA_TRY(cudaFuncSetAttribute(k_kmer_groups<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MET_SMEM_WORDS * 8));
      configured = true;
    }
    LAUNCH(k_kmer_groups<true>, grid, MET_THREADS, smem, st, d_flank, n_cols, d, d_meta, d_prob, n_class, region_size, G, copies, tab,
           d_err);
  } else {
    LAUNCH(k_kmer_groups<false>, grid, MET_THREADS, 0, st, d_flank, n_cols, d, d_meta, d_prob, n_class, region_size, G, 1, tab, d_err);
  }
  CUDA_TRY(cudaGetLastError());
  int h_err = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_err, d_err, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaFreeAsync(d_err, st));
  MURAL_CHECK(h_err == 0, "ValueError: local base code outside 0..4");
  return 0;
}

extern "C" int mural_window_runs(const int32_t* d_meta, const int32_t* d_start, const int64_t* d_order, const double* d_prob, int64_t n,
                                 int32_t n_class, int32_t window, int64_t* h_n_runs, int64_t* d_rows, int64_t max_runs, void* stream) {
  MURAL_CHECK(d_meta && d_start && h_n_runs, "NULL argument");
  MURAL_CHECK(n_class >= 1 && n_class <= MET_MAXK, "n_class out of range");
  MURAL_CHECK(window >= 1, "ValueError: window must be positive");
  MURAL_CHECK(n >= 0 && n <= (int64_t(1) << 27), "ValueError: more than 2^27 sites per call (fixed-point sums)");
  *h_n_runs = 0;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_blocks = cdiv(n, int64_t(MET_THREADS) * MET_PER_THREAD);
  unsigned long long* d_blk = nullptr;
  CUDA_TRY(cudaMallocAsync((void**)&d_blk, size_t(n_blocks + 1) * 8, st));
  LAUNCH(k_run_count, (unsigned)n_blocks, MET_THREADS, 0, st, d_meta, d_start, d_order, n, window, d_blk);
  LAUNCH(k_run_scan, 1, 1024, 0, st, d_blk, n_blocks);
  unsigned long long total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, d_blk + n_blocks, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  *h_n_runs = int64_t(total);
  int rc = 0;
  if (d_rows) {
    if (!d_prob) rc = fail(__FILE__, __LINE__, "NULL argument");
    else if (int64_t(total) > max_runs) rc = fail(__FILE__, __LINE__, "ValueError: more window runs than the output holds");
    else {
      const int W = 1 + 2 * n_class;
      cudaMemsetAsync(d_rows, 0, size_t(total) * W * 8, st);
      LAUNCH(k_run_sums, (unsigned)n_blocks, MET_THREADS, 0, st, d_meta, d_start, d_order, d_prob, n_class, n, window, d_blk,
             reinterpret_cast<unsigned long long*>(d_rows));
      if (cudaGetLastError() != cudaSuccess) rc = fail(__FILE__, __LINE__, "k_run_sums launch failed");
    }
  }
  cudaFreeAsync(d_blk, st);
  return rc;
}
