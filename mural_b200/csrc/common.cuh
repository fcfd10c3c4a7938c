// Shared declarations of libmural_b200.so (internal; the public surface is include/mural_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/mural_b200.h"

namespace mural {

// ---- error plumbing: C ABI returns codes, message kept per thread -------------------------------
void set_error(const std::string& msg);
int fail(const char* file, int line, const std::string& msg);
extern int64_t g_launches;
extern bool g_prof;
void prof_pre(const char* name, cudaStream_t st);
void prof_post(cudaStream_t st);

#define MURAL_FAIL(msg) return ::mural::fail(__FILE__, __LINE__, (msg))
#define MURAL_CHECK(cond, msg) \
  do {                         \
    if (!(cond)) MURAL_FAIL(msg); \
  } while (0)
#define CUDA_TRY(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return ::mural::fail(__FILE__, __LINE__, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
// every kernel launch of this library goes through LAUNCH so bench.py can report gpu_launches
// and the optional per-kernel CUDA-event profile (mural_profile_begin/end) sees it
#define LAUNCH(kernel, grid, block, smem, stream, ...)          \
  do {                                                          \
    if (::mural::g_prof) ::mural::prof_pre(#kernel, (stream));  \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
    if (::mural::g_prof) ::mural::prof_post((stream));          \
    ++::mural::g_launches;                                      \
  } while (0)

// same, with an explicit profile name (template kernels launched in several roles)
#define LAUNCH_N(name, kernel, grid, block, smem, stream, ...)  \
  do {                                                          \
    if (::mural::g_prof) ::mural::prof_pre((name), (stream));   \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
    if (::mural::g_prof) ::mural::prof_post((stream));          \
    ++::mural::g_launches;                                      \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- symbols -------------------------------------------------------------------------------------
// 0..3 A C G T, 4..14 R Y M S W K B D H V N (reference tables, MuRaL/data/preprocessing.py:655-666, 758-772)
constexpr int SYM_N = 14;
constexpr int N_SYM = 15;
constexpr int SYM_PAD = 15;  // "no base": a conv tap that falls outside the window (zero padding)

// ---- packed genome (device view) -----------------------------------------------------------------
struct GenomeView {
  const uint32_t* bits2;     // 16 bases / word
  const uint32_t* mask;      // 32 bases / word, 1 = not ACGT
  const int64_t* chrom_off;  // global base offset of each chromosome (multiple of 64)
  const int64_t* chrom_len;
  const int64_t* exc_start;  // sorted run table over global coordinates
  const int64_t* exc_end;
  const uint8_t* exc_sym;
  int32_t n_chrom;
  int32_t n_exc;
};

}  // namespace mural

struct mural_genome {
  mural::GenomeView view;  // device pointers
  int device;
  std::vector<int64_t> h_off, h_len;
  int64_t total_bases;  // padded
  int64_t device_bytes;
  void* d_block;  // single allocation backing every array of the view
};

#ifdef __CUDACC__
namespace mural {

// complement symbol (A<->T, C<->G, R<->Y, M<->K, B<->V, D<->H; S W N PAD fixed) as a nibble LUT.
// Channel-flipping a reference one-hot column (one_hot_encoder_rc, preprocessing.py:774-788) maps
// the IUPAC vectors onto each other exactly this way.
__host__ __device__ __forceinline__ int comp_sym(int s) { return int((0xFEABCD6879450123ull >> (4 * s)) & 15ull); }

// symbol of global base g (already known to be inside its chromosome)
__device__ __forceinline__ int genome_symbol(const GenomeView& G, int64_t g) {
  uint32_t w = __ldg(G.bits2 + (g >> 4));
  int code = (w >> ((int(g) & 15) * 2)) & 3;
  uint32_t m = __ldg(G.mask + (g >> 5));
  if ((m >> (int(g) & 31)) & 1u) {
    // rare: binary search the run table for the run containing g
    int lo = 0, hi = G.n_exc - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(G.exc_start + mid) <= g) lo = mid; else hi = mid - 1;
    }
    code = G.exc_sym[lo];
  }
  return code;
}

// Cooperative (whole CTA) load of one site's oriented symbol window into shared memory.
//   sym[i], i in [0,W): symbol at oriented window position i ('-' strand: reversed + complemented);
//   positions outside the chromosome read as N (imputation, preprocessing.py:681-695).
// One thread decodes one 16-base word.  Caller must __syncthreads() afterwards.
__device__ __forceinline__ void load_window(const GenomeView& G, int chrom, int64_t wstart, int W, int strand,
                                            uint8_t* sym) {
  const int64_t len = G.chrom_len[chrom];
  const int64_t off = G.chrom_off[chrom];
  // first word-aligned base at or below wstart (chrom_off is a multiple of 64, so alignment of
  // chromosome coordinates equals alignment of global coordinates)
  const int64_t a0 = wstart & ~int64_t(15);  // floor to a multiple of 16 (two's complement, also for wstart < 0)
  const int nwords = int((wstart + W - a0 + 15) >> 4);
  for (int w = threadIdx.x; w < nwords; w += blockDim.x) {
    const int64_t q0 = a0 + int64_t(w) * 16;  // chromosome coordinate of base 0 of this word
    uint32_t bits = 0, msk = 0;
    const bool any_inside = (q0 + 16 > 0) && (q0 < len);
    if (any_inside) {
      const int64_t g0 = off + q0;  // q0 may be negative only when any_inside is false (q0 multiple of 16)
      bits = __ldg(G.bits2 + (g0 >> 4));
      msk = (__ldg(G.mask + (g0 >> 5)) >> (int(g0) & 31)) & 0xffffu;
    }
#pragma unroll
    for (int b = 0; b < 16; ++b) {
      const int64_t q = q0 + b;
      const int i = int(q - wstart);
      if (i < 0 || i >= W) continue;
      int s;
      if (q < 0 || q >= len) s = SYM_N;
      else if ((msk >> b) & 1u) s = genome_symbol(G, off + q);
      else s = (bits >> (2 * b)) & 3;
      if (strand) sym[W - 1 - i] = uint8_t(comp_sym(s));
      else sym[i] = uint8_t(s);
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace mural
#endif
