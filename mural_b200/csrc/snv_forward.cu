// MuRaL-snv Network2 eval forward, fp32 CUDA-core path ("fp32-equivalent" mode) + shared stem/head kernels.
// Reference forward: MuRaL/model/model_snv.py:439-525; loop body of model_predict_m: MuRaL/model/nn_utils.py:48-65.
//
// Pipeline per chunk of sites (activations channels-last fp32 [site][pos][C] in a reusable workspace):
//   k_stem      gather 2-bit window -> smem symbols -> (BN(4)+Conv1d(4->C)) as table lookups -> max-pool 1,
//               for both CNN branches; also emits the local k-mer indices             (E4,E5,M3,M4 of SURVEY §8a)
//   k_local_mlp embedding gather + 3 Linear layers with BN folded forward               (M1,M2)
//   k_conv      Conv1d(C,C,ks) with ReLU/BN prologue, zero padding per site, bias, up to two residuals (M5,M6)
//   k_pool      MaxPool1d (-inf padding)                                                (M4)
//   k_head      global max, BN+Linear heads, 3 softmaxes, average, clamp, log           (M7,M8)
#include <cuda_bf16.h>
#include <float.h>
#include <string.h>

#include "snv_model.cuh"

namespace mural {

// ------------------------------------------------------------------------------------------------ stem
struct StemBranch {
  const float* T;     // [ks][16][C]
  const float* bias;  // [C]
  const float* T4;    // [256][C] pair table of the fast stem (ks == 3), may be NULL
  float* out;         // [n][L1][C], or fp32 planes [C/4][rows_alloc][4] with row(s,p) = 1 + s*(L1+1) + p
  int64_t rows_alloc;  // 0: dense site-major layout
  int out_bf16;        // planes of 8 bf16 channels (16 B per row) instead of 4 fp32 channels
  int L0, off0, L1, pk, ps, pp;
};

template <int C>
__global__ void __launch_bounds__(128) k_stem(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta,
                                              const uint8_t* __restrict__ sym_in, int R, int L, int ks, StemBranch b0,
                                              StemBranch b1, int local_R, int order, int n_cat, int32_t* __restrict__ cat_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sT = reinterpret_cast<float*>(smem_raw);  // [2][ks][16][C]
  float* sB = sT + 2 * ks * 16 * C;                // [2][C]
  uint8_t* sym = reinterpret_cast<uint8_t*>(sB + 2 * C);
  const int64_t site = blockIdx.x;
  const int tid = threadIdx.x;
  for (int e = tid; e < ks * 16 * C; e += blockDim.x) {
    sT[e] = b0.T[e];
    sT[ks * 16 * C + e] = b1.T[e];
  }
  for (int e = tid; e < C; e += blockDim.x) {
    sB[e] = b0.bias[e];
    sB[C + e] = b1.bias[e];
  }
  if (sym_in) {
    for (int i = tid; i < L; i += blockDim.x) sym[i] = sym_in[site * L + i];
  } else {
    const int m = meta[site];
    load_window(G, int(uint32_t(m) >> 8), int64_t(pos[site]) - R, L, m & 1, sym);
  }
  __syncthreads();
  const int half = ks / 2;
#pragma unroll 1
  for (int br = 0; br < 2; ++br) {
    const StemBranch& B = br ? b1 : b0;
    const float* T = sT + br * ks * 16 * C;
    const uint8_t* s0 = sym + B.off0;
    float* out = B.out + site * int64_t(B.L1) * C;
    for (int e = tid; e < B.L1 * C; e += blockDim.x) {
      const int j = e / C, c = e - j * C;
      int lo = j * B.ps - B.pp, hi = lo + B.pk;
      lo = lo < 0 ? 0 : lo;
      hi = hi > B.L0 ? B.L0 : hi;
      float mx = -FLT_MAX;
      for (int p = lo; p < hi; ++p) {
        float v = sB[br * C + c];
        for (int t = 0; t < ks; ++t) {
          const int q = p + t - half;
          const int s = (q >= 0 && q < B.L0) ? s0[q] : SYM_PAD;
          v += T[(t * 16 + s) * C + c];
        }
        mx = fmaxf(mx, v);
      }
      if (B.rows_alloc && B.out_bf16)
        reinterpret_cast<__nv_bfloat16*>(B.out)[(int64_t(c >> 3) * B.rows_alloc + 1 + site * int64_t(B.L1 + 1) + j) * 8 + (c & 7)] =
            __float2bfloat16(mx);
      else if (B.rows_alloc) B.out[(int64_t(c >> 2) * B.rows_alloc + 1 + site * int64_t(B.L1 + 1) + j) * 4 + (c & 3)] = mx;
      else out[e] = mx;
    }
  }
  // local k-mer indices from the centre of the already oriented window (seq_digit_encoder semantics)
  if (cat_out) {
    for (int j = tid; j < n_cat; j += blockDim.x) {
      int idx = 0;
      bool bad = false;
      for (int d = 0; d < order; ++d) {
        const int s = sym[R - local_R + j + d];
        bad |= s > 3;
        idx = idx * 4 + (s & 3);
      }
      cat_out[site * n_cat + j] = bad ? (1 << (2 * order)) : idx;
    }
  }
}

// Fast stem (ks == 3): persistent CTAs keep both branches' 4-mer pair tables in shared memory.  For an
// all-ACGT window a pooled bin of 15 positions is the max of 8 table rows (2 rows for the 3-wide pool of the
// middle branch); bins touching the window edge (missing tap = zero padding) and windows holding N/IUPAC
// symbols or chromosome overhang go through the exact per-tap tables.  Lane mapping: C/4 lanes share a bin and
// read one contiguous table row (conflict-free LDS.128), and write one 16-byte chunk each.
template <int C>
__device__ __forceinline__ float4 stem_bin_generic(const float* __restrict__ T, const float* __restrict__ bias,
                                                   const uint8_t* __restrict__ s0, int L0, int lo, int hi, int q) {
  float4 mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
  const float4 b = *reinterpret_cast<const float4*>(bias + 4 * q);
  for (int p = lo; p < hi; ++p) {
    float4 v = b;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int x = p + t - 1;
      const int sy = (x >= 0 && x < L0) ? s0[x] : SYM_PAD;
      const float4 w = *reinterpret_cast<const float4*>(T + (t * 16 + sy) * C + 4 * q);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
  }
  return mx;
}

template <int C>
__global__ void __launch_bounds__(256, 2) k_stem_fast(GenomeView G, const int32_t* __restrict__ pos,
                                                      const int32_t* __restrict__ meta, const uint8_t* __restrict__ sym_in,
                                                      int64_t ns, int R, int L, StemBranch b0, StemBranch b1, int local_R,
                                                      int order, int n_cat, int32_t* __restrict__ cat_out) {
  constexpr int CG = C / 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sT4 = reinterpret_cast<float*>(smem_raw);  // [2][256][C]
  float* sT = sT4 + 2 * 256 * C;                    // [2][3][16][C]
  float* sB = sT + 2 * 3 * 16 * C;                  // [2][C]
  uint8_t* sym = reinterpret_cast<uint8_t*>(sB + 2 * C);
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  for (int e = tid * 4; e < 256 * C; e += 256 * 4) {
    *reinterpret_cast<float4*>(sT4 + e) = *reinterpret_cast<const float4*>(b0.T4 + e);
    *reinterpret_cast<float4*>(sT4 + 256 * C + e) = *reinterpret_cast<const float4*>(b1.T4 + e);
  }
  for (int e = tid; e < 3 * 16 * C; e += 256) {
    sT[e] = b0.T[e];
    sT[3 * 16 * C + e] = b1.T[e];
  }
  for (int e = tid; e < C; e += 256) {
    sB[e] = b0.bias[e];
    sB[C + e] = b1.bias[e];
  }
  const int Lw = (L + 3) >> 2;
  for (int64_t site = blockIdx.x; site < ns; site += gridDim.x) {
    __syncthreads();  // previous site's readers are done with sym / s_bad; tables visible on the first pass
    if (tid == 0) s_bad = 0;
    if (sym_in) {
      for (int i = tid; i < L; i += 256) sym[i] = sym_in[site * L + i];
    } else {
      const int m = meta[site];
      load_window(G, int(uint32_t(m) >> 8), int64_t(pos[site]) - R, L, m & 1, sym);
    }
    if (tid < 4) sym[L + tid] = 0;  // tail of the last word
    __syncthreads();
    {
      bool bad = false;
      for (int i = tid; i < Lw; i += 256) bad |= (reinterpret_cast<const uint32_t*>(sym)[i] & 0xFCFCFCFCu) != 0u;
      if (bad) s_bad = 1;
    }
    __syncthreads();
    const bool slow = s_bad != 0;
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
      const StemBranch& B = br ? b1 : b0;
      const float* T4 = sT4 + br * 256 * C;
      const uint8_t* s0 = sym + B.off0;
      for (int item = tid; item < B.L1 * CG; item += 256) {
        const int j = item / CG, q = item - j * CG;
        int lo = j * B.ps - B.pp, hi = lo + B.pk;
        lo = lo < 0 ? 0 : lo;
        hi = hi > B.L0 ? B.L0 : hi;
        float4 mx;
        if (slow || lo == 0 || hi == B.L0 || hi - lo < 2) {
          mx = stem_bin_generic<C>(sT + br * 3 * 16 * C, sB + br * C, s0, B.L0, lo, hi, q);
        } else {
          mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
          for (int p = lo; p < hi; p += 2) {
            const int pp = (p + 1 < hi) ? p : hi - 2;  // odd tail: overlap the last pair (max is idempotent)
            const int k4 = s0[pp - 1] | (s0[pp] << 2) | (s0[pp + 1] << 4) | (s0[pp + 2] << 6);
            const float4 w = *reinterpret_cast<const float4*>(T4 + k4 * C + 4 * q);
            mx.x = fmaxf(mx.x, w.x); mx.y = fmaxf(mx.y, w.y); mx.z = fmaxf(mx.z, w.z); mx.w = fmaxf(mx.w, w.w);
          }
        }
        if (B.rows_alloc && B.out_bf16) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(mx.x, mx.y), p1 = __floats2bfloat162_rn(mx.z, mx.w);
          uint2 pk2 = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
          *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(B.out) +
                                    (int64_t(q >> 1) * B.rows_alloc + 1 + site * int64_t(B.L1 + 1) + j) * 16 + (q & 1) * 8) = pk2;
        } else if (B.rows_alloc) *(reinterpret_cast<float4*>(B.out) + int64_t(q) * B.rows_alloc + 1 + site * int64_t(B.L1 + 1) + j) = mx;
        else *reinterpret_cast<float4*>(B.out + (site * B.L1 + j) * int64_t(C) + 4 * q) = mx;
      }
    }
    if (cat_out) {
      for (int j = tid; j < n_cat; j += 256) {
        int idx = 0;
        bool bad = false;
        for (int d = 0; d < order; ++d) {
          const int sy = sym[R - local_R + j + d];
          bad |= sy > 3;
          idx = idx * 4 + (sy & 3);
        }
        cat_out[site * n_cat + j] = bad ? (1 << (2 * order)) : idx;
      }
    }
  }
}

// Packed fast stem (genome path): the oriented window is kept as 2-bit codes (16 bases per word) in shared
// memory — '+' strand by a funnel shift of the genome words, '-' strand by reversing the 2-bit groups of the mirrored
// words and complementing (~) — and the 4-mer index of a position pair is 8 consecutive bits of that stream, so the 8
// lookups of a 15-wide pool bin need three word loads and eight funnel shifts.  The next site's words are fetched
// into registers while the current site's bins are computed (one __syncthreads per site).  Windows touching a
// chromosome end or any non-ACGT base fall back to the exact byte path.
template <int C, int NP>  // NP: lookups of an interior bin (8 for the 15-wide pool, 2 for the 3-wide pool)
__device__ __forceinline__ float4 stem_bin_packed(const float* __restrict__ T4, const uint32_t* __restrict__ pk, int i0, int q) {
  const int w = i0 >> 4, sh = 2 * (i0 & 15);
  const uint32_t w0 = pk[w], w1 = pk[w + 1], w2 = pk[w + 2];
  const uint32_t xl = __funnelshift_r(w0, w1, sh), xh = __funnelshift_r(w1, w2, sh);
  float4 mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int rel = (NP == 8) ? (i < 7 ? 4 * i : 26) : 2 * i;  // pairs at +0,+2,..,+12 and the overlapping tail pair at +13
    const int k4 = __funnelshift_r(xl, xh, rel) & 0xFF;
    const float4 wv = *reinterpret_cast<const float4*>(T4 + k4 * C + 4 * q);
    mx.x = fmaxf(mx.x, wv.x); mx.y = fmaxf(mx.y, wv.y); mx.z = fmaxf(mx.z, wv.z); mx.w = fmaxf(mx.w, wv.w);
  }
  return mx;
}

template <int C>
__global__ void __launch_bounds__(256, 2) k_stem_pk(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta,
                                                    int64_t ns, int R, int L, StemBranch b0, StemBranch b1, int local_R, int order,
                                                    int n_cat, int32_t* __restrict__ cat_out, const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag) return;  // the dense-site stem (snv_dense_stem.cu) already produced this chunk
  constexpr int CG = C / 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sT4 = reinterpret_cast<float*>(smem_raw);  // [2][256][C]
  float* sT = sT4 + 2 * 256 * C;                    // [2][3][16][C]
  float* sB = sT + 2 * 3 * 16 * C;                  // [2][C]
  const int nW = (L + 15) / 16 + 3;
  uint32_t* pkbuf = reinterpret_cast<uint32_t*>(sB + 2 * C);  // [2][nW]
  uint8_t* sym = reinterpret_cast<uint8_t*>(pkbuf + 2 * nW);  // [L+4], slow path only
  const int tid = threadIdx.x;
  for (int e = tid * 4; e < 256 * C; e += 256 * 4) {
    *reinterpret_cast<float4*>(sT4 + e) = *reinterpret_cast<const float4*>(b0.T4 + e);
    *reinterpret_cast<float4*>(sT4 + 256 * C + e) = *reinterpret_cast<const float4*>(b1.T4 + e);
  }
  for (int e = tid; e < 3 * 16 * C; e += 256) { sT[e] = b0.T[e]; sT[3 * 16 * C + e] = b1.T[e]; }
  for (int e = tid; e < C; e += 256) { sB[e] = b0.bias[e]; sB[C + e] = b1.bias[e]; }

  // Window fetch in two halves so the global-load latency overlaps the bin computation of the current site:
  // issue_fetch() starts the loads into registers, commit_fetch() shifts/reverses them into pk afterwards.
  // (windows longer than 4080 bp need more than one word per thread and are fetched synchronously in commit_fetch)
  const bool one_word = nW <= 256;
  struct Pre { uint32_t wa, wb, mk; int sh, strand; bool ok, bad; int64_t g0; } pre;
  auto issue_fetch = [&](int64_t site) {
    const int m = meta[site];
    const int chrom = int(uint32_t(m) >> 8);
    pre.strand = m & 1;
    const int64_t wstart = int64_t(pos[site]) - R;
    const int64_t len = G.chrom_len[chrom];
    pre.g0 = G.chrom_off[chrom] + wstart;
    pre.ok = wstart >= 16 && wstart + L + 16 <= len;  // otherwise: chromosome overhang -> N imputation -> slow path
    pre.wa = pre.wb = pre.mk = 0;
    pre.sh = 0;
    if (!pre.ok) return;
    const int64_t m0 = pre.g0 >> 5, m1 = (pre.g0 + L - 1) >> 5;
    if (one_word) {
      if (m0 + tid <= m1) {
        uint32_t mk = __ldg(G.mask + m0 + tid);
        if (tid == 0) mk &= 0xFFFFFFFFu << int(pre.g0 & 31);
        if (m0 + tid == m1) mk &= 0xFFFFFFFFu >> (31 - int((pre.g0 + L - 1) & 31));
        pre.mk = mk;
      }
      if (tid < nW) {
        const int64_t base = pre.strand ? (pre.g0 + L - 1 - 16 * int64_t(tid) - 15) : (pre.g0 + 16 * int64_t(tid));
        if (base >= 0) {
          pre.wa = __ldg(G.bits2 + (base >> 4));
          pre.wb = __ldg(G.bits2 + (base >> 4) + 1);
          pre.sh = 2 * int(base & 15);
        }
      }
    }
  };
  auto commit_fetch = [&](uint32_t* pk) -> bool {
    if (!pre.ok) return true;
    bool bad = false;
    if (one_word) {
      bad = pre.mk != 0u;
      if (tid < nW) {
        uint32_t v = __funnelshift_r(pre.wa, pre.wb, pre.sh);
        if (pre.strand) {
          uint32_t r = __brev(v);
          r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
          v = ~r;
        }
        pk[tid] = v;
      }
    } else {
      const int64_t g0 = pre.g0, m0 = g0 >> 5, m1 = (g0 + L - 1) >> 5;
      for (int64_t w = m0 + tid; w <= m1; w += 256) {
        uint32_t mk = __ldg(G.mask + w);
        if (w == m0) mk &= 0xFFFFFFFFu << int(g0 & 31);
        if (w == m1) mk &= 0xFFFFFFFFu >> (31 - int((g0 + L - 1) & 31));
        bad |= mk != 0u;
      }
      for (int t = tid; t < nW; t += 256) {
        const int64_t base = pre.strand ? (g0 + L - 1 - 16 * int64_t(t) - 15) : (g0 + 16 * int64_t(t));
        uint32_t v = 0;
        if (base >= 0) {
          v = __funnelshift_r(__ldg(G.bits2 + (base >> 4)), __ldg(G.bits2 + (base >> 4) + 1), 2 * int(base & 15));
          if (pre.strand) {
            uint32_t r = __brev(v);
            r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
            v = ~r;
          }
        }
        pk[t] = v;
      }
    }
    return bad;
  };

  int cur = 0;
  int64_t site = blockIdx.x;
  bool slow = false;
  if (site < ns) {
    issue_fetch(site);
    slow = commit_fetch(pkbuf);
  }
  slow = __syncthreads_or(slow) != 0;  // also publishes the tables and pk[0]
  for (; site < ns; site += gridDim.x) {
    const uint32_t* pk = pkbuf + cur * nW;
    const int64_t nxt = site + gridDim.x;
    if (nxt < ns) issue_fetch(nxt);    // loads in flight while this site's bins are computed
    if (slow) {  // block-uniform
      const int m = meta[site];
      load_window(G, int(uint32_t(m) >> 8), int64_t(pos[site]) - R, L, m & 1, sym);
      __syncthreads();
    }
    auto sym_at = [&](int i) -> int { return slow ? int(sym[i]) : int((pk[i >> 4] >> (2 * (i & 15))) & 3u); };
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
      const StemBranch& B = br ? b1 : b0;
      const float* T4 = sT4 + br * 256 * C;
      const float* Tt = sT + br * 3 * 16 * C;
      const float* bias = sB + br * C;
      const int n_items = B.L1 * CG;
      const int row0 = 1 + int(site) * (B.L1 + 1);  // fits int: chunk * (L1+1) < 2^31
      for (int item = tid; item < n_items; item += 256) {
        const int j = item / CG, q = item - j * CG;
        int lo = j * B.ps - B.pp, hi = lo + B.pk;
        lo = lo < 0 ? 0 : lo;
        hi = hi > B.L0 ? B.L0 : hi;
        float4 mx;
        if (!slow && lo > 0 && hi < B.L0 && hi - lo == 15) {
          mx = stem_bin_packed<C, 8>(T4, pk, B.off0 + lo - 1, q);
        } else if (!slow && lo > 0 && hi < B.L0 && hi - lo == 3) {
          mx = stem_bin_packed<C, 2>(T4, pk, B.off0 + lo - 1, q);
        } else {  // window edges (missing tap = zero padding), odd pool shapes, N / IUPAC / overhang windows
          mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
          const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * q);
          int glo = lo, ghi = hi;  // positions still to be done with the exact per-tap tables
          if (!slow) {
            const int ilo = lo < 1 ? 1 : lo, ihi = hi > B.L0 - 1 ? B.L0 - 1 : hi;  // interior positions [ilo, ihi)
            if (ihi - ilo >= 2) {
              for (int p = ilo; p < ihi; p += 2) {
                const int pp = (p + 1 < ihi) ? p : ihi - 2;
                const int i0 = B.off0 + pp - 1;
                const uint32_t w0 = pk[i0 >> 4], w1 = pk[(i0 >> 4) + 1];
                const int k4 = __funnelshift_r(w0, w1, 2 * (i0 & 15)) & 0xFF;
                const float4 wv = *reinterpret_cast<const float4*>(T4 + k4 * C + 4 * q);
                mx.x = fmaxf(mx.x, wv.x); mx.y = fmaxf(mx.y, wv.y); mx.z = fmaxf(mx.z, wv.z); mx.w = fmaxf(mx.w, wv.w);
              }
              // what is left: position 0 and/or position L0-1 (if inside the bin)
              if (lo == 0) { glo = 0; ghi = 1; } else { glo = ghi = 0; }
              if (hi == B.L0) {
                if (glo == ghi) { glo = B.L0 - 1; ghi = B.L0; }
                else if (B.L0 - 1 > 0) {  // both edges in one bin (tiny windows): do the right edge here
                  float4 v = bb;
#pragma unroll
                  for (int t = 0; t < 3; ++t) {
                    const int x = B.L0 - 1 + t - 1;
                    const int sy = (x >= 0 && x < B.L0) ? sym_at(B.off0 + x) : SYM_PAD;
                    const float4 w = *reinterpret_cast<const float4*>(Tt + (t * 16 + sy) * C + 4 * q);
                    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                  }
                  mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
                }
              }
            }
          }
          for (int p = glo; p < ghi; ++p) {
            float4 v = bb;
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int x = p + t - 1;
              const int sy = (x >= 0 && x < B.L0) ? sym_at(B.off0 + x) : SYM_PAD;
              const float4 w = *reinterpret_cast<const float4*>(Tt + (t * 16 + sy) * C + 4 * q);
              v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
            }
            mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
          }
        }
        if (B.rows_alloc && B.out_bf16) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(mx.x, mx.y), p1 = __floats2bfloat162_rn(mx.z, mx.w);
          uint2 pk2 = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
          *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(B.out) + (int64_t(q >> 1) * B.rows_alloc + row0 + j) * 16 +
                                    (q & 1) * 8) = pk2;
        } else if (B.rows_alloc) {
          *(reinterpret_cast<float4*>(B.out) + int64_t(q) * B.rows_alloc + row0 + j) = mx;
        } else {
          *reinterpret_cast<float4*>(B.out + (site * B.L1 + j) * int64_t(C) + 4 * q) = mx;
        }
      }
    }
    if (cat_out) {
      for (int j = tid; j < n_cat; j += 256) {
        int idx = 0;
        bool bad = false;
        for (int d = 0; d < order; ++d) {
          const int sy = sym_at(R - local_R + j + d);
          bad |= sy > 3;
          idx = idx * 4 + (sy & 3);
        }
        cat_out[site * n_cat + j] = bad ? (1 << (2 * order)) : idx;
      }
    }
    // next site's window into the other buffer; one barrier publishes it and retires this site's readers
    bool nslow = false;
    if (nxt < ns) nslow = commit_fetch(pkbuf + (cur ^ 1) * nW);
    slow = __syncthreads_or(nslow) != 0;
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------ local MLP
constexpr int MLP_TS = 32;       // sites per CTA
constexpr int MLP_THREADS = 160;

// Activations are kept transposed in shared memory ([feature][site]) so one LDS.128 feeds four FMAs of the
// same weight; each thread owns one output neuron for all (or half) of the CTA's sites.
__global__ void __launch_bounds__(MLP_THREADS) k_local_mlp(LocalDev P, const int32_t* __restrict__ cat32,
                                                           const int64_t* __restrict__ cat64, int64_t n, int n_cat,
                                                           int emb_rows, int K1, int H1, int H2, int NC,
                                                           float* __restrict__ logits, int* __restrict__ err_flag,
                                                           const float* __restrict__ cont, int n_cont) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* xT = reinterpret_cast<float*>(smem_raw);  // [K1][TS]
  float* h1T = xT + K1 * MLP_TS;                   // [H1][TS]
  float* h2T = h1T + H1 * MLP_TS;                  // [H2][TS]
  const int64_t site0 = int64_t(blockIdx.x) * MLP_TS;
  const int tid = threadIdx.x;
  for (int e = tid; e < MLP_TS * K1; e += MLP_THREADS) {
    const int k = e / MLP_TS, s = e - k * MLP_TS;
    float v = 0.f;
    if (site0 + s < n && k >= 5 * n_cat) {
      v = cont[(site0 + s) * n_cont + (k - 5 * n_cat)];   // continuous features follow the embeddings (model_snv.py:460); BN folded into W1
    } else if (site0 + s < n) {
      const int64_t ci = (site0 + s) * n_cat + k / 5;
      int64_t idx = cat64 ? cat64[ci] : int64_t(cat32[ci]);
      if (idx < 0 || idx >= emb_rows) {  // nn.Embedding would raise IndexError
        if (err_flag) atomicOr(err_flag, 2);
        idx = 0;
      }
      v = P.emb[idx * 5 + k % 5];
    }
    xT[e] = v;
  }
  __syncthreads();
  for (int o = tid; o < H1; o += MLP_THREADS) {
    float acc[MLP_TS];
    const float b = P.b1[o];
#pragma unroll
    for (int s = 0; s < MLP_TS; ++s) acc[s] = b;
#pragma unroll 4
    for (int k = 0; k < K1; ++k) {
      const float w = __ldg(P.W1t + k * H1 + o);
#pragma unroll
      for (int s4 = 0; s4 < MLP_TS / 4; ++s4) {
        const float4 x = *reinterpret_cast<const float4*>(xT + k * MLP_TS + 4 * s4);
        acc[4 * s4] = fmaf(x.x, w, acc[4 * s4]); acc[4 * s4 + 1] = fmaf(x.y, w, acc[4 * s4 + 1]);
        acc[4 * s4 + 2] = fmaf(x.z, w, acc[4 * s4 + 2]); acc[4 * s4 + 3] = fmaf(x.w, w, acc[4 * s4 + 3]);
      }
    }
#pragma unroll
    for (int s4 = 0; s4 < MLP_TS / 4; ++s4)
      *reinterpret_cast<float4*>(h1T + o * MLP_TS + 4 * s4) =
          make_float4(fmaxf(acc[4 * s4], 0.f), fmaxf(acc[4 * s4 + 1], 0.f), fmaxf(acc[4 * s4 + 2], 0.f), fmaxf(acc[4 * s4 + 3], 0.f));
  }
  __syncthreads();
  for (int item = tid; item < 2 * H2; item += MLP_THREADS) {  // (output neuron, half of the sites)
    const int o = item % H2, hs = (item / H2) * (MLP_TS / 2);
    float acc[MLP_TS / 2];
    const float b = P.b2[o];
#pragma unroll
    for (int s = 0; s < MLP_TS / 2; ++s) acc[s] = b;
#pragma unroll 4
    for (int k = 0; k < H1; ++k) {
      const float w = __ldg(P.W2t + k * H2 + o);
#pragma unroll
      for (int s4 = 0; s4 < MLP_TS / 8; ++s4) {
        const float4 x = *reinterpret_cast<const float4*>(h1T + k * MLP_TS + hs + 4 * s4);
        acc[4 * s4] = fmaf(x.x, w, acc[4 * s4]); acc[4 * s4 + 1] = fmaf(x.y, w, acc[4 * s4 + 1]);
        acc[4 * s4 + 2] = fmaf(x.z, w, acc[4 * s4 + 2]); acc[4 * s4 + 3] = fmaf(x.w, w, acc[4 * s4 + 3]);
      }
    }
#pragma unroll
    for (int s4 = 0; s4 < MLP_TS / 8; ++s4)
      *reinterpret_cast<float4*>(h2T + o * MLP_TS + hs + 4 * s4) =
          make_float4(fmaxf(acc[4 * s4], 0.f), fmaxf(acc[4 * s4 + 1], 0.f), fmaxf(acc[4 * s4 + 2], 0.f), fmaxf(acc[4 * s4 + 3], 0.f));
  }
  __syncthreads();
  for (int e = tid; e < MLP_TS * NC; e += MLP_THREADS) {
    const int s = e / NC, o = e - s * NC;
    if (site0 + s >= n) continue;
    float acc = P.b3[o];
    for (int k = 0; k < H2; ++k) acc = fmaf(h2T[k * MLP_TS + s], __ldg(P.W3t + k * NC + o), acc);
    logits[(site0 + s) * NC + o] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ conv
// rows = n_sites*L flattened; a CTA computes TP consecutive rows x C channels.  Zero padding is per site:
// the contribution of tap t to output row r is dropped when (r % L) + t - ks/2 falls outside [0, L).
template <int C>
struct ConvTile {
  static constexpr int CG = C / 4;     // channel groups (4 channels each)
  static constexpr int RG = 128 / CG;  // row groups (4 rows each)
  static constexpr int TP = RG * 4;    // rows per CTA
  static constexpr int XS = C + 1;     // padded smem row stride
};

template <int C>
__global__ void __launch_bounds__(128) k_conv(const float* __restrict__ in, float* __restrict__ out,
                                              const float* __restrict__ res1, const float* __restrict__ res2,
                                              int64_t rows, int L, ConvLayerDev P, int relu_out) {
  using T = ConvTile<C>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* ws = reinterpret_cast<float*>(smem_raw);  // [ks][C][C]
  float* xs = ws + P.ks * C * C;                   // [TP+ks-1][C+1]
  const int ks = P.ks, half = ks / 2;
  const int tid = threadIdx.x;
  const int64_t r0 = int64_t(blockIdx.x) * T::TP;
  for (int e = tid * 4; e < ks * C * C; e += 128 * 4) *reinterpret_cast<float4*>(ws + e) = *reinterpret_cast<const float4*>(P.Wt + e);
  for (int e = tid; e < (T::TP + ks - 1) * C; e += 128) {
    const int k = e / C, ci = e - k * C;
    const int64_t r = r0 - half + k;
    float v = 0.f;
    if (r >= 0 && r < rows) {
      v = in[r * C + ci];
      if (P.relu_in) v = fmaxf(v, 0.f);
      v = fmaf(v, P.a[ci], P.b[ci]);
    }
    xs[k * T::XS + ci] = v;
  }
  __syncthreads();
  const int tc = tid % T::CG, tr = tid / T::CG;
  int pin[4];  // position of each of this thread's rows inside its site
#pragma unroll
  for (int i = 0; i < 4; ++i) pin[i] = int((r0 + tr * 4 + i) % L);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int t = 0; t < ks; ++t) {
    float tmp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) tmp[i][j] = 0.f;
    const float* wt = ws + t * C * C + 4 * tc;
    const float* xt = xs + (tr * 4 + t) * T::XS;
#pragma unroll 8
    for (int ci = 0; ci < C; ++ci) {
      const float4 w = *reinterpret_cast<const float4*>(wt + ci * C);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = xt[i * T::XS + ci];
        tmp[i][0] = fmaf(x, w.x, tmp[i][0]);
        tmp[i][1] = fmaf(x, w.y, tmp[i][1]);
        tmp[i][2] = fmaf(x, w.z, tmp[i][2]);
        tmp[i][3] = fmaf(x, w.w, tmp[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = pin[i] + t - half;
      if (q >= 0 && q < L) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += tmp[i][j];
      }
    }
  }
  const float4 bias = *reinterpret_cast<const float4*>(P.bias + 4 * tc);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + tr * 4 + i;
    if (r >= rows) continue;
    float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
    if (res1) {
      const float4 q = *reinterpret_cast<const float4*>(res1 + r * C + 4 * tc);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (res2) {
      const float4 q = *reinterpret_cast<const float4*>(res2 + r * C + 4 * tc);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (relu_out) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    *reinterpret_cast<float4*>(out + r * C + 4 * tc) = v;
  }
}

// ------------------------------------------------------------------------------------------------ pool
__global__ void k_pool(const float* __restrict__ in, float* __restrict__ out, int64_t n, int Lin, int Lout, int C, int pk,
                       int ps, int pp) {
  const int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (e >= n * Lout * C) return;
  const int c = int(e % C);
  const int64_t sj = e / C;
  const int j = int(sj % Lout);
  const int64_t site = sj / Lout;
  int lo = j * ps - pp, hi = lo + pk;
  lo = lo < 0 ? 0 : lo;
  hi = hi > Lin ? Lin : hi;
  float mx = -FLT_MAX;
  for (int p = lo; p < hi; ++p) mx = fmaxf(mx, in[(site * Lin + p) * C + c]);
  out[e] = mx;
}

// ------------------------------------------------------------------------------------------------ head
struct HeadBranch {
  const float* x;  // [n][L3][C]  conv3 output after ReLU
  const float* Wfc;
  const float* bfc;
  int L3;
};

__global__ void __launch_bounds__(128) k_head(HeadBranch b0, HeadBranch b1, const float* __restrict__ local_logits, int64_t n,
                                              int C, int NC, float* __restrict__ logp, float* __restrict__ tap_gmax0,
                                              float* __restrict__ tap_gmax1, float* __restrict__ tap_logit0,
                                              float* __restrict__ tap_logit1) {
  const int lane = threadIdx.x & 31;
  const int64_t site = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
  if (site >= n) return;
  float lg[2][16];
#pragma unroll 1
  for (int br = 0; br < 2; ++br) {
    const HeadBranch& B = br ? b1 : b0;
    float part[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) part[o] = 0.f;
    for (int c = lane; c < C; c += 32) {
      float mx = -FLT_MAX;
      for (int p = 0; p < B.L3; ++p) mx = fmaxf(mx, B.x[(site * B.L3 + p) * C + c]);  // torch.max(dim=2), model_snv.py:489,511
      float* tg = br ? tap_gmax1 : tap_gmax0;
      if (tg) tg[site * C + c] = mx;
#pragma unroll
      for (int o = 0; o < 16; ++o)
        if (o < NC) part[o] = fmaf(mx, B.Wfc[c * NC + o], part[o]);
    }
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (o < NC) lg[br][o] = warp_sum(part[o]) + B.bfc[o];
  }
  if (lane != 0) return;
  float sm[3][16];
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    float mx = -FLT_MAX, sum = 0.f;
    for (int o = 0; o < NC; ++o) {
      const float v = k == 2 ? local_logits[site * NC + o] : lg[k][o];
      sm[k][o] = v;
      mx = fmaxf(mx, v);
    }
    for (int o = 0; o < NC; ++o) {
      sm[k][o] = expf(sm[k][o] - mx);
      sum += sm[k][o];
    }
    for (int o = 0; o < NC; ++o) sm[k][o] /= sum;
  }
  for (int o = 0; o < NC; ++o) {
    if (tap_logit0) tap_logit0[site * NC + o] = lg[0][o];
    if (tap_logit1) tap_logit1[site * NC + o] = lg[1][o];
    const float distal = (sm[0][o] + sm[1][o]) / 2.f;                  // model_snv.py:515
    logp[site * NC + o] = logf(fmaxf((sm[2][o] + distal) / 2.f, 1e-9f));  // :523
  }
}

// ------------------------------------------------------------------------------------------------ loss
__global__ void k_ce_sum(const float* __restrict__ logp, const int32_t* __restrict__ meta, int64_t n, int NC,
                         double* __restrict__ loss) {
  // CrossEntropyLoss(reduction='sum') applied to log-probs: -log_softmax(logp)[y]  (nn_utils.py:64)
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  float v = 0.f;
  if (i < n) {
    const float* p = logp + i * NC;
    float mx = -FLT_MAX;
    for (int o = 0; o < NC; ++o) mx = fmaxf(mx, p[o]);
    float s = 0.f;
    for (int o = 0; o < NC; ++o) s += expf(p[o] - mx);
    const int y = (meta[i] >> 1) & 0x7f;
    v = -(p[y < NC ? y : 0] - mx - logf(s));
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(loss, double(v));
}

// ------------------------------------------------------------------------------------------------ host
template <int C>
static int conv_launch(const float* in, float* out, const float* r1, const float* r2, int64_t n, int L, const ConvLayerDev& P,
                       int relu_out, cudaStream_t st) {
  using T = ConvTile<C>;
  const int64_t rows = n * L;
  const size_t smem = sizeof(float) * (size_t(P.ks) * C * C + size_t(T::TP + P.ks - 1) * T::XS);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CUDA_TRY(cudaFuncSetAttribute(k_conv<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  LAUNCH(k_conv<C>, (unsigned)cdiv(rows, T::TP), 128, smem, st, in, out, r1, r2, rows, L, P, relu_out);
  return 0;
}

int conv_any(int C, const float* in, float* out, const float* r1, const float* r2, int64_t n, int L,
                    const ConvLayerDev& P, int relu_out, cudaStream_t st) {
  switch (C) {
    case 16: return conv_launch<16>(in, out, r1, r2, n, L, P, relu_out, st);
    case 32: {
      // tensor-core conv at fp32-equivalent precision (snv_conv_mma.cu); MURAL_NO_CONV_MMA=1 keeps the fp32 FMA kernel (parity switch)
      static const bool use_mma = getenv("MURAL_NO_CONV_MMA") == nullptr;
      // P.precise (training forward): the fp32 FMA kernel — a third split level costs more mma.sync issue slots than 96 FMAs
      // (measured 174 us vs 136 us per 549k rows, scratch/mb_conv_mma.py), and forward activations at fp32 precision keep
      // ReLU-kink flips against an fp64 reference rare
      if (P.ks == 3 && use_mma && !P.precise) return conv32_mma(in, out, r1, r2, n, L, P, relu_out, st);
      return conv_launch<32>(in, out, r1, r2, n, L, P, relu_out, st);
    }
    case 64: return conv_launch<64>(in, out, r1, r2, n, L, P, relu_out, st);
  }
  MURAL_FAIL("unsupported channel count");
}

static int pool_launch(const float* in, float* out, int64_t n, int Lin, int Lout, int C, const int* p, cudaStream_t st) {
  const int64_t tot = n * Lout * C;
  LAUNCH(k_pool, (unsigned)cdiv(tot, 256), 256, 0, st, in, out, n, Lin, Lout, C, p[0], p[1], p[2]);
  return 0;
}

static int save_tap(mural_snv_model* m, const char* name, const float* d, int64_t floats, cudaStream_t st) {
  if (!m->debug) return 0;
  std::vector<float>& v = m->tap_store[name];
  v.resize(floats);
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaMemcpy(v.data(), d, floats * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int snv_stem_launch(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                    const uint8_t* d_sym, int64_t ns, float* mid_out, float* large_out, int32_t* cat_out, cudaStream_t st) {
  return snv_stem_launch_planes(m, G, d_pos, d_meta, d_sym, ns, mid_out, 0, large_out, 0, cat_out, st);
}

int snv_stem_launch_planes(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                           const uint8_t* d_sym, int64_t ns, float* mid_out, int64_t mid_rows_alloc, float* large_out,
                           int64_t large_rows_alloc, int32_t* cat_out, cudaStream_t st, bool out_bf16, const int* skip_flag) {
  const int C = m->cfg.channels, ks = m->cfg.kernel_size, L = m->L, R = m->cfg.distal_radius;
  StemBranch sb[2];
  for (int br = 0; br < 2; ++br) {
    const BranchDev& B = m->br[br];
    sb[br] = StemBranch{B.T, B.bias1, B.T4, br ? large_out : mid_out, br ? large_rows_alloc : mid_rows_alloc, out_bf16 ? 1 : 0, B.L0,
                        br ? 0 : L / 2 - 100, B.L1, B.pool[0][0], B.pool[0][1], B.pool[0][2]};
  }
  GenomeView gvf = G ? *G : GenomeView{};
  if (ks == 3 && !m->slow_stem) {
    const size_t smem_f = sizeof(float) * (2 * size_t(256) * C + 2 * 3 * 16 * size_t(C) + 2 * C) + ((size_t(L) + 4 + 15) & ~size_t(15));
    int grid = int(ns < 148 * 2 ? ns : 148 * 2);
    if (!d_sym) {  // genome path: packed 2-bit stem
      const size_t smem_p = smem_f + 8 * (size_t(L + 15) / 16 + 3);
#define STEMP_CASE(CC)                                                                                                     \
  case CC: {                                                                                                              \
    static size_t configured = 0;                                                                                         \
    if (smem_p > configured) {                                                                                            \
      CUDA_TRY(cudaFuncSetAttribute(k_stem_pk<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));            \
      configured = smem_p;                                                                                                \
    }                                                                                                                     \
    LAUNCH(k_stem_pk<CC>, grid, 256, smem_p, st, gvf, d_pos, d_meta, ns, R, L, sb[0], sb[1], m->cfg.local_radius,         \
           m->cfg.local_order, m->n_cat, cat_out, skip_flag);                                                             \
  } break;
      switch (C) {
        STEMP_CASE(16)
        STEMP_CASE(32)
        STEMP_CASE(64)
        default: MURAL_FAIL("unsupported channel count");
      }
#undef STEMP_CASE
      return 0;
    }
#define STEMF_CASE(CC)                                                                                                     \
  case CC: {                                                                                                              \
    static size_t configured = 0;                                                                                         \
    if (smem_f > configured) {                                                                                            \
      CUDA_TRY(cudaFuncSetAttribute(k_stem_fast<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));          \
      configured = smem_f;                                                                                                \
    }                                                                                                                     \
    LAUNCH(k_stem_fast<CC>, grid, 256, smem_f, st, gvf, d_pos, d_meta, d_sym, ns, R, L, sb[0], sb[1], m->cfg.local_radius, \
           m->cfg.local_order, m->n_cat, cat_out);                                                                        \
  } break;
    switch (C) {
      STEMF_CASE(16)
      STEMF_CASE(32)
      STEMF_CASE(64)
      default: MURAL_FAIL("unsupported channel count");
    }
#undef STEMF_CASE
    return 0;
  }
  const size_t smem = sizeof(float) * (2 * size_t(ks) * 16 * C + 2 * C) + ((size_t(L) + 15) & ~size_t(15));
  GenomeView gv = G ? *G : GenomeView{};
#define STEM_CASE(CC)                                                                                                    \
  case CC: {                                                                                                             \
    static size_t configured = 0;                                                                                        \
    if (smem > 48 * 1024 && smem > configured) {                                                                         \
      CUDA_TRY(cudaFuncSetAttribute(k_stem<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                \
      configured = smem;                                                                                                 \
    }                                                                                                                    \
    LAUNCH(k_stem<CC>, (unsigned)ns, 128, smem, st, gv, d_pos, d_meta, d_sym, R, L, ks, sb[0], sb[1], m->cfg.local_radius, \
           m->cfg.local_order, m->n_cat, cat_out);                                                                       \
  } break;
  switch (C) {
    STEM_CASE(16)
    STEM_CASE(32)
    STEM_CASE(64)
    default: MURAL_FAIL("unsupported channel count");
  }
#undef STEM_CASE
  return 0;
}

int snv_local_launch(mural_snv_model* m, const int32_t* cat32, const int64_t* cat64, int64_t ns, float* logits, int* err_flag,
                     cudaStream_t st) {
  const int K1 = m->k1c, H1 = m->cfg.hidden1, H2 = m->cfg.hidden2;   // embeddings + continuous features
  MURAL_CHECK(m->cfg.n_cont == 0 || m->d_cont_cur != nullptr, "this model has continuous features: call mural_snv_set_cont before the forward");
  const size_t smem = sizeof(float) * MLP_TS * size_t(K1 + H1 + H2);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    CUDA_TRY(cudaFuncSetAttribute(k_local_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  LAUNCH(k_local_mlp, (unsigned)cdiv(ns, MLP_TS), MLP_THREADS, smem, st, m->local, cat32, cat64, ns, m->n_cat, m->emb_rows, K1, H1, H2,
         m->cfg.n_class, logits, err_flag, m->d_cont_cur, m->cfg.n_cont);
  return 0;
}

int snv_head_launch(mural_snv_model* m, const float* h_mid, const float* h_large, const float* local_logits, int64_t ns,
                    float* logp, float* tg0, float* tg1, float* tl0, float* tl1, cudaStream_t st) {
  HeadBranch hb[2] = {{h_mid, m->br[0].Wfc, m->br[0].bfc, m->br[0].L3}, {h_large, m->br[1].Wfc, m->br[1].bfc, m->br[1].L3}};
  LAUNCH(k_head, (unsigned)cdiv(ns * 32, 128), 128, 0, st, hb[0], hb[1], local_logits, ns, m->cfg.channels, m->cfg.n_class, logp,
         tg0, tg1, tl0, tl1);
  return 0;
}

int snv_ensure_workspace(mural_snv_model* m, int64_t bytes) {
  if (m->ws_bytes >= bytes) return 0;
  cudaFree(m->d_ws);
  m->d_ws = nullptr;
  m->ws_bytes = 0;
  CUDA_TRY(cudaMalloc(&m->d_ws, bytes));
  m->ws_bytes = bytes;
  return 0;
}

int snv_forward_fp32(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                     const uint8_t* d_sym, const int64_t* d_cat, int64_t n, float* d_logp, cudaStream_t st, bool aux_ws) {
  const int C = m->cfg.channels, NC = m->cfg.n_class;
  const BranchDev &Bm = m->br[0], &Bl = m->br[1];
  const int64_t Lmax = Bl.L1 > Bm.L1 ? Bl.L1 : Bm.L1;
  // floats per site: 4 rotating activation buffers at the widest length, the mid branch pool-1 output,
  // both conv3 outputs, local logits, tap scratch, k-mer indices
  const int64_t per_site = 4 * Lmax * C + int64_t(Bm.L1) * C + int64_t(Bm.L3 + Bl.L3) * C + 3 * NC + 2 * C + m->n_cat;
  int64_t chunk = m->chunk_sites > 0 ? m->chunk_sites : 2048;
  if (chunk > n) chunk = n;
  float* w;
  if (aux_ws) {
    const int64_t bytes = chunk * per_site * 4 + 256;
    if (m->ws2_bytes < bytes) {
      cudaFree(m->d_ws2);
      m->d_ws2 = nullptr;
      m->ws2_bytes = 0;
      CUDA_TRY(cudaMalloc(&m->d_ws2, bytes));
      m->ws2_bytes = bytes;
    }
    w = (float*)m->d_ws2;
  } else {
    if (int rc = snv_ensure_workspace(m, chunk * per_site * 4 + 256)) return rc;
    w = (float*)m->d_ws;
  }
  float* buf[4];
  for (int i = 0; i < 4; ++i) { buf[i] = w; w += chunk * Lmax * C; }
  float* mid0 = w; w += chunk * int64_t(Bm.L1) * C;
  float* hmid = w; w += chunk * int64_t(Bm.L3) * C;
  float* hlarge = w; w += chunk * int64_t(Bl.L3) * C;
  float* llog = w; w += chunk * NC;
  float* tl0 = w; w += chunk * NC;
  float* tl1 = w; w += chunk * NC;
  float* tg0 = w; w += chunk * C;
  float* tg1 = w; w += chunk * C;
  int32_t* cat32 = (int32_t*)w; w += chunk * m->n_cat;
  int* err_flag = (int*)w;
  CUDA_TRY(cudaMemsetAsync(err_flag, 0, 4, st));

  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int64_t ns = (n - s0 < chunk) ? (n - s0) : chunk;
    if (int rc = snv_stem_launch(m, G, d_pos ? d_pos + s0 : nullptr, d_meta ? d_meta + s0 : nullptr,
                                 d_sym ? d_sym + s0 * m->L : nullptr, ns, mid0, buf[0], d_cat ? nullptr : cat32, st))
      return rc;
    m->d_cont_cur = m->d_cont ? m->d_cont + s0 * m->cfg.n_cont : nullptr;
    if (int rc = snv_local_launch(m, d_cat ? nullptr : cat32, d_cat ? d_cat + s0 * m->n_cat : nullptr, ns, llog, err_flag, st))
      return rc;
    for (int br = 1; br >= 0; --br) {  // large first (its pool-1 output sits in buf[0]), then mid
      const BranchDev& B = m->br[br];
      float* P = br ? buf[0] : mid0;
      float *Q = buf[1], *Rb = buf[2], *U = buf[3];
      const char* sfx = br ? "_2" : "";
      if (int rc = save_tap(m, (std::string("pool1") + sfx).c_str(), P, ns * B.L1 * C, st)) return rc;
      if (!m->debug) {   // the whole branch in one kernel, activations in shared memory (the parity taps need the per-layer path)
        const int rc = snv_site_chain_launch(m, br, P, br ? hlarge : hmid, ns, st);
        if (rc > 0) return rc;
        if (rc == 0) continue;
      }
      // ---- stage 1 at length L1: two ResBlocks + outer skip (model_snv.py:477-479 / 499-501)
      if (int rc = conv_any(C, P, Q, nullptr, nullptr, ns, B.L1, B.rb1[0], 0, st)) return rc;
      if (int rc = conv_any(C, Q, Rb, P, nullptr, ns, B.L1, B.rb1[1], 0, st)) return rc;   // y1 = x0 + f(x0)
      if (int rc = conv_any(C, Rb, Q, nullptr, nullptr, ns, B.L1, B.rb1[2], 0, st)) return rc;
      if (int rc = conv_any(C, Q, U, Rb, P, ns, B.L1, B.rb1[3], 0, st)) return rc;          // y2 + jump = y1 + f(y1) + x0
      if (int rc = save_tap(m, (std::string("rb1") + sfx).c_str(), U, ns * B.L1 * C, st)) return rc;
      // ---- pool 2, conv2, stage 2 (:480-485 / 502-507)
      float* X2 = br ? P : buf[0];  // mid0 is smaller than needed? no: reuse buf[0] for the mid branch
      if (int rc = pool_launch(U, X2, ns, B.L1, B.L2, C, B.pool[1], st)) return rc;
      if (int rc = conv_any(C, X2, Q, nullptr, nullptr, ns, B.L2, B.conv2, 0, st)) return rc;  // jump2
      if (int rc = save_tap(m, (std::string("conv2") + sfx).c_str(), Q, ns * B.L2 * C, st)) return rc;
      if (int rc = conv_any(C, Q, Rb, nullptr, nullptr, ns, B.L2, B.rb2[0], 0, st)) return rc;
      if (int rc = conv_any(C, Rb, U, Q, nullptr, ns, B.L2, B.rb2[1], 0, st)) return rc;
      if (int rc = conv_any(C, U, Rb, nullptr, nullptr, ns, B.L2, B.rb2[2], 0, st)) return rc;
      if (int rc = conv_any(C, Rb, X2, U, Q, ns, B.L2, B.rb2[3], 0, st)) return rc;
      if (int rc = save_tap(m, (std::string("rb2") + sfx).c_str(), X2, ns * B.L2 * C, st)) return rc;
      // ---- pool 3, conv3 + ReLU (:486-488 / 508-510)
      if (int rc = pool_launch(X2, Rb, ns, B.L2, B.L3, C, B.pool[2], st)) return rc;
      if (int rc = conv_any(C, Rb, br ? hlarge : hmid, nullptr, nullptr, ns, B.L3, B.conv3, 1, st)) return rc;
    }
    if (int rc = snv_head_launch(m, hmid, hlarge, llog, ns, d_logp + s0 * NC, m->debug ? tg0 : nullptr, m->debug ? tg1 : nullptr,
                                 m->debug ? tl0 : nullptr, m->debug ? tl1 : nullptr, st))
      return rc;
    if (m->debug) {
      if (int rc = save_tap(m, "gmax", tg0, ns * C, st)) return rc;
      if (int rc = save_tap(m, "gmax_2", tg1, ns * C, st)) return rc;
      if (int rc = save_tap(m, "logit_mid", tl0, ns * NC, st)) return rc;
      if (int rc = save_tap(m, "logit_large", tl1, ns * NC, st)) return rc;
      if (int rc = save_tap(m, "logit_local", llog, ns * NC, st)) return rc;
    }
  }
  CUDA_TRY(cudaGetLastError());
  if (d_cat) {  // tensor path: surface out-of-range embedding indices like nn.Embedding would
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, err_flag, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    MURAL_CHECK((flag & 2) == 0, "IndexError: index out of range in embedding lookup");
  }
  return 0;
}

}  // namespace mural

using namespace mural;

static int check_model(const mural_snv_model_t* m) {
  MURAL_CHECK(m != nullptr, "model is NULL");
  MURAL_CHECK(m->loaded, "model weights not loaded (call mural_snv_model_load first)");
  return 0;
}

extern "C" int mural_snv_set_debug(mural_snv_model_t* m, int32_t on) {
  MURAL_CHECK(m != nullptr, "model is NULL");
  m->debug = (on & 1) != 0;      // bit 0: keep parity taps of the last chunk
  m->slow_stem = (on & 2) != 0;  // bit 1: force the generic per-tap stem kernel
  m->tap_store.clear();
  return 0;
}
extern "C" int mural_snv_set_chunk(mural_snv_model_t* m, int64_t chunk_sites) {
  MURAL_CHECK(m != nullptr && chunk_sites >= 0, "bad argument");
  m->chunk_sites = chunk_sites;
  return 0;
}

// ---- MURAL_MODE_AUTO: the bf16 tensor-core path for every site, and the fp32-equivalent path again for the sites whose
// expanded window holds a non-ACGT symbol or overhangs its chromosome (inputs far from what the network was trained on can
// drive activations an order of magnitude up, where bf16's relative rounding no longer stays inside the 5e-3 gate;
// DESIGN.md 5).  The site list is built on the device; the host learns its length from a 4-byte copy that completes long
// before the bf16 pass it is enqueued in front of, so the call stays asynchronous for the bulk of the work.
namespace mural {
__global__ void __launch_bounds__(256) k_exception_sites(GenomeView G, const int32_t* __restrict__ pos, const int32_t* __restrict__ meta,
                                                         int64_t n, int R, int* __restrict__ count, int32_t* __restrict__ idx) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int chrom = meta[i] >> 8;
  const int64_t lo = int64_t(pos[i]) - R, hi = int64_t(pos[i]) + R;  // inclusive window (preprocessing.py:524-567, snv)
  bool exc = lo < 0 || hi >= G.chrom_len[chrom];
  if (!exc && G.n_exc > 0) {
    const int64_t glo = G.chrom_off[chrom] + lo, ghi = G.chrom_off[chrom] + hi;
    int a = 0, b = G.n_exc;  // first run whose (exclusive) end lies beyond the window start
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (__ldg(G.exc_end + mid) > glo) b = mid; else a = mid + 1;
    }
    exc = a < G.n_exc && __ldg(G.exc_start + a) <= ghi;
  }
  if (exc) idx[atomicAdd(count, 1)] = int32_t(i);
}
__global__ void __launch_bounds__(256) k_exception_windows(const uint8_t* __restrict__ sym, int64_t n, int L, int* __restrict__ count,
                                                           int32_t* __restrict__ idx) {
  const int64_t site = blockIdx.x * int64_t(blockDim.x / 32) + (threadIdx.x >> 5);
  if (site >= n) return;
  bool exc = false;
  for (int p = threadIdx.x & 31; p < L; p += 32) exc |= sym[site * L + p] >= 4;
  if (__any_sync(0xffffffffu, exc) && (threadIdx.x & 31) == 0) idx[atomicAdd(count, 1)] = int32_t(site);
}
__global__ void __launch_bounds__(256) k_gather_sites(const int32_t* __restrict__ pos, const int32_t* __restrict__ meta,
                                                      const int32_t* __restrict__ idx, int cnt, int32_t* __restrict__ pos2,
                                                      int32_t* __restrict__ meta2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cnt) return;
  pos2[j] = pos[idx[j]];
  meta2[j] = meta[idx[j]];
}
__global__ void __launch_bounds__(256) k_gather_windows(const uint8_t* __restrict__ sym, const int64_t* __restrict__ cat, const int32_t* __restrict__ idx,
                                                        int cnt, int L, int n_cat, uint8_t* __restrict__ sym2, int64_t* __restrict__ cat2) {
  const int j = blockIdx.x;
  if (j >= cnt) return;
  const int64_t s = idx[j];
  for (int p = threadIdx.x; p < L; p += blockDim.x) sym2[int64_t(j) * L + p] = sym[s * L + p];
  for (int c = threadIdx.x; c < n_cat; c += blockDim.x) cat2[int64_t(j) * n_cat + c] = cat[s * n_cat + c];
}
__global__ void __launch_bounds__(256) k_scatter_rows(const float* __restrict__ src, const int32_t* __restrict__ idx, int cnt, int NC,
                                                      float* __restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cnt * NC) return;
  dst[int64_t(idx[e / NC]) * NC + e % NC] = src[e];
}

// scratch of the auto mode: [count | idx n | pos2 n | meta2 n | logp2 n*NC] (+ windows / k-mer rows on the tensor route)
static int auto_scratch(mural_snv_model* m, int64_t bytes) {
  if (m->auto_bytes < bytes) {
    cudaFree(m->d_auto);
    m->d_auto = nullptr;
    m->auto_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_auto, bytes));
    m->auto_bytes = bytes;
  }
  if (!m->h_auto) CUDA_TRY(cudaMallocHost(&m->h_auto, 64));
  if (!m->auto_ev) CUDA_TRY(cudaEventCreateWithFlags((cudaEvent_t*)&m->auto_ev, cudaEventDisableTiming));
  if (!m->aux_ev) CUDA_TRY(cudaEventCreateWithFlags((cudaEvent_t*)&m->aux_ev, cudaEventDisableTiming));
  if (!m->aux_stream) CUDA_TRY(cudaStreamCreateWithFlags((cudaStream_t*)&m->aux_stream, cudaStreamNonBlocking));
  return 0;
}

int snv_forward_auto(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta, const uint8_t* d_sym,
                     const int64_t* d_cat, int64_t n, float* d_logp, cudaStream_t st) {
  if (!m->tc) return snv_forward_fp32(m, G, d_pos, d_meta, d_sym, d_cat, n, d_logp, st);  // shape without a tcgen05 path
  MURAL_CHECK(n < (int64_t(1) << 31), "more than 2^31 sites in one call");
  const int NC = m->cfg.n_class;
  if (int rc = auto_scratch(m, 256 + n * (12 + 4 * NC))) return rc;
  int* d_count = (int*)m->d_auto;
  int32_t* idx = (int32_t*)((char*)m->d_auto + 256);
  int32_t* pos2 = idx + n;
  int32_t* meta2 = pos2 + n;
  float* logp2 = (float*)(meta2 + n);
  CUDA_TRY(cudaMemsetAsync(d_count, 0, 4, st));
  if (G) LAUNCH(k_exception_sites, (unsigned)cdiv(n, 256), 256, 0, st, *G, d_pos, d_meta, n, m->cfg.distal_radius, d_count, idx);
  else LAUNCH(k_exception_windows, (unsigned)cdiv(n, 8), 256, 0, st, d_sym, n, m->L, d_count, idx);
  int* h_count = (int*)m->h_auto;
  CUDA_TRY(cudaMemcpyAsync(h_count, d_count, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord((cudaEvent_t)m->auto_ev, st));
  if (int rc = snv_forward_tc(m, G, d_pos, d_meta, d_sym, d_cat, n, d_logp, st)) return rc;
  CUDA_TRY(cudaEventSynchronize((cudaEvent_t)m->auto_ev));  // completed while the bf16 pass was being enqueued
  const int cnt = *h_count;
  m->last_auto_sites = cnt;
  if (cnt == 0) return 0;
  if (G) {
    // The recompute is a few thousand sites in ~50 small kernels: on the caller's stream it would run after the bf16 pass and
    // cost its full latency-bound time (0.9 ms per 2^20-site call).  The site list is final (the host has just read its
    // length), so it goes to a side stream with its own workspace and fills idle SM resources while the bf16 pass, whose
    // enqueue has only just finished, executes; the caller's stream joins before the scatter.
    cudaStream_t ax = (cudaStream_t)m->aux_stream;
    LAUNCH(k_gather_sites, (unsigned)cdiv(cnt, 256), 256, 0, ax, d_pos, d_meta, idx, cnt, pos2, meta2);
    if (int rc = snv_forward_fp32(m, G, pos2, meta2, nullptr, nullptr, cnt, logp2, ax, true)) return rc;
    CUDA_TRY(cudaEventRecord((cudaEvent_t)m->aux_ev, ax));
    CUDA_TRY(cudaStreamWaitEvent(st, (cudaEvent_t)m->aux_ev, 0));
  } else {
    uint8_t* sym2 = nullptr;  // tensor route (parity surface): windows and k-mer rows of the flagged sites, packed
    const int64_t sb = (int64_t(cnt) * m->L + 255) & ~int64_t(255);
    CUDA_TRY(cudaMalloc((void**)&sym2, sb + int64_t(cnt) * m->n_cat * 8));
    int64_t* cat2 = (int64_t*)(sym2 + sb);
    LAUNCH(k_gather_windows, (unsigned)cnt, 128, 0, st, d_sym, d_cat, idx, cnt, m->L, m->n_cat, sym2, cat2);
    int rc = snv_forward_fp32(m, nullptr, nullptr, nullptr, sym2, cat2, cnt, logp2, st);
    cudaStreamSynchronize(st);
    cudaFree(sym2);
    if (rc) return rc;
  }
  LAUNCH(k_scatter_rows, (unsigned)cdiv(int64_t(cnt) * NC, 256), 256, 0, st, logp2, idx, cnt, NC, d_logp);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace mural

extern "C" int mural_snv_forward(mural_snv_model_t* m, const mural_genome_t* g, const int32_t* d_pos, const int32_t* d_meta,
                                 int64_t n, int32_t mode, float* d_logp, void* stream) {
  if (int rc = check_model(m)) return rc;
  MURAL_CHECK(g && (n == 0 || (d_pos && d_meta && d_logp)), "NULL argument");
  MURAL_CHECK(g->device == m->device, "genome and model live on different devices");
  if (n == 0) return 0;
  if (m->cfg.n_cont > 0) {   // continuous-feature models: fp32-equivalent path, cont_x handed over by mural_snv_set_cont
    MURAL_CHECK(mode == MURAL_MODE_FP32, "models with continuous features (n_cont > 0) run in MURAL_MODE_FP32");
    MURAL_CHECK(m->d_cont != nullptr, "this model has continuous features: call mural_snv_set_cont before the forward");
    const int rc = snv_forward_fp32(m, &g->view, d_pos, d_meta, nullptr, nullptr, n, d_logp, (cudaStream_t)stream);
    m->d_cont = m->d_cont_cur = nullptr;
    return rc;
  }
  if (mode == MURAL_MODE_FP32) return snv_forward_fp32(m, &g->view, d_pos, d_meta, nullptr, nullptr, n, d_logp, (cudaStream_t)stream);
  if (mode == MURAL_MODE_BF16) return snv_forward_tc(m, &g->view, d_pos, d_meta, nullptr, nullptr, n, d_logp, (cudaStream_t)stream);
  if (mode == MURAL_MODE_AUTO) return snv_forward_auto(m, &g->view, d_pos, d_meta, nullptr, nullptr, n, d_logp, (cudaStream_t)stream);
  MURAL_FAIL("unknown compute mode");
}

extern "C" int mural_snv_set_cont(mural_snv_model_t* m, const float* d_cont) {
  MURAL_CHECK(m != nullptr, "model is NULL");
  MURAL_CHECK(m->cfg.n_cont > 0 || d_cont == nullptr, "this model has no continuous features (n_cont == 0)");
  m->d_cont = d_cont;
  return 0;
}

extern "C" int mural_conv32_layer(const float* d_in, float* d_out, const float* d_res1, const float* d_res2, int64_t n, int32_t L,
                                  const float* d_Wt, const float* d_bias, const float* d_a, const float* d_b, int32_t relu_in,
                                  int32_t relu_out, int32_t impl, void* stream) {
  MURAL_CHECK(d_in && d_out && d_Wt && d_bias && d_a && d_b && n > 0 && L > 0, "bad argument");
  MURAL_CHECK(impl >= 0 && impl <= 2, "impl must be 0 (fp32 FMA), 1 (two-level split MMA) or 2 (three-level split MMA)");
  ConvLayerDev P{d_Wt, d_bias, d_a, d_b, 3, relu_in, impl == 2 ? 1 : 0};
  int rc = impl == 0 ? conv_launch<32>(d_in, d_out, d_res1, d_res2, n, L, P, relu_out, (cudaStream_t)stream)
                     : conv32_mma(d_in, d_out, d_res1, d_res2, n, L, P, relu_out, (cudaStream_t)stream);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int64_t mural_snv_last_auto_sites(const mural_snv_model_t* m) { return m ? m->last_auto_sites : -1; }

extern "C" int mural_snv_forward_tensors(mural_snv_model_t* m, const int64_t* d_cat, const float* d_distal, int64_t n, int32_t L,
                                         int32_t mode, float* d_logp, void* stream) {
  if (int rc = check_model(m)) return rc;
  MURAL_CHECK(n == 0 || (d_cat && d_distal && d_logp), "NULL argument");
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
    // any column that is not one of the reference's 15 one-hot vectors cannot go through the table stem
    int bad = 0;
    cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (bad) rc = fail(__FILE__, __LINE__, "distal_x holds columns that are not reference one-hot vectors (bigWig channels / arbitrary floats are not supported by the table stem)");
  }
  if (rc == 0 && m->cfg.n_cont > 0) {
    if (mode != MURAL_MODE_FP32) rc = fail(__FILE__, __LINE__, "models with continuous features (n_cont > 0) run in MURAL_MODE_FP32");
    else if (!m->d_cont) rc = fail(__FILE__, __LINE__, "this model has continuous features: call mural_snv_set_cont before the forward");
    else rc = snv_forward_fp32(m, nullptr, nullptr, nullptr, d_sym, d_cat, n, d_logp, st);
    m->d_cont = m->d_cont_cur = nullptr;
  } else if (rc == 0) {
    rc = mode == MURAL_MODE_BF16   ? snv_forward_tc(m, nullptr, nullptr, nullptr, d_sym, d_cat, n, d_logp, st)
         : mode == MURAL_MODE_AUTO ? snv_forward_auto(m, nullptr, nullptr, nullptr, d_sym, d_cat, n, d_logp, st)
                                   : snv_forward_fp32(m, nullptr, nullptr, nullptr, d_sym, d_cat, n, d_logp, st);
  }
  cudaStreamSynchronize(st);
  cudaFree(d_sym);
  return rc;
}

extern "C" int mural_snv_predict_host(mural_snv_model_t* m, const mural_genome_t* g, const int32_t* h_pos,
                                      const int32_t* h_meta, int64_t n, int32_t mode, float* h_logp, void* stream) {
  if (int rc = check_model(m)) return rc;
  MURAL_CHECK(g && (n == 0 || (h_pos && h_meta && h_logp)), "NULL argument");
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(m->device));
  const int NC = m->cfg.n_class;
  // persistent staging buffers (grown on demand) so repeated calls do not pay cudaMalloc
  const int64_t need = n * (8 + 4 * NC);
  if (m->io_bytes < need) {
    cudaFree(m->d_io);
    m->d_io = nullptr;
    m->io_bytes = 0;
    CUDA_TRY(cudaMalloc(&m->d_io, need));
    m->io_bytes = need;
  }
  int32_t* d_pos = (int32_t*)m->d_io;
  int32_t* d_meta = d_pos + n;
  float* d_logp = (float*)(d_meta + n);
  CUDA_TRY(cudaMemcpyAsync(d_pos, h_pos, n * 4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(d_meta, h_meta, n * 4, cudaMemcpyHostToDevice, st));
  if (int rc = mural_snv_forward(m, g, d_pos, d_meta, n, mode, d_logp, stream)) return rc;
  CUDA_TRY(cudaMemcpyAsync(h_logp, d_logp, n * NC * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int mural_ce_sum(const float* d_logp, const int32_t* d_meta, int64_t n, int32_t n_class, double* d_loss,
                            void* stream) {
  MURAL_CHECK(d_logp && d_meta && d_loss && n_class >= 2, "bad argument");
  if (n == 0) return 0;
  LAUNCH(k_ce_sum, (unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream, d_logp, d_meta, n, n_class, d_loss);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int mural_snv_debug_tap(mural_snv_model_t* m, const char* name, float* h_out, int64_t max_floats,
                                   int64_t* n_written) {
  MURAL_CHECK(m && name && n_written, "NULL argument");
  auto it = m->tap_store.find(name);
  MURAL_CHECK(it != m->tap_store.end(), std::string("no such tap (enable mural_snv_set_debug first): ") + name);
  const int64_t k = (int64_t)it->second.size() < max_floats ? (int64_t)it->second.size() : max_floats;
  if (h_out && k) memcpy(h_out, it->second.data(), k * sizeof(float));
  *n_written = (int64_t)it->second.size();
  return 0;
}
