// Dense-site stem for genome-wide prediction.
//
// The first Conv1d of a branch (BN(4) + Conv1d(4->C, k=3) on a one-hot window) at window position i depends only on
// the three genome bases around genomic position g and on the strand, never on the site; only the max-pool bins are
// anchored at the site.  When the sites of a chunk are dense on one chromosome (genome-wide `predict`: every A/T,
// ~2 bp apart) the per-site work  L x C x 3 table lookups  collapses to
//   (1) k_dense_tables : conv output c[g] once per genomic position and strand, and its sliding-window maxima
//                        W_w[g] = max(c[g .. g+w-1]) for the three bin widths a window has (full bin, clipped first
//                        bin, clipped last bin), stored as bf16 rows;
//   (2) k_stem_gather  : per site and bin ONE 64-byte row copy (plus the two window-edge positions whose conv misses a
//                        tap, evaluated exactly from the per-tap tables).
// Values are bit-identical to the per-site stem (same fp32 sums in the same order; max and bf16 rounding commute).
// Sparse chunks (training sets, several chromosomes) keep using the per-site kernel: the decision is taken on the
// device (k_chunk_span) so no host synchronisation is needed.
//
// Stage-1 lattice (same idea one level up).  An interior pooled bin j of a site is W_pk[g0 + j*ps] with g0 fixed by the
// site, so the first ResBlock pair (4 convs at stride-ps1 spacing) of every row whose receptive field stays clear of the
// window ends is a function of the genomic position alone: Y1[g] = RB(W[g-4ps], ..., W[g+4ps]).  We evaluate it ONCE
// per genomic position and strand on 2*ps1 phase-major pseudo-sites (the stage kernel's lattice loader reads the full-bin table rows in place),
// and per site only the LAT_EO rows at each window end, as one 18-row edge pseudo-site whose rows 1..16 are table rows
// read in place by the stage kernel and whose rows 0 and 17 are written by k_stem_gather in edge mode.
// The stage-2 loader (snv_tc.cu) max-pools across lattice rows and edge rows.  Every row goes through the same
// tcgen05 arithmetic as in the per-site path, so the result is bit-identical.
#include <cuda_bf16.h>
#include <float.h>
#include <limits.h>

#include "snv_model.cuh"

namespace mural {

constexpr int DT_POS = 256;   // table positions per CTA
constexpr int DT_RUN = 16;    // consecutive positions per thread (16 threads x 2 channels per run)
constexpr int DT_MAXW = 16;   // widest pool window supported by the table kernel

__device__ __forceinline__ int sym_genomic(const GenomeView& G, int chrom, long long q) {
  return (q >= 0 && q < G.chrom_len[chrom]) ? genome_symbol(G, G.chrom_off[chrom] + q) : SYM_N;
}

// Multi-CTA span detection; the last CTA to finish writes ChunkInfo and re-zeroes the scratch words (self-cleaning:
// scr[] must be zero before the first launch).  scr: [0] max(INT_MAX - pos), [1] max(pos + 1), [2] mixed, [3] has '+',
// [4] has '-', [5] finished-CTA counter.
__global__ void __launch_bounds__(256) k_chunk_span(const int32_t* __restrict__ pos, const int32_t* __restrict__ meta, int64_t ns, int R,
                                                    int cap, int ps_mid, int ps_large, int tile_stride, ChunkInfo* __restrict__ info,
                                                    int* __restrict__ scr) {
  __shared__ int s_min, s_max, s_mixed, s_has[2], s_last;
  if (threadIdx.x == 0) { s_min = INT_MAX; s_max = INT_MIN; s_mixed = 0; s_has[0] = s_has[1] = 0; }
  __syncthreads();
  const int chrom0 = int(uint32_t(meta[0]) >> 8);
  int mn = INT_MAX, mx = INT_MIN, mixed = 0, h0 = 0, h1 = 0;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < ns; i += int64_t(gridDim.x) * blockDim.x) {
    const int p = pos[i], m = meta[i];
    mn = min(mn, p);
    mx = max(mx, p);
    mixed |= int(uint32_t(m) >> 8) != chrom0;
    if (m & 1) h1 = 1; else h0 = 1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mixed |= __shfl_xor_sync(0xffffffffu, mixed, o);
    h0 |= __shfl_xor_sync(0xffffffffu, h0, o);
    h1 |= __shfl_xor_sync(0xffffffffu, h1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_min, mn);
    atomicMax(&s_max, mx);
    if (mixed) s_mixed = 1;
    if (h0) s_has[0] = 1;
    if (h1) s_has[1] = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_min != INT_MAX) { atomicMax(&scr[0], INT_MAX - s_min); atomicMax(&scr[1], s_max + 1); }
    if (s_mixed) atomicOr(&scr[2], 1);
    if (s_has[0]) atomicOr(&scr[3], 1);
    if (s_has[1]) atomicOr(&scr[4], 1);
    __threadfence();
    s_last = atomicAdd(&scr[5], 1) == int(gridDim.x) - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  const int g_min = INT_MAX - atomicExch(&scr[0], 0), g_max = atomicExch(&scr[1], 0) - 1;
  const int g_mixed = atomicExch(&scr[2], 0), g_h0 = atomicExch(&scr[3], 0), g_h1 = atomicExch(&scr[4], 0);
  atomicExch(&scr[5], 0);
  const long long span = (long long)g_max - g_min + 2LL * R + 64;
  info->g_lo = (long long)g_min - R - 24;
  info->n_pos = int(span < cap ? span : cap);
  info->dense = (!g_mixed && span <= cap) ? 1 : 0;
  info->chrom = chrom0;
  info->has[0] = g_h0;
  info->has[1] = g_h1;
  for (int br = 0; br < 2; ++br) {
    const int ps = br ? ps_large : ps_mid;
    const int M = (info->n_pos + ps - 1) / ps;
    info->M[br] = M;
    info->lat_rows[br] = 2 * ps * (M + 1) + 1;
    info->lat_tiles[br] = (info->lat_rows[br] + tile_stride - 1) / tile_stride;
  }
}

struct DenseBranch {
  const float* T;     // [3][16][C] per-tap table
  const float* bias;  // [C]
  int w[3];           // sliding-window widths: full bin, first-bin interior, last-bin interior (0: unused)
};

// o[i] = max(v[i .. i+W-1]) for i < 16 by doubling: m_p[i] = max(m_{p/2}[i], m_{p/2}[i+p/2]), then two overlapping
// power-of-two windows.  max is exact, so the evaluation order does not matter for bit-identity with the per-site stem.
template <int W>
__device__ __forceinline__ void win_max16(const float (&v)[DT_RUN + DT_MAXW - 1], float (&o)[DT_RUN]) {
  constexpr int P = W >= 16 ? 16 : (W >= 8 ? 8 : (W >= 4 ? 4 : (W >= 2 ? 2 : 1)));
  float m[DT_RUN + DT_MAXW - 1];
#pragma unroll
  for (int i = 0; i < DT_RUN + DT_MAXW - 1; ++i) m[i] = v[i];
  constexpr int NV = DT_RUN + DT_MAXW - 1;
  // ascending i: m[i+p] is still the previous level when m[i] is updated
  if constexpr (P >= 2) {
#pragma unroll
    for (int i = 0; i + 1 < NV; ++i) m[i] = fmaxf(m[i], m[i + 1]);
  }
  if constexpr (P >= 4) {
#pragma unroll
    for (int i = 0; i + 2 < NV; ++i) m[i] = fmaxf(m[i], m[i + 2]);
  }
  if constexpr (P >= 8) {
#pragma unroll
    for (int i = 0; i + 4 < NV; ++i) m[i] = fmaxf(m[i], m[i + 4]);
  }
  if constexpr (P >= 16) {
#pragma unroll
    for (int i = 0; i + 8 < NV; ++i) m[i] = fmaxf(m[i], m[i + 8]);
  }
#pragma unroll
  for (int i = 0; i < DT_RUN; ++i) o[i] = fmaxf(m[i], m[i + W - P]);
}

__device__ __forceinline__ void win_max_any(int w, const float (&v)[DT_RUN + DT_MAXW - 1], float (&o)[DT_RUN]) {
  switch (w) {
    case 1: win_max16<1>(v, o); break;
    case 2: win_max16<2>(v, o); break;
    case 3: win_max16<3>(v, o); break;
    case 4: win_max16<4>(v, o); break;
    case 5: win_max16<5>(v, o); break;
    case 6: win_max16<6>(v, o); break;
    case 7: win_max16<7>(v, o); break;
    case 8: win_max16<8>(v, o); break;
    case 9: win_max16<9>(v, o); break;
    case 10: win_max16<10>(v, o); break;
    case 11: win_max16<11>(v, o); break;
    case 12: win_max16<12>(v, o); break;
    case 13: win_max16<13>(v, o); break;
    case 14: win_max16<14>(v, o); break;
    case 15: win_max16<15>(v, o); break;
    default: win_max16<16>(v, o); break;
  }
}

// tables: [strand][branch][3 widths][cap][C] bf16
// One thread owns a channel pair and a run of DT_RUN positions: the DT_RUN + 15 conv values it needs stay in registers
// (3 table rows summed per position, same order as the per-site stem), the window maxima are computed with static
// indexing per width, and every store is one packed bf16 pair.
template <int C>
#ifndef MURAL_DT_MINB
#define MURAL_DT_MINB 2
#endif
__global__ void __launch_bounds__(256, MURAL_DT_MINB) k_dense_tables(GenomeView G, const ChunkInfo* __restrict__ info, DenseBranch b0, DenseBranch b1,
                                                      int cap, __nv_bfloat16* __restrict__ tables) {
  static_assert(C == 32, "channel-pair mapping assumes 16 lanes x 2 channels");
  if (!info->dense) return;
  const int p0 = blockIdx.x * DT_POS;
  if (p0 >= info->n_pos) return;
  __shared__ uint8_t sym[DT_POS + DT_MAXW + 4];
  extern __shared__ __align__(16) float dsm[];
  float* sT = dsm;                       // [2][3][16][C]
  float* sB = sT + 2 * 3 * 16 * C;       // [2][C]
  const int tid = threadIdx.x;
  const int chrom = info->chrom;
  const long long g0 = info->g_lo + p0;
  for (int e = tid; e < 3 * 16 * C; e += 256) { sT[e] = b0.T[e]; sT[3 * 16 * C + e] = b1.T[e]; }
  for (int e = tid; e < C; e += 256) { sB[e] = b0.bias[e]; sB[C + e] = b1.bias[e]; }
  for (int k = tid; k < DT_POS + DT_MAXW + 2; k += 256) sym[k] = uint8_t(sym_genomic(G, chrom, g0 - 1 + k));  // sym[k] = base g0-1+k
  __syncthreads();
  const int c2 = (tid & 15) * 2;         // channels c2, c2+1
  const int k0 = (tid >> 4) * DT_RUN;    // first position of this thread's run (16 runs x DT_RUN = DT_POS)
  const int n_pos = info->n_pos;
#pragma unroll 1
  for (int strand = 0; strand < 2; ++strand) {
    if (!info->has[strand]) continue;
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
      const int bw[3] = {br ? b1.w[0] : b0.w[0], br ? b1.w[1] : b0.w[1], br ? b1.w[2] : b0.w[2]};
      const float* T = sT + br * 3 * 16 * C + c2;
      const float2 bias = *reinterpret_cast<const float2*>(sB + br * C + c2);
      float v0[DT_RUN + DT_MAXW - 1], v1[DT_RUN + DT_MAXW - 1];
      // conv output of every position in the oriented sequence of this strand (tap order as in the per-site stem)
#pragma unroll
      for (int i = 0; i < DT_RUN + DT_MAXW - 1; ++i) {
        const int k = k0 + i;
        int s0 = sym[k], s1 = sym[k + 1], s2 = sym[k + 2];
        if (strand) {  // oriented neighbours of genomic g are comp(g+1), comp(g), comp(g-1)
          const int t = comp_sym(s0);
          s0 = comp_sym(s2); s1 = comp_sym(s1); s2 = t;
        }
        const float2 t0 = *reinterpret_cast<const float2*>(T + (0 * 16 + s0) * C);
        const float2 t1 = *reinterpret_cast<const float2*>(T + (1 * 16 + s1) * C);
        const float2 t2 = *reinterpret_cast<const float2*>(T + (2 * 16 + s2) * C);
        v0[i] = ((bias.x + t0.x) + t1.x) + t2.x;
        v1[i] = ((bias.y + t0.y) + t1.y) + t2.y;
      }
      __nv_bfloat16* tb = tables + (size_t(strand * 2 + br) * 3) * size_t(cap) * C;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int w = bw[t];
        if (w <= 0) continue;
        float o0[DT_RUN], o1[DT_RUN];
        win_max_any(w, v0, o0);
        win_max_any(w, v1, o1);
#pragma unroll
        for (int i = 0; i < DT_RUN; ++i)
          if (p0 + k0 + i < n_pos)
            *reinterpret_cast<__nv_bfloat162*>(tb + (size_t(t) * cap + p0 + k0 + i) * C + c2) = __floats2bfloat162_rn(o0[i], o1[i]);
      }
    }
  }
}

struct GatherBranch {
  void* out;           // bf16 planes [C/8][rows_alloc][8]
  int64_t rows_alloc;
  const float* T;      // per-tap table (edge positions)
  const float* bias;
  int L0, off0, L1, pk, ps, pp;
  int w[3];
  int edge;            // 1: only the first and last row of a site's edge pseudo-site (the stage kernel reads the 16 rows in
                       //    between straight from the tables), written as row 2*site + (last ? 1 : 0)
};

// one warp-quarter (8 lanes x 16 B... here: C/8 lanes, 16 B each) copies one table row into one output row
template <int C>
__global__ void __launch_bounds__(256) k_stem_gather(GenomeView G, const ChunkInfo* __restrict__ info, const int32_t* __restrict__ pos,
                                                     const int32_t* __restrict__ meta, int64_t ns, int R, GatherBranch b0, GatherBranch b1,
                                                     int cap, const __nv_bfloat16* __restrict__ tables) {
  if (!info->dense) return;
  constexpr int PL = C / 8;  // 16-byte chunks per row
  const long long g_lo = info->g_lo;
  const int chrom = info->chrom;
#pragma unroll 1
  for (int br = 0; br < 2; ++br) {
    const GatherBranch& B = br ? b1 : b0;
    const int rows_per_site = B.edge ? LAT_EL : B.L1;
    // two passes so that the lanes of a warp do the same kind of work: pass 0 = the rows strictly inside the (pseudo-)site,
    // plain 64-byte row copies; pass 1 = its first and last row, which carry the exactly evaluated window-edge positions
    // (symbol lookups + 3 table rows per channel, ~10x the work of a copy)
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
    const int rows_pass = pass ? (rows_per_site > 1 ? 2 : 1) : ((rows_per_site > 2 && !B.edge) ? rows_per_site - 2 : 0);
    const uint32_t total = uint32_t(ns) * uint32_t(rows_pass) * PL;  // < 2^31 (host-checked): 32-bit index arithmetic only
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int q = int(e % PL);
      const uint32_t sj = e / PL;
      const uint32_t site = sj / uint32_t(rows_pass);
      const int jr = int(sj - site * uint32_t(rows_pass));
      const int jj = pass ? (jr ? rows_per_site - 1 : 0) : jr + 1;
      const int j = (B.edge && jj >= LAT_EI) ? B.L1 - LAT_EL + jj : jj;
      const int s = pos[site], strand = meta[site] & 1;
      int lo = j * B.ps - B.pp, hi = lo + B.pk;
      lo = lo < 0 ? 0 : lo;
      hi = hi > B.L0 ? B.L0 : hi;
      const int ilo = lo < 1 ? 1 : lo, ihi = hi > B.L0 - 1 ? B.L0 - 1 : hi;  // interior positions [ilo, ihi)
      const int wn = ihi - ilo;
      uint4 v = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);  // bf16 -inf
      if (wn > 0) {
        const int t = wn == B.w[0] ? 0 : (wn == B.w[1] ? 1 : 2);
        // oriented window index i_w = off0 + i  <->  genomic  s - R + i_w ('+')  /  s + R - i_w ('-')
        const long long gs = strand ? (long long)s + R - (B.off0 + ihi - 1) : (long long)s - R + B.off0 + ilo;
        const __nv_bfloat16* row = tables + ((size_t(strand * 2 + br) * 3 + t) * size_t(cap) + size_t(gs - g_lo)) * C;
        v = *reinterpret_cast<const uint4*>(row + 8 * q);
      }
      if (lo == 0 || hi == B.L0) {  // window-edge position(s): the tap outside the window contributes 0 (zero padding)
        float ev[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) ev[c] = -FLT_MAX;
        for (int side = 0; side < 2; ++side) {
          const int p = side ? B.L0 - 1 : 0;
          if (side ? (hi != B.L0) : (lo != 0)) continue;
          if (side && B.L0 - 1 == 0) continue;
          float a[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) a[c] = B.bias[8 * q + c];
          for (int t = 0; t < 3; ++t) {
            const int x = p + t - 1;
            if (x < 0 || x >= B.L0) continue;
            const int iw = B.off0 + x;
            const long long g = strand ? (long long)s + R - iw : (long long)s - R + iw;
            int sy = sym_genomic(G, chrom, g);
            if (strand) sy = comp_sym(sy);
#pragma unroll
            for (int c = 0; c < 8; ++c) a[c] += B.T[(t * 16 + sy) * C + 8 * q + c];
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) ev[c] = fmaxf(ev[c], a[c]);
        }
        uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          __nv_bfloat162 e2 = __floats2bfloat162_rn(ev[2 * c], ev[2 * c + 1]);
          __nv_bfloat162 m2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&w[c]), e2);
          w[c] = *reinterpret_cast<uint32_t*>(&m2);
        }
      }
      const int64_t orow = B.edge ? 2 * int64_t(site) + (jj ? 1 : 0) : 1 + site * int64_t(rows_per_site + 1) + jj;
      *(reinterpret_cast<uint4*>(B.out) + int64_t(q) * B.rows_alloc + orow) = v;
    }
    }
  }
}

static void bin_widths(const BranchDev& B, int w[3]) {
  // interior positions (conv with all three taps inside the window) of a full bin, the first bin and the last bin
  w[0] = B.pool[0][0];
  int lo = 0, hi = B.pool[0][0] - B.pool[0][2];
  hi = hi > B.L0 ? B.L0 : hi;
  w[1] = (hi > B.L0 - 1 ? B.L0 - 1 : hi) - (lo < 1 ? 1 : lo);
  lo = (B.L1 - 1) * B.pool[0][1] - B.pool[0][2];
  hi = lo + B.pool[0][0];
  lo = lo < 0 ? 0 : lo;
  hi = hi > B.L0 ? B.L0 : hi;
  w[2] = (hi > B.L0 - 1 ? B.L0 - 1 : hi) - (lo < 1 ? 1 : lo);
  for (int t = 1; t < 3; ++t)
    if (w[t] < 0) w[t] = 0;
}

int64_t snv_dense_cap(int64_t chunk) { return 8 * chunk + 4096; }

size_t snv_dense_bytes(const mural_snv_model* m, int64_t chunk) {
  const size_t cap = size_t(snv_dense_cap(chunk));
  return 256 + size_t(12) * cap * m->cfg.channels * 2;
}

// The lattice needs: both branches long enough to have interior rows, and the bins of those rows to be full-width
// bins made only of interior conv positions (so that they equal the full-bin sliding maximum), and the lattice
// neighbourhood of every interior row to lie inside the table span [s_min - R - 24, s_max + R + 39].
bool snv_lattice_supported(const mural_snv_model* m) {
  if (m->cfg.channels != 32 || m->cfg.kernel_size != 3) return false;
  const int R = m->cfg.distal_radius;
  for (int br = 0; br < 2; ++br) {
    const BranchDev& B = m->br[br];
    const int pk = B.pool[0][0], ps = B.pool[0][1], pp = B.pool[0][2], off0 = br ? 0 : m->L / 2 - 100;
    if (B.L1 < 2 * LAT_EI + 6 || pk > DT_MAXW) return false;
    // rows [LAT_EO-4, L1-LAT_EO+3] are read by the lattice chain of the interior rows [LAT_EO, L1-LAT_EO-1]
    const int r_lo = LAT_EO - 4, r_hi = B.L1 - LAT_EO + 3;
    if (r_lo * ps - pp < 1 || r_hi * ps - pp + pk > B.L0 - 1) return false;
    if (off0 + r_lo * ps - pp + 24 < 0 || off0 + r_hi * ps - pp + pk - 1 > 2 * R + 24 + 39) return false;
    if (2 * R + 24 - off0 - (r_hi * ps - pp) - pk + 1 < 0) return false;
  }
  return true;
}

// Enqueues span detection, table build and gather.  d_scratch: snv_dense_bytes() bytes.  Returns the device flag
// (int*, 1 = the dense path produced the stem output) so that the per-site kernel can skip itself.
int snv_dense_stem_launch(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta, int64_t ns,
                          int64_t chunk, void* mid_out, int64_t mid_ra, void* large_out, int64_t large_ra, void* d_scratch,
                          const int** d_flag, cudaStream_t st, const LatticeBufs* lattice, const ChunkInfo** d_info) {
  const int C = m->cfg.channels;
  MURAL_CHECK(C == 32, "dense stem is built for C == 32");
  const int cap = int(snv_dense_cap(chunk));
  ChunkInfo* info = reinterpret_cast<ChunkInfo*>(d_scratch);
  __nv_bfloat16* tables = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<char*>(d_scratch) + 256);
  DenseBranch db[2];
  GatherBranch gb[2];
  for (int br = 0; br < 2; ++br) {
    const BranchDev& B = m->br[br];
    db[br].T = B.T;
    db[br].bias = B.bias1;
    bin_widths(B, db[br].w);
    for (int t = 0; t < 3; ++t) MURAL_CHECK(db[br].w[t] <= DT_MAXW, "pool window too wide for the dense stem");
    gb[br] = GatherBranch{br ? large_out : mid_out, br ? large_ra : mid_ra, B.T, B.bias1, B.L0, br ? 0 : m->L / 2 - 100, B.L1,
                          B.pool[0][0], B.pool[0][1], B.pool[0][2], {db[br].w[0], db[br].w[1], db[br].w[2]}, 0};
    if (lattice) {
      gb[br].out = lattice[br].edge_in;
      gb[br].rows_alloc = lattice[br].edge_ra;
      gb[br].edge = 1;
    }
  }
  int* scr = reinterpret_cast<int*>(reinterpret_cast<char*>(d_scratch) + 128);  // zeroed once per forward call by the caller
  LAUNCH(k_chunk_span, 296, 256, 0, st, d_pos, d_meta, ns, m->cfg.distal_radius, cap, m->br[0].pool[0][1], m->br[1].pool[0][1],
         128 - 2 * 4, info, scr);
  const size_t smem = sizeof(float) * (2 * 3 * 16 * C + 2 * C);
  static bool conf = false;
  if (!conf) {
    CUDA_TRY(cudaFuncSetAttribute(k_dense_tables<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = true;
  }
  LAUNCH(k_dense_tables<32>, (unsigned)cdiv(cap, DT_POS), 256, smem, st, *G, info, db[0], db[1], cap, tables);
  MURAL_CHECK(ns * int64_t(m->br[1].L1 > LAT_EL ? m->br[1].L1 : LAT_EL) * 4 < (int64_t(1) << 31), "chunk too large for the stem gather's 32-bit indices");
  LAUNCH(k_stem_gather<32>, 148 * 8, 256, 0, st, *G, info, d_pos, d_meta, ns, m->cfg.distal_radius, gb[0], gb[1], cap, tables);
  *d_flag = &info->dense;
  if (d_info) *d_info = info;
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace mural
