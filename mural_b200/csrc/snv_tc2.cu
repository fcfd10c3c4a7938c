// Stage kernels of the bf16 tcgen05 path, second generation: TWO rows per TMEM lane.
//
// The first-generation kernel (snv_tc.cu) maps one row of the row space to one TMEM lane and evaluates a 3-tap conv as
// 3 taps x 2 K-halves of M=128 N=32 K=16 MMAs whose A operand costs 40 cycles each against 16 of math, and it pays one
// commit / barrier / epilogue round per 128 rows and layer (profiles/r01_stage_tc_phase_timing.txt).  Here a lane owns
// the row pair (2m, 2m+1) of a 256-row tile and one MMA chain produces both:
//     D[m, (e, co)] = sum_{j=0..3} sum_ci  X[2m - 1 + j][ci] * Wp[(j, ci), (e, co)],   Wp = W[tap = j - e] (0 otherwise)
// i.e. K = 4 x 32, N = 2 x 32: 8 MMAs of M=128 N=64 K=16 (48 cycles each, 75 % useful MACs) + one constant-column MMA
// for bias / site-edge corrections = 432 tensor cycles per 256 rows (216 per 128 against 280), and HALF the number of
// synchronisation rounds per row.  The activations live in shared memory split by row parity (even rows / odd rows, each
// as K-major SWIZZLE_NONE planes), so that "row 2m - 1 + j" is again a plain 16-byte-pitch operand:
//     j=0: odd[m-1]   j=1: even[m]   j=2: odd[m]   j=3: even[m+1]
// Everything else (row space with zero separators, residual stream in TMEM, BN folding with hi/lo bias and edge
// corrections, loaders for per-site / lattice / edge / pre-pooled rows) is the arithmetic of snv_tc.cu, and the results
// are bit-identical to it: the same bf16 products are accumulated in fp32 in the same K order per output.
// Reference arithmetic: MuRaL/model/model_snv.py:475-488 / 497-510 and ResBlock :794-812.
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "snv_tc_args.cuh"

namespace mural {
namespace tc2 {

using tc::C_RB4;
using tc::RB4;
using tc::StageArgs;

constexpr int TILE = 256;                 // rows per tile
constexpr int HALF = 128;                 // lanes = row pairs
constexpr int PLANE = (HALF + 1) * 16;    // bytes of one (parity, 8-channel) plane: 128 entries + 1 halo entry
constexpr int PAR_BYTES = 4 * PLANE;      // one parity: 4 planes
constexpr int A_BYTES = 2 * PAR_BYTES;    // [odd planes: stored index i = entry i-1][even planes: stored index i = entry i]
constexpr int AC_BYTES = 2 * HALF * 16;   // constant-column operand: [2 core-matrix columns][128 lanes][16 B]
constexpr int SLOT_BYTES = A_BYTES + AC_BYTES;
constexpr int W_CONV = 16 * 64 * 16;      // bf16 [k/8 = (j, ci/8)][n = (e, co)][8] = 16384 bytes
constexpr int W_BIAS = 2 * 64 * 16;       // bf16 [k/8][n][8] of the constant-column MMA = 2048 bytes
constexpr int W_LAYER = W_CONV + W_BIAS;
constexpr int NSLOT = 4;                  // one slot per thread group: 4 x (R, T) regions of 64 TMEM columns = 512
constexpr int THREADS = NSLOT * HALF;
static_assert(A_BYTES % 128 == 0, "slot alignment");
// instruction descriptor, kind::f16: D=F32, A=B=BF16, K-major, N=64, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__host__ __device__ constexpr int n_layers(int mode) { return mode == RB4 ? 4 : 5; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t w) {
  __nv_bfloat162 v = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&w), __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 v = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

#define TMEM_LD32(r, taddr)                                                                                          \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"                    \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
      : "r"(taddr)                                                                                                   \
      : "memory")
#define TMEM_ST32(taddr, r)                                                                                          \
  asm volatile(                                                                                                      \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                     \
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),             \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),  \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),    \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),    \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                 \
      : "memory")

// lattice row of stage-1 row 0 of a site (row rr is lat_base + rr for either strand; see snv_dense_stem.cu)
__device__ __forceinline__ int lattice_base(const ChunkInfo* info, int br, int s, int strand, int R, int off0, int ps1, int pp1, int pk1) {
  const int M = info->M[br];
  const int g_lo = int(info->g_lo);
  const int x0 = strand ? s + R - off0 + pp1 - pk1 + 1 - g_lo : s - R + off0 - pp1 - g_lo;
  const int m0 = x0 / ps1, phase = x0 - m0 * ps1;
  return strand ? 1 + (ps1 + phase) * (M + 1) + (M - 1 - m0) : 1 + phase * (M + 1) + m0;
}

// One stage of one branch, persistent: one CTA per SM, 4 thread groups, one 256-row tile in flight per group.
// FM: 0 plain rows, 1 lattice (RB4: device-side geometry; C_RB4: pre-pooled single rows), 2 edge gather (RB4).
template <int MODE, int FM>
__global__ void __launch_bounds__(THREADS, 1) k_stage2_tc(StageArgs a) {
  constexpr bool LAT = FM == 1;
  constexpr bool EDGE = (MODE == RB4) && FM == 2;
  constexpr bool DYN = (MODE == RB4) && LAT;
  if (a.info && a.info->dense != a.want) return;  // uniform over the grid; nothing allocated yet
  constexpr int NL = n_layers(MODE);
  constexpr int STRIDE = TILE - 2 * NL;  // valid output rows per tile (the chain eats NL rows on each side)
  const int L_ = DYN ? a.info->M[a.lat_branch] : a.L;
  const int rows_ = DYN ? a.info->lat_rows[a.lat_branch] : int(a.rows);
  const int n_tiles_ = DYN ? (rows_ + STRIDE - 1) / STRIDE : a.n_tiles;
  if (int(blockIdx.x) * NSLOT >= n_tiles_) return;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sW = smem;                              // [NL][W_LAYER], shared by all slots
  unsigned char* sSlots = sW + NL * W_LAYER;             // [NSLOT][SLOT_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSlots + NSLOT * SLOT_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NSLOT);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = tid >> 7, lt = tid & 127;

  // ---- one-time setup
  {
    constexpr int NV = NL * W_LAYER / 16;
    const uint4* src = reinterpret_cast<const uint4*>(a.wblob);
    uint4* dst = reinterpret_cast<uint4*>(sW);
    for (int e = tid; e < NV; e += THREADS) dst[e] = __ldg(src + e);
  }
  for (int e = tid; e < NSLOT * 8; e += THREADS) {  // halo entries: odd stored index 0 (row -1), even stored index 128 (row 256)
    const int sl = e >> 3, par = (e >> 2) & 1, plane = e & 3;
    *reinterpret_cast<uint4*>(sSlots + sl * SLOT_BYTES + par * PAR_BYTES + plane * PLANE + (par ? HALF * 16 : 0)) = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
  const int Lp1 = L_ + 1;
  const int tile_step = gridDim.x * NSLOT;
  unsigned char* const sA = sSlots + g * SLOT_BYTES;        // this group's slot: [odd planes][even planes][const operand]
  const uint32_t bar = smem_u32(bars + g);
  const uint32_t dreg = tmem_base + g * 128;                 // R = columns [0, 64), T = [64, 128) of the slot
  const uint64_t dW0 = umma_desc(smem_u32(sW), 64 * 16, 128);
  const uint64_t dA0 = umma_desc(smem_u32(sA), PLANE, 128);
  const uint64_t dC0 = umma_desc(smem_u32(sA) + A_BYTES, HALF * 16, 128);
  uint32_t phase = 0;

  // publish the slot's operands to the tensor core: three warps arrive and run ahead, the issuing warp (rotating with
  // the layer) waits for all 128 threads and its elected lane issues the 9 MMAs of layer l, then commits.
  auto sync_and_issue = [&](int l) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if ((lt >> 5) != ((l + g) & 3)) {
      asm volatile("bar.arrive %0, 128;" ::"r"(1 + g) : "memory");
    } else {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if ((lt & 31) == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // R-type layers (second conv of a ResBlock, conv2) accumulate into / create region R, others use T
        const bool rtype = (MODE == RB4) ? (l & 1) : !(l & 1);
        const bool acc_first = (MODE == RB4) ? rtype : (rtype && l > 0);
        const uint32_t d = dreg + (rtype ? 0 : 64);
        const uint64_t dW = dW0 + uint64_t((l * W_LAYER) >> 4);
        umma_bf16(d, dC0, dW + (W_CONV >> 4), acc_first ? 1u : 0u);
#pragma unroll
        for (int j = 0; j < 4; ++j)  // X[2m - 1 + j]: odd[m-1], even[m], odd[m], even[m+1]
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int a_off = ((j & 1) ? PAR_BYTES : 0) + 2 * h * PLANE + ((j >> 1) ? 16 : 0);
            umma_bf16(d, dA0 + uint64_t(a_off >> 4), dW + uint64_t(((j * 4 + 2 * h) * 64 * 16) >> 4), 1u);
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      }
      __syncwarp();
    }
  };

  // position of row r inside its site (-1 = separator / outside) and the chain input of that row
  auto fetch_row = [&](int r, int& p, uint4 (&x)[4]) {
    p = -1;
    int site = 0;
    if (r > 0 && r < rows_) {
      site = r / Lp1;
      p = r - site * Lp1 - 1;
    }
    const bool live = p >= 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) x[q] = make_uint4(0, 0, 0, 0);
    if (!live) return;
    if (EDGE) {
      if (p == 0 || p == LAT_EL - 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = __ldg(a.special + q * a.special_ra + 2 * int64_t(site) + (p ? 1 : 0));
      } else {
        const int s = __ldg(a.pos + site), strand = __ldg(a.meta + site) & 1;
        const int j = p < LAT_EI ? p : a.L1real - LAT_EL + p;
        const int lo = j * a.ps1 - a.pp1, g_lo = int(a.info->g_lo);
        const int idx = strand ? s + a.R - a.off0 - lo - a.pk1 + 1 - g_lo : s - a.R + a.off0 + lo - g_lo;
        const uint4* row = a.tab[strand] + int64_t(idx) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = __ldg(row + q);
      }
    } else if (MODE == RB4) {
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = __ldg(a.in + q * a.in_rows_alloc + r);
    } else if (LAT) {
      // one pre-pooled row per bin: k_lattice_pool (bins on the lattice) or k_edge_pool (bins touching edge rows)
      if (p >= a.nlo && p < L_ - a.nhi) {
        const int s = __ldg(a.pos + site), strand = __ldg(a.meta + site) & 1;
        const int lat_base = lattice_base(a.info, a.br, s, strand, a.R, a.off0, a.ps1, a.pp1, a.pk1);
        const int lo = p * a.ps - a.pp;
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = __ldg(a.lat2 + q * a.lat_ra + lat_base + lo);
      } else {
        const int64_t idx = int64_t(site) * (a.nlo + a.nhi) + (p < a.nlo ? p : p - (L_ - a.nhi) + a.nlo);
#pragma unroll
        for (int q = 0; q < 4; ++q) x[q] = __ldg(a.epool + q * a.epool_ra + idx);
      }
    } else {  // max-pool fused into the loader (model_snv.py:361,404), -inf padding
      int lo = p * a.ps - a.pp, hi = lo + a.pk;
      lo = lo < 0 ? 0 : lo;
      hi = hi > a.Lin ? a.Lin : hi;
      const int base = 1 + site * (a.Lin + 1);
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
#pragma unroll
      for (int u = 0; u < 7; ++u) {  // pk <= 7 for every pool of Network2
        if (lo + u < hi) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = __ldg(a.in + q * a.in_rows_alloc + base + lo + u);
            x[q].x = max_bf16x2(x[q].x, v.x); x[q].y = max_bf16x2(x[q].y, v.y);
            x[q].z = max_bf16x2(x[q].z, v.z); x[q].w = max_bf16x2(x[q].w, v.w);
          }
        }
      }
    }
  };

  // first A operand / constant operand / residual region from the fetched row pair, then layer 0
  auto begin_tile = [&](int r0, const int (&p)[2], const uint4 (&x)[2][4]) {
    uint32_t cw[8];  // constant-column operand row: per row e {live, live, pos==0, pos==0, pos==L-1, pos==L-1} as bf16 1.0 / 0
    const uint32_t one2 = 0x3F803F80u;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const bool live = p[e] >= 0;
      const int i = 2 * lt + e;                      // row index inside the tile
      unsigned char* dstp = sA + (e ? 0 : PAR_BYTES) + (lt + (e ? 1 : 0)) * 16;  // even rows -> even planes, odd rows -> odd planes (+1 halo)
      if (EDGE && live && i >= NL && i < TILE - NL) {  // gathered input row, needed again for the outer skip: parked in the output row
        uint4* out4 = reinterpret_cast<uint4*>(a.out);
#pragma unroll
        for (int q = 0; q < 4; ++q) out4[q * a.out_rows_alloc + r0 + e] = x[e][q];
      }
      if (MODE == RB4) {
        uint32_t f[32];  // residual region R <- x0 (fp32)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t* w = reinterpret_cast<const uint32_t*>(&x[e][q]);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            f[8 * q + 2 * jj] = w[jj] << 16;
            f[8 * q + 2 * jj + 1] = w[jj] & 0xFFFF0000u;
          }
        }
        TMEM_ST32(dreg + lane_off + 32 * e, f);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = relu_bf16x2(x[e][q].x); o.y = relu_bf16x2(x[e][q].y); o.z = relu_bf16x2(x[e][q].z); o.w = relu_bf16x2(x[e][q].w);
          *reinterpret_cast<uint4*>(dstp + q * PLANE) = o;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dstp + q * PLANE) = x[e][q];
      }
      cw[3 * e] = live ? one2 : 0u;
      cw[3 * e + 1] = p[e] == 0 ? one2 : 0u;
      cw[3 * e + 2] = (live && p[e] == L_ - 1) ? one2 : 0u;
    }
    cw[6] = 0u; cw[7] = 0u;
    if (MODE == RB4) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    *reinterpret_cast<uint4*>(sA + A_BYTES + lt * 16) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
    *reinterpret_cast<uint4*>(sA + A_BYTES + HALF * 16 + lt * 16) = make_uint4(cw[4], cw[5], cw[6], cw[7]);
    sync_and_issue(0);
  };

  int tile = blockIdx.x * NSLOT + g;
  int pn[2];
  uint4 xn[2][4];
  pn[0] = pn[1] = -1;
  if (tile < n_tiles_) {
    const int rb = tile * STRIDE - NL + 2 * lt;
    fetch_row(rb, pn[0], xn[0]);
    fetch_row(rb + 1, pn[1], xn[1]);
  }
  for (; tile < n_tiles_; tile += tile_step) {
    const int r0 = tile * STRIDE - NL + 2 * lt;  // rows r0 (even slot) and r0 + 1 (odd slot) of this lane
    int p[2] = {pn[0], pn[1]};
    begin_tile(r0, p, xn);
    if (tile + tile_step < n_tiles_) {
      const int rb = (tile + tile_step) * STRIDE - NL + 2 * lt;
      fetch_row(rb, pn[0], xn[0]);
      fetch_row(rb + 1, pn[1], xn[1]);
    }
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      mbar_wait(bar, phase);
      phase ^= 1;
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool rtype = (MODE == RB4) ? (l & 1) : !(l & 1);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool live = p[e] >= 0;
        const int i = 2 * lt + e;
        const bool valid = live && i >= NL && i < TILE - NL;
        const int r = r0 + e;
        uint32_t acc[32];
        TMEM_LD32(acc, dreg + lane_off + (rtype ? 0 : 64) + 32 * e);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (l < NL - 1) {
          uint4 o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[q].x = pack_bf16(__uint_as_float(acc[8 * q]), __uint_as_float(acc[8 * q + 1]));
            o[q].y = pack_bf16(__uint_as_float(acc[8 * q + 2]), __uint_as_float(acc[8 * q + 3]));
            o[q].z = pack_bf16(__uint_as_float(acc[8 * q + 4]), __uint_as_float(acc[8 * q + 5]));
            o[q].w = pack_bf16(__uint_as_float(acc[8 * q + 6]), __uint_as_float(acc[8 * q + 7]));
          }
          if (MODE == C_RB4 && l == 0 && valid) {  // jump = conv2 output: parked (bf16) in the output row, re-read at the end
            uint4* out4 = reinterpret_cast<uint4*>(a.out);
#pragma unroll
            for (int q = 0; q < 4; ++q) out4[q * a.out_rows_alloc + r] = o[q];
          }
          if (live) {  // separator rows were zeroed by begin_tile and are never rewritten
            unsigned char* dstp = sA + (e ? 0 : PAR_BYTES) + (lt + (e ? 1 : 0)) * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = relu_bf16x2(o[q].x); w.y = relu_bf16x2(o[q].y); w.z = relu_bf16x2(o[q].z); w.w = relu_bf16x2(o[q].w);
              *reinterpret_cast<uint4*>(dstp + q * PLANE) = w;
            }
          }
        } else if (valid) {  // outer skip: + x0 (RB4, re-read from the input) or + jump / parked x0 (C_RB4, EDGE)
          uint4* out4 = reinterpret_cast<uint4*>(a.out);
          uint4 xr[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            xr[q] = (MODE == RB4 && !EDGE) ? __ldg(a.in + q * a.in_rows_alloc + r) : out4[q * a.out_rows_alloc + r];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(acc[8 * q]) + bf16_lo(xr[q].x), __uint_as_float(acc[8 * q + 1]) + bf16_hi(xr[q].x));
            o.y = pack_bf16(__uint_as_float(acc[8 * q + 2]) + bf16_lo(xr[q].y), __uint_as_float(acc[8 * q + 3]) + bf16_hi(xr[q].y));
            o.z = pack_bf16(__uint_as_float(acc[8 * q + 4]) + bf16_lo(xr[q].z), __uint_as_float(acc[8 * q + 5]) + bf16_hi(xr[q].z));
            o.w = pack_bf16(__uint_as_float(acc[8 * q + 6]) + bf16_lo(xr[q].w), __uint_as_float(acc[8 * q + 7]) + bf16_hi(xr[q].w));
            out4[q * a.out_rows_alloc + r] = o;
          }
        }
      }
      if (l < NL - 1) sync_and_issue(l + 1);
    }
  }
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

struct State {
  uint8_t* d_w = nullptr;
  const uint8_t* blob[2][2] = {};  // [branch][stage 0 (RB4), stage 1 (C_RB4)]
};

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int MODE, int FM>
static int launch(const StageArgs& a, cudaStream_t st, const char* role) {
  constexpr int NL = n_layers(MODE);
  const size_t smem = size_t(NL) * W_LAYER + size_t(NSLOT) * SLOT_BYTES + NSLOT * 8 + 16;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY((cudaFuncSetAttribute(k_stage2_tc<MODE, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    configured = true;
  }
  int grid = (a.n_tiles + NSLOT - 1) / NSLOT;
  if (grid > sm_count()) grid = sm_count();
  if (grid < 1) grid = 1;
  static std::map<std::string, std::string> names;  // profile names are interned per (mode, role): prof_pre keeps the pointer
  const std::string key = std::string(MODE == RB4 ? "k_stage2_tc<RB4>" : "k_stage2_tc<C_RB4>") + role;
  const char* nm = names.emplace(key, key).first->second.c_str();
  LAUNCH_N(nm, (k_stage2_tc<MODE, FM>), grid, THREADS, smem, st, a);
  return 0;
}

}  // namespace tc2

int snv_tc2_stride(int mode) { return tc2::TILE - 2 * tc2::n_layers(mode); }

const uint8_t* snv_tc2_blob(const mural_snv_model* m, int br, int stage) {
  return m->tc2 ? ((const tc2::State*)m->tc2)->blob[br][stage] : nullptr;
}

int snv_tc2_launch(int mode, int fm, const tc::StageArgs& a, cudaStream_t st, const char* role) {
  using namespace tc2;
  if (mode == RB4 && fm == 0) return launch<RB4, 0>(a, st, role);
  if (mode == RB4 && fm == 1) return launch<RB4, 1>(a, st, role);
  if (mode == RB4 && fm == 2) return launch<RB4, 2>(a, st, role);
  if (mode == C_RB4 && fm == 0) return launch<C_RB4, 0>(a, st, role);
  if (mode == C_RB4 && fm == 1) return launch<C_RB4, 1>(a, st, role);
  return -1;
}

void snv_tc2_destroy(mural_snv_model* m) {
  if (!m->tc2) return;
  tc2::State* S = (tc2::State*)m->tc2;
  cudaFree(S->d_w);
  delete S;
  m->tc2 = nullptr;
}

// Row-pair weight blobs: per layer  conv [k/8 = (j, ci/8)][n = (e, co)][8]  with Wp = bf16(W[co][ci][j - e] * a[ci]) for
// 0 <= j - e <= 2 and 0 elsewhere, then the constant-column operand [k/8][n][8]: k 0..5 feed row e=0, k 6..11 row e=1 with
// {bias_hi, bias_lo, -e_left_hi, -e_left_lo, -e_right_hi, -e_right_lo} exactly as snv_tc.cu splits them.
int snv_tc2_prepare(mural_snv_model* m, const float* h_blob) {
  snv_tc2_destroy(m);
  if (m->cfg.channels != 32 || m->cfg.kernel_size != 3) return 0;
  using namespace tc2;
  auto T = [&](const std::string& n) { return h_blob + m->layout[m->index.at(n)].offset; };
  std::vector<uint8_t> all;
  size_t offs[2][2];
  for (int br = 0; br < 2; ++br) {
    const std::string s = br ? "_2" : "";
    std::vector<std::pair<std::string, std::string>> chains[2];  // (bn, conv) per layer
    for (int gg = 1; gg <= 2; ++gg) {
      if (gg == 2) chains[1].push_back({"conv2" + s + ".0", "conv2" + s + ".1"});
      for (int i = 0; i < 2; ++i) {
        const std::string p = "RBs" + std::to_string(gg) + s + "." + std::to_string(i);
        chains[gg - 1].push_back({p + ".bn1", p + ".conv1"});
        chains[gg - 1].push_back({p + ".bn2", p + ".conv2"});
      }
    }
    for (int stg = 0; stg < 2; ++stg) {
      const int NL = (int)chains[stg].size();
      while (all.size() % 256) all.push_back(0);
      offs[br][stg] = all.size();
      std::vector<uint8_t> wb(size_t(NL) * W_LAYER, 0);
      for (int l = 0; l < NL; ++l) {
        const std::string &bn = chains[stg][l].first, &cv = chains[stg][l].second;
        const float *gm = T(bn + ".weight"), *be = T(bn + ".bias"), *mu = T(bn + ".running_mean"), *var = T(bn + ".running_var");
        const float *W = T(cv + ".weight"), *bi = T(cv + ".bias");  // [co][ci][tap]
        double a[32], b[32];
        for (int c = 0; c < 32; ++c) {
          a[c] = double(gm[c]) / sqrt(double(var[c]) + 1e-5);
          b[c] = double(be[c]) - double(mu[c]) * a[c];
        }
        __nv_bfloat16* wl = reinterpret_cast<__nv_bfloat16*>(wb.data() + size_t(l) * W_LAYER);
        for (int j = 0; j < 4; ++j)
          for (int ci = 0; ci < 32; ++ci)
            for (int e = 0; e < 2; ++e)
              for (int co = 0; co < 32; ++co) {
                const int t = j - e;
                const float v = (t >= 0 && t <= 2) ? float(double(W[(co * 32 + ci) * 3 + t]) * a[ci]) : 0.f;
                wl[((j * 4 + ci / 8) * 64 + e * 32 + co) * 8 + (ci % 8)] = __float2bfloat16(v);
              }
        __nv_bfloat16* cl = reinterpret_cast<__nv_bfloat16*>(wb.data() + size_t(l) * W_LAYER + W_CONV);
        auto split = [](double x, __nv_bfloat16* hi, __nv_bfloat16* lo) {
          *hi = __float2bfloat16(float(x));
          *lo = __float2bfloat16(float(x - double(__bfloat162float(*hi))));
        };
        for (int co = 0; co < 32; ++co) {
          double ee[3] = {0, 0, 0};
          for (int t = 0; t < 3; ++t)
            for (int ci = 0; ci < 32; ++ci) ee[t] += double(W[(co * 32 + ci) * 3 + t]) * b[ci];
          __nv_bfloat16 v[6];
          split(double(bi[co]) + ee[0] + ee[1] + ee[2], &v[0], &v[1]);
          split(-ee[0], &v[2], &v[3]);
          split(-ee[2], &v[4], &v[5]);
          for (int e = 0; e < 2; ++e)
            for (int k6 = 0; k6 < 6; ++k6) {
              const int k = 6 * e + k6;  // k-group k / 8, element k % 8
              cl[((k / 8) * 64 + e * 32 + co) * 8 + (k % 8)] = v[k6];
            }
        }
      }
      all.insert(all.end(), wb.begin(), wb.end());
    }
  }
  State* S = new State();
  if (cudaMalloc((void**)&S->d_w, all.size()) != cudaSuccess) {
    delete S;
    MURAL_FAIL("cudaMalloc of the row-pair tcgen05 weight blob failed");
  }
  cudaMemcpy(S->d_w, all.data(), all.size(), cudaMemcpyHostToDevice);
  for (int br = 0; br < 2; ++br)
    for (int stg = 0; stg < 2; ++stg) S->blob[br][stg] = S->d_w + offs[br][stg];
  m->tc2 = S;
  return 0;
}

}  // namespace mural
