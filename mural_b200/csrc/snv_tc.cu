// bf16 tcgen05 path of the Network2 conv stack (MURAL_MODE_BF16): one kernel per *stage* of a branch, the whole
// chain of Conv1d(32,32,3) layers of that stage executed back to back on 128-row tiles with every
// intermediate activation kept on chip (TMEM accumulator -> registers -> bf16 A operand in shared memory).
// Reference arithmetic: MuRaL/model/model_snv.py:475-488 / 497-510 and ResBlock :794-812.
//
// Row space of a stage: the L positions of every site of the chunk laid end to end with ONE all-zero
// separator row between sites (and before the first / after the last):  row(s, p) = 1 + s*(L+1) + p.
// The separator is the conv's zero padding, so a 3-tap convolution is three MMAs whose A operand is the
// same shared-memory tile shifted by one 16-byte row (SWIZZLE_NONE K-major core matrices are row-linear):
//     D[128 x 32] (TMEM, fp32) += A_tap[128 x 32ci] (smem bf16) * W_tap[32ci x 32co] (smem bf16),  tap = 0,1,2
// i.e. 6 tcgen05.mma (M=128, N=32, K=16) per layer per tile, issued by one thread, completion via
// tcgen05.commit -> mbarrier.  BatchNorm (eval) in front of each conv is folded into the bf16 weights and a
// bias; because the reference pads AFTER the BN, the two rows at the site edges get a per-edge correction.
// Global activations are fp32 in a plane-major layout  buf[plane = c/4][row][4 floats]  so that a warp
// whose lanes own consecutive rows reads/writes 512 contiguous bytes per instruction.
#include <cuda_bf16.h>
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "snv_model.cuh"
#include "snv_tc_args.cuh"

namespace mural {
namespace tc {

constexpr int TILE = 128;
constexpr int A_ROWS = TILE + 2;         // one zero halo row above and below the tile
constexpr int A_PLANE = A_ROWS * 16;     // bytes per 8-channel plane
constexpr int A_BYTES = 4 * A_PLANE;     // 8320
constexpr int A_SLOT = (A_BYTES + 127) & ~127;
constexpr int AC_PLANE = TILE * 16;      // constant-column operand: [2 planes][128 rows][16 B]
constexpr int AC_BYTES = 2 * AC_PLANE;
constexpr int PARK_BYTES = 4 * TILE * 16;  // outer-skip operand of the tile (x0 / conv2 output), bf16 planes, private to its thread
constexpr int SLOT_BYTES = A_SLOT + AC_BYTES + PARK_BYTES;
constexpr int W_CONV = 3 * 4 * 32 * 16;  // bf16 [tap][ci/8][co][8] = 6144 bytes
constexpr int W_BIAS = 2 * 32 * 16;      // bf16 [k/8][co][8]: bias / edge-correction rows of the 7th MMA = 1024 bytes
constexpr int W_LAYER = W_CONV + W_BIAS;
#ifndef MURAL_TC_NGROUP
#define MURAL_TC_NGROUP 4
#endif
constexpr int NGROUP = MURAL_TC_NGROUP;  // thread groups (128 threads = one tile's rows) per CTA
constexpr int NINFL = 8 / NGROUP;        // tiles in flight per group
constexpr int NSLOT = NGROUP * NINFL;    // 8 slots x 64 TMEM columns = all 512 columns of the SM
constexpr int THREADS = NGROUP * 128;
// instruction descriptor, kind::f16: D=F32 (bit4), A=BF16 (bit7), B=BF16 (bit10), K-major A and B, N=32, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

// Optional in-kernel phase timing (-DMURAL_TC_TIMING, scratch builds only): per-thread clock() deltas accumulated
// per phase, dumped for lane 0 of warps 0 (issuer) and 1 of every group by mural_tc_timing_dump().
enum { T_WAIT = 0, T_LD, T_EPI, T_FENCE, T_BAR, T_ISSUE, T_BEGIN, T_FETCH, T_FINAL, T_OTHER, T_N };
#ifdef MURAL_TC_TIMING
__device__ unsigned long long g_tc_timing[3][148][NGROUP][2][T_N + 2];
#define TT(cat)                         \
  do {                                  \
    const uint32_t _n = clock();        \
    tacc[cat] += _n - tlast;            \
    tlast = _n;                         \
  } while (0)
#else
#define TT(cat) do { } while (0)
#endif
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SWIZZLE_NONE, K-major UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout)
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

// One stage of one branch, persistent: one CTA per SM, 4 thread groups x 2 tiles in flight.
//
// A group's 128 threads own the 128 rows of a tile (thread = row = TMEM lane).  Per layer one warp of the group (rotating)
// issues 7 tcgen05.mma (constant-column MMA for bias + site-edge corrections, then 3 taps x 2 K-halves) and
// commits to the slot's mbarrier; while those run, the group serves its other in-flight tile.  The other three warps only
// bar.arrive on the slot's named barrier and run ahead.  The residual stream never leaves the tensor core:
// TMEM region R (32 columns) is pre-loaded with x0 (tcgen05.st) and the second conv of each ResBlock
// ACCUMULATES onto it, the first conv of each ResBlock goes to a scratch region T.  The epilogue of a layer is
// therefore only  tcgen05.ld -> bf16 -> ReLU -> st.shared  (the next layer's A operand).
template <int MODE, int FM>  // FM: 0 plain rows, 1 lattice (RB4: device-side geometry; C_RB4: pre-pooled single rows), 2 edge gather (RB4)
__global__ void __launch_bounds__(THREADS, 1) k_stage_tc(StageArgs a) {
  constexpr bool LAT = FM == 1;
  constexpr bool EDGE = (MODE == RB4) && FM == 2;
  if (a.info && a.info->dense != a.want) return;  // uniform over the grid; nothing allocated yet
  // RB4 with LAT: stage-1 lattice launch, geometry known only on the device
  constexpr bool DYN = (MODE == RB4) && LAT;
  const int L_ = DYN ? a.info->M[a.lat_branch] : a.L;
  const int rows_ = DYN ? a.info->lat_rows[a.lat_branch] : int(a.rows);
  const int n_tiles_ = DYN ? a.info->lat_tiles[a.lat_branch] : a.n_tiles;
  if (DYN && int(blockIdx.x) * NSLOT >= n_tiles_) return;
  constexpr int NL = n_layers(MODE);
  constexpr int STRIDE = TILE - 2 * NL;  // valid output rows per tile (the chain eats NL rows on each side)
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sW = smem;                              // [NL][W_LAYER], shared by all slots
  unsigned char* sSlots = sW + NL * W_LAYER;             // [NSLOT][SLOT_BYTES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSlots + NSLOT * SLOT_BYTES);   // [NSLOT] per-slot MMA-done barriers + 1 weight-arrival barrier
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NSLOT + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int g = tid >> 7, lt = tid & 127;
#ifdef MURAL_TC_TIMING
  uint32_t tacc[T_N] = {0};
  uint32_t tlast = clock(), tlayers = 0;
  const uint32_t tstart = tlast;
#endif

  // ---- one-time setup
  // the stage's weight blob (NL x 7 KB, contiguous, 16-byte aligned on both sides) arrives by ONE bulk async copy (TMA engine,
  // cp.async.bulk global -> shared, completion on an mbarrier) while the threads clear the slots' halo rows
  uint64_t* wbar = reinterpret_cast<uint64_t*>(sSlots + NSLOT * SLOT_BYTES) + NSLOT;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(wbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    constexpr uint32_t WB = NL * W_LAYER;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(wbar)), "r"(WB) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sW)),
                 "l"(a.wblob), "r"(WB), "r"(smem_u32(wbar))
                 : "memory");
  }
  for (int e = tid; e < NSLOT * 8; e += THREADS) {  // halo rows 0 and 129 of the 4 planes of every slot
    const int sl = e >> 3, plane = (e >> 1) & 3, row = (e & 1) ? (A_ROWS - 1) : 0;
    *reinterpret_cast<uint4*>(sSlots + sl * SLOT_BYTES + plane * A_PLANE + row * 16) = make_uint4(0, 0, 0, 0);
  }
  for (int e = tid; e < NSLOT * TILE; e += THREADS) {  // K columns 8..15 of the constant operand are unused
    const int sl = e >> 7, row = e & 127;
    *reinterpret_cast<uint4*>(sSlots + sl * SLOT_BYTES + A_SLOT + AC_PLANE + row * 16) = make_uint4(0, 0, 0, 0);
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
  mbar_wait(smem_u32(wbar), 0);  // weights have landed (written by the async proxy: directly visible to tcgen05.mma)
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;  // this warp's TMEM lane quarter
  const int Lp1 = L_ + 1;
  const int tile_step = gridDim.x * NSLOT;
  // descriptor bases (start-address field is the low 14 bits: offsets below never carry out of it)
  const uint64_t dW0 = umma_desc(smem_u32(sW), 512, 128);
  const uint64_t dA0 = umma_desc(smem_u32(sSlots), A_PLANE, 128);
  const uint64_t dC0 = umma_desc(smem_u32(sSlots) + A_SLOT, AC_PLANE, 128);

  // ---- per-slot constants of this group (slot k of group g = g*NINFL + k)
  unsigned char* const sA0 = sSlots + (g * NINFL) * SLOT_BYTES;
  const uint32_t bar0 = smem_u32(bars + g * NINFL);
  uint32_t phase[NINFL] = {};

  // publish the slot's operands to the tensor core: three warps of the group arrive and run ahead, the issuing warp waits for
  // all 128 threads and its elected lane issues the 7 MMAs of layer l, then commits to the slot's mbarrier.
  // One named barrier per slot so that a warp running one step ahead never double-arrives.
  auto sync_and_issue = [&](int k, int l) {
    TT(T_EPI);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    TT(T_FENCE);
    const int sl = g * NINFL + k;
    // the issuing warp rotates with (layer, slot): UTCHMMA issue blocks while the tensor pipe's queue is full, and no
    // single warp should carry all of that back-pressure
    if ((lt >> 5) != ((l + k) & 3)) {
      asm volatile("bar.arrive %0, 128;" ::"r"(1 + sl) : "memory");
      TT(T_BAR);
    } else {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + sl) : "memory");
      TT(T_BAR);
      if ((lt & 31) == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // R-type layers (second conv of a ResBlock, conv2, conv3) accumulate into / create region R, others use T
        const bool rtype = (MODE == RB4) ? (l & 1) : (MODE == C_RB4 ? !(l & 1) : true);
        const bool acc_first = (MODE == RB4) ? rtype : (MODE == C_RB4 ? (rtype && l > 0) : false);
        const uint32_t d = tmem_base + sl * 64 + (rtype ? 0 : 32);
        const uint64_t dA = dA0 + uint64_t((sl * SLOT_BYTES) >> 4), dC = dC0 + uint64_t((sl * SLOT_BYTES) >> 4);
        const uint64_t dW = dW0 + uint64_t((l * W_LAYER) >> 4);
        umma_bf16(d, dC, dW + (W_CONV >> 4), acc_first ? 1u : 0u);
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            umma_bf16(d, dA + uint64_t((2 * h * A_PLANE + t * 16) >> 4), dW + uint64_t(((t * 4 + 2 * h) * 512) >> 4), 1u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * k) : "memory");
      }
      __syncwarp();
      TT(T_ISSUE);
    }
  };

  // row of this thread in tile `tile_idx`, its position inside the site (-1 = separator / outside) and the chain
  // input of that row (RB4: the row itself; pooled modes: max over the pool window, model_snv.py:361,371,404,414)
  auto fetch = [&](int tile_idx, int& r, int& p, uint4 (&x)[4]) {
    r = tile_idx * STRIDE - NL + lt;
    p = -1;
    int site = 0;
    if (r > 0 && r < rows_) {
      site = r / Lp1;
      p = r - site * Lp1 - 1;
    }
    bool live = p >= 0;
#ifdef MURAL_TC_TIMING
    if (a.ablate & 1) live = false;  // ablation (scratch timing builds only): no loads at all
#endif
    if (EDGE) {
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = make_uint4(0, 0, 0, 0);
      if (live) {
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
      }
    } else if (MODE == RB4 && LAT) {
      // stage-1 lattice: pseudo-site u = strand*ps1 + phase walks the full-bin table of its strand in steps of ps1 ('-' strand
      // against the genome, so that the oriented conv taps line up with the '+' ones); rows are read in place from the stem tables
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = make_uint4(0, 0, 0, 0);
      if (live) {
        const int strand = site >= a.ps1, phase = site - strand * a.ps1;
        const int xx = (strand ? L_ - 1 - p : p) * a.ps1 + phase;
        if (xx < a.info->n_pos && a.info->has[strand]) {
          const uint4* row = a.tab[strand] + int64_t(xx) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) x[q] = __ldg(row + q);
        }
      }
    } else if (MODE == RB4) {
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = live ? __ldg(a.in + q * a.in_rows_alloc + r) : make_uint4(0, 0, 0, 0);
    } else if (MODE != RB4) {
      int lo = p * a.ps - a.pp, hi = lo + a.pk;
      lo = lo < 0 ? 0 : lo;
      hi = hi > a.Lin ? a.Lin : hi;
      if (!live) hi = lo;
      const uint32_t ninf = live ? 0xFF80FF80u : 0u;  // bf16 -inf pair; separator rows stay zero
#pragma unroll
      for (int q = 0; q < 4; ++q) x[q] = make_uint4(ninf, ninf, ninf, ninf);
      if (!LAT || MODE != C_RB4) {
        const int base = 1 + site * (a.Lin + 1);
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
      } else {
        // every stage-2 input row is ONE pre-pooled row: bins whose pool window lies on the stage-1 lattice read the pooled
        // lattice (k_lattice_pool) at genomic bin start x(rr) = x(0) +- rr*ps1, i.e. row lat_base + rr for either strand;
        // the nlo + nhi bins per site that touch edge rows read the per-site pooled rows of k_edge_pool.  Plain loads
        // into the prefetch registers: nothing here waits for memory.
        if (live) {
          if (p >= a.nlo && p < L_ - a.nhi) {
            const int s = __ldg(a.pos + site), strand = __ldg(a.meta + site) & 1;
            const int lat_base = lattice_base(a.info, a.br, s, strand, a.R, a.off0, a.ps1, a.pp1, a.pk1);
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = __ldg(a.lat2 + q * a.lat_ra + lat_base + lo);
          } else {
            const int64_t idx = int64_t(site) * (a.nlo + a.nhi) + (p < a.nlo ? p : p - (L_ - a.nhi) + a.nlo);
#pragma unroll
            for (int q = 0; q < 4; ++q) x[q] = __ldg(a.epool + q * a.epool_ra + idx);
          }
        }
      }
    }
  };

  // first A operand / constant operand / residual region of slot k from the fetched row, then layer 0
  auto begin_tile = [&](int k, int r, int p, const uint4 (&x)[4]) {
    unsigned char* sA = sA0 + k * SLOT_BYTES;
    const bool live = p >= 0;
    if (MODE == RB4) {  // x0 is needed again for the outer skip after the last layer: parked in the slot (same thread reads it back)
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(sA + A_SLOT + AC_BYTES + q * (TILE * 16) + lt * 16) = x[q];
    }
    if (MODE == RB4) {
      uint32_t f[32];  // residual region R <- x0 (fp32)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&x[q]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[8 * q + 2 * j] = w[j] << 16;
          f[8 * q + 2 * j + 1] = w[j] & 0xFFFF0000u;
        }
      }
      TMEM_ST32(tmem_base + lane_off + (g * NINFL + k) * 64, f);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = relu_bf16x2(x[q].x); o.y = relu_bf16x2(x[q].y); o.z = relu_bf16x2(x[q].z); o.w = relu_bf16x2(x[q].w);
        *reinterpret_cast<uint4*>(sA + q * A_PLANE + (lt + 1) * 16) = o;
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(sA + q * A_PLANE + (lt + 1) * 16) = x[q];
    }
    // constant-column operand row: {1, 1, [pos==0], [pos==0], [pos==L-1], [pos==L-1], 0, 0} in bf16 (1.0 = 0x3F80)
    const uint32_t one2 = 0x3F803F80u;
    *reinterpret_cast<uint4*>(sA + A_SLOT + lt * 16) =
        make_uint4(live ? one2 : 0u, p == 0 ? one2 : 0u, (live && p == L_ - 1) ? one2 : 0u, 0u);
    TT(T_BEGIN);
    sync_and_issue(k, 0);
  };

  // static schedule: both slots of the group walk the layer chain in lockstep (layer index is a compile-time
  // constant after unrolling); the NEXT pair of tiles is fetched into registers while the chain runs.
  int tile0 = blockIdx.x * NSLOT + g * NINFL;
  int rn[NINFL], pn[NINFL];
  uint4 xn[NINFL][4];
#pragma unroll
  for (int k = 0; k < NINFL; ++k) {
    rn[k] = 0; pn[k] = -1;
    if (tile0 + k < n_tiles_) fetch(tile0 + k, rn[k], pn[k], xn[k]);
  }
  for (; tile0 < n_tiles_; tile0 += tile_step) {
    bool act[NINFL];
    int r[NINFL], p[NINFL];
#pragma unroll
    for (int k = 0; k < NINFL; ++k) {
      act[k] = tile0 + k < n_tiles_;
      r[k] = rn[k];
      p[k] = pn[k];
      TT(T_OTHER);
      if (act[k]) begin_tile(k, r[k], p[k], xn[k]);
    }
#pragma unroll
    for (int k = 0; k < NINFL; ++k)
      if (tile0 + tile_step + k < n_tiles_) fetch(tile0 + tile_step + k, rn[k], pn[k], xn[k]);
    TT(T_FETCH);
#pragma unroll
    for (int l = 0; l < NL; ++l) {
#pragma unroll
      for (int k = 0; k < NINFL; ++k) {
        if (!act[k]) continue;
#ifdef MURAL_TC_TIMING
        ++tlayers;
#endif
        TT(T_OTHER);
        const bool live = p[k] >= 0;
        const bool valid = live && lt >= NL && lt < TILE - NL;
        uint4 xr[4];
        mbar_wait(bar0 + 8 * k, phase[k]);
        phase[k] ^= 1;
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool rtype = (MODE == RB4) ? (l & 1) : (MODE == C_RB4 ? !(l & 1) : true);
        uint32_t acc[32];
        TT(T_WAIT);
        TMEM_LD32(acc, tmem_base + lane_off + (g * NINFL + k) * 64 + (rtype ? 0 : 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        TT(T_LD);
        if (l < NL - 1) {
          uint4 o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o[q].x = pack_bf16(__uint_as_float(acc[8 * q]), __uint_as_float(acc[8 * q + 1]));
            o[q].y = pack_bf16(__uint_as_float(acc[8 * q + 2]), __uint_as_float(acc[8 * q + 3]));
            o[q].z = pack_bf16(__uint_as_float(acc[8 * q + 4]), __uint_as_float(acc[8 * q + 5]));
            o[q].w = pack_bf16(__uint_as_float(acc[8 * q + 6]), __uint_as_float(acc[8 * q + 7]));
          }
          if (MODE == C_RB4 && l == 0) {  // jump = conv2 output: parked (bf16) in the slot, re-read by the same thread at the end
            unsigned char* pk = sA0 + k * SLOT_BYTES + A_SLOT + AC_BYTES + lt * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(pk + q * (TILE * 16)) = o[q];
          }
          if (live) {  // separator rows were zeroed by begin_tile and are never rewritten
            unsigned char* sA = sA0 + k * SLOT_BYTES;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = relu_bf16x2(o[q].x); w.y = relu_bf16x2(o[q].y); w.z = relu_bf16x2(o[q].z); w.w = relu_bf16x2(o[q].w);
              *reinterpret_cast<uint4*>(sA + q * A_PLANE + (lt + 1) * 16) = w;
            }
          }
          sync_and_issue(k, l + 1);
        } else if (valid) {
          if (MODE == SINGLE) {  // conv3 + ReLU -> fp32 planes for the head
            float4* out4 = reinterpret_cast<float4*>(a.out);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              out4[q * a.out_rows_alloc + r[k]] =
                  make_float4(fmaxf(__uint_as_float(acc[4 * q]), 0.f), fmaxf(__uint_as_float(acc[4 * q + 1]), 0.f),
                              fmaxf(__uint_as_float(acc[4 * q + 2]), 0.f), fmaxf(__uint_as_float(acc[4 * q + 3]), 0.f));
          } else {  // outer skip: + x0 (RB4, re-read from the input) or + jump (C_RB4, parked in the output row)
            uint4* out4 = reinterpret_cast<uint4*>(a.out);
            {
              const unsigned char* pk = sA0 + k * SLOT_BYTES + A_SLOT + AC_BYTES + lt * 16;
#pragma unroll
              for (int q = 0; q < 4; ++q) xr[q] = *reinterpret_cast<const uint4*>(pk + q * (TILE * 16));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 o;
              o.x = pack_bf16(__uint_as_float(acc[8 * q]) + bf16_lo(xr[q].x), __uint_as_float(acc[8 * q + 1]) + bf16_hi(xr[q].x));
              o.y = pack_bf16(__uint_as_float(acc[8 * q + 2]) + bf16_lo(xr[q].y), __uint_as_float(acc[8 * q + 3]) + bf16_hi(xr[q].y));
              o.z = pack_bf16(__uint_as_float(acc[8 * q + 4]) + bf16_lo(xr[q].z), __uint_as_float(acc[8 * q + 5]) + bf16_hi(xr[q].z));
              o.w = pack_bf16(__uint_as_float(acc[8 * q + 6]) + bf16_lo(xr[q].w), __uint_as_float(acc[8 * q + 7]) + bf16_hi(xr[q].w));
              out4[q * a.out_rows_alloc + r[k]] = o;
            }
          }
        }
        if (l == NL - 1) TT(T_FINAL);
      }
    }
  }
#ifdef MURAL_TC_TIMING
  if ((lt & 31) == 0 && (lt >> 5) < 2) {
    unsigned long long* o = g_tc_timing[MODE][blockIdx.x % 148][g][lt >> 5];
    for (int i = 0; i < T_N; ++i) atomicAdd(&o[i], (unsigned long long)tacc[i]);
    atomicAdd(&o[T_N], (unsigned long long)tlayers);
    atomicAdd(&o[T_N + 1], (unsigned long long)(clock() - tstart));
  }
#endif
  // ---- teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

// second-level reuse on the lattice: the stage-2 pool window of an interior bin covers pk consecutive lattice rows of
// one pseudo-site, so its maximum is again a function of the genomic position only
__global__ void __launch_bounds__(256) k_lattice_pool(const ChunkInfo* __restrict__ info, int br, int pk, const uint4* __restrict__ y1,
                                                      uint4* __restrict__ p2, int64_t ra) {
  if (!info->dense) return;
  const int M = info->M[br], rows = info->lat_rows[br];
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < int64_t(rows) * 4; e += int64_t(gridDim.x) * blockDim.x) {
    const int q = int(e & 3), row = int(e >> 2);
    if (row < 1) continue;
    const int mm = (row - 1) % (M + 1);
    if (mm == M) continue;  // separator
    const int n = pk < M - mm ? pk : M - mm;
    uint4 v = __ldg(y1 + q * ra + row);
    for (int k = 1; k < n; ++k) {
      const uint4 w = __ldg(y1 + q * ra + row + k);
      v.x = max_bf16x2(v.x, w.x); v.y = max_bf16x2(v.y, w.y); v.z = max_bf16x2(v.z, w.z); v.w = max_bf16x2(v.w, w.w);
    }
    p2[q * ra + row] = v;
  }
}

// pool-2 bins of a site that touch its edge rows (the first nlo and last nhi bins): max over the window's stage-1 rows,
// each taken from the site's edge pseudo-site (rr < LAT_EO or rr >= L1 - LAT_EO) or from the lattice
struct EdgePool {
  const uint4* lat; int64_t lat_ra;
  const uint4* edge; int64_t edge_ra;
  uint4* out; int64_t out_ra;
  const int32_t* pos; const int32_t* meta;
  int64_t ns;
  int br, L1, L2, pk, ps, pp, nlo, nhi, ps1, pp1, pk1, off0, R;
};
__global__ void __launch_bounds__(256) k_edge_pool(const ChunkInfo* __restrict__ info, EdgePool a) {
  if (!info->dense) return;
  const int neb = a.nlo + a.nhi;
  const int64_t total = a.ns * neb * 4;
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
    const int q = int(e & 3);
    const int64_t sb = e >> 2;
    const int64_t site = sb / neb;
    const int b = int(sb - site * neb);
    const int p = b < a.nlo ? b : a.L2 - a.nhi + (b - a.nlo);
    int lo = p * a.ps - a.pp, hi = lo + a.pk;
    lo = lo < 0 ? 0 : lo;
    hi = hi > a.L1 ? a.L1 : hi;
    const int s = a.pos[site], strand = a.meta[site] & 1;
    const int lat_base = lattice_base(info, a.br, s, strand, a.R, a.off0, a.ps1, a.pp1, a.pk1);
    const int64_t ebase = 1 + site * (LAT_EL + 1);
    const uint4 ninf = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);  // bf16 -inf
    uint4 w[7];  // pk <= 7 for every pool of Network2; all loads are issued before the first max
#pragma unroll
    for (int u = 0; u < 7; ++u) {
      const int rr = lo + u;
      const bool e_lo = rr < LAT_EO, e_hi = rr >= a.L1 - LAT_EO;
      const uint4* src = (e_lo || e_hi) ? a.edge + q * a.edge_ra + (e_lo ? ebase + rr : ebase + rr - a.L1 + LAT_EL)
                                        : a.lat + q * a.lat_ra + lat_base + rr;
      w[u] = rr < hi ? __ldg(src) : ninf;
    }
    uint4 v = w[0];
#pragma unroll
    for (int u = 1; u < 7; ++u) {
      v.x = max_bf16x2(v.x, w[u].x); v.y = max_bf16x2(v.y, w[u].y); v.z = max_bf16x2(v.z, w[u].z); v.w = max_bf16x2(v.w, w[u].w);
    }
    a.out[q * a.out_ra + sb] = v;
  }
}

// head for the plane layout: global max over the site's L3 rows of both branches, folded BN+Linear, combine
struct HeadTc {
  const float* x;  // fp32 planes [8][rows_alloc][4], conv3 output after ReLU
  int64_t rows_alloc;
  const float* Wfc;
  const float* bfc;
  int L3;
};

// 8 threads per site, one per 4-channel plane: each reads its plane's L3 consecutive float4 rows (coalesced across the
// sites of a warp), reduces the max, multiplies its 4 channels into the folded BN+Linear and the 8 partial sums are
// combined with three xor-shuffles.
__global__ void __launch_bounds__(256) k_head_tc(HeadTc b0, HeadTc b1, const float* __restrict__ local_logits, int64_t n, int NC,
                                                 float* __restrict__ logp, float* __restrict__ tg0, float* __restrict__ tg1,
                                                 float* __restrict__ tl0, float* __restrict__ tl1) {
  const int64_t gt = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int64_t site = gt >> 3;
  const int q = int(gt & 7);
  const bool ok = site < n;  // out-of-range lanes stay for the shuffles
  float lg[2][16];
#pragma unroll 1
  for (int br = 0; br < 2; ++br) {
    const HeadTc& B = br ? b1 : b0;
    float4 mx = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (ok) {
      const float4* x = reinterpret_cast<const float4*>(B.x) + int64_t(q) * B.rows_alloc + 1 + site * (B.L3 + 1);
      for (int p = 0; p < B.L3; ++p) {
        const float4 v = __ldg(x + p);
        mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
      }
      float* tg = br ? tg1 : tg0;
      if (tg) *reinterpret_cast<float4*>(tg + site * 32 + 4 * q) = mx;
    }
    const float* W = B.Wfc + 4 * q * NC;
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (o < NC) {
        float part = mx.x * W[o] + mx.y * W[NC + o] + mx.z * W[2 * NC + o] + mx.w * W[3 * NC + o];
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        lg[br][o] = part + B.bfc[o];
      }
  }
  if (q != 0 || !ok) return;
  // combine (model_snv.py:515-523); static indexing keeps everything in registers
  float pr[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) pr[o] = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v[16], mx = -FLT_MAX, sum = 0.f;
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (o < NC) {
        v[o] = k == 2 ? local_logits[site * NC + o] : lg[k][o];
        mx = fmaxf(mx, v[o]);
      }
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (o < NC) { v[o] = expf(v[o] - mx); sum += v[o]; }
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (o < NC) pr[o] += (k == 2 ? 1.f : 0.5f) * (v[o] / sum);
  }
#pragma unroll
  for (int o = 0; o < 16; ++o)
    if (o < NC) {
      if (tl0) tl0[site * NC + o] = lg[0][o];
      if (tl1) tl1[site * NC + o] = lg[1][o];
      logp[site * NC + o] = logf(fmaxf(pr[o] / 2.f, 1e-9f));
    }
}

constexpr int N_SIDE = 1, N_SIDE_EV = 8;
struct TcState {
  uint8_t* d_w = nullptr;         // all stage blobs back to back
  const uint8_t* blob[2][3] = {};  // [branch][stage]
  // side stream of the dense path (snv_forward_tc): the local branch runs beside the stem tables of the first chunk, the tail of
  // chunk c beside the stem of chunk c+1
  cudaStream_t side[N_SIDE] = {};
  cudaEvent_t ev[N_SIDE_EV] = {};
  unsigned ev_next = 0;   // ring position (unsigned: wraps cleanly)
};

static inline int64_t rows_of(int64_t ns, int L) { return ns * (L + 1) + 1; }
static inline int64_t rows_alloc(int64_t ns, int L) { return (rows_of(ns, L) + 7 + 8) & ~int64_t(7); }

static int m_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int MODE, int FM = 0>
static int launch_stage(const StageArgs& a, cudaStream_t st, const char* role = "") {
  constexpr int NL = n_layers(MODE);
  const size_t smem = size_t(NL) * W_LAYER + size_t(NSLOT) * SLOT_BYTES + (NSLOT + 1) * 8 + 16;
  static bool configured = false;
  if (!configured) {
    CUDA_TRY((cudaFuncSetAttribute(k_stage_tc<MODE, FM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    configured = true;
  }
  int grid = (a.n_tiles + NSLOT - 1) / NSLOT;
  if (grid > m_sm_count()) grid = m_sm_count();
#ifdef MURAL_TC_TIMING
  StageArgs a2 = a;
  { const char* e = getenv("MURAL_TC_ABLATE"); a2.ablate = e ? atoi(e) : 0; }
#else
  const StageArgs& a2 = a;
#endif
  // profile names are interned per (mode, role): prof_pre keeps the pointer
  static std::map<std::string, std::string> names;
  const std::string key = std::string(MODE == RB4 ? "k_stage_tc<RB4>" : (MODE == C_RB4 ? "k_stage_tc<C_RB4>" : "k_stage_tc<SINGLE>")) + role;
  const char* nm = names.emplace(key, key).first->second.c_str();
  if (MODE == RB4 && FM == 2) LAUNCH_N(nm, (k_stage_tc<RB4, 2>), grid, THREADS, smem, st, a2);
  else if (MODE == RB4 && FM == 1) LAUNCH_N(nm, (k_stage_tc<RB4, 1>), grid, THREADS, smem, st, a2);
  else if (MODE == RB4) LAUNCH_N(nm, (k_stage_tc<RB4, 0>), grid, THREADS, smem, st, a2);
  else if (MODE == C_RB4 && FM == 1) LAUNCH_N(nm, (k_stage_tc<C_RB4, 1>), grid, THREADS, smem, st, a2);
  else if (MODE == C_RB4) LAUNCH_N(nm, (k_stage_tc<C_RB4, 0>), grid, THREADS, smem, st, a2);
  else LAUNCH_N(nm, (k_stage_tc<SINGLE, 0>), grid, THREADS, smem, st, a2);
  return 0;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------------- host
int snv_tc_prepare(mural_snv_model* m, const float* h_blob) {
  snv_tc_destroy(m);
  if (m->cfg.channels != 32 || m->cfg.kernel_size != 3 || m->cfg.n_cont > 0) return 0;  // tcgen05 path is specialised; fp32 path serves the rest
  using namespace tc;
  auto T = [&](const std::string& n) { return h_blob + m->layout[m->index.at(n)].offset; };
  std::vector<uint8_t> all;
  size_t offs[2][3];
  for (int br = 0; br < 2; ++br) {
    const std::string s = br ? "_2" : "";
    std::vector<std::pair<std::string, std::string>> chains[3];  // (bn, conv) per layer
    for (int g = 1; g <= 2; ++g) {
      if (g == 2) chains[1].push_back({"conv2" + s + ".0", "conv2" + s + ".1"});
      for (int i = 0; i < 2; ++i) {
        const std::string p = "RBs" + std::to_string(g) + s + "." + std::to_string(i);
        chains[g - 1].push_back({p + ".bn1", p + ".conv1"});
        chains[g - 1].push_back({p + ".bn2", p + ".conv2"});
      }
    }
    chains[2].push_back({"conv3" + s + ".0", "conv3" + s + ".1"});
    for (int stg = 0; stg < 3; ++stg) {
      const int NL = (int)chains[stg].size();
      while (all.size() % 256) all.push_back(0);
      offs[br][stg] = all.size();
      std::vector<uint8_t> wb(size_t(NL) * W_LAYER, 0);
      for (int l = 0; l < NL; ++l) {
        const std::string &bn = chains[stg][l].first, &cv = chains[stg][l].second;
        const float *g = T(bn + ".weight"), *be = T(bn + ".bias"), *mu = T(bn + ".running_mean"), *var = T(bn + ".running_var");
        const float *W = T(cv + ".weight"), *bi = T(cv + ".bias");  // [co][ci][tap]
        double a[32], b[32];
        for (int c = 0; c < 32; ++c) {
          a[c] = double(g[c]) / sqrt(double(var[c]) + 1e-5);
          b[c] = double(be[c]) - double(mu[c]) * a[c];
        }
        __nv_bfloat16* wl = reinterpret_cast<__nv_bfloat16*>(wb.data() + size_t(l) * W_LAYER);
        for (int t = 0; t < 3; ++t)
          for (int ci = 0; ci < 32; ++ci)
            for (int co = 0; co < 32; ++co)
              wl[((t * 4 + ci / 8) * 32 + co) * 8 + (ci % 8)] = __float2bfloat16(float(double(W[(co * 32 + ci) * 3 + t]) * a[ci]));
        // constant-column MMA operand: row co = {bias_hi, bias_lo, -e_left_hi, -e_left_lo, -e_right_hi, -e_right_lo, 0, 0}
        __nv_bfloat16* cl = reinterpret_cast<__nv_bfloat16*>(wb.data() + size_t(l) * W_LAYER + W_CONV);
        auto split = [](double x, __nv_bfloat16* hi, __nv_bfloat16* lo) {
          *hi = __float2bfloat16(float(x));
          *lo = __float2bfloat16(float(x - double(__bfloat162float(*hi))));
        };
        for (int co = 0; co < 32; ++co) {
          double e[3] = {0, 0, 0};
          for (int t = 0; t < 3; ++t)
            for (int ci = 0; ci < 32; ++ci) e[t] += double(W[(co * 32 + ci) * 3 + t]) * b[ci];
          split(double(bi[co]) + e[0] + e[1] + e[2], &cl[co * 8 + 0], &cl[co * 8 + 1]);
          split(-e[0], &cl[co * 8 + 2], &cl[co * 8 + 3]);
          split(-e[2], &cl[co * 8 + 4], &cl[co * 8 + 5]);
        }
      }
      all.insert(all.end(), wb.begin(), wb.end());
    }
  }
  TcState* S = new TcState();
  if (cudaMalloc((void**)&S->d_w, all.size()) != cudaSuccess) {
    delete S;
    MURAL_FAIL("cudaMalloc of the tcgen05 weight blob failed");
  }
  cudaMemcpy(S->d_w, all.data(), all.size(), cudaMemcpyHostToDevice);
  for (int br = 0; br < 2; ++br)
    for (int stg = 0; stg < 3; ++stg) S->blob[br][stg] = S->d_w + offs[br][stg];
  m->tc = S;
  if (int rc = snv_mlp_tc_prepare(m)) return rc;
  return snv_tail_prepare(m, h_blob);
}

void snv_tc_destroy(mural_snv_model* m) {
  snv_mlp_tc_destroy(m);
  snv_tail_destroy(m);
  if (!m->tc) return;
  tc::TcState* S = (tc::TcState*)m->tc;
  cudaFree(S->d_w);
  for (cudaStream_t q : S->side)
    if (q) cudaStreamDestroy(q);
  for (cudaEvent_t e : S->ev)
    if (e) cudaEventDestroy(e);
  delete S;
  m->tc = nullptr;
}

// de-plane a bf16 [4][rows_alloc][8] (or fp32 [8][rows_alloc][4]) buffer into host [site][L][32] for the parity taps
static int save_tap_planes(mural_snv_model* m, const char* name, const void* d, bool is_bf16, int64_t ralloc, int64_t ns, int L,
                           cudaStream_t st) {
  if (!m->debug) return 0;
  std::vector<uint8_t> raw(size_t(ralloc) * (is_bf16 ? 64 : 128));
  CUDA_TRY(cudaStreamSynchronize(st));
  CUDA_TRY(cudaMemcpy(raw.data(), d, raw.size(), cudaMemcpyDeviceToHost));
  std::vector<float>& v = m->tap_store[name];
  v.resize(size_t(ns) * L * 32);
  for (int64_t s = 0; s < ns; ++s)
    for (int p = 0; p < L; ++p)
      for (int c = 0; c < 32; ++c) {
        const int64_t r = 1 + s * (L + 1) + p;
        float x;
        if (is_bf16) {
          const uint16_t h = reinterpret_cast<const uint16_t*>(raw.data())[((c / 8) * ralloc + r) * 8 + (c % 8)];
          const uint32_t u = uint32_t(h) << 16;
          memcpy(&x, &u, 4);
        } else {
          x = reinterpret_cast<const float*>(raw.data())[((c / 4) * ralloc + r) * 4 + (c % 4)];
        }
        v[(s * L + p) * 32 + c] = x;
      }
  return 0;
}

int snv_forward_tc(mural_snv_model* m, const GenomeView* G, const int32_t* d_pos, const int32_t* d_meta,
                   const uint8_t* d_sym, const int64_t* d_cat, int64_t n, float* d_logp, cudaStream_t st) {
  using namespace tc;
  MURAL_CHECK(m->tc != nullptr, "MURAL_MODE_BF16 needs CNN_out_channels == 32 and CNN_kernel_size == 3");
  TcState* S = (TcState*)m->tc;
  const int NC = m->cfg.n_class;
  static int64_t env_chunk = -1;
  if (env_chunk < 0) { const char* e = getenv("MURAL_TC_CHUNK"); env_chunk = e ? atoll(e) : 0; }
  int64_t chunk = m->chunk_sites > 0 ? m->chunk_sites : (env_chunk > 0 ? env_chunk : 524288);
  if (m->chunk_sites <= 0 || (int64_t(1) << 20) % chunk != 0) {  // keep chunks aligned inside super-chunks
    int64_t c2 = 1;
    while (c2 * 2 <= chunk) c2 *= 2;
    chunk = c2;
  }
  if (m->chunk_sites <= 0) {
    // the per-site row buffers grow with the window (stage-1 rows L1 of both branches, 64 B each, stem out + stage-1 out +
    // ~0.3 of that for the later stages): halve the chunk until they fit 16 GiB (524 288 sites at the shipped 1 Kb radius)
    const double per_site = 64.0 * 2.3 * double(m->br[0].L1 + m->br[1].L1);
    static double ws_gib = -1;
    if (ws_gib < 0) { const char* e = getenv("MURAL_TC_WS_GIB"); ws_gib = e ? atof(e) : 16.0; }
    while (chunk > 4096 && double(chunk) * per_site > ws_gib * 1024 * 1024 * 1024) chunk /= 2;
  }
  if (chunk > n) chunk = n;
  // workspace: per branch X0 (stem out), Z1, Z2 as bf16 planes, H as fp32 planes; + local logits, taps, k-mer indices.
  // On the dense path X0 / Z1 hold the stage-1 lattice and the edge pseudo-sites instead of per-site rows.
  const bool use_dense = G != nullptr && !m->slow_stem && getenv("MURAL_NO_DENSE_STEM") == nullptr;
  const bool use_mlp_tc = getenv("MURAL_NO_MLP_TC") == nullptr;
  const bool use_tail = m->tail != nullptr && getenv("MURAL_NO_TAIL") == nullptr;
  const int stride_rb4 = TILE - 2 * 4, stride_crb4 = TILE - 2 * 5;
  auto stage = [&](int mode, int fm, const StageArgs& sa, const char* role) -> int {
    if (mode == RB4) return fm == 2 ? launch_stage<RB4, 2>(sa, st, role) : (fm == 1 ? launch_stage<RB4, 1>(sa, st, role) : launch_stage<RB4, 0>(sa, st, role));
    return fm == 1 ? launch_stage<C_RB4, 1>(sa, st, role) : launch_stage<C_RB4, 0>(sa, st, role);
  };
  const bool use_lat = use_dense && !m->debug && snv_lattice_supported(m) && getenv("MURAL_NO_LATTICE") == nullptr;
  // Dense path: everything up to stage 2 stays in order on the caller's stream; the local branch and the tail (latency-bound, and
  // not on the path to the next chunk's stage kernels) go to a side stream: the local branch runs beside the stem tables, the tail
  // of chunk c beside the stem of chunk c+1.  Measured and not adopted (profiles/r02_side_stream_experiments.txt): the pool passes
  // on side streams next to the stage kernels - a persistent stage CTA that starts late behind a pool CTA holds its whole static
  // share of tiles back.
  static int env_side = -1;
  if (env_side < 0) { const char* e = getenv("MURAL_TC_SIDE_STREAMS"); env_side = e ? atoi(e) : 1; }
  // (the per-kernel profile of bench.py times every kernel alone: one stream there)
  const bool multi = use_lat && use_tail && env_side != 0 && !g_prof;
  static int pool_grid = -1;
  if (pool_grid < 0) { const char* e = getenv("MURAL_POOL_GRID"); pool_grid = e ? atoi(e) : 6; }
  if (multi && !S->side[0]) {
    for (int i = 0; i < N_SIDE; ++i) CUDA_TRY(cudaStreamCreateWithFlags(&S->side[i], cudaStreamNonBlocking));
    for (int i = 0; i < N_SIDE_EV; ++i) CUDA_TRY(cudaEventCreateWithFlags(&S->ev[i], cudaEventDisableTiming));
  }
  cudaStream_t s_loc = multi ? S->side[0] : st;
  // `to` continues after everything queued on `from` so far (a wait captures the record that precedes it, so the ring is safe)
  auto dep = [&](cudaStream_t from, cudaStream_t to) -> int {
    if (from == to) return 0;
    cudaEvent_t e = S->ev[S->ev_next++ % N_SIDE_EV];
    CUDA_TRY(cudaEventRecord(e, from));
    CUDA_TRY(cudaStreamWaitEvent(to, e, 0));
    return 0;
  };
  int64_t floats = 0;
  int64_t ra[2][4], lat_ra[2] = {0, 0}, edge_ra[2] = {0, 0}, epool_ra[2] = {0, 0};
  int nlo[2] = {0, 0}, nhi[2] = {0, 0};
  for (int br = 0; br < 2; ++br) {
    const BranchDev& B = m->br[br];
    ra[br][0] = rows_alloc(chunk, B.L1);
    if (use_lat) {
      const int ps = B.pool[0][1];
      lat_ra[br] = (2 * ps * (cdiv(snv_dense_cap(chunk), ps) + 1) + 1 + 15) & ~int64_t(7);
      edge_ra[br] = rows_alloc(chunk, LAT_EL);
      if (lat_ra[br] + edge_ra[br] > ra[br][0]) ra[br][0] = lat_ra[br] + edge_ra[br];
      // pool-2 bins that touch edge rows: p < nlo or p >= L2 - nhi (a bin is lattice-only iff its unclipped window lies
      // inside stage-1 rows [LAT_EO, L1 - LAT_EO))
      const int pk2 = B.pool[1][0], ps2 = B.pool[1][1], pp2 = B.pool[1][2];
      nlo[br] = 0;
      while (nlo[br] < B.L2 && nlo[br] * ps2 - pp2 < LAT_EO) ++nlo[br];
      nhi[br] = 0;
      while (nhi[br] < B.L2 - nlo[br] && (B.L2 - 1 - nhi[br]) * ps2 - pp2 + pk2 > B.L1 - LAT_EO) ++nhi[br];
      epool_ra[br] = (chunk * (nlo[br] + nhi[br]) + 15) & ~int64_t(7);
      floats += 16 * epool_ra[br];
    }
    ra[br][1] = ra[br][0];
    ra[br][2] = rows_alloc(chunk, B.L2);
    ra[br][3] = rows_alloc(chunk, B.L3);
    for (int k = 0; k < 4; ++k) floats += (k < 3 ? 16 : 32) * ra[br][k];   // X0, Z1, Z2 are bf16 planes, H is fp32 planes
  }
  // local branch runs as a pre-pass over super-chunks (one launch fills the GPU; per 4096-site chunk it cannot)
  const int64_t super = n < (int64_t(1) << 20) ? n : (int64_t(1) << 20);
  floats += chunk * (2 * NC + 64) + super * (NC + m->n_cat) + 64;
  const int64_t dense_floats = use_dense ? int64_t((snv_dense_bytes(m, chunk) + 255) / 4 + 192) : 0;
  floats += dense_floats;
  if (int rc = snv_ensure_workspace(m, floats * 4 + 256)) return rc;
  float* w = (float*)m->d_ws;
  float* bufs[2][4];
  for (int br = 0; br < 2; ++br)
    for (int k = 0; k < 4; ++k) { bufs[br][k] = w; w += (k < 3 ? 16 : 32) * ra[br][k]; }
  // lattice / edge sub-buffers: plane stride is the sub-buffer's own row count, carved out of X0 (inputs) and Z1 (outputs)
  LatticeBufs lb[2];
  for (int br = 0; br < 2; ++br) {
    lb[br].lat_in = bufs[br][0];
    lb[br].lat_out = bufs[br][1];
    lb[br].lat_ra = lat_ra[br];
    lb[br].edge_in = bufs[br][0] + 16 * lat_ra[br];
    lb[br].edge_out = bufs[br][1] + 16 * lat_ra[br];
    lb[br].edge_ra = edge_ra[br];
  }
  float* epool[2];
  for (int br = 0; br < 2; ++br) { epool[br] = w; w += 16 * epool_ra[br]; }
  float* llog = w; w += super * NC;
  float* tl0 = w; w += chunk * NC;
  float* tl1 = w; w += chunk * NC;
  float* tg0 = w; w += chunk * 32;
  float* tg1 = w; w += chunk * 32;
  int32_t* cat32 = (int32_t*)w; w += super * m->n_cat;
  int* err_flag = (int*)w; w += 64;
  void* dense_scratch = use_dense ? (void*)((uintptr_t(w) + 255) & ~uintptr_t(255)) : nullptr;  // uint4 rows / int64 header
  CUDA_TRY(cudaMemsetAsync(err_flag, 0, 4, st));
  if (use_dense) CUDA_TRY(cudaMemsetAsync(dense_scratch, 0, 256, st));  // ChunkInfo + k_chunk_span scratch words

  if (int rc = dep(st, s_loc)) return rc;  // the local / tail stream starts behind the caller's queue and the memsets above
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int64_t ns = (n - s0 < chunk) ? (n - s0) : chunk;
    if (s0 % super == 0) {  // (chunk divides 2^20 or n <= 2^20, so chunks never straddle a super-chunk)
      // s_loc runs in order behind the tails that read the previous super-chunk's logits, and beside the stem of this chunk
      const int64_t nsup = (n - s0 < super) ? (n - s0) : super;
      if (!d_cat)
        if (int rc = snv_local_idx_launch(m, G, d_pos + s0, d_meta + s0, nsup, cat32, s_loc)) return rc;
      int rc = use_mlp_tc ? snv_local_launch_tc(m, d_cat ? nullptr : cat32, d_cat ? d_cat + s0 * m->n_cat : nullptr, nsup, llog, err_flag, s_loc) : -1;
      if (rc < 0) rc = snv_local_launch(m, d_cat ? nullptr : cat32, d_cat ? d_cat + s0 * m->n_cat : nullptr, nsup, llog, err_flag, s_loc);
      if (rc) return rc;
    }
    const float* llog_c = llog + (s0 % super) * NC;
    const int* dense_flag = nullptr;
    const ChunkInfo* info = nullptr;
    if (use_dense)
      if (int rc = snv_dense_stem_launch(m, G, d_pos + s0, d_meta + s0, ns, chunk, bufs[0][0], ra[0][0], bufs[1][0], ra[1][0],
                                         dense_scratch, &dense_flag, st, use_lat ? lb : nullptr, &info))
        return rc;
    if (int rc = snv_stem_launch_planes(m, G, d_pos ? d_pos + s0 : nullptr, d_meta ? d_meta + s0 : nullptr,
                                        d_sym ? d_sym + s0 * m->L : nullptr, ns, bufs[0][0], ra[0][0], bufs[1][0], ra[1][0],
                                        nullptr, st, /*out_bf16=*/true, dense_flag))
      return rc;
    // per-site arguments of stage 1 (two ResBlocks + outer skip at length L1) and stage 2 (pool2 fused in the loader + conv2 + two
    // ResBlocks + skip at length L2); on a dense chunk the per-site kernels exit on the device and the lattice / edge roles run
    StageArgs a1[2], a2[2];
    for (int br = 0; br < 2; ++br) {
      const BranchDev& B = m->br[br];
      StageArgs a{};
      a.lat_branch = -1;
      a.in = reinterpret_cast<const uint4*>(bufs[br][0]); a.out = bufs[br][1]; a.wblob = S->blob[br][0];
      a.in_rows_alloc = ra[br][0]; a.out_rows_alloc = ra[br][1];
      a.rows = rows_of(ns, B.L1); a.L = B.L1; a.Lin = B.L1; a.pk = 0; a.ps = 1; a.pp = 0;
      a.n_tiles = (int)cdiv(a.rows, stride_rb4);
      if (use_lat) { a.info = info; a.want = 0; }
      a1[br] = a;
      a.in = reinterpret_cast<const uint4*>(bufs[br][1]); a.out = bufs[br][2]; a.wblob = S->blob[br][1];
      a.in_rows_alloc = ra[br][1]; a.out_rows_alloc = ra[br][2];
      a.rows = rows_of(ns, B.L2); a.L = B.L2; a.Lin = B.L1; a.pk = B.pool[1][0]; a.ps = B.pool[1][1]; a.pp = B.pool[1][2];
      a.n_tiles = (int)cdiv(a.rows, stride_crb4);
      a2[br] = a;
    }
    // stage 3 (pool3 + conv3 + ReLU at length L3) as a stage kernel, for shapes the warp-level tail kernel does not take
    auto stage3 = [&](int br) -> int {
      if (use_tail) return 0;
      const BranchDev& B = m->br[br];
      StageArgs a = a2[br];
      a.info = nullptr;
      a.in = reinterpret_cast<const uint4*>(bufs[br][2]); a.out = bufs[br][3]; a.wblob = S->blob[br][2];
      a.in_rows_alloc = ra[br][2]; a.out_rows_alloc = ra[br][3];
      a.rows = rows_of(ns, B.L3); a.L = B.L3; a.Lin = B.L2; a.pk = B.pool[2][0]; a.ps = B.pool[2][1]; a.pp = B.pool[2][2];
      a.n_tiles = (int)cdiv(a.rows, TILE - 2 * 1);
      return launch_stage<SINGLE>(a, st);
    };
    if (use_lat) {
      StageArgs l2[2];
      EdgePool ep[2];
      const uint4* tab0 = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(dense_scratch) + 256);
      for (int br = 1; br >= 0; --br) {
        const BranchDev& B = m->br[br];
        if (int rc = stage(RB4, 0, a1[br], "/site")) return rc;
        // stage 1 on the lattice (geometry read from info on the device; grid sized for the largest lattice) ...
        StageArgs l = a1[br];
        l.want = 1; l.lat_branch = br;
        l.in = nullptr; l.out = lb[br].lat_out;   // input rows: the full-bin stem tables, read in place by the loader
        l.in_rows_alloc = l.out_rows_alloc = lb[br].lat_ra;
        l.n_tiles = (int)cdiv(lb[br].lat_ra, stride_rb4);
        for (int sd = 0; sd < 2; ++sd) l.tab[sd] = tab0 + int64_t((sd * 2 + br) * 3) * snv_dense_cap(chunk) * 4;
        l.ps1 = B.pool[0][1];
        if (int rc = stage(RB4, 1, l, "/lattice")) return rc;
        // ... whose pool-2 maxima go to the lattice-shaped part of the stem buffer
        LAUNCH(k_lattice_pool, 148 * pool_grid, 256, 0, st, info, br, B.pool[1][0], reinterpret_cast<const uint4*>(lb[br].lat_out),
               reinterpret_cast<uint4*>(lb[br].lat_in), lb[br].lat_ra);
      }
      for (int br = 1; br >= 0; --br) {
        const BranchDev& B = m->br[br];
        // stage 1 on the per-site edge pseudo-sites
        StageArgs e = a1[br];
        e.want = 1;
        e.in = nullptr; e.out = lb[br].edge_out;
        e.in_rows_alloc = e.out_rows_alloc = lb[br].edge_ra;
        e.rows = rows_of(ns, LAT_EL); e.L = e.Lin = LAT_EL;
        e.n_tiles = (int)cdiv(e.rows, stride_rb4);
        for (int sd = 0; sd < 2; ++sd) e.tab[sd] = tab0 + int64_t((sd * 2 + br) * 3) * snv_dense_cap(chunk) * 4;
        e.special = reinterpret_cast<const uint4*>(lb[br].edge_in); e.special_ra = lb[br].edge_ra;
        e.L1real = B.L1;
        e.pos = d_pos + s0; e.meta = d_meta + s0;
        e.ps1 = B.pool[0][1]; e.pp1 = B.pool[0][2]; e.pk1 = B.pool[0][0];
        e.off0 = br ? 0 : m->L / 2 - 100; e.R = m->cfg.distal_radius; e.br = br;
        if (int rc = stage(RB4, 2, e, "/edge")) return rc;
        // pool-2 maxima of the bins that touch edge rows
        StageArgs l = a2[br];
        l.want = 1;
        l.lat = reinterpret_cast<const uint4*>(lb[br].lat_out); l.lat_ra = lb[br].lat_ra;
        l.lat2 = reinterpret_cast<const uint4*>(lb[br].lat_in);
        l.edge = reinterpret_cast<const uint4*>(lb[br].edge_out); l.edge_ra = lb[br].edge_ra;
        l.pos = d_pos + s0; l.meta = d_meta + s0;
        l.ps1 = B.pool[0][1]; l.pp1 = B.pool[0][2]; l.pk1 = B.pool[0][0];
        l.off0 = br ? 0 : m->L / 2 - 100; l.R = m->cfg.distal_radius; l.br = br;
        l.epool = reinterpret_cast<const uint4*>(epool[br]); l.epool_ra = epool_ra[br]; l.nlo = nlo[br]; l.nhi = nhi[br];
        l2[br] = l;
        ep[br] = EdgePool{l.lat, l.lat_ra, l.edge, l.edge_ra, reinterpret_cast<uint4*>(epool[br]), epool_ra[br], l.pos, l.meta, ns,
                          br, B.L1, B.L2, B.pool[1][0], B.pool[1][1], B.pool[1][2], nlo[br], nhi[br], l.ps1, l.pp1, l.pk1, l.off0, l.R};
        LAUNCH(k_edge_pool, 148 * pool_grid, 256, 0, st, info, ep[br]);
      }
      if (s0 > 0)
        if (int rc = dep(s_loc, st)) return rc;  // the previous chunk's tail still reads the stage-2 rows
      for (int br = 1; br >= 0; --br) {
        if (int rc = stage(C_RB4, 0, a2[br], "/site")) return rc;
        if (int rc = stage(C_RB4, 1, l2[br], "/lattice")) return rc;
        if (int rc = stage3(br)) return rc;
      }
    } else {
      for (int br = 1; br >= 0; --br) {
        const BranchDev& B = m->br[br];
        const char* sfx = br ? "_2" : "";
        if (int rc = save_tap_planes(m, (std::string("pool1") + sfx).c_str(), bufs[br][0], true, ra[br][0], ns, B.L1, st)) return rc;
        if (int rc = stage(RB4, 0, a1[br], "")) return rc;
        if (int rc = save_tap_planes(m, (std::string("rb1") + sfx).c_str(), bufs[br][1], true, ra[br][1], ns, B.L1, st)) return rc;
        if (int rc = stage(C_RB4, 0, a2[br], "")) return rc;
        if (int rc = save_tap_planes(m, (std::string("rb2") + sfx).c_str(), bufs[br][2], true, ra[br][2], ns, B.L2, st)) return rc;
        if (int rc = stage3(br)) return rc;
      }
    }
    if (int rc = dep(st, s_loc)) return rc;  // tail (and heads) behind the local branch, beside the next chunk's stem
    if (use_tail) {
      if (int rc = snv_tail_launch(m, bufs[0][2], ra[0][2], bufs[1][2], ra[1][2], llog_c, ns, d_logp + s0 * NC, m->debug ? tg0 : nullptr,
                                   m->debug ? tg1 : nullptr, m->debug ? tl0 : nullptr, m->debug ? tl1 : nullptr, s_loc))
        return rc;
    } else {
      HeadTc hb[2] = {{bufs[0][3], ra[0][3], m->br[0].Wfc, m->br[0].bfc, m->br[0].L3},
                      {bufs[1][3], ra[1][3], m->br[1].Wfc, m->br[1].bfc, m->br[1].L3}};
      LAUNCH(k_head_tc, (unsigned)cdiv(ns * 8, 256), 256, 0, st, hb[0], hb[1], llog_c, ns, NC, d_logp + s0 * NC,
             m->debug ? tg0 : nullptr, m->debug ? tg1 : nullptr, m->debug ? tl0 : nullptr, m->debug ? tl1 : nullptr);
    }
    if (m->debug) {
      auto flat = [&](const char* nm, const float* d, int64_t k) -> int {
        std::vector<float>& v = m->tap_store[nm];
        v.resize(k);
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(v.data(), d, k * 4, cudaMemcpyDeviceToHost));
        return 0;
      };
      if (int rc = flat("gmax", tg0, ns * 32)) return rc;
      if (int rc = flat("gmax_2", tg1, ns * 32)) return rc;
      if (int rc = flat("logit_mid", tl0, ns * NC)) return rc;
      if (int rc = flat("logit_large", tl1, ns * NC)) return rc;
      if (int rc = flat("logit_local", llog_c, ns * NC)) return rc;
    }
  }
  if (int rc = dep(s_loc, st)) return rc;
  CUDA_TRY(cudaGetLastError());
  if (d_cat) {
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, err_flag, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    MURAL_CHECK((flag & 2) == 0, "IndexError: index out of range in embedding lookup");
  }
  return 0;
}

}  // namespace mural

// scratch builds with -DMURAL_TC_TIMING: prints the per-phase cycle breakdown of the stage kernels and clears it
extern "C" int mural_tc_timing_dump(void) {
#ifdef MURAL_TC_TIMING
  using namespace mural::tc;
  static unsigned long long h[3][148][NGROUP][2][T_N + 2];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_tc_timing, sizeof(h));
  const char* nm[T_N] = {"wait", "ld", "epi", "fence", "bar", "issue", "begin", "fetch", "final", "other"};
  for (int mode = 0; mode < 3; ++mode)
    for (int w = 0; w < 2; ++w) {
      double tot[T_N + 2] = {0};
      for (int b = 0; b < 148; ++b)
        for (int g = 0; g < NGROUP; ++g)
          for (int i = 0; i < T_N + 2; ++i) tot[i] += double(h[mode][b][g][w][i]);
      if (tot[T_N] == 0) continue;
      fprintf(stderr, "mode %d warp %d: layers/thread-sum %.0f, cycles per tile-layer (per group):", mode, w, tot[T_N]);
      for (int i = 0; i < T_N; ++i) fprintf(stderr, " %s=%.0f", nm[i], tot[i] / tot[T_N]);
      fprintf(stderr, " | total=%.0f\n", tot[T_N + 1] / tot[T_N]);
    }
  static unsigned long long z[3][148][NGROUP][2][T_N + 2];
  cudaMemcpyToSymbol(g_tc_timing, z, sizeof(z));
  return 1;
#else
  return 0;
#endif
}

extern "C" int mural_snv_tc_available(const mural_snv_model_t* m) { return (m && m->tc) ? 1 : 0; }
