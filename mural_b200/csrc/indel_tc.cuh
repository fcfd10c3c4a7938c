// MuRaL-indel UNet_Small eval forward on tensor cores: one kernel per U-Net level (MuRaL/model/model_indel.py:158-172).
//
// A level is  lconv (Conv1d k=ks, stride s | nearest upsampling by u, BatchNorm folded)  ->  ConvBlock
// (model_indel.py:6-19:  x + BN(Conv1x1(SiLU(BN(Conv5(x))))))  [+ encoder skip in the decoder], and for the last decoder level
// also out_conv (two 1x1 convs, ReLU / Softplus) and the max over positions (model_indel.py:172-173).  The fp32 path
// (indel.cu, k_conv_gen) runs these as 3-5 kernels that each write and re-read [site][L][C] fp32 tensors; here a CTA takes
// a tile of positions of one site through the whole level:
//
//   phase 1  lconv as an implicit GEMM  A[RA x C] = X[RA x ks*Cin] * Wl  (mma.sync m16n8k16): the input rows are staged
//            channels-last in shared memory, so the K index  k = tap*Cin + ci  of output row m is the contiguous run starting
//            at input row m*s — one ldmatrix per (m-tile, 16 k) with no im2col;  A goes to shared memory (fp32 for the
//            residual, split bf16 for the next conv);
//   phase 2  Conv5 from shared memory, then SiLU, Conv1x1, residual, skip and (last level) both out_conv layers chained in
//            REGISTERS: the m16n8 accumulator fragments of two neighbouring column tiles are exactly the A fragment of the
//            next 16-wide k chunk, so the hidden tensor H [L x 2C] never exists outside the register file;
//            max over positions: softplus is monotonic, so the kernel reduces the pre-activation and the head applies it.
//
// Precision: every operand is split  x = hi + lo  into two bf16 values and multiplied as  lo*hi + hi*lo + hi*hi  with fp32
// accumulation (16 mantissa bits per operand): "fp32-equivalent", measured 3e-5 .. 1.5e-4 of the output scale over the 13
// golden checkpoints in a CPU emulation of exactly this data flow (scratch/indel_precision.py) — a single bf16 or fp16
// product per term does NOT hold the 1e-3 gate (1.4e-2 / 5e-3 on hs_del_start).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace mural {
namespace indel_tc {

#ifndef MURAL_INDEL_MINB1
#define MURAL_INDEL_MINB1 4
#endif
#ifndef MURAL_INDEL_MINB2
#define MURAL_INDEL_MINB2 2
#endif
constexpr int RA_MAX = 256;     // A rows (lconv outputs incl. the +-2 halo of Conv5) per tile

// pitch (elements) of a channels-last shared-memory row of c elements (c % 8 == 0) such that pitch/8 is odd:
// 8 consecutive rows of 16 bytes (ldmatrix) / of a float2 quad land in distinct bank groups
__host__ __device__ constexpr int pitch8(int c) { return ((c >> 3) & 1) ? c : c + 8; }

struct LevelParams {
  const float* in;    // [site][Lin][Cin] fp32
  const float* skip;  // [site][Lout][C] encoder output added at the end (decoder) or NULL
  float* out;         // [site][Lout][C] (not TAIL)
  float* gmax;        // [site][C] running max of the pre-Softplus out_conv output (TAIL), pre-filled with -inf
  const uint4* Wl;    // B fragments [F phases][KCl][NC8][32 lanes] {b0 hi, b1 hi, b0 lo, b1 lo}
  const uint4* W5;    // [KC5][2*NC8][32]
  const uint4* W1;    // [NC8][NC8][32]
  const uint4* Wo0;   // [KCo][NC8][32] (TAIL)
  const uint4* Wo1;
  const float* bias;  // bl[C] | b5[2C] | b1[C] | bo0[C] | bo1[C]
  int Cin, CinP, ks, stride, up;
  uint32_t magic_stride, magic_up;  // ceil(2^32 / d): x / d == __umulhi(x, magic) for 0 <= x < 2^20, d <= 16 (d == 1 handled apart)
  int Lin, Lout;
  int KCl, KC5, KCo;
  int TP, RA, n_tiles;
  int F, HA, dmin;    // decoder levels fold the nearest upsampling by F into the weights (see phase 1): F phases of RA/F rows; HA = rows
                      // of A in front of the tile's first output (2 = the Conv5 halo, or F when folded); dmin = first low-res tap offset
  int SG;             // sites per work item (> 1 only when n_tiles == 1: short levels share a CTA so that every warp has a tile)
  int rows_in;        // staged input rows per tile and site: (RA-1)*stride + ks + 1
  int RS;             // rows per stride phase of the staged input: ceil(rows_in / stride)
  int64_t n_sites;
  int64_t n_items;    // ceil(n_sites / SG) * n_tiles
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t a) {  // a: shared-space byte address
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xFFFF0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - hx, y - hy);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float silu(float x) {  // x * rcp(1 + 2^(-x log2 e)): two MUFU ops, ~3 ulp (x -> -inf: x * 0)
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ int div_magic(int x, int d, uint32_t magic) { return d == 1 ? x : int(__umulhi(uint32_t(x), magic)); }

// The three products of the two-level split for MT row tiles x one column tile, small terms first, issued term-major across
// the row tiles so that an MMA does not wait on the one just before it.  (Separate accumulators for the cross terms were
// measured: +80 registers, occupancy 40 -> 12 warps per SM, 35 % slower.)
template <int MT>
__device__ __forceinline__ void mma3(float (&d)[MT][4], const uint32_t (&ah)[MT][4], const uint32_t (&al)[MT][4], const uint4& b) {
#ifdef MURAL_INDEL_MMA_CHAIN
#pragma unroll
  for (int i = 0; i < MT; ++i) { mma_bf16(d[i], al[i], b.x, b.y); mma_bf16(d[i], ah[i], b.z, b.w); mma_bf16(d[i], ah[i], b.x, b.y); }
#else
#pragma unroll
  for (int i = 0; i < MT; ++i) mma_bf16(d[i], al[i], b.x, b.y);
#pragma unroll
  for (int i = 0; i < MT; ++i) mma_bf16(d[i], ah[i], b.z, b.w);
#pragma unroll
  for (int i = 0; i < MT; ++i) mma_bf16(d[i], ah[i], b.x, b.y);
#endif
}

// accumulator fragments of column tiles (2j, 2j+1) -> split A fragment of k chunk j after f(acc + bias)
template <int ACT>  // 0 none, 1 SiLU, 2 ReLU
__device__ __forceinline__ void chain_frag(const float* c0, const float* c1, const float* bias0, const float* bias1, int q, uint32_t (&ah)[4],
                                           uint32_t (&al)[4]) {
  float v[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = c0[e] + (bias0 ? bias0[2 * q + (e & 1)] : 0.f);
#pragma unroll
  for (int e = 0; e < 4; ++e) v[4 + e] = c1 ? c1[e] + (bias1 ? bias1[2 * q + (e & 1)] : 0.f) : 0.f;
  if (ACT == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = silu(v[e]);
  } else if (ACT == 2) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  split2(v[0], v[1], ah[0], al[0]);  // row g,   k 2q..2q+1
  split2(v[2], v[3], ah[1], al[1]);  // row g+8
  split2(v[4], v[5], ah[2], al[2]);  // row g,   k 2q+8..
  split2(v[6], v[7], ah[3], al[3]);  // row g+8
}

// shared-memory carve-up, identical on host and device
struct Smem {
  int oWl, oW5, oW1, oWo0, oWo1, oBias, oOff, oXhi, oXlo, oAf, oAhi, oAlo, total;
  int PinP, PA, PC, xs, as;  // xs / as: elements per site of the X / A operand arrays
};
__host__ __device__ inline Smem smem_layout(int NC8, bool tail, const LevelParams& P) {
  Smem s;
  const int C = 8 * NC8;
  int o = 0;
  s.oWl = o; o += P.F * P.KCl * NC8 * 512;
  s.oW5 = o; o += P.KC5 * 2 * NC8 * 512;
  s.oW1 = o; o += NC8 * NC8 * 512;
  s.oWo0 = o; o += tail ? P.KCo * NC8 * 512 : 0;
  s.oWo1 = o; o += tail ? P.KCo * NC8 * 512 : 0;
  s.oBias = o; o += 6 * C * 4;
  s.oOff = o; o += (P.KCl + P.KC5) * 2 * 4;  // element offset of (k chunk, k half) inside the X / A operand: no div / mod in the k loops
  s.PinP = pitch8(P.CinP);
  s.PA = pitch8(C);
  s.PC = pitch8(C);
  s.xs = P.RS * P.stride * s.PinP;  // stride phases of RS rows each
  s.as = (P.RA + 8 + P.HA) * s.PC;  // + zero rows behind the tile: k overrun of the last taps
  o = (o + 15) & ~15;
  s.oXhi = o; o += P.SG * s.xs * 2;
  s.oXlo = o; o += P.SG * s.xs * 2;
  s.oAf = o; o += P.SG * P.RA * s.PA * 4;
  s.oAhi = o; o += P.SG * s.as * 2;
  s.oAlo = o; o += P.SG * s.as * 2;
  s.total = o;
  return s;
}

template <int NC8, int MT, int NW, bool TAIL>
__global__ void __launch_bounds__(NW * 32, NC8 == 1 ? MURAL_INDEL_MINB1 : NC8 == 2 ? MURAL_INDEL_MINB2 : 1) k_unet_level(const LevelParams P) {
  constexpr int C = 8 * NC8, NH8 = 2 * NC8, THREADS = NW * 32;
  extern __shared__ __align__(16) unsigned char smraw[];
  const Smem S = smem_layout(NC8, TAIL, P);
  const uint4* sWl = reinterpret_cast<const uint4*>(smraw + S.oWl);
  const uint4* sW5 = reinterpret_cast<const uint4*>(smraw + S.oW5);
  const uint4* sW1 = reinterpret_cast<const uint4*>(smraw + S.oW1);
  const uint4* sWo0 = reinterpret_cast<const uint4*>(smraw + S.oWo0);
  const uint4* sWo1 = reinterpret_cast<const uint4*>(smraw + S.oWo1);
  const float* sB = reinterpret_cast<const float*>(smraw + S.oBias);
  const float *bl = sB, *b5 = sB + C, *b1 = sB + 3 * C, *bo0 = sB + 4 * C, *bo1 = sB + 5 * C;
  __nv_bfloat16* Xhi = reinterpret_cast<__nv_bfloat16*>(smraw + S.oXhi);
  __nv_bfloat16* Xlo = reinterpret_cast<__nv_bfloat16*>(smraw + S.oXlo);
  float* Af = reinterpret_cast<float*>(smraw + S.oAf);
  __nv_bfloat16* Ahi = reinterpret_cast<__nv_bfloat16*>(smraw + S.oAhi);
  __nv_bfloat16* Alo = reinterpret_cast<__nv_bfloat16*>(smraw + S.oAlo);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lkh = lane >> 4;  // ldmatrix: row of this lane's address, k half
  const int PinP = S.PinP, PA = S.PA, PC = S.PC;
  const int mt_site = P.RA >> 4;  // row tiles per site

  // ---- once per CTA: weights (pre-split B fragments, straight copy), biases, zero tail rows of the A operand
  {
    uint4* d = reinterpret_cast<uint4*>(smraw);
    const int nl = P.F * P.KCl * NC8 * 32, n5 = P.KC5 * NH8 * 32, n1 = NC8 * NC8 * 32, no = TAIL ? P.KCo * NC8 * 32 : 0;
    for (int e = tid; e < nl; e += THREADS) d[(S.oWl >> 4) + e] = __ldg(P.Wl + e);
    for (int e = tid; e < n5; e += THREADS) d[(S.oW5 >> 4) + e] = __ldg(P.W5 + e);
    for (int e = tid; e < n1; e += THREADS) d[(S.oW1 >> 4) + e] = __ldg(P.W1 + e);
    for (int e = tid; e < no; e += THREADS) { d[(S.oWo0 >> 4) + e] = __ldg(P.Wo0 + e); d[(S.oWo1 >> 4) + e] = __ldg(P.Wo1 + e); }
    float* db = reinterpret_cast<float*>(smraw + S.oBias);
    for (int e = tid; e < 6 * C; e += THREADS) db[e] = (e < (TAIL ? 6 : 4) * C) ? __ldg(P.bias + e) : 0.f;
    for (int sg = 0; sg < P.SG; ++sg)
      for (int e = tid; e < (8 + P.HA) * PC; e += THREADS) {
        Ahi[sg * S.as + P.RA * PC + e] = __float2bfloat16(0.f);
        Alo[sg * S.as + P.RA * PC + e] = __float2bfloat16(0.f);
      }
    int* doff = reinterpret_cast<int*>(smraw + S.oOff);
    for (int e = tid; e < 2 * (P.KCl + P.KC5); e += THREADS) {
      const bool l = e < 2 * P.KCl;
      const int k = (l ? e : e - 2 * P.KCl) * 8, cw = l ? P.CinP : C;
      const int t = k / cw, c = k - t * cw;
      const int td = l ? t / P.stride : 0;
      doff[e] = l ? ((t - td * P.stride) * P.RS + td) * PinP + c : t * PC + c;
    }
  }
  const int* offL = reinterpret_cast<const int*>(smraw + S.oOff);
  const int* off5 = offL + 2 * P.KCl;
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(smraw);
  const int half = P.ks >> 1;
  const int Lv = P.Lin * P.up;
  const int cg_in = P.CinP >> 2;  // 4-channel groups per staged row
  const int n_stage = P.RS * P.stride * cg_in;  // every slot of the stride-phase layout (a bijection of j < RS*stride)
  const int st_dj = THREADS / cg_in, st_dc = THREADS - st_dj * cg_in;  // (row, group) advance of one staging step

  int64_t grp = blockIdx.x / P.n_tiles;
  int tile = blockIdx.x - int(grp) * P.n_tiles;
  const int d_grp = gridDim.x / P.n_tiles, d_tile = gridDim.x - d_grp * P.n_tiles;
  for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x, grp += d_grp, tile += d_tile) {
    if (tile >= P.n_tiles) { tile -= P.n_tiles; ++grp; }
    const int64_t site0 = grp * P.SG;
    const int p0 = tile * P.TP;
    __syncthreads();  // previous item's readers of X / A are done (first pass: orders the weight staging)
    // ---- stage the input rows of this tile: virtual (upsampled) positions v0 + j, split to bf16 hi / lo.  Row j is stored at
    //      slot (j % stride) * RS + j / stride: the rows that consecutive outputs read for one tap are consecutive 16-byte-pitch
    //      slots (conflict-free ldmatrix for any stride).
    {
      const int v0 = P.F > 1 ? p0 / P.F - 1 + P.dmin : (p0 - P.HA) * P.stride - half;   // folded levels stage low-resolution rows
      for (int sg = 0; sg < P.SG; ++sg) {
        const int64_t site = site0 + sg;
        const bool site_ok = site < P.n_sites;
        const float* ins = P.in + site * int64_t(P.Lin) * P.Cin;
        __nv_bfloat16 *xh = Xhi + sg * S.xs, *xl = Xlo + sg * S.xs;
        int j = tid / cg_in, cgi = tid - j * cg_in;
        for (int e = tid; e < n_stage; e += THREADS) {
          const int v = v0 + j, c4 = cgi * 4;
          float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
          if (site_ok && v >= 0 && v < Lv && c4 < P.Cin)
            x = __ldg(reinterpret_cast<const float4*>(ins + int64_t(div_magic(v, P.up, P.magic_up)) * P.Cin + c4));
          const int jd = div_magic(j, P.stride, P.magic_stride);
          const int slot = (j - jd * P.stride) * P.RS + jd;
          uint2 h, l;
          split2(x.x, x.y, h.x, l.x);
          split2(x.z, x.w, h.y, l.y);
          *reinterpret_cast<uint2*>(xh + slot * PinP + c4) = h;
          *reinterpret_cast<uint2*>(xl + slot * PinP + c4) = l;
          j += st_dj; cgi += st_dc;
          if (cgi >= cg_in) { cgi -= cg_in; ++j; }
        }
      }
    }
    __syncthreads();
    // ---- phase 1: A = lconv(X) for rows [p0-HA, p0-HA+RA) of every site of the item.
    //      Decoder levels (nn.Upsample(scale_factor=F) in front of the conv, model_indel.py:90-117): output position F*m + ph reads
    //      the low-resolution rows m + floor((ph + t - half) / F), t = 0..ks-1 — at most ceil((ks-1)/F)+1 distinct rows — so the
    //      taps that hit the same row are summed on the host into one weight per (phase ph, row offset): K shrinks from ks*Cin to
    //      ND*Cin (112 -> 48 at the top level) and nothing is replicated in shared memory.  Rows of one phase form the row tiles
    //      (tile index = ph * RP/16 + ...); the zero padding of the upsampled signal coincides with that of the low-res rows.
    const int mt_phase = mt_site / P.F;  // row tiles per phase
    for (int mt0 = warp * MT; mt0 < P.SG * mt_site; mt0 += NW * MT) {
      const int sg = P.SG == 1 ? 0 : mt0 / mt_site, ml = mt0 - sg * mt_site;
      const int ph = P.F == 1 ? 0 : ml / mt_phase, mi0 = (ml - ph * mt_phase) * 16;
      const uint32_t xh = sm_base + S.oXhi + 2 * (sg * S.xs + (mi0 + lrow) * PinP), xl = xh + (S.oXlo - S.oXhi);
      const uint4* wl = sWl + ph * P.KCl * NC8 * 32;
      float accm[NC8][MT][4];
#pragma unroll
      for (int n = 0; n < NC8; ++n)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) accm[n][i][e] = 0.f;
      for (int kc = 0; kc < P.KCl; ++kc) {
        uint32_t ah[MT][4], al[MT][4];
        const int base = 2 * offL[2 * kc + lkh];  // (tap, channel) of k = kc*16 + 8*lkh as a byte offset
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          ldsm4(ah[i], xh + base + i * 32 * PinP);
          ldsm4(al[i], xl + base + i * 32 * PinP);
        }
#pragma unroll
        for (int n = 0; n < NC8; ++n) {
          const uint4 b = wl[(kc * NC8 + n) * 32 + lane];
          mma3<MT>(accm[n], ah, al, b);
        }
      }
      float* af = Af + sg * P.RA * PA;
      __nv_bfloat16 *ahs = Ahi + sg * S.as, *als = Alo + sg * S.as;
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int r = (mi0 + i * 16 + g + 8 * hrow) * P.F + ph;
          const int pos = p0 - P.HA + r;
          const bool valid = pos >= 0 && pos < P.Lout;  // Conv5 pads A with zeros outside the sequence
#pragma unroll
          for (int n = 0; n < NC8; ++n) {
            const int col = n * 8 + 2 * q;
            const float v0 = valid ? accm[n][i][2 * hrow] + bl[col] : 0.f;
            const float v1 = valid ? accm[n][i][2 * hrow + 1] + bl[col + 1] : 0.f;
            *reinterpret_cast<float2*>(af + r * PA + col) = make_float2(v0, v1);
            uint32_t h, l;
            split2(v0, v1, h, l);
            *reinterpret_cast<uint32_t*>(ahs + r * PC + col) = h;
            *reinterpret_cast<uint32_t*>(als + r * PC + col) = l;
          }
        }
    }
    __syncthreads();
    // ---- phase 2: Conv5 -> SiLU -> Conv1x1 -> + A (+ skip) [-> out_conv -> max]
    for (int mt0 = warp * MT; mt0 < P.SG * mt_site; mt0 += NW * MT) {
      const int sg = P.SG == 1 ? 0 : mt0 / mt_site, ml = mt0 - sg * mt_site;
      if (ml * 16 >= P.TP) continue;
      const int64_t site = site0 + sg;
      const bool site_ok = site < P.n_sites;
      const uint32_t ahs = sm_base + S.oAhi + 2 * (sg * S.as + (ml * 16 + lrow + P.HA - 2) * PC), als = ahs + (S.oAlo - S.oAhi);
      const float* af = Af + sg * P.RA * PA;
      const float* skipp = P.skip ? P.skip + site * int64_t(P.Lout) * C : nullptr;
      float* outp = TAIL ? nullptr : P.out + site * int64_t(P.Lout) * C;
      float acc5[NH8][MT][4];
#pragma unroll
      for (int n = 0; n < NH8; ++n)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc5[n][i][e] = 0.f;
      for (int kc = 0; kc < P.KC5; ++kc) {
        uint32_t ah[MT][4], al[MT][4];
        const int base = 2 * off5[2 * kc + lkh];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          ldsm4(ah[i], ahs + base + i * 32 * PC);
          ldsm4(al[i], als + base + i * 32 * PC);
        }
#pragma unroll
        for (int n = 0; n < NH8; ++n) {
          const uint4 b = sW5[(kc * NH8 + n) * 32 + lane];
          mma3<MT>(acc5[n], ah, al, b);
        }
      }
      float acc1[NC8][MT][4];
#pragma unroll
      for (int n = 0; n < NC8; ++n)
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc1[n][i][e] = 0.f;
#pragma unroll
      for (int j = 0; j < NC8; ++j) {  // k chunk j of Conv1x1 = hidden channels 16j .. 16j+15 = column tiles 2j, 2j+1 of Conv5
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) chain_frag<1>(acc5[2 * j][i], acc5[2 * j + 1][i], b5 + 16 * j, b5 + 16 * j + 8, q, ah[i], al[i]);
#pragma unroll
        for (int n = 0; n < NC8; ++n) {
          const uint4 b = sW1[(j * NC8 + n) * 32 + lane];
          mma3<MT>(acc1[n], ah, al, b);
        }
      }
      // epilogue: + bias + A (residual of the ConvBlock) + encoder skip
      bool valid[MT][2];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int m = (ml + i) * 16 + g + 8 * hrow;
          const int pos = p0 + m;
          valid[i][hrow] = site_ok && m < P.TP && pos < P.Lout;
          const int o = pos * C;
#pragma unroll
          for (int n = 0; n < NC8; ++n) {
            const int col = n * 8 + 2 * q;
            float2 v = make_float2(acc1[n][i][2 * hrow] + b1[col], acc1[n][i][2 * hrow + 1] + b1[col + 1]);
            if (m + P.HA < P.RA) {
              const float2 a = *reinterpret_cast<const float2*>(af + (m + P.HA) * PA + col);
              v.x += a.x; v.y += a.y;
            }
            if (valid[i][hrow]) {
              if (P.skip) { const float2 sk = __ldg(reinterpret_cast<const float2*>(skipp + o + col)); v.x += sk.x; v.y += sk.y; }
              if (!TAIL) *reinterpret_cast<float2*>(outp + o + col) = v;
            }
            acc1[n][i][2 * hrow] = v.x;
            acc1[n][i][2 * hrow + 1] = v.y;
          }
        }
      if (TAIL) {  // out_conv.0 (+BN, ReLU) and out_conv.3 chained in registers; max of the pre-Softplus value over the positions
        float a0[NC8][MT][4], a1[NC8][MT][4];
#pragma unroll
        for (int n = 0; n < NC8; ++n)
#pragma unroll
          for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) a0[n][i][e] = a1[n][i][e] = 0.f;
#pragma unroll
        for (int j = 0; j < (NC8 + 1) / 2; ++j) {
          uint32_t ah[MT][4], al[MT][4];
#pragma unroll
          for (int i = 0; i < MT; ++i)
            chain_frag<0>(acc1[2 * j][i], (2 * j + 1 < NC8) ? acc1[(2 * j + 1 < NC8) ? 2 * j + 1 : 0][i] : nullptr, nullptr, nullptr, q, ah[i], al[i]);
#pragma unroll
          for (int n = 0; n < NC8; ++n) mma3<MT>(a0[n], ah, al, sWo0[(j * NC8 + n) * 32 + lane]);
        }
#pragma unroll
        for (int j = 0; j < (NC8 + 1) / 2; ++j) {
          uint32_t ah[MT][4], al[MT][4];
#pragma unroll
          for (int i = 0; i < MT; ++i)
            chain_frag<2>(a0[2 * j][i], (2 * j + 1 < NC8) ? a0[(2 * j + 1 < NC8) ? 2 * j + 1 : 0][i] : nullptr, bo0 + 16 * j, bo0 + 16 * j + 8, q, ah[i], al[i]);
#pragma unroll
          for (int n = 0; n < NC8; ++n) mma3<MT>(a1[n], ah, al, sWo1[(j * NC8 + n) * 32 + lane]);
        }
#pragma unroll
        for (int n = 0; n < NC8; ++n)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = -INFINITY;
#pragma unroll
            for (int i = 0; i < MT; ++i) {
              if (valid[i][0]) v = fmaxf(v, a1[n][i][e]);
              if (valid[i][1]) v = fmaxf(v, a1[n][i][2 + e]);
            }
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
            const int col = n * 8 + 2 * q + e;
            if (g == 0 && v > -INFINITY) atomic_max_float(P.gmax + site * C + col, v + bo1[col]);
          }
      }
    }
  }
}

__global__ void k_fill_f32(float* p, int64_t n, float v) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) p[i] = v;
}

// Softplus of the position max (== max of Softplus) -> BN(out_fc.0) folded into Linear(out_fc.2) -> Softplus
__global__ void __launch_bounds__(128) k_indel_head_tc(const float* __restrict__ gmax, int64_t n, int C, const float* __restrict__ Wfc,
                                                       const float* __restrict__ bfc, int NC, float* __restrict__ out) {
  const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (i >= n * NC) return;
  const int64_t site = i / NC;
  const int o = int(i - site * NC);
  float acc = bfc[o];
  for (int k = 0; k < C; ++k) {
    const float x = gmax[site * C + k];
    const float sp = x > 20.f ? x : log1pf(expf(x));
    acc = fmaf(sp, Wfc[k * NC + o], acc);
  }
  out[i] = acc > 20.f ? acc : log1pf(expf(acc));
}

}  // namespace indel_tc
}  // namespace mural
