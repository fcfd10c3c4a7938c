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

constexpr int NW = 8;           // warps per CTA
constexpr int THREADS = NW * 32;
constexpr int RA_MAX = 256;     // A rows (lconv outputs incl. the +-2 halo of Conv5) per tile

// pitch (elements) of a channels-last shared-memory row of c elements (c % 8 == 0) such that pitch/8 is odd:
// 8 consecutive rows of 16 bytes (ldmatrix) / of a float2 quad land in distinct bank groups
__host__ __device__ constexpr int pitch8(int c) { return ((c >> 3) & 1) ? c : c + 8; }

struct LevelParams {
  const float* in;    // [site][Lin][Cin] fp32
  const float* skip;  // [site][Lout][C] encoder output added at the end (decoder) or NULL
  float* out;         // [site][Lout][C] (not TAIL)
  float* gmax;        // [site][C] running max of the pre-Softplus out_conv output (TAIL), pre-filled with -inf
  const uint4* Wl;    // B fragments [KCl][NC8][32 lanes] {b0 hi, b1 hi, b0 lo, b1 lo}
  const uint4* W5;    // [KC5][2*NC8][32]
  const uint4* W1;    // [NC8][NC8][32]
  const uint4* Wo0;   // [KCo][NC8][32] (TAIL)
  const uint4* Wo1;
  const float* bias;  // bl[C] | b5[2C] | b1[C] | bo0[C] | bo1[C]
  int Cin, CinP, ks, stride, up;
  int Lin, Lout;
  int KCl, KC5, KCo;
  int TP, RA, n_tiles;
  int rows_in;        // staged input rows per tile: (RA-1)*stride + ks + 1
  int64_t n_items;    // n_sites * n_tiles
};

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// three products of the two-level split, small terms first
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], const uint4& b) {
  mma_bf16(d, al, b.x, b.y);
  mma_bf16(d, ah, b.z, b.w);
  mma_bf16(d, ah, b.x, b.y);
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xFFFF0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(x - hx, y - hy);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ float silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }  // as indel.cu apply_act
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// accumulator fragments of column tiles (2j, 2j+1) -> split A fragment of k chunk j after f(acc + bias)
template <int ACT>  // 0 none, 1 SiLU, 2 ReLU
__device__ __forceinline__ void chain_frag(const float (&c0)[4], const float* c1, const float* bias0, const float* bias1, int q,
                                           uint32_t (&ah)[4], uint32_t (&al)[4]) {
  float v[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = c0[e] + bias0[2 * q + (e & 1)];
  if (c1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[4 + e] = c1[e] + bias1[2 * q + (e & 1)];
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[4 + e] = 0.f;
  }
  if (ACT == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = silu(v[e]);
  } else if (ACT == 2) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  if (!c1) {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[4 + e] = 0.f;
  }
  split2(v[0], v[1], ah[0], al[0]);  // row g,   k 2q..2q+1
  split2(v[2], v[3], ah[1], al[1]);  // row g+8
  split2(v[4], v[5], ah[2], al[2]);  // row g,   k 2q+8..
  split2(v[6], v[7], ah[3], al[3]);  // row g+8
}

// shared-memory carve-up, identical on host and device
struct Smem {
  int oWl, oW5, oW1, oWo0, oWo1, oBias, oXhi, oXlo, oAf, oAhi, oAlo, total;
  int PinP, PA, PC;
};
__host__ __device__ inline Smem smem_layout(int NC8, bool tail, const LevelParams& P) {
  Smem s;
  const int C = 8 * NC8;
  int o = 0;
  s.oWl = o; o += P.KCl * NC8 * 512;
  s.oW5 = o; o += P.KC5 * 2 * NC8 * 512;
  s.oW1 = o; o += NC8 * NC8 * 512;
  s.oWo0 = o; o += tail ? P.KCo * NC8 * 512 : 0;
  s.oWo1 = o; o += tail ? P.KCo * NC8 * 512 : 0;
  s.oBias = o; o += 6 * C * 4;
  s.PinP = pitch8(P.CinP);
  s.PA = pitch8(C);
  s.PC = pitch8(C);
  o = (o + 15) & ~15;
  s.oXhi = o; o += ((P.rows_in * s.PinP * 2 + 15) & ~15);
  s.oXlo = o; o += ((P.rows_in * s.PinP * 2 + 15) & ~15);
  s.oAf = o; o += P.RA * s.PA * 4;
  s.oAhi = o; o += (P.RA + 8) * s.PC * 2;
  s.oAlo = o; o += (P.RA + 8) * s.PC * 2;
  s.total = o;
  return s;
}

template <int NC8, int MT, bool TAIL>
__global__ void __launch_bounds__(THREADS) k_unet_level(const LevelParams P) {
  constexpr int C = 8 * NC8, NH8 = 2 * NC8;
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

  // ---- once per CTA: weights (pre-split B fragments, straight copy), biases, zero tail rows of the A operand
  {
    uint4* d = reinterpret_cast<uint4*>(smraw);
    const int nl = P.KCl * NC8 * 32, n5 = P.KC5 * NH8 * 32, n1 = NC8 * NC8 * 32, no = TAIL ? P.KCo * NC8 * 32 : 0;
    for (int e = tid; e < nl; e += THREADS) d[(S.oWl >> 4) + e] = __ldg(P.Wl + e);
    for (int e = tid; e < n5; e += THREADS) d[(S.oW5 >> 4) + e] = __ldg(P.W5 + e);
    for (int e = tid; e < n1; e += THREADS) d[(S.oW1 >> 4) + e] = __ldg(P.W1 + e);
    for (int e = tid; e < no; e += THREADS) { d[(S.oWo0 >> 4) + e] = __ldg(P.Wo0 + e); d[(S.oWo1 >> 4) + e] = __ldg(P.Wo1 + e); }
    float* db = reinterpret_cast<float*>(smraw + S.oBias);
    for (int e = tid; e < 6 * C; e += THREADS) db[e] = (e < (TAIL ? 6 : 4) * C) ? __ldg(P.bias + e) : 0.f;
    for (int e = tid; e < 8 * PC; e += THREADS) { Ahi[P.RA * PC + e] = __float2bfloat16(0.f); Alo[P.RA * PC + e] = __float2bfloat16(0.f); }
  }
  const int half = P.ks >> 1;
  const int Lv = P.Lin * P.up;
  const int cg_in = P.CinP >> 2;  // 4-channel groups per staged row

  for (int64_t item = blockIdx.x; item < P.n_items; item += gridDim.x) {
    const int64_t site = item / P.n_tiles;
    const int tile = int(item - site * P.n_tiles);
    const int p0 = tile * P.TP;
    __syncthreads();  // previous item's readers of X / A are done (first pass: orders the weight staging)
    // ---- stage the input rows of this tile: virtual (upsampled) positions v0 .. v0 + rows_in - 1, split to bf16 hi / lo
    {
      const int v0 = (p0 - 2) * P.stride - half;
      const float* ins = P.in + site * int64_t(P.Lin) * P.Cin;
      for (int e = tid; e < P.rows_in * cg_in; e += THREADS) {
        const int j = e / cg_in, c4 = (e - j * cg_in) * 4;
        const int v = v0 + j;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v >= 0 && v < Lv && c4 < P.Cin) x = __ldg(reinterpret_cast<const float4*>(ins + int64_t(P.up == 1 ? v : v / P.up) * P.Cin + c4));
        uint2 h, l;
        split2(x.x, x.y, h.x, l.x);
        split2(x.z, x.w, h.y, l.y);
        *reinterpret_cast<uint2*>(Xhi + j * PinP + c4) = h;
        *reinterpret_cast<uint2*>(Xlo + j * PinP + c4) = l;
      }
    }
    __syncthreads();
    // ---- phase 1: A = lconv(X) for rows [p0-2, p0-2+RA)
    for (int mt0 = warp * MT; mt0 * 16 < P.RA; mt0 += NW * MT) {
      float acc[MT][NC8][4];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int n = 0; n < NC8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.f;
      int t = 0, c = 8 * lkh;  // (tap, channel) of k = kc*16 + 8*lkh
      while (c >= P.CinP) { c -= P.CinP; ++t; }
      for (int kc = 0; kc < P.KCl; ++kc) {
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int off = (((mt0 + i) * 16 + lrow) * P.stride + t) * PinP + c;
          ldsm4(ah[i], Xhi + off);
          ldsm4(al[i], Xlo + off);
        }
#pragma unroll
        for (int n = 0; n < NC8; ++n) {
          const uint4 b = sWl[(kc * NC8 + n) * 32 + lane];
#pragma unroll
          for (int i = 0; i < MT; ++i) mma3(acc[i][n], ah[i], al[i], b);
        }
        c += 16;
        while (c >= P.CinP) { c -= P.CinP; ++t; }
      }
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int r = (mt0 + i) * 16 + g + 8 * hrow;
          const int pos = p0 - 2 + r;
          const bool valid = pos >= 0 && pos < P.Lout;  // Conv5 pads A with zeros outside the sequence
#pragma unroll
          for (int n = 0; n < NC8; ++n) {
            const int col = n * 8 + 2 * q;
            const float v0 = valid ? acc[i][n][2 * hrow] + bl[col] : 0.f;
            const float v1 = valid ? acc[i][n][2 * hrow + 1] + bl[col + 1] : 0.f;
            *reinterpret_cast<float2*>(Af + r * PA + col) = make_float2(v0, v1);
            uint32_t h, l;
            split2(v0, v1, h, l);
            *reinterpret_cast<uint32_t*>(Ahi + r * PC + col) = h;
            *reinterpret_cast<uint32_t*>(Alo + r * PC + col) = l;
          }
        }
    }
    __syncthreads();
    // ---- phase 2: Conv5 -> SiLU -> Conv1x1 -> + A (+ skip) [-> out_conv -> max]
    float tmax[NC8][2];
    if (TAIL) {
#pragma unroll
      for (int n = 0; n < NC8; ++n) tmax[n][0] = tmax[n][1] = -INFINITY;
    }
    for (int mt0 = warp * MT; mt0 * 16 < P.TP; mt0 += NW * MT) {
      float acc5[MT][NH8][4];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int n = 0; n < NH8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc5[i][n][e] = 0.f;
      int t = 0, c = 8 * lkh;
      while (c >= C) { c -= C; ++t; }
      for (int kc = 0; kc < P.KC5; ++kc) {
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int off = ((mt0 + i) * 16 + lrow + t) * PC + c;
          ldsm4(ah[i], Ahi + off);
          ldsm4(al[i], Alo + off);
        }
#pragma unroll
        for (int n = 0; n < NH8; ++n) {
          const uint4 b = sW5[(kc * NH8 + n) * 32 + lane];
#pragma unroll
          for (int i = 0; i < MT; ++i) mma3(acc5[i][n], ah[i], al[i], b);
        }
        c += 16;
        while (c >= C) { c -= C; ++t; }
      }
      float acc1[MT][NC8][4];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int n = 0; n < NC8; ++n)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc1[i][n][e] = 0.f;
#pragma unroll
      for (int j = 0; j < NC8; ++j) {  // k chunk j of Conv1x1 = hidden channels 16j .. 16j+15 = column tiles 2j, 2j+1 of Conv5
        uint32_t ah[MT][4], al[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) chain_frag<1>(acc5[i][2 * j], acc5[i][2 * j + 1], b5 + 16 * j, b5 + 16 * j + 8, q, ah[i], al[i]);
#pragma unroll
        for (int n = 0; n < NC8; ++n) {
          const uint4 b = sW1[(j * NC8 + n) * 32 + lane];
#pragma unroll
          for (int i = 0; i < MT; ++i) mma3(acc1[i][n], ah[i], al[i], b);
        }
      }
      // epilogue: + bias + A (residual of the ConvBlock) + encoder skip
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        bool valid[2];
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
          const int m = (mt0 + i) * 16 + g + 8 * hrow;
          const int pos = p0 + m;
          valid[hrow] = m < P.TP && pos < P.Lout;
          const int64_t o = (site * P.Lout + pos) * int64_t(C);
#pragma unroll
          for (int n = 0; n < NC8; ++n) {
            const int col = n * 8 + 2 * q;
            float2 v = make_float2(acc1[i][n][2 * hrow] + b1[col], acc1[i][n][2 * hrow + 1] + b1[col + 1]);
            if (m + 2 < P.RA) {
              const float2 a = *reinterpret_cast<const float2*>(Af + (m + 2) * PA + col);
              v.x += a.x; v.y += a.y;
            }
            if (valid[hrow]) {
              if (P.skip) { const float2 s = __ldg(reinterpret_cast<const float2*>(P.skip + o + col)); v.x += s.x; v.y += s.y; }
              if (!TAIL) *reinterpret_cast<float2*>(P.out + o + col) = v;
            }
            acc1[i][n][2 * hrow] = v.x;
            acc1[i][n][2 * hrow + 1] = v.y;
          }
        }
        if (TAIL) {  // out_conv.0 (+BN, ReLU) and out_conv.3 chained in registers; running max of the pre-Softplus value
          float zero4[4] = {0.f, 0.f, 0.f, 0.f};
          float a0[NC8][4], a1[NC8][4];
#pragma unroll
          for (int n = 0; n < NC8; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) a0[n][e] = a1[n][e] = 0.f;
#pragma unroll
          for (int j = 0; j < (NC8 + 1) / 2; ++j) {
            uint32_t ah[4], al[4];
            chain_frag<0>(acc1[i][2 * j], (2 * j + 1 < NC8) ? acc1[i][(2 * j + 1 < NC8) ? 2 * j + 1 : 0] : nullptr, zero4, zero4, 0, ah, al);
#pragma unroll
            for (int n = 0; n < NC8; ++n) {
              const uint4 b = sWo0[(j * NC8 + n) * 32 + lane];
              mma3(a0[n], ah, al, b);
            }
          }
#pragma unroll
          for (int j = 0; j < (NC8 + 1) / 2; ++j) {
            uint32_t ah[4], al[4];
            chain_frag<2>(a0[2 * j], (2 * j + 1 < NC8) ? a0[(2 * j + 1 < NC8) ? 2 * j + 1 : 0] : nullptr, bo0 + 16 * j, bo0 + 16 * j + 8, q, ah, al);
#pragma unroll
            for (int n = 0; n < NC8; ++n) {
              const uint4 b = sWo1[(j * NC8 + n) * 32 + lane];
              mma3(a1[n], ah, al, b);
            }
          }
#pragma unroll
          for (int n = 0; n < NC8; ++n) {
            const int col = n * 8 + 2 * q;
            if (valid[0]) { tmax[n][0] = fmaxf(tmax[n][0], a1[n][0] + bo1[col]); tmax[n][1] = fmaxf(tmax[n][1], a1[n][1] + bo1[col + 1]); }
            if (valid[1]) { tmax[n][0] = fmaxf(tmax[n][0], a1[n][2] + bo1[col]); tmax[n][1] = fmaxf(tmax[n][1], a1[n][3] + bo1[col + 1]); }
          }
        }
      }
    }
    if (TAIL) {
#pragma unroll
      for (int n = 0; n < NC8; ++n)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float v = tmax[n][e];
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
          v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
          if (g == 0 && v > -INFINITY) atomic_max_float(P.gmax + site * C + n * 8 + 2 * q + e, v);
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
