// Shared-memory-tiled kernels for the three convolution ops of the MuRaL-indel training tape (stride-1 layers, with or
// without the decoder's nearest upsampling in front: > 95 % of UNet_Small's FLOPs; MuRaL/model/model_indel.py:6-19,
// 151-176).  CUDA build only — the host emulation (tests/emu) keeps the work-item functors of indel_train_core.cuh, which
// remain the definition of the arithmetic and the fallback for strided / short / very wide layers.
//
//   k_conv_tiled<K>   y[b][o][p] (+)= bias[o] + sum_i sum_t w(o,i,t) * xv[b][i][p + t - pad]      xv = x read through the upsample
//       forward:        w(o,i,t) = W[o][i][t]
//       input gradient: the same correlation with the roles of the channel axes swapped and the taps flipped
//                       (w(o,i,t) = W[i][o][K-1-t]); with an upsample in front the result is the gradient of the virtual
//                       upsampled input, summed over groups of `up` by UpReduce afterwards.
//     CTA = (512-position tile, sample), thread = 4 consecutive positions x 8 output channels at a time; per input channel a
//     thread loads its 12-float window as three aligned LDS.128 and the 8 x K weights as broadcast LDS.128: 17 shared loads
//     feed 224 FMAs at K = 7.
//   k_wgrad_tiled<K>  dW[o][i][t] += sum_b sum_p dy[b][o][p] * xv[b][i][p*stride + t - pad],  db[o] += sum dy   (any stride / length)
//     CTA = (group of 512-position tiles, sample); thread = one (o, i) pair (x a slice of the tile when there are fewer than
//     256 pairs) holding its K accumulators in registers over all tiles of the CTA; one atomicAdd per element and CTA.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <stdint.h>

namespace indel_train {
namespace tiled {

constexpr int TL = 512;   // positions per tile
constexpr int XO = 4;     // xs index of tile-local virtual position 0 (>= pad, multiple of 4: windows start 16-byte aligned)
constexpr int XS = TL + 8;

struct ConvT {
  const float* x; const float* W; const float* bias; float* y;
  int Cin, Lin, Cout, Lout, pad, up;   // Cin / Cout: input / output channels of THIS correlation (already swapped for the input gradient)
  int so, si, flip;                    // w(o, i, t) = W[o*so + i*si + (flip ? K-1-t : t)]
  int accumulate;                      // y += instead of y =
};

template <int K>
__global__ void __launch_bounds__(128) k_conv_tiled(ConvT a) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                          // [Cin][XS]
  float* ws = sm + size_t(a.Cin) * XS;     // [Cin][K][8]
  const int tid = threadIdx.x, b = blockIdx.y, p0 = blockIdx.x * TL;
  const int Lv = a.Lin * a.up;
  const float* xb = a.x + int64_t(b) * a.Cin * a.Lin;
  for (int e = tid; e < a.Cin * XS; e += 128) {
    const int i = e / XS, q = e - i * XS;
    const int j = p0 + q - XO;
    xs[e] = (j >= 0 && j < Lv) ? xb[int64_t(i) * a.Lin + (a.up == 1 ? j : j / a.up)] : 0.f;
  }
  for (int o0 = blockIdx.z * 8; o0 < a.Cout; o0 += 8 * gridDim.z) {   // output-channel chunks may be split over grid.z (small grids)
    __syncthreads();                       // xs ready (first pass) / previous chunk's weight reads done
    for (int e = tid; e < a.Cin * K * 8; e += 128) {
      const int c = e & 7, t = (e >> 3) % K, i = e / (8 * K);
      ws[e] = (o0 + c < a.Cout) ? a.W[int64_t(o0 + c) * a.so + int64_t(i) * a.si + (a.flip ? K - 1 - t : t)] : 0.f;
    }
    __syncthreads();
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[p][c] = 0.f;
    for (int i = 0; i < a.Cin; ++i) {
      float w[12];
      const float4* xr = reinterpret_cast<const float4*>(xs + i * XS + 4 * tid);   // virtual positions 4*tid - 4 .. 4*tid + 7
#pragma unroll
      for (int q = 0; q < 3; ++q) { const float4 v = xr[q]; w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
      const float4* wr = reinterpret_cast<const float4*>(ws + i * K * 8);
#pragma unroll
      for (int t = 0; t < K; ++t) {
        const float4 w0 = wr[2 * t], w1 = wr[2 * t + 1];
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float xv = w[XO + p + t - (K - 1) / 2];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(xv, wv[c], acc[p][c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int o = o0 + c;
      if (o >= a.Cout) break;
      const float bv = a.bias ? a.bias[o] : 0.f;
      float* yr = a.y + (int64_t(b) * a.Cout + o) * a.Lout;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int pos = p0 + 4 * tid + p;
        if (pos < a.Lout) yr[pos] = (a.accumulate ? yr[pos] : 0.f) + acc[p][c] + bv;
      }
    }
  }
}

// Strided encoder convs (stride 4 / 5 / 5 / 5 / 2; < 5 % of the FLOPs) and short rows: same tile structure, a thread takes the
// output positions tid, tid + 128, ... of the tile and reads its taps straight from the shared tile.
constexpr int TLS = 256;   // output positions per tile
template <int K>
__global__ void __launch_bounds__(128) k_conv_strided(ConvT a, int stride) {
  extern __shared__ __align__(16) float sm[];
  const int XR = (TLS - 1) * stride + K;
  float* xs = sm;                              // [Cin][XR]: index q <-> virtual position p0*stride + q - pad
  float* ws = sm + size_t(a.Cin) * XR;         // [Cin][K][8]
  const int tid = threadIdx.x, b = blockIdx.y, p0 = blockIdx.x * TLS;
  const int Lv = a.Lin * a.up;
  const float* xb = a.x + int64_t(b) * a.Cin * a.Lin;
  for (int e = tid; e < a.Cin * XR; e += 128) {
    const int i = e / XR, q = e - i * XR;
    const int j = p0 * stride + q - a.pad;
    xs[e] = (j >= 0 && j < Lv) ? xb[int64_t(i) * a.Lin + (a.up == 1 ? j : j / a.up)] : 0.f;
  }
  for (int o0 = blockIdx.z * 8; o0 < a.Cout; o0 += 8 * gridDim.z) {   // output-channel chunks may be split over grid.z (small grids)
    __syncthreads();
    for (int e = tid; e < a.Cin * K * 8; e += 128) {
      const int c = e & 7, t = (e >> 3) % K, i = e / (8 * K);
      ws[e] = (o0 + c < a.Cout) ? a.W[int64_t(o0 + c) * a.so + int64_t(i) * a.si + (a.flip ? K - 1 - t : t)] : 0.f;
    }
    __syncthreads();
    float acc[2][8];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[p][c] = 0.f;
    for (int i = 0; i < a.Cin; ++i) {
      const float* xr = xs + i * XR;
      const float4* wr = reinterpret_cast<const float4*>(ws + i * K * 8);
#pragma unroll
      for (int t = 0; t < K; ++t) {
        const float4 w0 = wr[2 * t], w1 = wr[2 * t + 1];
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float xv = xr[(tid + 128 * p) * stride + t];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(xv, wv[c], acc[p][c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int o = o0 + c;
      if (o >= a.Cout) break;
      const float bv = a.bias ? a.bias[o] : 0.f;
      float* yr = a.y + (int64_t(b) * a.Cout + o) * a.Lout;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int pos = p0 + tid + 128 * p;
        if (pos < a.Lout) yr[pos] = (a.accumulate ? yr[pos] : 0.f) + acc[p][c] + bv;
      }
    }
  }
}

struct WgradT {
  const float* x; const float* dy; float* dW; float* db;
  int Cin, Lin, Cout, Lout, pad, up, stride, n_tiles;
  int TLe, RSx, RSd;   // positions per tile (min(TL, Lout)); odd row strides of the shared x / dy tiles
};

template <int K>
__global__ void __launch_bounds__(256) k_wgrad_tiled(WgradT a) {
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;                                  // [Cin][RSx]: index q <-> virtual position p0*stride + q - pad
  float* ds = sm + size_t(a.Cin) * a.RSx;          // [Cout][RSd]
  const int TLe = a.TLe, XR = (TLe - 1) * a.stride + K;
  const int tid = threadIdx.x, b = blockIdx.y;
  const int pairs = a.Cin * a.Cout;
  const int slices = pairs >= 256 ? 1 : 256 / pairs;
  const int chunk = (TLe + slices - 1) / slices;
  const int Lv = a.Lin * a.up;
  const float* xb = a.x + int64_t(b) * a.Cin * a.Lin;
  const float* db_ = a.dy + int64_t(b) * a.Cout * a.Lout;
  // this thread's pairs: tid % pairs (+ 256, ...) when slices == 1, else pair = tid % pairs, slice = tid / pairs
  constexpr int MAXP = 8;                          // pairs per thread when there are more than 256 (checked by the launcher)
  float acc[MAXP][K];
  float accb[MAXP];
#pragma unroll
  for (int m = 0; m < MAXP; ++m) {
    accb[m] = 0.f;
#pragma unroll
    for (int t = 0; t < K; ++t) acc[m][t] = 0.f;
  }
  const int my_slice = slices == 1 ? 0 : tid / pairs;
  const bool active = slices == 1 || my_slice < slices;
  // more than 256*MAXP pairs (the two deepest levels, rows of 8-16 positions): batches of pairs, each re-staging its (tiny) tiles
  for (int pbase = 0; pbase < pairs; pbase += 256 * MAXP) {
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int p0 = tile * TLe;
    __syncthreads();
    for (int e = tid; e < a.Cin * XR; e += 256) {
      const int i = e / XR, q = e - i * XR;
      const int j = p0 * a.stride + q - a.pad;
      xs[i * a.RSx + q] = (j >= 0 && j < Lv) ? xb[int64_t(i) * a.Lin + (a.up == 1 ? j : j / a.up)] : 0.f;
    }
    for (int e = tid; e < a.Cout * TLe; e += 256) {
      const int o = e / TLe, q = e - o * TLe;
      ds[o * a.RSd + q] = (p0 + q < a.Lout) ? db_[int64_t(o) * a.Lout + p0 + q] : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    const int l0 = my_slice * chunk, l1 = (l0 + chunk < TLe) ? l0 + chunk : TLe;
#pragma unroll
    for (int m = 0; m < MAXP; ++m) {
      const int pr = (slices == 1 ? pbase + tid + m * 256 : tid - my_slice * pairs);
      if (pr >= pairs || (slices > 1 && m > 0)) break;
      const int o = pr / a.Cin, i = pr - o * a.Cin;
      const float* xr = xs + i * a.RSx;
      const float* dr = ds + o * a.RSd;
      float sb = 0.f;
      for (int l = l0; l < l1; ++l) {
        const float d = dr[l];
        sb += d;
#pragma unroll
        for (int t = 0; t < K; ++t) acc[m][t] = fmaf(d, xr[l * a.stride + t], acc[m][t]);
      }
      if (i == 0) accb[m] += sb;
    }
  }
  if (active) {
#pragma unroll
    for (int m = 0; m < MAXP; ++m) {
      const int pr = (slices == 1 ? pbase + tid + m * 256 : tid - my_slice * pairs);
      if (pr >= pairs || (slices > 1 && m > 0)) break;
      const int o = pr / a.Cin, i = pr - o * a.Cin;
#pragma unroll
      for (int t = 0; t < K; ++t) { atomicAdd(a.dW + (int64_t(o) * a.Cin + i) * K + t, acc[m][t]); acc[m][t] = 0.f; }
      if (i == 0 && a.db) atomicAdd(a.db + o, accb[m]);
      accb[m] = 0.f;
    }
  }
  }  // pair batches
}

// y[r] = max over the row, arg[r] = first index of the maximum (torch.max semantics); one warp per row
__global__ void __launch_bounds__(256) k_max_rows(const float* __restrict__ x, float* __restrict__ y, int32_t* __restrict__ arg, int64_t rows, int L) {
  const int64_t r = blockIdx.x * int64_t(blockDim.x / 32) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = x + r * L;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int l = lane; l < L; l += 32) {
    const float v = p[l];
    if (v > m) { m = v; am = l; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
    if (m2 > m || (m2 == m && a2 < am)) { m = m2; am = a2; }
  }
  if (lane == 0) { y[r] = m; arg[r] = am; }
}

}  // namespace tiled
}  // namespace indel_train
#endif
