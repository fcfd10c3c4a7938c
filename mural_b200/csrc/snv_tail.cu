// Tail of both CNN branches of Network2 in one warp-level kernel (MURAL_MODE_BF16):
//   MaxPool1d(pool3) -> BN·Conv1d(32,32,3)·ReLU (conv3) -> max over length -> BN·Linear(32,n_class) per branch, then
//   log(clamp((softmax(local) + (softmax(mid) + softmax(large))/2)/2, 1e-9))      (MuRaL/model/model_snv.py:371-376,
//   414-419, 489-493, 511-523).
// After pool 3 a site is 7-8 rows long: a 128-row tcgen05 tile with its barriers and TMEM round trips costs far more
// than the 49k MACs it carries (profiles/r01_stage_tc_phase_timing.txt, mode 2).  Here one warp owns two sites at a
// time as the 16 rows of mma.sync.m16n8k16 (bf16 in, fp32 accumulate): the pooled rows are built in registers from the
// stage-2 output planes, the taps come from two shuffles, conv3's weights stay in registers as B fragments for a whole
// batch of sites, and ReLU / global max / the folded BN+Linear head / the three-softmax combine are shuffle reductions.
// No shared-memory tile, no block barrier, no intermediate written to HBM.
#include <cuda_bf16.h>
#include <float.h>
#include <math.h>

#include <vector>

#include "snv_model.cuh"

namespace mural {
namespace tail {

constexpr int BATCH = 16;    // sites per warp batch (8 two-site tiles per branch with the B fragments resident)
constexpr int WARPS = 4;

struct Branch {
  const uint32_t* z2;     // stage-2 output, bf16 planes [4][ra][8] viewed as 32-bit words [4][ra][4]
  int64_t ra;
  int L2, L3, pk, ps, pp;  // pool 3
  const uint32_t* wfrag;  // [6 k-chunks][4 n-tiles][32 lanes][2] B fragments of conv3 (BN folded, bf16)
  const float* cvec;      // [3][32]: bias + e0 + e1 + e2, e0 (left-edge correction), e2 (right-edge correction)
  const float* Wfc;       // [32][NC]
  const float* bfc;       // [NC]
};

struct State {
  uint32_t* d_wfrag[2] = {nullptr, nullptr};
  float* d_cvec[2] = {nullptr, nullptr};
};

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 v = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Lane (j = lane / 4, i = lane % 4) of a two-site tile owns pooled row j of both sites, 32-bit word i of each plane.
#ifndef MURAL_TAIL_MINB
#define MURAL_TAIL_MINB 3
#endif
__global__ void __launch_bounds__(WARPS * 32, MURAL_TAIL_MINB) k_tail(Branch b0, Branch b1, const float* __restrict__ local_logits, int64_t n, int NC,
                                                     float* __restrict__ logp, float* __restrict__ tg0, float* __restrict__ tg1,
                                                     float* __restrict__ tl0, float* __restrict__ tl1) {
  __shared__ float s_lg[WARPS][2][BATCH][16];
  __shared__ uint2 s_w[2][6 * 4 * 32];  // conv3 B fragments of both branches: [k-chunk][n-tile][lane]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = lane >> 2, i = lane & 3;
  for (int e = threadIdx.x; e < 2 * 6 * 4 * 32; e += WARPS * 32)
    s_w[e / (6 * 4 * 32)][e % (6 * 4 * 32)] = __ldg(reinterpret_cast<const uint2*>(e < 6 * 4 * 32 ? b0.wfrag : b1.wfrag) + e % (6 * 4 * 32));
  __syncthreads();
  const int64_t n_batches = (n + BATCH - 1) / BATCH;
  for (int64_t batch = int64_t(blockIdx.x) * WARPS + warp; batch < n_batches; batch += int64_t(gridDim.x) * WARPS) {
    const int64_t site_b = batch * BATCH;
#pragma unroll 1
    for (int br = 0; br < 2; ++br) {
      const Branch& B = br ? b1 : b0;
      // per-lane output-channel constants (channels 8*nt + 2*i + e): bias with the site-edge corrections of this lane's row
      // already applied (the padded tap contributes 0, not BN(0))
      float cb[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int co = 8 * nt + 2 * i + e;
          float v = __ldg(B.cvec + co);
          if (j == 0) v -= __ldg(B.cvec + 32 + co);
          if (j == B.L3 - 1) v -= __ldg(B.cvec + 64 + co);
          cb[nt][e] = v;
        }
      // head weights: row group j produces logit o = j (and o = j + 8) from the 8 channels this lane holds
      float wh[2][4][2];
#pragma unroll
      for (int oo = 0; oo < 2; ++oo)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int o = j + 8 * oo;
            wh[oo][nt][e] = o < NC ? __ldg(B.Wfc + (8 * nt + 2 * i + e) * NC + o) : 0.f;
          }
      const float bh0 = j < NC ? __ldg(B.bfc + j) : 0.f, bh1 = j + 8 < NC ? __ldg(B.bfc + j + 8) : 0.f;
      float* tg = br ? tg1 : tg0;
      float* tl = br ? tl1 : tl0;
      // pool window of this lane's row and a one-tile-ahead register prefetch of its stage-2 words (3 rows x 4 planes per
      // site), so the loads of tile t+1 are in flight while tile t computes
      int lo = j * B.ps - B.pp, hi = lo + B.pk;
      lo = lo < 0 ? 0 : lo;
      hi = hi > B.L2 ? B.L2 : hi;
      const bool fast = B.pk <= 3;  // every third-stage pool of Network2
      uint32_t nxt[2][3][4];
      auto prefetch = [&](int t) {
#pragma unroll
        for (int ss = 0; ss < 2; ++ss) {
          const int64_t site = site_b + 2 * t + ss;
          const bool ok = fast && site < n && j < B.L3 && t < BATCH / 2;
          const uint32_t* zr = B.z2 + (1 + site * int64_t(B.L2 + 1) + lo) * 4 + i;
#pragma unroll
          for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) nxt[ss][u][q] = (ok && lo + u < hi) ? __ldg(zr + (q * B.ra + u) * 4) : 0xFF80FF80u;
        }
      };
      prefetch(0);
#pragma unroll 1
      for (int t = 0; t < BATCH / 2; ++t) {
        const int64_t s0 = site_b + 2 * t;
        if (s0 >= n) break;  // warp-uniform
        // ---- pooled row j of both sites: max over the pool window of stage-2 rows (bf16, -inf padding)
        uint32_t pr[2][4], up[2][4], dn[2][4];
#pragma unroll
        for (int ss = 0; ss < 2; ++ss) {
          const bool ok = (s0 + ss < n) && j < B.L3;
          if (fast) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              pr[ss][q] = ok ? max_bf16x2(max_bf16x2(nxt[ss][0][q], nxt[ss][1][q]), nxt[ss][2][q]) : 0u;  // rows >= L3: zero padding
          } else {
            const uint32_t* zr = B.z2 + (1 + (s0 + ss) * int64_t(B.L2 + 1) + lo) * 4 + i;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t mx = 0xFF80FF80u;
              for (int u = 0; u < 7; ++u)
                if (ok && lo + u < hi) mx = max_bf16x2(mx, __ldg(zr + (q * B.ra + u) * 4));
              pr[ss][q] = ok ? mx : 0u;
            }
          }
        }
        prefetch(t + 1);
        // ---- taps: row j-1 from lane - 4, row j+1 from lane + 4 (zero padding outside [0, L3))
#pragma unroll
        for (int ss = 0; ss < 2; ++ss)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, pr[ss][q], 4), b = __shfl_down_sync(0xffffffffu, pr[ss][q], 4);
            up[ss][q] = j == 0 ? 0u : a;
            dn[ss][q] = j == 7 ? 0u : b;
          }
        // ---- conv3 as 6 k-chunks (tap, channel half) x 4 n-tiles of m16n8k16; rows 0-7 = site 0, rows 8-15 = site 1
        float d[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) d[nt][e] = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const int tp = c >> 1, h = c & 1;
          uint32_t a[4];
          a[0] = tp == 0 ? up[0][2 * h] : (tp == 1 ? pr[0][2 * h] : dn[0][2 * h]);
          a[1] = tp == 0 ? up[1][2 * h] : (tp == 1 ? pr[1][2 * h] : dn[1][2 * h]);
          a[2] = tp == 0 ? up[0][2 * h + 1] : (tp == 1 ? pr[0][2 * h + 1] : dn[0][2 * h + 1]);
          a[3] = tp == 0 ? up[1][2 * h + 1] : (tp == 1 ? pr[1][2 * h + 1] : dn[1][2 * h + 1]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const uint2 w = s_w[br][(c * 4 + nt) * 32 + lane];
            mma_bf16(d[nt], a, w.x, w.y);
          }
        }
        // ---- bias (+ site-edge corrections: the padded tap contributes 0, not BN(0)), ReLU, max over the L3 rows
        float g[2][4][2];
#pragma unroll
        for (int ss = 0; ss < 2; ++ss)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float v = d[nt][2 * ss + e] + cb[nt][e];
              v = j < B.L3 ? fmaxf(v, 0.f) : 0.f;  // ReLU output >= 0, so 0 is neutral for the max
              v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
              v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
              v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
              g[ss][nt][e] = v;
            }
        // ---- head: folded BN + Linear over the 8 channels this lane holds, reduced over i
#pragma unroll
        for (int ss = 0; ss < 2; ++ss) {
          if (tg && j == 0 && s0 + ss < n) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) tg[(s0 + ss) * 32 + 8 * nt + 2 * i + e] = g[ss][nt][e];
          }
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              p0 = fmaf(g[ss][nt][e], wh[0][nt][e], p0);
              p1 = fmaf(g[ss][nt][e], wh[1][nt][e], p1);
            }
          p0 += __shfl_xor_sync(0xffffffffu, p0, 1);
          p0 += __shfl_xor_sync(0xffffffffu, p0, 2);
          if (NC > 8) {
            p1 += __shfl_xor_sync(0xffffffffu, p1, 1);
            p1 += __shfl_xor_sync(0xffffffffu, p1, 2);
          }
          if (i == 0) {
            if (j < NC) {
              s_lg[warp][br][2 * t + ss][j] = p0 + bh0;
              if (tl && s0 + ss < n) tl[(s0 + ss) * NC + j] = p0 + bh0;
            }
            if (j + 8 < NC) {
              s_lg[warp][br][2 * t + ss][j + 8] = p1 + bh1;
              if (tl && s0 + ss < n) tl[(s0 + ss) * NC + j + 8] = p1 + bh1;
            }
          }
        }
      }
    }
    __syncwarp();
    // ---- combine (model_snv.py:515-523): one lane per site of the batch
    if (lane < BATCH && site_b + lane < n) {
      const int64_t site = site_b + lane;
      float pr_[16];
#pragma unroll
      for (int o = 0; o < 16; ++o) pr_[o] = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float v[16], mx = -FLT_MAX, sum = 0.f;
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (o < NC) {
            v[o] = k == 2 ? local_logits[site * NC + o] : s_lg[warp][k][lane][o];
            mx = fmaxf(mx, v[o]);
          }
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (o < NC) { v[o] = expf(v[o] - mx); sum += v[o]; }
#pragma unroll
        for (int o = 0; o < 16; ++o)
          if (o < NC) pr_[o] += (k == 2 ? 1.f : 0.5f) * (v[o] / sum);
      }
#pragma unroll
      for (int o = 0; o < 16; ++o)
        if (o < NC) logp[site * NC + o] = logf(fmaxf(pr_[o] / 2.f, 1e-9f));
    }
    __syncwarp();
  }
}

}  // namespace tail

void snv_tail_destroy(mural_snv_model* m) {
  if (!m->tail) return;
  tail::State* S = (tail::State*)m->tail;
  for (int br = 0; br < 2; ++br) { cudaFree(S->d_wfrag[br]); cudaFree(S->d_cvec[br]); }
  delete S;
  m->tail = nullptr;
}

// conv3 of both branches as mma.sync B fragments (BN scale folded into bf16 weights exactly as the tcgen05 blob does)
// plus the fp32 bias / edge-correction vectors.  Needs C == 32, ks == 3 and at most 8 rows after pool 3.
int snv_tail_prepare(mural_snv_model* m, const float* h_blob) {
  snv_tail_destroy(m);
  if (m->cfg.channels != 32 || m->cfg.kernel_size != 3 || m->cfg.n_class > 16) return 0;
  for (int br = 0; br < 2; ++br)
    if (m->br[br].L3 > 8 || m->br[br].L3 < 1 || m->br[br].pool[2][0] > 7) return 0;
  auto T = [&](const std::string& nm) { return h_blob + m->layout[m->index.at(nm)].offset; };
  tail::State* S = new tail::State();
  for (int br = 0; br < 2; ++br) {
    const std::string s = br ? "_2" : "";
    const std::string bn = "conv3" + s + ".0", cv = "conv3" + s + ".1";
    const float *g = T(bn + ".weight"), *be = T(bn + ".bias"), *mu = T(bn + ".running_mean"), *var = T(bn + ".running_var");
    const float *W = T(cv + ".weight"), *bi = T(cv + ".bias");  // [co][ci][tap]
    double a[32], b[32];
    for (int c = 0; c < 32; ++c) {
      a[c] = double(g[c]) / sqrt(double(var[c]) + 1e-5);
      b[c] = double(be[c]) - double(mu[c]) * a[c];
    }
    // B fragment of m16n8k16 (col-major 16x8): lane holds k = 2*(lane%4) + {0,1} (reg 0) and + 8 (reg 1), column n = lane/4
    std::vector<uint32_t> wf(6 * 4 * 32 * 2);
    for (int c = 0; c < 6; ++c)
      for (int nt = 0; nt < 4; ++nt)
        for (int lane = 0; lane < 32; ++lane)
          for (int r = 0; r < 2; ++r) {
            uint32_t word = 0;
            for (int e = 0; e < 2; ++e) {
              const int k = 2 * (lane % 4) + e + 8 * r, tap = c / 2, ci = (c % 2) * 16 + k, co = 8 * nt + lane / 4;
              const __nv_bfloat16 hv = __float2bfloat16(float(double(W[(co * 32 + ci) * 3 + tap]) * a[ci]));
              uint16_t bits;
              memcpy(&bits, &hv, 2);
              word |= uint32_t(bits) << (16 * e);
            }
            wf[((c * 4 + nt) * 32 + lane) * 2 + r] = word;
          }
    std::vector<float> cvec(96);
    for (int co = 0; co < 32; ++co) {
      double e[3] = {0, 0, 0};
      for (int t = 0; t < 3; ++t)
        for (int ci = 0; ci < 32; ++ci) e[t] += double(W[(co * 32 + ci) * 3 + t]) * b[ci];
      cvec[co] = float(double(bi[co]) + e[0] + e[1] + e[2]);
      cvec[32 + co] = float(e[0]);
      cvec[64 + co] = float(e[2]);
    }
    if (cudaMalloc((void**)&S->d_wfrag[br], wf.size() * 4) != cudaSuccess || cudaMalloc((void**)&S->d_cvec[br], cvec.size() * 4) != cudaSuccess) {
      m->tail = S;
      snv_tail_destroy(m);
      MURAL_FAIL("cudaMalloc of the tail-kernel weights failed");
    }
    cudaMemcpy(S->d_wfrag[br], wf.data(), wf.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(S->d_cvec[br], cvec.data(), cvec.size() * 4, cudaMemcpyHostToDevice);
  }
  m->tail = S;
  return 0;
}

// returns -1 when unavailable for this model (caller runs the SINGLE stage kernels + k_head_tc instead)
int snv_tail_launch(mural_snv_model* m, const void* z2_mid, int64_t ra_mid, const void* z2_large, int64_t ra_large,
                    const float* local_logits, int64_t ns, float* logp, float* tg0, float* tg1, float* tl0, float* tl1, cudaStream_t st) {
  if (!m->tail) return -1;
  tail::State* S = (tail::State*)m->tail;
  tail::Branch b[2];
  for (int br = 0; br < 2; ++br) {
    const BranchDev& B = m->br[br];
    b[br] = tail::Branch{reinterpret_cast<const uint32_t*>(br ? z2_large : z2_mid), br ? ra_large : ra_mid, B.L2, B.L3, B.pool[2][0],
                         B.pool[2][1], B.pool[2][2], S->d_wfrag[br], S->d_cvec[br], B.Wfc, B.bfc};
  }
  const int64_t batches = cdiv(ns, tail::BATCH);
  int64_t grid = cdiv(batches, tail::WARPS);
  if (grid > 148 * 16) grid = 148 * 16;
  LAUNCH((tail::k_tail), (unsigned)grid, tail::WARPS * 32, 0, st, b[0], b[1], local_logits, ns, m->cfg.n_class, logp, tg0, tg1, tl0, tl1);
  return 0;
}

}  // namespace mural
