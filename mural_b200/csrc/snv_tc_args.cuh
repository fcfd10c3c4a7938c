// Arguments of the tcgen05 stage kernels (snv_tc.cu: 128-row tiles, one row per TMEM lane).
#pragma once
#include "snv_model.cuh"

namespace mural {
namespace tc {

enum Mode { RB4 = 0, C_RB4 = 1, SINGLE = 2 };
__host__ __device__ constexpr int n_layers(int mode) { return mode == RB4 ? 4 : (mode == C_RB4 ? 5 : 1); }

struct StageArgs {
  const uint4* in;       // bf16 planes [4][in_rows_alloc] (one uint4 = 8 channels of one row)
  void* out;             // bf16 planes [4][out_rows_alloc] (RB4, C_RB4) or fp32 planes [8][out_rows_alloc] float4 (SINGLE)
  const uint8_t* wblob;  // n_layers * W_LAYER bytes
  int64_t in_rows_alloc, out_rows_alloc;
  int64_t rows;          // n_sites*(L+1)+1 rows of this stage
  int L;                 // site length at this stage
  int Lin;               // site length of the input buffer (== L when not pooled)
  int pk, ps, pp;        // max-pool fused into the loader (RB4: none)
  int n_tiles;
  // dense-site dispatch (snv_dense_stem.cu): all decided on the device, no host synchronisation
  const ChunkInfo* info;  // nullptr: unconditional launch
  int want;               // run only if info->dense == want
  int lat_branch;         // >= 0: this launch is the stage-1 lattice of that branch; L / rows / n_tiles come from info
  // LAT loader (C_RB4 only): stage-1 rows of a site come from the lattice (interior) and the edge pseudo-site (ends)
  const uint4* lat;       // stage-1 lattice output planes [4][lat_ra]
  const uint4* lat2;      // the same max-pooled along the lattice: lat2[row] = max(lat[row .. row+pk-1]) (k_lattice_pool)
  int64_t lat_ra;
  const uint4* edge;      // stage-1 edge output planes [4][edge_ra], LAT_EL rows per site
  int64_t edge_ra;
  const uint4* epool;     // pooled rows of the bins that touch edge rows (k_edge_pool): planes [4][epool_ra], row = site*(nlo+nhi) + b
  int64_t epool_ra;
  int nlo, nhi;           // such bins at the low / high end of a site
  // EDGE loader (RB4 on the edge pseudo-sites): rows 1..16 of a pseudo-site are full-bin table rows read in place
  // (snv_dense_stem.cu tables, width-0 table of this branch per strand), rows 0 and 17 come from the special-row buffer
  const uint4* tab[2];    // [strand]: rows of 4 uint4 (32 bf16) per genomic position
  const uint4* special;   // planes [4][special_ra], row = 2*site + (last ? 1 : 0)
  int64_t special_ra;
  int L1real;             // stage-1 length of a real site (pseudo-site row jj >= LAT_EI is real row L1real - LAT_EL + jj)
  const int32_t* pos;
  const int32_t* meta;
  int ps1, pp1, pk1, off0, R, br;
  int ablate;             // scratch timing builds only (-DMURAL_TC_TIMING): bit 0 = loaders fetch nothing
};


}  // namespace tc

}  // namespace mural
