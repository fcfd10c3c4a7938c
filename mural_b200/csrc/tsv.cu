// Prediction TSV writer (host code): the tail of MuRaL/scripts/run_predict.py:228-239 —
//   pred_df.to_csv(pred_file, sep='\t', float_format='%.4g', index=False)
// with columns chrom, start, end, strand, mut_type, prob0..prob{k-1}.  At genome-wide scale (5e7 rows) pandas' writer
// takes minutes while the network takes a second; here rows are formatted by a pool of threads into per-thread buffers
// (printf's "%.4g" is exactly what pandas applies per value) and written in order.
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

extern "C" int mural_write_tsv(const char* path, int64_t n, int32_t n_class, const char* const* chrom_names, const int32_t* chrom_idx,
                               const int64_t* start, const int64_t* end, const char* strand, const double* mut_type, const double* prob,
                               int32_t n_threads) {
  MURAL_CHECK(path && (n == 0 || (chrom_names && chrom_idx && start && end && strand && mut_type && prob)), "NULL argument");
  MURAL_CHECK(n_class >= 1 && n_class <= 64, "n_class out of range");
  FILE* f = fopen(path, "wb");
  MURAL_CHECK(f != nullptr, std::string("cannot open ") + path);
  std::string head = "chrom\tstart\tend\tstrand\tmut_type";
  for (int i = 0; i < n_class; ++i) head += "\tprob" + std::to_string(i);
  head += "\n";
  bool ok = fwrite(head.data(), 1, head.size(), f) == head.size();
  if (n_threads < 1) n_threads = 1;
  const int64_t block = 1 << 16;  // rows per work item; blocks are written in order, a batch of n_threads at a time
  std::vector<std::string> bufs(n_threads);
  for (int64_t b0 = 0; b0 < n && ok; b0 += block * n_threads) {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) {
      const int64_t lo = b0 + t * block, hi = lo + block < n ? lo + block : n;
      bufs[t].clear();
      if (lo >= n) continue;
      pool.emplace_back([&, t, lo, hi] {
        std::string& s = bufs[t];
        s.reserve(size_t(hi - lo) * (40 + 12 * n_class));
        char tmp[64];
        for (int64_t i = lo; i < hi; ++i) {
          s += chrom_names[chrom_idx[i]];
          int k = snprintf(tmp, sizeof tmp, "\t%lld\t%lld\t%c\t%.4g", (long long)start[i], (long long)end[i], strand[i], mut_type[i]);
          s.append(tmp, k);
          for (int c = 0; c < n_class; ++c) {
            k = snprintf(tmp, sizeof tmp, "\t%.4g", prob[i * n_class + c]);
            s.append(tmp, k);
          }
          s += '\n';
        }
      });
    }
    for (auto& th : pool) th.join();
    for (int t = 0; t < n_threads && ok; ++t)
      if (!bufs[t].empty()) ok = fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
  }
  ok = (fclose(f) == 0) && ok;
  MURAL_CHECK(ok, std::string("write error on ") + path);
  return 0;
}
