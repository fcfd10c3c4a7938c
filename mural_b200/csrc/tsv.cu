// Prediction TSV writer (host code): the tail of MuRaL/scripts/run_predict.py:228-239 —
//   pred_df.to_csv(pred_file, sep='\t', float_format='%.4g', index=False)
// with columns chrom, start, end, strand, mut_type, prob0..prob{k-1}.  At genome-wide scale (5e7 rows) pandas' writer
// takes minutes while the network takes a second; here rows are formatted by a pool of threads into per-thread buffers
// (printf's "%.4g" is exactly what pandas applies per value) and written in order.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <cmath>

#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {
// printf("%.4g") for the values a prediction table holds, without printf: 4 significant digits, correctly rounded, trailing zeros
// (and a bare point) stripped, scientific form when the decimal exponent is < -4 or >= 4 — byte for byte what pandas' float_format
// applies.  The digits come from one multiplication by a power of ten; a value whose scaled form lies within 1e-9 of a rounding
// boundary (where that product's last-bit error could matter), a non-finite value or an exponent outside the table goes through
// snprintf.  Returns the number of characters written to out (no terminator).
const double kPow10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
inline int fmt_g4(double v, char* out) {
  if (v == 0.0) {
    if (std::signbit(v)) { out[0] = '-'; out[1] = '0'; return 2; }
    out[0] = '0';
    return 1;
  }
  const double a = std::fabs(v);
  if (!(a >= 1e-19 && a < 1e19)) return snprintf(out, 32, "%.4g", v);
  int e = int(std::floor(std::log10(a)));           // decimal exponent estimate, corrected below
  const int sh = 3 - e;                              // scaled = a * 10^sh in [1000, 10000)
  double scaled = sh >= 0 ? a * kPow10[sh] : a / kPow10[-sh];
  if (scaled < 1000.0) { --e; scaled *= 10.0; }
  else if (scaled >= 10000.0) { ++e; scaled /= 10.0; }
  const double fl = std::floor(scaled), fr = scaled - fl;
  if (std::fabs(fr - 0.5) < 1e-9) return snprintf(out, 32, "%.4g", v);   // too close to a tie to call here (near an integer either way agrees)
  int m = int(fl) + (fr > 0.5 ? 1 : 0);
  if (m >= 10000) { m = 1000; ++e; }
  char d[4] = {char('0' + m / 1000), char('0' + (m / 100) % 10), char('0' + (m / 10) % 10), char('0' + m % 10)};
  int nd = 4;
  while (nd > 1 && d[nd - 1] == '0') --nd;          // significant digits left after stripping trailing zeros
  int k = 0;
  if (v < 0) out[k++] = '-';
  if (e < -4 || e >= 4) {
    out[k++] = d[0];
    if (nd > 1) { out[k++] = '.'; for (int i = 1; i < nd; ++i) out[k++] = d[i]; }
    out[k++] = 'e';
    int ae = e;
    if (ae < 0) { out[k++] = '-'; ae = -ae; } else out[k++] = '+';
    if (ae >= 100) { out[k++] = char('0' + ae / 100); ae %= 100; }
    out[k++] = char('0' + ae / 10);
    out[k++] = char('0' + ae % 10);
  } else if (e >= 0) {
    for (int i = 0; i <= e; ++i) out[k++] = i < 4 ? d[i] : '0';
    if (nd > e + 1) { out[k++] = '.'; for (int i = e + 1; i < nd; ++i) out[k++] = d[i]; }
  } else {
    out[k++] = '0'; out[k++] = '.';
    for (int i = 0; i < -e - 1; ++i) out[k++] = '0';
    for (int i = 0; i < nd; ++i) out[k++] = d[i];
  }
  return k;
}
inline int fmt_i64(long long x, char* out) {
  char t[24];
  int n = 0;
  unsigned long long u = x < 0 ? 0ull - (unsigned long long)x : (unsigned long long)x;
  do { t[n++] = char('0' + u % 10); u /= 10; } while (u);
  int k = 0;
  if (x < 0) out[k++] = '-';
  while (n) out[k++] = t[--n];
  return k;
}
}  // namespace

// parity hook of the formatter (tests/test_host_logic.py compares it with printf over random and boundary values)
extern "C" int mural_format_g4(double v, char* out32) {
  const int k = fmt_g4(v, out32);
  out32[k] = 0;
  return k;
}

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
  // two buffer sets: while this thread writes batch k to the file, the pool formats batch k + 1
  std::vector<std::string> bufs[2] = {std::vector<std::string>(n_threads), std::vector<std::string>(n_threads)};
  std::vector<std::thread> pool[2];
  auto launch = [&](int slot, int64_t b0) {
    for (int t = 0; t < n_threads; ++t) {
      const int64_t lo = b0 + t * block, hi = lo + block < n ? lo + block : n;
      bufs[slot][t].clear();
      if (lo >= n) continue;
      pool[slot].emplace_back([&, slot, t, lo, hi] {
        std::string& s = bufs[slot][t];
        s.reserve(size_t(hi - lo) * (40 + 12 * n_class));
        char tmp[128 + 34 * 64];
        for (int64_t i = lo; i < hi; ++i) {
          s += chrom_names[chrom_idx[i]];
          int k = 0;
          tmp[k++] = '\t'; k += fmt_i64((long long)start[i], tmp + k);
          tmp[k++] = '\t'; k += fmt_i64((long long)end[i], tmp + k);
          tmp[k++] = '\t'; tmp[k++] = strand[i];
          tmp[k++] = '\t'; k += fmt_g4(mut_type[i], tmp + k);
          for (int c = 0; c < n_class; ++c) { tmp[k++] = '\t'; k += fmt_g4(prob[i * n_class + c], tmp + k); }
          tmp[k++] = '\n';
          s.append(tmp, k);
        }
      });
    }
  };
  const int64_t step = block * n_threads;
  int cur = 0;
  if (n > 0) launch(cur, 0);
  for (int64_t b0 = 0; b0 < n; b0 += step) {
    for (auto& th : pool[cur]) th.join();
    pool[cur].clear();
    if (b0 + step < n) launch(1 - cur, b0 + step);
    for (int t = 0; t < n_threads && ok; ++t)
      if (!bufs[cur][t].empty()) ok = fwrite(bufs[cur][t].data(), 1, bufs[cur][t].size(), f) == bufs[cur][t].size();
    cur = 1 - cur;
  }
  ok = (fclose(f) == 0) && ok;
  MURAL_CHECK(ok, std::string("write error on ") + path);
  return 0;
}
