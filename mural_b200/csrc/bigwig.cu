// bigWig reader (host code) for the continuous features of MuRaL-snv: replaces pyBigWig in
// get_mean_bw_for_bed (MuRaL/data/preprocessing.py:725-750) — per site and track, the mean over the expanded window of
// np.nan_to_num(bw.values(chrom, start1, stop1)) (bases without data count as 0).
//
// Format (Kent et al. 2010, "BigWig and BigBed", supplementary tables): 64-byte header (magic 0x888FFC26), zoom headers,
// chromosome B+ tree (magic 0x78CA8C91: name -> id, size), data sections (zlib-deflated when uncompressBufSize > 0; 24-byte
// section header + bedGraph / variableStep / fixedStep items) and an R-tree over the sections (magic 0x2468ACE0).  A track is
// decoded once per chromosome into a dense per-base array + a prefix sum in double, so a window mean is two lookups.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

struct mural_bigwig {
  FILE* f = nullptr;
  bool swap = false;  // file written on a machine of the other endianness
  uint32_t uncompress_buf = 0;
  uint64_t chrom_tree = 0, full_index = 0;
  std::vector<std::string> names;      // by position in `ids`
  std::vector<uint32_t> ids, sizes;
  std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> blocks;  // chrom id -> (offset, size) of its data sections
  std::map<uint32_t, std::vector<double>> prefix;                         // chrom id -> prefix sums of nan_to_num(values), length size + 1
};

namespace {
using namespace mural;

template <typename T> T rd(const mural_bigwig* b, const unsigned char* p) {
  T v;
  memcpy(&v, p, sizeof(T));
  if (b->swap) {
    unsigned char* q = reinterpret_cast<unsigned char*>(&v);
    for (size_t i = 0; i < sizeof(T) / 2; ++i) { unsigned char t = q[i]; q[i] = q[sizeof(T) - 1 - i]; q[sizeof(T) - 1 - i] = t; }
  }
  return v;
}
bool read_at(mural_bigwig* b, uint64_t off, void* dst, size_t n) {
  return fseek(b->f, (long)off, SEEK_SET) == 0 && fread(dst, 1, n, b->f) == n;
}

int walk_chrom_tree(mural_bigwig* b, uint64_t off, uint32_t key_size) {
  unsigned char h[4];
  MURAL_CHECK(read_at(b, off, h, 4), "bigWig: truncated chromosome tree");
  const bool leaf = h[0] != 0;
  const uint16_t count = rd<uint16_t>(b, h + 2);
  const size_t item = key_size + 8;
  std::vector<unsigned char> buf(size_t(count) * item);
  MURAL_CHECK(read_at(b, off + 4, buf.data(), buf.size()), "bigWig: truncated chromosome tree node");
  for (uint16_t i = 0; i < count; ++i) {
    const unsigned char* p = buf.data() + size_t(i) * item;
    if (leaf) {
      std::string nm(reinterpret_cast<const char*>(p), strnlen(reinterpret_cast<const char*>(p), key_size));
      b->names.push_back(nm);
      b->ids.push_back(rd<uint32_t>(b, p + key_size));
      b->sizes.push_back(rd<uint32_t>(b, p + key_size + 4));
    } else if (int rc = walk_chrom_tree(b, rd<uint64_t>(b, p + key_size), key_size)) {
      return rc;
    }
  }
  return 0;
}

int walk_rtree(mural_bigwig* b, uint64_t off) {
  unsigned char h[4];
  MURAL_CHECK(read_at(b, off, h, 4), "bigWig: truncated R-tree");
  const bool leaf = h[0] != 0;
  const uint16_t count = rd<uint16_t>(b, h + 2);
  const size_t item = leaf ? 32 : 24;
  std::vector<unsigned char> buf(size_t(count) * item);
  MURAL_CHECK(read_at(b, off + 4, buf.data(), buf.size()), "bigWig: truncated R-tree node");
  for (uint16_t i = 0; i < count; ++i) {
    const unsigned char* p = buf.data() + size_t(i) * item;
    if (leaf) {
      const uint32_t c0 = rd<uint32_t>(b, p), c1 = rd<uint32_t>(b, p + 8);
      for (uint32_t c = c0; c <= c1; ++c) b->blocks[c].push_back({rd<uint64_t>(b, p + 16), rd<uint64_t>(b, p + 24)});
    } else if (int rc = walk_rtree(b, rd<uint64_t>(b, p + 16))) {
      return rc;
    }
  }
  return 0;
}

// dense values of one chromosome (NaN -> 0 like np.nan_to_num) as prefix sums
int decode_chrom(mural_bigwig* b, int ci) {
  const uint32_t id = b->ids[ci], size = b->sizes[ci];
  if (b->prefix.count(id)) return 0;
  std::vector<float> v(size, 0.f);
  std::vector<unsigned char> raw, buf;
  auto it = b->blocks.find(id);
  if (it != b->blocks.end()) {
    for (auto& blk : it->second) {
      raw.resize(blk.second);
      MURAL_CHECK(read_at(b, blk.first, raw.data(), raw.size()), "bigWig: truncated data section");
      const unsigned char* d = raw.data();
      size_t n = raw.size();
      if (b->uncompress_buf) {
        buf.resize(b->uncompress_buf);
        uLongf out_len = buf.size();
        MURAL_CHECK(uncompress(buf.data(), &out_len, raw.data(), raw.size()) == Z_OK, "bigWig: zlib error in a data section");
        d = buf.data();
        n = out_len;
      }
      MURAL_CHECK(n >= 24, "bigWig: data section shorter than its header");
      const uint32_t chrom = rd<uint32_t>(b, d), start = rd<uint32_t>(b, d + 4), step = rd<uint32_t>(b, d + 12), span = rd<uint32_t>(b, d + 16);
      const uint8_t type = d[20];
      const uint16_t cnt = rd<uint16_t>(b, d + 22);
      if (chrom != id) continue;  // a leaf spanning several chromosomes lists the section under each of them
      const unsigned char* p = d + 24;
      auto fill = [&](uint64_t s, uint64_t e, float val) {
        if (isnan(val) || isinf(val)) val = isnan(val) ? 0.f : (val > 0 ? 3.4028235e38f : -3.4028235e38f);  // np.nan_to_num
        for (uint64_t q = s; q < e && q < size; ++q) v[q] = val;
      };
      for (uint16_t i = 0; i < cnt; ++i) {
        if (type == 1) {         // bedGraph: start, end, value
          MURAL_CHECK(size_t(p - d) + 12 <= n, "bigWig: truncated bedGraph item");
          fill(rd<uint32_t>(b, p), rd<uint32_t>(b, p + 4), rd<float>(b, p + 8));
          p += 12;
        } else if (type == 2) {  // variableStep: start, value (span from the header)
          MURAL_CHECK(size_t(p - d) + 8 <= n, "bigWig: truncated variableStep item");
          const uint32_t s = rd<uint32_t>(b, p);
          fill(s, uint64_t(s) + span, rd<float>(b, p + 4));
          p += 8;
        } else if (type == 3) {  // fixedStep: value (start + i*step, span)
          MURAL_CHECK(size_t(p - d) + 4 <= n, "bigWig: truncated fixedStep item");
          const uint64_t s = uint64_t(start) + uint64_t(i) * step;
          fill(s, s + span, rd<float>(b, p));
          p += 4;
        } else {
          MURAL_FAIL("bigWig: unknown data section type");
        }
      }
    }
  }
  std::vector<double>& ps = b->prefix[id];
  ps.resize(size_t(size) + 1);
  ps[0] = 0;
  for (uint32_t q = 0; q < size; ++q) ps[q + 1] = ps[q] + double(v[q]);
  return 0;
}
}  // namespace

extern "C" int mural_bigwig_open(const char* path, mural_bigwig_t** out) {
  MURAL_CHECK(path && out, "NULL argument");
  *out = nullptr;
  mural_bigwig* b = new mural_bigwig();
  b->f = fopen(path, "rb");
  if (!b->f) { delete b; MURAL_FAIL(std::string("cannot open ") + path); }
  unsigned char h[64];
  auto bail = [&](const char* msg) { fclose(b->f); delete b; return fail(__FILE__, __LINE__, msg); };
  if (fread(h, 1, 64, b->f) != 64) return bail("bigWig: file shorter than its header");
  uint32_t magic;
  memcpy(&magic, h, 4);
  if (magic == 0x26FC8F88u) b->swap = true;
  else if (magic != 0x888FFC26u) return bail("not a bigWig file (bad magic)");
  b->chrom_tree = rd<uint64_t>(b, h + 8);
  b->full_index = rd<uint64_t>(b, h + 24);
  b->uncompress_buf = rd<uint32_t>(b, h + 52);
  unsigned char th[32];
  if (!read_at(b, b->chrom_tree, th, 32) || rd<uint32_t>(b, th) != 0x78CA8C91u) return bail("bigWig: bad chromosome tree");
  const uint32_t key_size = rd<uint32_t>(b, th + 8);
  if (walk_chrom_tree(b, b->chrom_tree + 32, key_size)) { fclose(b->f); delete b; return 1; }
  unsigned char rh[48];
  if (!read_at(b, b->full_index, rh, 48) || rd<uint32_t>(b, rh) != 0x2468ACE0u) return bail("bigWig: bad R-tree index");
  if (walk_rtree(b, b->full_index + 48)) { fclose(b->f); delete b; return 1; }
  *out = b;
  return 0;
}
extern "C" int32_t mural_bigwig_n_chrom(const mural_bigwig_t* b) { return b ? (int32_t)b->names.size() : 0; }
extern "C" const char* mural_bigwig_chrom_name(const mural_bigwig_t* b, int32_t i) {
  return (b && i >= 0 && i < (int32_t)b->names.size()) ? b->names[i].c_str() : "";
}
extern "C" int64_t mural_bigwig_chrom_len(const mural_bigwig_t* b, int32_t i) {
  return (b && i >= 0 && i < (int32_t)b->sizes.size()) ? (int64_t)b->sizes[i] : -1;
}
// out[i] = mean(nan_to_num(values(chrom, max(lo[i], 0), min(hi[i], chrom length)))) — lo/hi: the expanded window, end exclusive
extern "C" int mural_bigwig_window_means(mural_bigwig_t* b, const char* chrom, int64_t n, const int64_t* lo, const int64_t* hi, double* out) {
  MURAL_CHECK(b && chrom && (n == 0 || (lo && hi && out)), "NULL argument");
  int ci = -1;
  for (size_t i = 0; i < b->names.size(); ++i)
    if (b->names[i] == chrom) ci = (int)i;
  MURAL_CHECK(ci >= 0, std::string("bigWig: no chromosome named ") + chrom);
  if (int rc = decode_chrom(b, ci)) return rc;
  const std::vector<double>& ps = b->prefix[b->ids[ci]];
  const int64_t len = b->sizes[ci];
  for (int64_t i = 0; i < n; ++i) {
    const int64_t a = lo[i] < 0 ? 0 : lo[i], e = hi[i] > len ? len : hi[i];
    out[i] = e > a ? (ps[e] - ps[a]) / double(e - a) : NAN;   // np.mean of an empty array
  }
  return 0;
}
extern "C" void mural_bigwig_close(mural_bigwig_t* b) {
  if (!b) return;
  if (b->f) fclose(b->f);
  delete b;
}
