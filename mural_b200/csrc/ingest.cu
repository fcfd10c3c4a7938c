// Host-side ingest: streaming BED and FASTA readers replacing `BedTool(test_file)` iteration
// (MuRaL/data/preprocessing.py:39-106 reads .chrom/.start/.stop/.score/.strand per record through pybedtools) and
// `SeqIO.to_dict(SeqIO.parse(open(ref_genome), 'fasta'))` (preprocessing.py:836).  Plain or gzip files (zlib).
// At genome-wide scale (5e7 sites, 3e9 bases) the Python readers take minutes while the network takes a second.
#include <limits.h>
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

struct mural_bed {
  std::vector<std::string> names;  // chromosome index -> name, order of first appearance
  std::vector<int32_t> chrom;
  std::vector<int64_t> start, end, label;
  std::vector<int8_t> strand;      // 0 '+', 1 anything else (bed_reader :97-101)
};
struct mural_fasta {
  std::vector<std::string> names;  // record id = first word of the header (Bio.SeqIO record.id)
  std::vector<std::string> seqs;
};

namespace {

// calls fn(line, len) for every line of a (possibly gzip-compressed) file; the trailing '\n' / '\r\n' is removed
template <class F>
int for_each_line(const char* path, F&& fn) {
  gzFile f = gzopen(path, "rb");
  if (!f) return 1;
  gzbuffer(f, 1 << 20);
  std::vector<char> buf(1 << 22);
  std::string carry;
  int n;
  while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) {
    const char* p = buf.data();
    const char* e = p + n;
    while (p < e) {
      const char* nl = (const char*)memchr(p, '\n', size_t(e - p));
      if (!nl) {
        carry.append(p, size_t(e - p));
        break;
      }
      if (!carry.empty()) {
        carry.append(p, size_t(nl - p));
        size_t len = carry.size();
        if (len && carry[len - 1] == '\r') --len;
        fn(carry.data(), len);
        carry.clear();
      } else {
        size_t len = size_t(nl - p);
        if (len && p[len - 1] == '\r') --len;
        fn(p, len);
      }
      p = nl + 1;
    }
  }
  if (!carry.empty()) {
    size_t len = carry.size();
    if (len && carry[len - 1] == '\r') --len;
    fn(carry.data(), len);
  }
  const bool bad = n < 0;
  gzclose(f);
  return bad ? 2 : 0;
}

}  // namespace

namespace {

// whole (inflated) file in memory; 0 ok, 1 cannot open, 2 read error
int slurp(const char* path, std::vector<char>& data) {
  gzFile f = gzopen(path, "rb");
  if (!f) return 1;
  gzbuffer(f, 1 << 20);
  size_t used = 0;
  data.resize(size_t(1) << 24);
  int n;
  while ((n = gzread(f, data.data() + used, (unsigned)std::min<size_t>(data.size() - used, size_t(1) << 30))) > 0) {
    used += size_t(n);
    if (data.size() - used < (size_t(1) << 22)) data.resize(data.size() * 2);
  }
  const bool bad = n < 0;
  gzclose(f);
  data.resize(used);
  return bad ? 2 : 0;
}

struct BedPart {                     // what one thread parsed from its slice of the file
  std::vector<std::string> names;   // first-appearance order inside the slice
  std::map<std::string, int32_t> index;
  std::vector<int32_t> chrom;
  std::vector<int64_t> start, end, label;
  std::vector<int8_t> strand;
  int64_t lines = 0, err_line = 0;  // err_line: 1-based line inside the slice of the first malformed record
  std::string err;
};

// BED6: chrom start end name score strand; score = label (preprocessing.py:752-754).  Same rules as
// mural_b200.data.SiteTable.from_bed: blank lines and '#', 'track', 'browser' lines are skipped; fields split on tabs,
// or on any whitespace when a line has fewer than three tab-separated fields; label = int(float(score)), '.'/'' -> 0.
void parse_bed_line(const char* s, size_t len, BedPart& b) {
  ++b.lines;
  if (!b.err.empty()) return;
  auto fail_line = [&](const char* what) { b.err = what; b.err_line = b.lines; };
  size_t i = 0;
  while (i < len && isspace((unsigned char)s[i])) ++i;
  if (i == len) return;                                    // blank
  if (s[0] == '#' || (len >= 5 && !memcmp(s, "track", 5)) || (len >= 7 && !memcmp(s, "browser", 7))) return;
  const char* f[6];
  size_t fl[6];
  int nf = 0;
  {
    size_t a = 0;
    for (size_t k = 0; k <= len && nf < 6; ++k)
      if (k == len || s[k] == '\t') { f[nf] = s + a; fl[nf] = k - a; ++nf; a = k + 1; }
  }
  if (nf < 3) {                                            // whitespace-separated
    nf = 0;
    size_t k = 0;
    while (k < len && nf < 6) {
      while (k < len && isspace((unsigned char)s[k])) ++k;
      if (k == len) break;
      const size_t a = k;
      while (k < len && !isspace((unsigned char)s[k])) ++k;
      f[nf] = s + a; fl[nf] = k - a; ++nf;
    }
  }
  if (nf < 3) return fail_line("fewer than 3 fields");
  char tmp[64];
  auto to_ll = [&](int j, long long& v) {
    if (fl[j] == 0 || fl[j] >= sizeof tmp) return false;
    memcpy(tmp, f[j], fl[j]); tmp[fl[j]] = 0;
    char* endp = nullptr;
    v = strtoll(tmp, &endp, 10);
    return endp && *endp == 0;
  };
  long long st, en;
  if (!to_ll(1, st) || !to_ll(2, en)) return fail_line("start/end are not integers");
  long long lab = 0;
  if (nf > 4 && !(fl[4] == 0 || (fl[4] == 1 && f[4][0] == '.'))) {
    if (fl[4] >= sizeof tmp) return fail_line("bad score");
    memcpy(tmp, f[4], fl[4]); tmp[fl[4]] = 0;
    char* endp = nullptr;
    const double d = strtod(tmp, &endp);
    if (!endp || *endp != 0) return fail_line("bad score");
    lab = (long long)d;                                    // int(float(score)) truncates toward zero
  }
  const std::string name(f[0], fl[0]);
  auto it = b.index.find(name);
  if (it == b.index.end()) {
    it = b.index.emplace(name, (int32_t)b.names.size()).first;
    b.names.push_back(name);
  }
  b.chrom.push_back(it->second);
  b.start.push_back(st);
  b.end.push_back(en);
  b.label.push_back(lab);
  b.strand.push_back((nf > 5 && fl[5] == 1 && f[5][0] == '+') ? 0 : 1);
}

void parse_bed_slice(const char* p, const char* e, BedPart& b) {
  while (p < e) {
    const char* nl = (const char*)memchr(p, '\n', size_t(e - p));
    const char* le = nl ? nl : e;
    size_t len = size_t(le - p);
    if (len && p[len - 1] == '\r') --len;
    parse_bed_line(p, len, b);
    p = le + 1;
  }
}

}  // namespace

// The file is inflated into memory once, cut into slices at line boundaries, parsed by a pool of threads and merged in file
// order (chromosome indices follow the order of first appearance in the FILE, as iterating a BedTool does).
extern "C" int mural_bed_read(const char* path, mural_bed_t** out) {
  MURAL_CHECK(path && out, "NULL argument");
  std::vector<char> data;
  const int rc = slurp(path, data);
  if (rc) MURAL_FAIL(rc == 1 ? std::string("cannot open ") + path : std::string("read error on ") + path);
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 4;
  size_t n_parts = std::min<size_t>(std::min<unsigned>(hw, 32u), data.size() / (size_t(4) << 20) + 1);
  std::vector<const char*> cut(n_parts + 1);
  const char* base = data.data();
  const char* endp = base + data.size();
  cut[0] = base;
  cut[n_parts] = endp;
  for (size_t k = 1; k < n_parts; ++k) {
    const char* q = base + data.size() * k / n_parts;
    if (q < cut[k - 1]) q = cut[k - 1];
    const char* nl = q < endp ? (const char*)memchr(q, '\n', size_t(endp - q)) : nullptr;
    cut[k] = nl ? nl + 1 : endp;
  }
  std::vector<BedPart> parts(n_parts);
  {
    std::vector<std::thread> pool;
    for (size_t k = 1; k < n_parts; ++k) pool.emplace_back([&, k] { parse_bed_slice(cut[k], cut[k + 1], parts[k]); });
    parse_bed_slice(cut[0], cut[1], parts[0]);
    for (auto& t : pool) t.join();
  }
  int64_t line0 = 0, total = 0;
  for (const BedPart& pt : parts) {
    if (!pt.err.empty())
      MURAL_FAIL("ValueError: " + std::string(path) + " line " + std::to_string(line0 + pt.err_line) + ": " + pt.err);
    line0 += pt.lines;
    total += (int64_t)pt.start.size();
  }
  mural_bed* b = new mural_bed();
  std::map<std::string, int32_t> index;
  b->chrom.resize(total); b->start.resize(total); b->end.resize(total); b->label.resize(total); b->strand.resize(total);
  std::vector<std::vector<int32_t>> remap(n_parts);
  std::vector<int64_t> off(n_parts + 1, 0);
  for (size_t k = 0; k < n_parts; ++k) {
    for (const std::string& nm : parts[k].names) {
      auto it = index.find(nm);
      if (it == index.end()) {
        it = index.emplace(nm, (int32_t)b->names.size()).first;
        b->names.push_back(nm);
      }
      remap[k].push_back(it->second);
    }
    off[k + 1] = off[k] + (int64_t)parts[k].start.size();
  }
  {
    auto copy_part = [&](size_t k) {
      const BedPart& pt = parts[k];
      const int64_t o = off[k], m = (int64_t)pt.start.size();
      for (int64_t r = 0; r < m; ++r) b->chrom[o + r] = remap[k][pt.chrom[r]];
      if (m) {
        memcpy(b->start.data() + o, pt.start.data(), m * sizeof(int64_t));
        memcpy(b->end.data() + o, pt.end.data(), m * sizeof(int64_t));
        memcpy(b->label.data() + o, pt.label.data(), m * sizeof(int64_t));
        memcpy(b->strand.data() + o, pt.strand.data(), m);
      }
    };
    std::vector<std::thread> pool;
    for (size_t k = 1; k < n_parts; ++k) pool.emplace_back(copy_part, k);
    copy_part(0);
    for (auto& t : pool) t.join();
  }
  *out = b;
  return 0;
}
extern "C" int64_t mural_bed_n(const mural_bed_t* b) { return b ? (int64_t)b->start.size() : 0; }
extern "C" int32_t mural_bed_n_chrom(const mural_bed_t* b) { return b ? (int32_t)b->names.size() : 0; }
extern "C" const char* mural_bed_chrom_name(const mural_bed_t* b, int32_t i) {
  return (b && i >= 0 && i < (int32_t)b->names.size()) ? b->names[i].c_str() : nullptr;
}
extern "C" int mural_bed_columns(const mural_bed_t* b, int32_t* chrom, int64_t* start, int64_t* end, int8_t* strand, int64_t* label) {
  MURAL_CHECK(b && chrom && start && end && strand && label, "NULL argument");
  const size_t n = b->start.size();
  if (n) {
    memcpy(chrom, b->chrom.data(), n * sizeof(int32_t));
    memcpy(start, b->start.data(), n * sizeof(int64_t));
    memcpy(end, b->end.data(), n * sizeof(int64_t));
    memcpy(strand, b->strand.data(), n);
    memcpy(label, b->label.data(), n * sizeof(int64_t));
  }
  return 0;
}
extern "C" void mural_bed_destroy(mural_bed_t* b) { delete b; }

// Site records in emission order from the BED columns in file order (the gathers + pack_meta of PackedSiteDataset in one pass):
// pos = start[perm] (int32), strand, label, chrom = chrom_map[chrom[perm]] (genome index of the BED's chromosome), and
// meta = strand | label << 1 | chrom << 8 (MURAL_META).  Labels outside [0, 127] or starts beyond int32 are an error.
extern "C" int mural_pack_sites(const int64_t* perm, int64_t n, const int32_t* chrom, const int64_t* start, const int8_t* strand,
                                const int64_t* label, const int64_t* chrom_map, int32_t n_chrom, int32_t* pos_out, int8_t* strand_out,
                                int64_t* label_out, int64_t* chrom_out, int32_t* meta_out) {
  MURAL_CHECK(n == 0 || (perm && chrom && start && strand && label && chrom_map && pos_out && strand_out && label_out && chrom_out && meta_out),
              "NULL argument");
  for (int64_t k = 0; k < n; ++k) {
    const int64_t i = perm[k];
    const int64_t lb = label[i], st = start[i];
    const int32_t c = chrom[i];
    if (lb < 0 || lb > 0x7f) MURAL_FAIL("BED score column (label) must be in [0, 127]; got " + std::to_string(lb));
    if (st < INT32_MIN || st > INT32_MAX) MURAL_FAIL("site position does not fit 32 bits: " + std::to_string(st));
    if (c < 0 || c >= n_chrom) MURAL_FAIL("chromosome index out of range");
    const int64_t gc = chrom_map[c];
    const int8_t sd = strand[i];
    pos_out[k] = (int32_t)st;
    strand_out[k] = sd;
    label_out[k] = lb;
    chrom_out[k] = gc;
    meta_out[k] = (int32_t)((gc << 8) | ((lb & 0x7f) << 1) | (sd & 1));
  }
  return 0;
}

// Emission order of bed_reader (MuRaL/data/preprocessing.py:39-106).  The reader walks the BED in file order with a window
// [s, s + segment_center) anchored at the first site of the first chromosome block (at 1 for every later block; a chromosome
// that re-appears opens a new block), moves the window forward until it holds the site (it never moves back), and emits for
// every window first its '+' sites, then its '-' sites, each in file order.  Sites that share (block, window) are therefore
// consecutive in the file and the emission order is a stable partition of every such run — one pass, no sort.
// perm[k] = file index of the k-th emitted site; batch_sizes (capacity n) receives the sizes of the non-empty batches.
extern "C" int mural_segment_order(const int32_t* chrom, const int64_t* start, const int8_t* strand, int64_t n, int64_t segment_center,
                                   int64_t* perm, int64_t* batch_sizes, int64_t* n_batches) {
  MURAL_CHECK(n_batches != nullptr && (n == 0 || (chrom && start && strand && perm && batch_sizes)), "NULL argument");
  MURAL_CHECK(segment_center > 0, "segment_center must be positive");
  *n_batches = 0;
  if (n == 0) return 0;
  int64_t nb = 0, out = 0, run_start = 0;
  int64_t anchor = start[0], cur_win = -1;
  auto flush = [&](int64_t a, int64_t b) {   // file rows [a, b) share a window
    const int64_t o0 = out;
    for (int64_t i = a; i < b; ++i)
      if (strand[i] == 0) perm[out++] = i;
    const int64_t n_plus = out - o0;
    for (int64_t i = a; i < b; ++i)
      if (strand[i] != 0) perm[out++] = i;
    if (n_plus) batch_sizes[nb++] = n_plus;
    if (out - o0 - n_plus) batch_sizes[nb++] = out - o0 - n_plus;
  };
  for (int64_t i = 0; i < n; ++i) {
    const bool new_block = i > 0 && chrom[i] != chrom[i - 1];
    if (new_block) anchor = 1;
    const int64_t a = start[i] - anchor + segment_center - 1;
    int64_t q = a / segment_center;                    // floor division (a may be negative: a site in front of the anchor)
    if (a % segment_center != 0 && a < 0) --q;
    int64_t win = q - 1;                               // first window whose end is >= start
    if (win < 0) win = 0;
    if (new_block || i == 0) {
      if (i > 0) flush(run_start, i);
      run_start = i;
      cur_win = win;
    } else if (win > cur_win) {                        // the window only moves forward inside a block
      flush(run_start, i);
      run_start = i;
      cur_win = win;
    }
  }
  flush(run_start, n);
  *n_batches = nb;
  return 0;
}

// FASTA -> records in file order; sequence lines are stripped of surrounding whitespace and concatenated; a repeated
// record id is an error (SeqIO.to_dict raises ValueError("Duplicate key ...")).
extern "C" int mural_fasta_read(const char* path, mural_fasta_t** out) {
  MURAL_CHECK(path && out, "NULL argument");
  mural_fasta* fa = new mural_fasta();
  std::map<std::string, int> seen;
  std::string err;
  const int rc = for_each_line(path, [&](const char* s, size_t len) {
    if (!err.empty()) return;
    if (len && s[0] == '>') {
      size_t a = 1;
      while (a < len && isspace((unsigned char)s[a])) ++a;
      size_t e = a;
      while (e < len && !isspace((unsigned char)s[e])) ++e;
      std::string name(s + a, e - a);
      if (seen.count(name)) { err = "ValueError: Duplicate key '" + name + "'"; return; }
      seen[name] = 1;
      fa->names.push_back(name);
      fa->seqs.emplace_back();
      return;
    }
    if (fa->seqs.empty()) return;                            // text before the first header is ignored
    size_t a = 0, e = len;
    while (a < e && isspace((unsigned char)s[a])) ++a;
    while (e > a && isspace((unsigned char)s[e - 1])) --e;
    fa->seqs.back().append(s + a, e - a);
  });
  if (rc || !err.empty()) {
    delete fa;
    MURAL_FAIL(rc == 1 ? std::string("cannot open ") + path : (rc == 2 ? std::string("read error on ") + path : err));
  }
  *out = fa;
  return 0;
}
extern "C" int32_t mural_fasta_n(const mural_fasta_t* f) { return f ? (int32_t)f->names.size() : 0; }
extern "C" const char* mural_fasta_name(const mural_fasta_t* f, int32_t i) {
  return (f && i >= 0 && i < (int32_t)f->names.size()) ? f->names[i].c_str() : nullptr;
}
extern "C" const char* mural_fasta_seq(const mural_fasta_t* f, int32_t i) {
  return (f && i >= 0 && i < (int32_t)f->seqs.size()) ? f->seqs[i].data() : nullptr;
}
extern "C" int64_t mural_fasta_len(const mural_fasta_t* f, int32_t i) {
  return (f && i >= 0 && i < (int32_t)f->seqs.size()) ? (int64_t)f->seqs[i].size() : -1;
}
extern "C" void mural_fasta_destroy(mural_fasta_t* f) { delete f; }
