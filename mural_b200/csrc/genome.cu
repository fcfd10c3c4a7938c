// Packed reference genome: host-side packer + device upload.
// Replaces the python-str genome of the reference (MuRaL/data/preprocessing.py:836, 458, 964).
#include <string.h>

#include <map>
#include <mutex>

#include "common.cuh"

namespace mural {

static thread_local std::string t_err;
int64_t g_launches = 0;

void set_error(const std::string& msg) { t_err = msg; }
int fail(const char* file, int line, const std::string& msg) {
  const char* base = strrchr(file, '/');
  t_err = std::string(base ? base + 1 : file) + ":" + std::to_string(line) + ": " + msg;
  return 1;
}

// ---- per-kernel CUDA-event profile (bench.py roofline leg) ------------------------------------
bool g_prof = false;
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_recs;
void prof_pre(const char* name, cudaStream_t st) {
  ProfRec r{name, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
}
void prof_post(cudaStream_t st) { cudaEventRecord(g_recs.back().b, st); }

// ASCII -> symbol (0..14), 255 = illegal.  Case-folded like .upper() (preprocessing.py:694,802).
static const uint8_t* ascii_table() {
  static uint8_t t[256];
  static std::once_flag once;
  std::call_once(once, [] {
    memset(t, 255, sizeof(t));
    const char* s = "ACGTRYMSWKBDHVN";
    for (int i = 0; s[i]; ++i) {
      t[(unsigned char)s[i]] = (uint8_t)i;
      t[(unsigned char)(s[i] + 32)] = (uint8_t)i;
    }
  });
  return t;
}

}  // namespace mural

using namespace mural;

extern "C" const char* mural_last_error(void) { return mural::t_err.c_str(); }
extern "C" int mural_abi_version(void) { return MURAL_ABI_VERSION; }
extern "C" int64_t mural_launch_count(void) { return mural::g_launches; }
extern "C" void mural_reset_launch_count(void) { mural::g_launches = 0; }

extern "C" void mural_profile_begin(void) {
  for (auto& r : mural::g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  mural::g_recs.clear();
  mural::g_prof = true;
}
// Writes {"kernel": {"count": n, "ms": total}, ...} (JSON) into buf; returns the number of records.
extern "C" int64_t mural_profile_end(char* buf, int64_t cap) {
  mural::g_prof = false;
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<int64_t, double>> agg;
  for (auto& r : mural::g_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  const int64_t n = (int64_t)mural::g_recs.size();
  mural::g_recs.clear();
  std::string js = "{";
  for (auto& kv : agg) {
    if (js.size() > 1) js += ", ";
    char tmp[256];
    snprintf(tmp, sizeof(tmp), "\"%s\": {\"count\": %lld, \"ms\": %.6f}", kv.first.c_str(), (long long)kv.second.first, kv.second.second);
    js += tmp;
  }
  js += "}";
  if (buf && cap > 0) {
    strncpy(buf, js.c_str(), (size_t)cap - 1);
    buf[cap - 1] = 0;
  }
  return n;
}

extern "C" int mural_genome_create(int32_t n_chrom, const char* const* h_seqs, const int64_t* h_lens, int device,
                                   mural_genome_t** out) {
  MURAL_CHECK(out != nullptr, "out is NULL");
  *out = nullptr;
  MURAL_CHECK(n_chrom > 0 && n_chrom < (1 << 23), "n_chrom out of range");
  const uint8_t* tab = ascii_table();
  std::vector<int64_t> off(n_chrom), len(n_chrom);
  int64_t total = 0;
  for (int c = 0; c < n_chrom; ++c) {
    MURAL_CHECK(h_lens[c] >= 0 && h_lens[c] < (int64_t(1) << 31), "chromosome length must fit int32 coordinates");
    off[c] = total;
    len[c] = h_lens[c];
    total += (h_lens[c] + 63) & ~int64_t(63);
  }
  total += 64;  // slack so word loads one past the end stay in bounds
  std::vector<uint32_t> bits(total / 16, 0u), mask(total / 32, 0u);
  std::vector<int64_t> es, ee;
  std::vector<uint8_t> ey;
  for (int c = 0; c < n_chrom; ++c) {
    const unsigned char* s = (const unsigned char*)h_seqs[c];
    int64_t run_start = -1;
    int run_sym = -1;
    for (int64_t i = 0; i < len[c]; ++i) {
      const uint8_t v = tab[s[i]];
      if (v == 255) {
        char buf[128];
        snprintf(buf, sizeof(buf), "KeyError: '%c' (chromosome %d, position %lld) is not an IUPAC nucleotide code",
                 s[i] >= 32 && s[i] < 127 ? s[i] : '?', c, (long long)i);
        MURAL_FAIL(buf);
      }
      const int64_t g = off[c] + i;
      if (v < 4) {
        bits[g >> 4] |= uint32_t(v) << ((g & 15) * 2);
        if (run_start >= 0) { es.push_back(run_start); ee.push_back(g); ey.push_back((uint8_t)run_sym); run_start = -1; }
      } else {
        mask[g >> 5] |= 1u << (g & 31);
        if (run_start >= 0 && run_sym != v) { es.push_back(run_start); ee.push_back(g); ey.push_back((uint8_t)run_sym); run_start = -1; }
        if (run_start < 0) { run_start = g; run_sym = v; }
      }
    }
    if (run_start >= 0) { es.push_back(run_start); ee.push_back(off[c] + len[c]); ey.push_back((uint8_t)run_sym); }
  }
  MURAL_CHECK(es.size() < (size_t(1) << 31), "too many non-ACGT runs");
  CUDA_TRY(cudaSetDevice(device));
  auto al = [](int64_t x) { return (x + 255) & ~int64_t(255); };
  const int64_t n_exc = (int64_t)es.size();
  const int64_t b_bits = al(bits.size() * 4), b_mask = al(mask.size() * 4), b_off = al(n_chrom * 8),
                b_exc8 = al((n_exc + 1) * 8), b_excs = al(n_exc + 1);
  const int64_t bytes = b_bits + b_mask + 2 * b_off + 2 * b_exc8 + b_excs;
  char* d = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d, bytes));
  mural_genome* G = new mural_genome();
  G->d_block = d;
  G->device = device;
  G->device_bytes = bytes;
  G->total_bases = total;
  G->h_off = off;
  G->h_len = len;
  char* p = d;
  auto put = [&](const void* src, int64_t nbytes, int64_t slot) -> const void* {
    const void* r = p;
    if (nbytes) cudaMemcpy(p, src, nbytes, cudaMemcpyHostToDevice);
    p += slot;
    return r;
  };
  G->view.bits2 = (const uint32_t*)put(bits.data(), bits.size() * 4, b_bits);
  G->view.mask = (const uint32_t*)put(mask.data(), mask.size() * 4, b_mask);
  G->view.chrom_off = (const int64_t*)put(off.data(), n_chrom * 8, b_off);
  G->view.chrom_len = (const int64_t*)put(len.data(), n_chrom * 8, b_off);
  G->view.exc_start = (const int64_t*)put(es.data(), n_exc * 8, b_exc8);
  G->view.exc_end = (const int64_t*)put(ee.data(), n_exc * 8, b_exc8);
  G->view.exc_sym = (const uint8_t*)put(ey.data(), n_exc, b_excs);
  G->view.n_chrom = n_chrom;
  G->view.n_exc = (int32_t)n_exc;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(d);
    delete G;
    MURAL_FAIL(std::string("genome upload: ") + cudaGetErrorString(e));
  }
  *out = G;
  return 0;
}

extern "C" void mural_genome_destroy(mural_genome_t* g) {
  if (!g) return;
  cudaFree(g->d_block);
  delete g;
}
extern "C" int32_t mural_genome_n_chrom(const mural_genome_t* g) { return g ? g->view.n_chrom : 0; }
extern "C" int64_t mural_genome_chrom_len(const mural_genome_t* g, int32_t c) {
  return (g && c >= 0 && c < g->view.n_chrom) ? g->h_len[c] : -1;
}
extern "C" int64_t mural_genome_device_bytes(const mural_genome_t* g) { return g ? g->device_bytes : 0; }
extern "C" int64_t mural_genome_n_exception_runs(const mural_genome_t* g) { return g ? g->view.n_exc : 0; }

// Host copy of the non-ACGT run table as (chromosome, start, end) in chromosome coordinates, end exclusive, sorted.
extern "C" int mural_genome_exception_runs(const mural_genome_t* g, int32_t* h_chrom, int64_t* h_start, int64_t* h_end) {
  MURAL_CHECK(g && (g->view.n_exc == 0 || (h_chrom && h_start && h_end)), "NULL argument");
  const int64_t n = g->view.n_exc;
  if (n == 0) return 0;
  CUDA_TRY(cudaSetDevice(g->device));
  CUDA_TRY(cudaMemcpy(h_start, g->view.exc_start, n * 8, cudaMemcpyDeviceToHost));
  CUDA_TRY(cudaMemcpy(h_end, g->view.exc_end, n * 8, cudaMemcpyDeviceToHost));
  for (int64_t i = 0; i < n; ++i) {  // global -> chromosome coordinates (runs never span chromosomes)
    int c = 0;
    while (c + 1 < (int)g->h_off.size() && g->h_off[c + 1] <= h_start[i]) ++c;
    h_chrom[i] = c;
    h_start[i] -= g->h_off[c];
    h_end[i] -= g->h_off[c];
  }
  return 0;
}
