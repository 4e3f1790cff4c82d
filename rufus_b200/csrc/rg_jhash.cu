// rg_jhash.cu — the on-disk `.Jhash` form of a counted table, byte-compatible
// with Jellyfish 2.2.5's binary/sorted format so that files written here can be read, queried and
// merged by the real tools (and `-e` exclude files written by real Jellyfish can be loaded here).
//
//   header  : generic_file_header.hpp:96-119 (9 decimal digits = length, terse JSON, NUL padding to 8)
//             fields of file_header.hpp:34-41 + count_main.cc:247-252
//   records : binary_dumper.hpp:44-48  key = ceil(2k/8) bytes LE, count = counter_len bytes LE, saturated
//   order   : sorted_dumper.hpp:80-111 / mer_heap.hpp:34-38: ascending (pos, key) with
//             pos = M * key & (size-1), M = the table's hash matrix (rectangular_binary_matrix.hpp:206-242)
//   matrix  : large_hash_array.hpp:948: the pseudo-inverse of the first pseudo-invertible random
//             r x 2k binary matrix drawn from glibc's UNSEEDED random() (lib/misc.cc:74-80), hence the
//             same for every run with the same (size, k); `jellyfish merge` refuses files whose
//             matrices differ (merge_files.cc:193-202), so it is regenerated here bit for bit from a
//             private random_r() state seeded like the default one.
//
// The ordering work runs on the GPU: pos kernel + two stable CUB radix sorts (by key, then by pos) +
// a pack kernel that lays the records out as file bytes; the host only writes the buffer.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <cub/cub.cuh>
#include <string>
#include <vector>

#include "rg_internal.cuh"

// ---------------------------------------------------------------------------
// hash matrix
// ---------------------------------------------------------------------------
namespace {

struct JfRng {
  struct random_data rd;
  char state[128];
  JfRng() {
    memset(&rd, 0, sizeof(rd));
    memset(state, 0, sizeof(state));
    initstate_r(1, state, sizeof(state), &rd);     // the state random() starts from when never seeded
  }
  // lib/misc.cc:74-80 with length 64 and RAND_MAX = 2^31-1: three draws, 30 bits apart
  uint64_t bits64() {
    uint64_t res = 0;
    for (int i = 0; i < 64; i += 30) {
      int32_t v;
      random_r(&rd, &v);
      res ^= (uint64_t)v << i;
    }
    return res;
  }
};

// pseudo-inverse of a r x c binary matrix given by columns (bit i of cols[j] = row i), viewed as square
// with an identity block above (rectangular_binary_matrix.cc:138-186). false if singular.
bool jf_pseudo_inverse(const uint64_t* cols, unsigned r, unsigned c, uint64_t* out) {
  std::vector<uint64_t> piv(cols, cols + c);
  const unsigned srow = std::min(r, c), scol = c - srow;
  for (unsigned i = 0; i < c; ++i) out[i] = 0;
  out[scol] = 1ull << (srow - 1);
  for (unsigned i = scol + 1; i < c; ++i) out[i] = out[i - 1] >> 1;
  uint64_t mask = 1ull << (srow - 1);
  for (unsigned i = scol; i < c; ++i, mask >>= 1) {
    if (!(piv[i] & mask)) {
      unsigned j = i + 1;
      while (j < c && !(piv[j] & mask)) ++j;
      if (j == c) return false;
      piv[i] ^= piv[j];
      out[i] ^= out[j];
    }
    for (unsigned j = i + 1; j < c; ++j)
      if (piv[j] & mask) {
        piv[j] ^= piv[i];
        out[j] ^= out[i];
      }
  }
  mask = 1ull << (srow - 1);
  for (unsigned i = scol; i < c; ++i, mask >>= 1)
    for (unsigned j = 0; j < i; ++j)
      if (piv[j] & mask) {
        piv[j] ^= piv[i];
        out[j] ^= out[i];
      }
  return true;
}

unsigned ceil_log2(uint64_t x) {
  unsigned l = 0;
  while (l < 63 && (1ull << l) < x) ++l;
  return l;
}

// quadratic reprobes of lib/storage.cc:21-50
uint64_t jf_reprobe(unsigned i) { return i == 0 ? 1 : (uint64_t)i * (i + 1) / 2; }

}  // namespace

struct JfGeometry {
  unsigned r = 0, c = 0;
  uint64_t size = 0;
  unsigned max_reprobe = 126;
  uint64_t columns[64];
};

static int jf_geometry(uint64_t size_arg, int k, JfGeometry& g) {
  if (k < 1 || k > 32 || size_arg == 0) {
    rg_set_error("jhash: k must be 1..32 and size > 0");
    return RG_ERR_ARG;
  }
  g.r = ceil_log2(size_arg);                   // large_hash_array.hpp:164
  if (g.r == 0) g.r = 1;
  g.c = 2 * k;
  g.size = 1ull << g.r;
  g.max_reprobe = 126;                         // count_main's default --reprobes
  while (jf_reprobe(g.max_reprobe) >= g.size && g.max_reprobe >= 1) g.max_reprobe--;   // large_hash_array.hpp:40-45
  JfRng rng;
  const uint64_t cmask = ~0ull >> (64 - g.r);
  uint64_t raw[64];
  for (;;) {
    for (unsigned i = 0; i < g.c; ++i) raw[i] = rng.bits64() & cmask;
    if (jf_pseudo_inverse(raw, g.r, g.c, g.columns)) return 0;
  }
}

// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
struct JfMatrixDev {
  uint64_t col[64];   // col[j] is XORed in when bit j of the key is set (= columns[c-1-j])
  unsigned c;
  uint64_t mask;
};

__global__ void jf_pos_kernel(const uint64_t* __restrict__ keys, uint64_t n, JfMatrixDev m, uint64_t* __restrict__ pos,
                              uint32_t* __restrict__ idx) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t x = keys[i], p = 0;
    for (unsigned j = 0; j < m.c; ++j, x >>= 1) p ^= (0ull - (x & 1ull)) & m.col[j];
    pos[i] = p & m.mask;
    if (idx) idx[i] = (uint32_t)i;
  }
}

__global__ void jf_gather_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts,
                                 const uint32_t* __restrict__ idx, uint64_t n, uint64_t* __restrict__ keys_out,
                                 uint32_t* __restrict__ counts_out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t s = idx[i];
    keys_out[i] = keys[s];
    counts_out[i] = counts[s];
  }
}

// records as file bytes: key_bytes little-endian key bytes then val_bytes little-endian count bytes
__global__ void jf_pack_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n,
                               unsigned key_bytes, unsigned val_bytes, unsigned char* __restrict__ out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    unsigned char* rec = out + i * (key_bytes + val_bytes);
    uint64_t kx = keys[i];
    for (unsigned b = 0; b < key_bytes; ++b, kx >>= 8) rec[b] = (unsigned char)kx;
    uint64_t v = counts[i];
    if (val_bytes < 8 && v > ((1ull << (8 * val_bytes)) - 1)) v = (1ull << (8 * val_bytes)) - 1;   // binary_dumper.hpp:46
    for (unsigned b = 0; b < val_bytes; ++b, v >>= 8) rec[key_bytes + b] = (unsigned char)v;
  }
}

__global__ void jf_unpack_kernel(const unsigned char* __restrict__ in, uint64_t n, unsigned key_bytes,
                                 unsigned val_bytes, uint64_t* __restrict__ keys, uint32_t* __restrict__ counts) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned char* rec = in + i * (key_bytes + val_bytes);
    uint64_t kx = 0, v = 0;
    for (unsigned b = 0; b < key_bytes; ++b) kx |= (uint64_t)rec[b] << (8 * b);
    for (unsigned b = 0; b < val_bytes; ++b) v |= (uint64_t)rec[key_bytes + b] << (8 * b);
    keys[i] = kx;
    counts[i] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)v;
  }
}

static unsigned jf_grid(rg_ctx* ctx, uint64_t n) {
  const uint64_t need = (n + 255) / 256;
  return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(need, rg_grid(ctx, 8)));
}

static JfMatrixDev jf_matrix_dev(const JfGeometry& g) {
  JfMatrixDev m;
  memset(&m, 0, sizeof(m));
  m.c = g.c;
  m.mask = g.size - 1;
  for (unsigned j = 0; j < g.c; ++j) m.col[j] = g.columns[g.c - 1 - j];
  return m;
}

// Reorder n (key, count) pairs on the device into ascending (pos, key). keys/counts are overwritten.
static int jf_sort_dev(rg_ctx* ctx, const JfGeometry& g, uint64_t* keys, uint32_t* counts, uint64_t n) {
  if (n < 2) return 0;
  if (n >= (1ull << 32)) {
    rg_set_error("jhash: more than 2^32 records in one piece");
    return RG_ERR_ARG;
  }
  uint64_t *k2 = nullptr, *pos = nullptr, *pos2 = nullptr;
  uint32_t *c2 = nullptr, *idx = nullptr, *idx2 = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, tb2 = 0;
  int rc = 0;
  auto fail = [&](cudaError_t e) {
    rg_set_error(std::string("jhash sort: ") + cudaGetErrorString(e));
    rc = RG_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaMalloc(&k2, n * 8)) != cudaSuccess || (e = cudaMalloc(&c2, n * 4)) != cudaSuccess ||
      (e = cudaMalloc(&pos, n * 8)) != cudaSuccess || (e = cudaMalloc(&pos2, n * 8)) != cudaSuccess ||
      (e = cudaMalloc(&idx, n * 4)) != cudaSuccess || (e = cudaMalloc(&idx2, n * 4)) != cudaSuccess) {
    fail(e);
  }
  if (rc == 0) {
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, k2, counts, c2, (int64_t)n, 0, (int)g.c, ctx->stream);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, pos, pos2, idx, idx2, (int64_t)n, 0, (int)g.r, ctx->stream);
    tmp_bytes = std::max(tmp_bytes, tb2);
    if ((e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16)) != cudaSuccess) fail(e);
  }
  if (rc == 0) {
    // 1. by key; 2. stable by pos (radix sort is stable) => (pos, key)
    ctx->launches++;
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, k2, counts, c2, (int64_t)n, 0, (int)g.c, ctx->stream);
    if (e != cudaSuccess) fail(e);
  }
  if (rc == 0) {
    {
      rg_launch_scope _ls(ctx, "jf_pos_kernel");
      jf_pos_kernel<<<jf_grid(ctx, n), 256, 0, ctx->stream>>>(k2, n, jf_matrix_dev(g), pos, idx);
    }
    ctx->launches++;
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, pos, pos2, idx, idx2, (int64_t)n, 0, (int)g.r, ctx->stream);
    if (e != cudaSuccess) fail(e);
  }
  if (rc == 0) {
    rg_launch_scope _ls(ctx, "jf_gather_kernel");
    jf_gather_kernel<<<jf_grid(ctx, n), 256, 0, ctx->stream>>>(k2, c2, idx2, n, keys, counts);
  }
  if (rc == 0 && (e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) fail(e);
  cudaFree(k2); cudaFree(c2); cudaFree(pos); cudaFree(pos2); cudaFree(idx); cudaFree(idx2); cudaFree(tmp);
  return rc;
}

// ---------------------------------------------------------------------------
// header text
// ---------------------------------------------------------------------------
static std::string json_escape(const std::string& s) {
  std::string o;
  for (unsigned char ch : s) {
    switch (ch) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\t': o += "\\t"; break;
      case '\r': o += "\\r"; break;
      default:
        if (ch < 0x20) {
          char b[8];
          snprintf(b, sizeof(b), "\\u%04x", ch);
          o += b;
        } else {
          o += (char)ch;
        }
    }
  }
  return o;
}

// the keys in the (alphabetical) order jsoncpp's FastWriter emits them
static std::string jf_header_text(const JfGeometry& g, int canonical, unsigned counter_len, unsigned val_len,
                                  const char* const* cmdline, int n_cmdline) {
  std::string h = "{\"alignment\":8,\"canonical\":";
  h += canonical ? "true" : "false";
  h += ",\"cmdline\":[";
  for (int i = 0; i < n_cmdline; ++i) {
    if (i) h += ",";
    h += "\"" + json_escape(cmdline[i]) + "\"";
  }
  h += "],\"counter_len\":" + std::to_string(counter_len);
  char buf[4096];
  ssize_t nexe = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
  h += ",\"exe_path\":\"" + json_escape(nexe > 0 ? std::string(buf, (size_t)nexe) : std::string("")) + "\"";
  h += ",\"format\":\"binary/sorted\"";
  if (gethostname(buf, sizeof(buf)) != 0) buf[0] = 0;
  buf[sizeof(buf) - 1] = 0;
  h += ",\"hostname\":\"" + json_escape(buf) + "\"";
  h += ",\"key_len\":" + std::to_string(g.c);
  h += ",\"matrix1\":{\"c\":" + std::to_string(g.c) + ",\"columns\":[";
  for (unsigned i = 0; i < g.c; ++i) {
    if (i) h += ",";
    h += std::to_string(g.columns[i]);
  }
  h += "],\"r\":" + std::to_string(g.r) + "}";
  h += ",\"max_reprobe\":" + std::to_string(g.max_reprobe);
  h += ",\"pwd\":\"" + json_escape(getcwd(buf, sizeof(buf)) ? buf : "") + "\"";
  h += ",\"reprobes\":[";
  for (unsigned i = 0; i <= g.max_reprobe; ++i) {
    if (i) h += ",";
    h += std::to_string(jf_reprobe(i));
  }
  h += "],\"size\":" + std::to_string(g.size);
  time_t now = time(nullptr);
  std::string ts = ctime_r(&now, buf) ? buf : "";
  while (!ts.empty() && ts.back() == '\n') ts.pop_back();
  h += ",\"time\":\"" + json_escape(ts) + "\"";
  h += ",\"val_len\":" + std::to_string(val_len) + "}";
  return h;
}

// ---- minimal JSON reader (objects, arrays, strings, numbers, literals): enough for the header ----
namespace {
struct JsonCursor {
  const char* p;
  const char* end;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool skip_string(std::string* out) {
    if (p >= end || *p != '"') return false;
    ++p;
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) {
        if (out) out->push_back(p[1]);
        p += 2;
      } else {
        if (out) out->push_back(*p);
        ++p;
      }
    }
    if (p >= end) return false;
    ++p;
    return true;
  }
  bool skip_value() {
    ws();
    if (p >= end) return false;
    if (*p == '"') return skip_string(nullptr);
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      ++p;
      ws();
      if (p < end && *p == close) { ++p; return true; }
      for (;;) {
        ws();
        if (open == '{') {
          if (!skip_string(nullptr)) return false;
          ws();
          if (p >= end || *p != ':') return false;
          ++p;
        }
        if (!skip_value()) return false;
        ws();
        if (p < end && *p == ',') { ++p; continue; }
        if (p < end && *p == close) { ++p; return true; }
        return false;
      }
    }
    while (p < end && *p != ',' && *p != '}' && *p != ']') ++p;   // number / true / false / null
    return true;
  }
};

// text of the value of top-level (or nested, for an object text) key `name`; empty if absent
std::string json_member(const std::string& obj, const char* name) {
  JsonCursor c{obj.data(), obj.data() + obj.size()};
  c.ws();
  if (c.p >= c.end || *c.p != '{') return "";
  ++c.p;
  for (;;) {
    c.ws();
    std::string key;
    if (!c.skip_string(&key)) return "";
    c.ws();
    if (c.p >= c.end || *c.p != ':') return "";
    ++c.p;
    c.ws();
    const char* v0 = c.p;
    if (!c.skip_value()) return "";
    if (key == name) return std::string(v0, c.p);
    c.ws();
    if (c.p < c.end && *c.p == ',') { ++c.p; continue; }
    return "";
  }
}
}  // namespace

struct JfFileHeader {
  JfGeometry g;
  int k = 0, canonical = 0;
  unsigned counter_len = 4, key_bytes = 0;
  uint64_t offset = 0;    // first record byte
  std::string format;
};

static int jf_read_header(FILE* f, const char* path, JfFileHeader& h) {
  char digits[10] = {0};
  if (fread(digits, 1, 9, f) != 9) {
    rg_set_error(std::string("jhash: cannot read header of ") + path);
    return RG_ERR_IO;
  }
  for (int i = 0; i < 9; ++i)
    if (digits[i] < '0' || digits[i] > '9') {
      rg_set_error(std::string("jhash: not a Jellyfish binary file: ") + path);
      return RG_ERR_IO;
    }
  const uint64_t hlen = strtoull(digits, nullptr, 10);
  std::string text(hlen, '\0');
  if (fread(&text[0], 1, hlen, f) != hlen) {
    rg_set_error(std::string("jhash: truncated header in ") + path);
    return RG_ERR_IO;
  }
  while (!text.empty() && text.back() == '\0') text.pop_back();
  h.offset = 9 + hlen;
  h.format = json_member(text, "format");
  if (h.format != "\"binary/sorted\"") {
    rg_set_error(std::string("jhash: unsupported format ") + h.format + " in " + path);
    return RG_ERR_IO;
  }
  const std::string key_len = json_member(text, "key_len");
  if (key_len.empty()) {
    rg_set_error(std::string("jhash: header without key_len in ") + path);
    return RG_ERR_IO;
  }
  h.g.c = (unsigned)strtoul(key_len.c_str(), nullptr, 10);
  h.k = (int)h.g.c / 2;
  h.key_bytes = (h.g.c + 7) / 8;
  h.canonical = json_member(text, "canonical") == "true";
  const std::string cl = json_member(text, "counter_len");
  h.counter_len = cl.empty() ? 4 : (unsigned)strtoul(cl.c_str(), nullptr, 10);
  h.g.size = strtoull(json_member(text, "size").c_str(), nullptr, 10);
  h.g.max_reprobe = (unsigned)strtoul(json_member(text, "max_reprobe").c_str(), nullptr, 10);
  const std::string m = json_member(text, "matrix1");
  h.g.r = (unsigned)strtoul(json_member(m, "r").c_str(), nullptr, 10);
  std::string cols = json_member(m, "columns");
  memset(h.g.columns, 0, sizeof(h.g.columns));
  unsigned nc = 0;
  for (const char* p = cols.c_str(); *p && nc < 64;) {
    if (*p >= '0' && *p <= '9') {
      char* e;
      h.g.columns[nc++] = strtoull(p, &e, 10);
      p = e;
    } else {
      ++p;
    }
  }
  if (h.k < 1 || h.k > 32 || nc != h.g.c || h.counter_len < 1 || h.counter_len > 8) {
    rg_set_error(std::string("jhash: unsupported header values in ") + path);
    return RG_ERR_IO;
  }
  return 0;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
// host bytes of a file, not zero-filled before they are read into
struct JfBytes {
  unsigned char* p = nullptr;
  size_t n = 0, cap = 0;
  JfBytes() = default;
  JfBytes(const JfBytes&) = delete;
  JfBytes& operator=(const JfBytes&) = delete;
  ~JfBytes() { free(p); }
  bool reserve(size_t want) {
    if (want <= cap) return true;
    unsigned char* q = (unsigned char*)realloc(p, want ? want : 1);
    if (!q) return false;
    p = q;
    cap = want ? want : 1;
    return true;
  }
  bool append(const unsigned char* src, size_t m) {
    if (n + m > cap && !reserve(std::max(n + m, cap + cap / 2 + (1u << 20)))) return false;
    memcpy(p + n, src, m);
    n += m;
    return true;
  }
  size_t size() const { return n; }
  const unsigned char* data() const { return p; }
};

extern "C" {

int rg_jf_matrix(uint64_t size, int k, uint64_t* columns_out, uint32_t* r_out, uint64_t* size_out) {
  JfGeometry g;
  const int rc = jf_geometry(size, k, g);
  if (rc) return rc;
  if (columns_out) memcpy(columns_out, g.columns, g.c * sizeof(uint64_t));
  if (r_out) *r_out = g.r;
  if (size_out) *size_out = g.size;
  return 0;
}

uint64_t rg_jf_pos(uint64_t jf_key, uint64_t size, int k) {
  JfGeometry g;
  if (jf_geometry(size, k, g)) return ~0ull;
  uint64_t p = 0;
  for (unsigned j = 0; j < g.c; ++j)
    if ((jf_key >> j) & 1ull) p ^= g.columns[g.c - 1 - j];
  return p & (g.size - 1);
}

int64_t rg_jf_sort(rg_ctx* ctx, uint64_t* keys_host, uint32_t* counts_host, uint64_t n, uint64_t size, int k) {
  if (!ctx || (n && (!keys_host || !counts_host))) return RG_ERR_ARG;
  JfGeometry g;
  int rc = jf_geometry(size, k, g);
  if (rc) return rc;
  if (n == 0) return 0;
  uint64_t* dk = nullptr;
  uint32_t* dc = nullptr;
  RG_CUDA_OK(cudaMalloc(&dk, n * 8));
  RG_CUDA_OK(cudaMalloc(&dc, n * 4));
  cudaMemcpyAsync(dk, keys_host, n * 8, cudaMemcpyHostToDevice, ctx->stream);
  cudaMemcpyAsync(dc, counts_host, n * 4, cudaMemcpyHostToDevice, ctx->stream);
  rc = jf_sort_dev(ctx, g, dk, dc, n);
  if (rc == 0) {
    cudaMemcpyAsync(keys_host, dk, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(counts_host, dc, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
      rg_set_error("rg_jf_sort: copy back failed");
      rc = RG_ERR_CUDA;
    }
  }
  cudaFree(dk);
  cudaFree(dc);
  return rc ? rc : (int64_t)n;
}

// One .Jhash from the hash-prefix shards of a table (n == 1: a plain table). Every shard is dumped on its own
// GPU; the pieces are gathered on the first shard's GPU (peer copies), which puts them in file order and
// packs the records.
int64_t rg_tables_write_jhash(rg_table* const* tabs, int n_tabs, const char* path, uint64_t min_count, uint64_t max_count,
                              uint64_t size, uint32_t counter_len, const char* const* cmdline, int n_cmdline) {
  if (!tabs || n_tabs < 1 || !tabs[0] || !path) return RG_ERR_ARG;
  if (counter_len == 0) counter_len = 4;
  if (counter_len > 8) {
    rg_set_error("jhash: counter_len must be 1..8");
    return RG_ERR_ARG;
  }
  rg_table* t = tabs[0];
  rg_ctx* ctx = t->ctx;
  JfGeometry g;
  int rc = jf_geometry(size, t->k, g);
  if (rc) return rc;
  // 1. the entries in range, on the first shard's device
  std::vector<uint64_t> n_shard(n_tabs, 0);
  uint64_t cap = 0;
  for (int i = 0; i < n_tabs; ++i) {
    if (!tabs[i] || tabs[i]->k != t->k || tabs[i]->canonical != t->canonical) {
      rg_set_error("jhash: shards must share k and the canonical flag");
      return RG_ERR_ARG;
    }
    RG_CUDA_OK(cudaSetDevice(tabs[i]->ctx->device));
    rc = rg_table_finish(tabs[i]);
    if (rc == 0) rc = rg_table_stats(tabs[i], &n_shard[i], nullptr, nullptr);
    if (rc) return rc;
    cap += n_shard[i];
  }
  if (cap == 0) cap = 1;
  RG_CUDA_OK(cudaSetDevice(ctx->device));
  uint64_t* dk = nullptr;
  uint32_t* dc = nullptr;
  unsigned char* drec = nullptr;
  RG_CUDA_OK(cudaMalloc(&dk, cap * 8));
  RG_CUDA_OK(cudaMalloc(&dc, cap * 4));
  int64_t result = 0;
  uint64_t n = 0;
  for (int i = 0; i < n_tabs && rc == 0; ++i) {
    rg_ctx* ci = tabs[i]->ctx;
    cudaSetDevice(ci->device);
    uint64_t *sk = dk + n, *tk = nullptr;
    uint32_t *sc = dc + n, *tc = nullptr;
    if (i > 0) {       // dump on the shard's own device, then one peer copy
      const uint64_t c = n_shard[i] ? n_shard[i] : 1;
      if (cudaMalloc(&tk, c * 8) != cudaSuccess || cudaMalloc(&tc, c * 4) != cudaSuccess) {
        rg_set_error("jhash: out of device memory");
        rc = RG_ERR_CUDA;
        cudaFree(tk);
        break;
      }
      sk = tk;
      sc = tc;
    }
    rc = rg_launch_dump(tabs[i], min_count ? min_count : 1, max_count ? max_count : ~0ull, sk, sc, n_shard[i], ci->d_scratch);
    uint64_t got = 0;
    if (rc == 0) {
      cudaError_t e = cudaMemcpyAsync(ci->h_scratch, ci->d_scratch, 8, cudaMemcpyDeviceToHost, ci->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ci->stream);
      if (e != cudaSuccess) { rg_set_error(cudaGetErrorString(e)); rc = RG_ERR_CUDA; }
      got = ci->h_scratch[0];
    }
    if (rc == 0 && i > 0 && got) {
      cudaError_t e = cudaMemcpyPeer(dk + n, ctx->device, tk, ci->device, got * 8);
      if (e == cudaSuccess) e = cudaMemcpyPeer(dc + n, ctx->device, tc, ci->device, got * 4);
      if (e != cudaSuccess) { rg_set_error(std::string("jhash: peer copy: ") + cudaGetErrorString(e)); rc = RG_ERR_CUDA; }
    }
    cudaFree(tk);
    cudaFree(tc);
    n += got;
  }
  cudaSetDevice(ctx->device);
  // 2. file order, 3. file bytes
  if (rc == 0) rc = jf_sort_dev(ctx, g, dk, dc, n);
  const unsigned key_bytes = (g.c + 7) / 8, rec = key_bytes + counter_len;
  std::vector<unsigned char> host;
  if (rc == 0 && n) {
    if (cudaMalloc(&drec, n * rec) != cudaSuccess) { rg_set_error("jhash: out of device memory"); rc = RG_ERR_CUDA; }
    if (rc == 0) {
      {
        rg_launch_scope _ls(ctx, "jf_pack_kernel");
        jf_pack_kernel<<<jf_grid(ctx, n), 256, 0, ctx->stream>>>(dk, dc, n, key_bytes, counter_len, drec);
      }
      host.resize(n * rec);
      cudaError_t e = cudaMemcpyAsync(host.data(), drec, n * rec, cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      if (e != cudaSuccess) { rg_set_error(cudaGetErrorString(e)); rc = RG_ERR_CUDA; }
    }
  }
  if (rc == 0) {
    FILE* f = fopen(path, "wb");
    if (!f) {
      rg_set_error(std::string("jhash: cannot open ") + path + " for writing");
      rc = RG_ERR_IO;
    } else {
      const std::string h = jf_header_text(g, t->canonical, counter_len, 7, cmdline, n_cmdline);
      uint64_t hlen = h.size();
      const uint64_t misalign = (9 + h.size()) % 8;
      const uint64_t pad = misalign ? 8 - misalign : 0;
      hlen += pad;
      char digits[16];
      snprintf(digits, sizeof(digits), "%09llu", (unsigned long long)hlen);
      const char zeros[8] = {0};
      bool ok = fwrite(digits, 1, 9, f) == 9 && fwrite(h.data(), 1, h.size(), f) == h.size() &&
                fwrite(zeros, 1, pad, f) == pad && (host.empty() || fwrite(host.data(), 1, host.size(), f) == host.size());
      ok = (fclose(f) == 0) && ok;
      if (!ok) {
        rg_set_error(std::string("jhash: write to ") + path + " failed");
        rc = RG_ERR_IO;
      }
      result = (int64_t)n;
    }
  }
  cudaFree(dk);
  cudaFree(dc);
  cudaFree(drec);
  return rc ? rc : result;
}

int64_t rg_table_write_jhash(rg_table* t, const char* path, uint64_t min_count, uint64_t max_count, uint64_t size,
                             uint32_t counter_len, const char* const* cmdline, int n_cmdline) {
  return rg_tables_write_jhash(&t, 1, path, min_count, max_count, size, counter_len, cmdline, n_cmdline);
}

rg_table* rg_table_read_jhash(rg_ctx* ctx, const char* path, rg_jf_info* info) {
  if (!ctx || !path) {
    rg_set_error("rg_table_read_jhash: null argument");
    return nullptr;
  }
  FILE* f = fopen(path, "rb");
  if (!f) {
    rg_set_error(std::string("jhash: cannot open ") + path);
    return nullptr;
  }
  JfFileHeader h;
  if (jf_read_header(f, path, h) != 0) {
    fclose(f);
    return nullptr;
  }
  if (h.k < 1 || h.k > 32 || (h.g.c & 1) || h.counter_len < 1 || h.counter_len > 8) {   // (a damaged header must not divide by zero)
    fclose(f);
    rg_set_error(std::string("jhash: unsupported key_len / counter_len in ") + path);
    return nullptr;
  }
  // the records: one read into a buffer of the file's size (a regular file; 64 KiB pieces appended to a growing
  // vector cost 0.20 s per 123 MB .Jhash on the host, this 0.07 s), pieces only for a pipe
  JfBytes bytes;
  {
    struct stat st;
    const long at = ftell(f);
    if (at >= 0 && fstat(fileno(f), &st) == 0 && S_ISREG(st.st_mode) && (uint64_t)st.st_size >= (uint64_t)at) {
      const size_t want = (size_t)((uint64_t)st.st_size - (uint64_t)at);
      if (!bytes.reserve(want)) {
        fclose(f);
        rg_set_error("jhash: out of host memory");
        return nullptr;
      }
      bytes.n = fread(bytes.p, 1, want, f);
    }
    unsigned char buf[1 << 16];     // (a pipe, or a file that grew since fstat)
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), f)) > 0)
      if (!bytes.append(buf, got)) {
        fclose(f);
        rg_set_error("jhash: out of host memory");
        return nullptr;
      }
  }
  fclose(f);
  const unsigned rec = h.key_bytes + h.counter_len;
  const uint64_t n = bytes.size() / rec;
  if (info) {
    memset(info, 0, sizeof(*info));
    info->k = h.k;
    info->canonical = h.canonical;
    info->size = h.g.size;
    info->counter_len = h.counter_len;
    info->n_records = n;
    info->matrix_r = h.g.r;
    info->max_reprobe_offset = jf_reprobe(h.g.max_reprobe);
    memcpy(info->columns, h.g.columns, sizeof(info->columns));
  }
  rg_table* t = rg_count_begin(ctx, h.k, h.canonical, n ? n : 1, RG_MODE_DIRECT);
  if (!t || n == 0) return t;
  unsigned char* dbytes = nullptr;
  uint64_t* dk = nullptr;
  uint32_t* dc = nullptr;
  bool ok = cudaMalloc(&dbytes, n * rec) == cudaSuccess && cudaMalloc(&dk, n * 8) == cudaSuccess &&
            cudaMalloc(&dc, n * 4) == cudaSuccess;
  if (ok) {
    cudaMemcpyAsync(dbytes, bytes.data(), n * rec, cudaMemcpyHostToDevice, ctx->stream);
    {
      rg_launch_scope _ls(ctx, "jf_unpack_kernel");
      jf_unpack_kernel<<<jf_grid(ctx, n), 256, 0, ctx->stream>>>(dbytes, n, h.key_bytes, h.counter_len, dk, dc);
    }
    ok = rg_launch_insert_keys(t, dk, n, dc) == 0;
    t->tiles_dirty = true;
    if (ok) ok = rg_count_end(t) == 0;
    cudaStreamSynchronize(ctx->stream);
  } else {
    rg_set_error("jhash: out of device memory");
  }
  cudaFree(dbytes);
  cudaFree(dk);
  cudaFree(dc);
  if (!ok) {
    rg_table_free(t);
    return nullptr;
  }
  return t;
}

}  // extern "C"
