// rg_jf_parts.hpp — the merge of the intermediate files of a `jellyfish count --disk` that did not fit the device
// (hash_counter.hpp:182-202 dumps the table to `<output><N>` and keeps counting; count_main.cc:316-338 then runs the
// STOCK merge — tar:jellyfish-2.2.5/jellyfish/merge_files.cc:78-79 — over the parts: counts of a k-mer are summed over
// the parts, THEN -L / -U apply, and the parts are unlinked). Every part is a complete binary/sorted .Jhash written by
// rg_table_write_jhash with the same -s and -m, hence the same hash matrix and the same (pos, key) record order, so the
// merge is a streaming file merge: read the heads, emit the smallest (pos, key) with the summed count. Host side and
// I/O-bound by nature (it touches every record of every part once, sequentially, and nothing else); the counting that
// produced the parts ran on the GPU.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

namespace rgparts {

[[noreturn]] void fail(const std::string& message);   // defined by the includer

// value of `"name":` inside the terse JSON header (a number, or the text between the brackets of an array)
static inline std::string member(const std::string& js, const char* name, size_t from = 0) {
  const std::string pat = std::string("\"") + name + "\":";
  const size_t at = js.find(pat, from);
  if (at == std::string::npos) return "";
  size_t p = at + pat.size();
  if (p < js.size() && js[p] == '[') {
    const size_t e = js.find(']', p);
    return e == std::string::npos ? "" : js.substr(p + 1, e - p - 1);
  }
  size_t e = p;
  while (e < js.size() && js[e] != ',' && js[e] != '}') ++e;
  return js.substr(p, e - p);
}

struct Part {
  std::string path, header;          // header = the 9 digits + the padded JSON, as they are in the file
  FILE* f = nullptr;
  unsigned key_len = 0, key_bytes = 0, val_bytes = 0;
  uint64_t size = 0;
  std::vector<uint64_t> columns;
  std::vector<unsigned char> buf;
  size_t at = 0, len = 0;
  bool live = false;
  uint64_t key = 0, pos = 0, count = 0;

  void open(const std::string& p) {
    path = p;
    f = fopen(p.c_str(), "rb");
    if (!f) fail("cannot open intermediate file '" + p + "'");
    char digits[10] = {0};
    if (fread(digits, 1, 9, f) != 9) fail("'" + p + "': no header");
    for (int i = 0; i < 9; ++i)
      if (digits[i] < '0' || digits[i] > '9') fail("'" + p + "' is not a Jellyfish binary file");
    const size_t hlen = (size_t)strtoull(digits, nullptr, 10);
    std::string js(hlen, '\0');
    if (fread(&js[0], 1, hlen, f) != hlen) fail("'" + p + "': truncated header");
    header = std::string(digits, 9) + js;
    if (member(js, "format") != "\"binary/sorted\"") fail("'" + p + "': not a binary/sorted file");
    key_len = (unsigned)strtoul(member(js, "key_len").c_str(), nullptr, 10);
    val_bytes = (unsigned)strtoul(member(js, "counter_len").c_str(), nullptr, 10);
    size = strtoull(member(js, "size").c_str(), nullptr, 10);
    key_bytes = (key_len + 7) / 8;
    if (key_len < 2 || key_len > 64 || (key_len & 1) || val_bytes < 1 || val_bytes > 8 || size == 0 || (size & (size - 1)))
      fail("'" + p + "': unsupported header (key_len / counter_len / size)");
    const size_t m = js.find("\"matrix1\":");
    const std::string cols = m == std::string::npos ? "" : member(js, "columns", m);
    for (const char* c = cols.c_str(); *c;) {
      char* e = nullptr;
      const unsigned long long v = strtoull(c, &e, 10);
      if (e == c) break;
      columns.push_back((uint64_t)v);
      c = *e == ',' ? e + 1 : e;
    }
    if (columns.size() != key_len) fail("'" + p + "': hash matrix with " + std::to_string(columns.size()) + " columns for key_len " + std::to_string(key_len));
    buf.resize((size_t)(key_bytes + val_bytes) << 16);
  }
  void close() {
    if (f) fclose(f);
    f = nullptr;
  }
};

// pos = (XOR over the set bits j of key of columns[c-1-j]) & (size-1) (rectangular_binary_matrix.hpp:206-242), through
// one 256-entry table per key byte
struct PosTable {
  std::vector<uint64_t> t;   // [key_bytes][256]
  unsigned key_bytes = 0;
  uint64_t mask = 0;
  void build(const Part& p) {
    key_bytes = p.key_bytes;
    mask = p.size - 1;
    t.assign((size_t)key_bytes * 256, 0);
    for (unsigned b = 0; b < key_bytes; ++b)
      for (unsigned v = 0; v < 256; ++v) {
        uint64_t x = 0;
        for (unsigned i = 0; i < 8; ++i) {
          const unsigned j = 8 * b + i;
          if (((v >> i) & 1u) && j < p.key_len) x ^= p.columns[p.key_len - 1 - j];
        }
        t[(size_t)b * 256 + v] = x;
      }
  }
  uint64_t pos(const unsigned char* key_le) const {
    uint64_t x = 0;
    for (unsigned b = 0; b < key_bytes; ++b) x ^= t[(size_t)b * 256 + key_le[b]];
    return x & mask;
  }
};

static inline bool advance(Part& p, const PosTable& pt) {
  const size_t rec = p.key_bytes + p.val_bytes;
  if (p.at + rec > p.len) {
    const size_t keep = p.len - p.at;
    if (keep) memmove(p.buf.data(), p.buf.data() + p.at, keep);
    p.at = 0;
    p.len = keep + fread(p.buf.data() + keep, 1, p.buf.size() - keep, p.f);
    if (p.len < rec) {
      if (p.len != 0) fail("'" + p.path + "': truncated record at the end of the file");
      p.live = false;
      return false;
    }
  }
  const unsigned char* r = p.buf.data() + p.at;
  uint64_t k = 0, c = 0;
  for (unsigned b = 0; b < p.key_bytes; ++b) k |= (uint64_t)r[b] << (8 * b);
  for (unsigned b = 0; b < p.val_bytes; ++b) c |= (uint64_t)r[p.key_bytes + b] << (8 * b);
  const uint64_t ps = pt.pos(r);
  if (p.live && (ps < p.pos || (ps == p.pos && k <= p.key))) fail("'" + p.path + "': records are not in (pos, key) order");
  p.key = k;
  p.count = c;
  p.pos = ps;
  p.at += rec;
  p.live = true;
  return true;
}

// parts -> out: header of the first part, then every k-mer once with the sum of its counts (saturating at 2^32-1 on
// disk, binary_dumper.hpp:44-48) when lower <= sum and (upper == 0 or sum <= upper). Returns the records written.
// `histo` (optional, R rows): the records written, by count — row min(count, R-1), the histogram `jellyfish histo -l 1
// -h R-2` computes from the file (histo_main.cc:64-84).
static inline uint64_t merge(const std::vector<std::string>& paths, const std::string& out, uint64_t lower, uint64_t upper,
                             std::vector<uint64_t>* histo = nullptr) {
  if (paths.empty()) fail("no intermediate files to merge");
  std::vector<Part> parts(paths.size());
  for (size_t i = 0; i < paths.size(); ++i) {
    parts[i].open(paths[i]);
    if (i && (parts[i].key_len != parts[0].key_len || parts[i].size != parts[0].size || parts[i].columns != parts[0].columns ||
              parts[i].val_bytes != parts[0].val_bytes))
      fail("'" + paths[i] + "' was not written with the size / mer length / hash matrix of '" + paths[0] + "'");   // merge_files.cc:193-202
  }
  if (parts[0].val_bytes != 4) fail("intermediate files must carry 4-byte counts");
  PosTable pt;
  pt.build(parts[0]);
  FILE* o = fopen(out.c_str(), "wb");
  if (!o) fail("Error opening output file '" + out + "'");
  std::vector<char> obuf(1 << 22);
  setvbuf(o, obuf.data(), _IOFBF, obuf.size());
  bool ok = fwrite(parts[0].header.data(), 1, parts[0].header.size(), o) == parts[0].header.size();
  for (Part& p : parts) advance(p, pt);
  const unsigned kb = parts[0].key_bytes;
  uint64_t written = 0;
  for (;;) {
    Part* best = nullptr;
    for (Part& p : parts)
      if (p.live && (!best || p.pos < best->pos || (p.pos == best->pos && p.key < best->key))) best = &p;
    if (!best) break;
    const uint64_t key = best->key;
    uint64_t sum = 0;
    for (Part& p : parts)
      if (p.live && p.key == key) {
        sum = sum + p.count < sum ? ~0ull : sum + p.count;
        advance(p, pt);
      }
    if (sum >= lower && (upper == 0 || sum <= upper)) {
      unsigned char r[12];
      for (unsigned b = 0; b < kb; ++b) r[b] = (unsigned char)(key >> (8 * b));
      const uint32_t c = sum > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)sum;
      for (unsigned b = 0; b < 4; ++b) r[kb + b] = (unsigned char)(c >> (8 * b));
      ok = ok && fwrite(r, 1, kb + 4, o) == kb + 4;
      ++written;
      if (histo && !histo->empty()) ++(*histo)[std::min<uint64_t>(c, histo->size() - 1)];
    }
  }
  for (Part& p : parts) p.close();
  ok = fclose(o) == 0 && ok;
  if (!ok) fail("write error on '" + out + "'");
  return written;
}

}  // namespace rgparts
