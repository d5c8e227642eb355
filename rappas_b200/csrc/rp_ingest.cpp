// rp_ingest.cpp -- host side either end of the hot path ("next" rows 2 and 3): query FASTA
// ingest with the duplicate structure the placement loop relies on, and the .jplace writer.  Plain C++,
// no CUDA; part of librappas_b200.so so that the Java caller (or the Python mirror) can hand whole files over.
//
//   R1  FASTAPointer.nextSequenceAsFastaObject   inputs/FASTAPointer.java:67-149, Fasta inputs/Fasta.java:21-39
//   R2  duplicate merge                          core/algos/PlacementProcess.java:591-629 (+ registration :1047)
//   R11 jplace rows + file shell                 core/algos/PlacementProcess.java:974-1047,
//                                                main_v2/Main_PLACEMENT_v07.java:224-234, 270-315
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rappas_b200.h"

namespace rp {
int set_error(int code, const char* fmt, ...);
}
using rp::set_error;

// byte buffer that is NOT zero-filled when sized (the parallel gathers overwrite every byte; a vector's
// resize would first memset hundreds of MB)
struct RawBytes {
  uint8_t* p = nullptr;
  size_t n = 0;
  RawBytes() = default;
  RawBytes(const RawBytes&) = delete;
  RawBytes& operator=(const RawBytes&) = delete;
  ~RawBytes() { free(p); }
  bool resize(size_t bytes) { free(p); p = (uint8_t*)malloc(bytes ? bytes : 1); n = p ? bytes : 0; return p != nullptr; }
  uint8_t* data() { return p; }
  const uint8_t* data() const { return p; }
  size_t size() const { return n; }
  uint8_t operator[](size_t i) const { return p[i]; }
};

struct rp_reads {
  // every FASTA record, in file order
  RawBytes hdr;                    // headers (first line without '>'), concatenated
  std::vector<uint64_t> hdr_off;   // [n_records + 1]
  std::vector<uint32_t> unique_of; // record -> index of its exact sequence among the unique ones
  std::vector<uint32_t> group_of;  // record -> id of its gap-stripped sequence (the reference's MD5 key)
  // distinct exact sequences (gaps kept: Main_PLACEMENT_v07.java:195), in order of first appearance
  RawBytes seq;
  std::vector<uint64_t> seq_off;   // [n_unique + 1]
  uint32_t n_groups = 0;
};

namespace {

unsigned host_threads(size_t work_items, size_t per_thread_min) {
  unsigned nt = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (const char* e = getenv("RP_HOST_THREADS")) return (unsigned)std::max(1, std::min(64, atoi(e)));  // tests: forced
  while (nt > 1 && work_items / nt < per_thread_min) nt--;
  return nt;
}

// 64-bit hash of a byte string, 8 bytes per step (it only buckets: equality is always checked on the bytes)
inline uint64_t mix64(uint64_t x) {
  x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32; x *= 0xd6e8feb86659fd93ull; x ^= x >> 32;
  return x;
}
inline uint64_t hash_bytes(const uint8_t* p, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xff51afd7ed558ccdull);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x9FB21C651E98DF25ull;
    h ^= h >> 29;
  }
  if (i < n) {
    uint64_t w = 0;
    memcpy(&w, p + i, n - i);
    h = (h ^ w) * 0x9FB21C651E98DF25ull;
    h ^= h >> 29;
  }
  return mix64(h);
}

// java.io.BufferedReader.readLine: a line ends at \n, \r or \r\n.  The scan is two memchr calls per line.
inline size_t next_line(const uint8_t* t, size_t n, size_t pos, size_t* line_end) {
  const uint8_t* nl = (const uint8_t*)memchr(t + pos, '\n', n - pos);
  size_t e = nl ? (size_t)(nl - t) : n;
  const uint8_t* cr = (const uint8_t*)memchr(t + pos, '\r', e - pos);
  if (cr) e = (size_t)(cr - t);
  *line_end = e;
  if (e < n && t[e] == '\r' && e + 1 < n && t[e + 1] == '\n') return e + 2;
  return e < n ? e + 1 : n;
}

// the records of one piece of the text (a piece starts at a header line, except possibly the first)
struct Piece {
  std::vector<uint8_t> hdr, seq;   // seq: only the sequences that had to be joined from several lines
  std::vector<uint32_t> hdr_len, seq_len;
  std::vector<uint64_t> seq_at;    // where the (joined + trimmed) sequence is: offset into the text, or kInJoined | offset into seq
  std::vector<uint64_t> hash;      // of the joined + trimmed sequence
  bool any_gap = false, data_before_header = false;
};
constexpr uint64_t kInJoined = 1ull << 63;

// A record whose sequence is ONE line -- nearly all reads -- is not copied here: it is hashed where the text has it
// and copied once, when the distinct sequences are gathered.  Only wrapped sequences are joined in a side buffer.
void parse_piece(const uint8_t* t, size_t begin, size_t end, Piece* P) {
  bool in_record = false;
  size_t one_at = 0, one_len = 0;  // the record's only sequence line so far (text offset, length)
  int n_lines = 0;                 // sequence lines of the record; from the second on they are joined in P->seq
  size_t rec_start = 0;            // of the current record in P->seq
  auto close_record = [&]() {
    // String.trim(): strip code points <= U+0020 from both ends of the joined sequence (FASTAPointer.java:143-145)
    const uint8_t* base = n_lines >= 2 ? P->seq.data() : t;
    size_t b = n_lines >= 2 ? rec_start : one_at, e = n_lines >= 2 ? P->seq.size() : one_at + one_len;
    while (e > b && base[e - 1] <= 0x20) e--;
    while (b < e && base[b] <= 0x20) b++;
    if (n_lines >= 2) {
      if (b > rec_start) memmove(&P->seq[rec_start], &P->seq[b], e - b);
      P->seq.resize(rec_start + (e - b));
      P->seq_at.push_back(kInJoined | rec_start);
      base = P->seq.data();
      e = rec_start + (e - b);
      b = rec_start;
    } else {
      P->seq_at.push_back(b);
    }
    P->seq_len.push_back((uint32_t)(e - b));
    P->hash.push_back(hash_bytes(base + b, e - b));
    if (!P->any_gap && e > b && memchr(base + b, '-', e - b)) P->any_gap = true;
  };
  P->hdr.reserve((end - begin) / 4);
  size_t pos = begin;
  while (pos < end) {
    size_t le;
    const size_t nxt = next_line(t, end, pos, &le);
    const uint8_t* line = t + pos;
    const size_t len = le - pos;
    const size_t at = pos;
    pos = nxt;
    if (len == 0 || line[0] == '#') continue;  // empty and '#' lines are skipped (FASTAPointer.java:82-87)
    if (line[0] == '>') {
      if (in_record) close_record();
      P->hdr.insert(P->hdr.end(), line + 1, line + len);
      P->hdr_len.push_back((uint32_t)(len - 1));
      in_record = true;
      n_lines = 0;
      one_at = one_len = 0;
      continue;
    }
    if (!in_record) { P->data_before_header = true; return; }
    if (n_lines == 0) {
      one_at = at; one_len = len;
    } else {
      if (n_lines == 1) {  // a second line: the sequence is wrapped, join it
        rec_start = P->seq.size();
        P->seq.insert(P->seq.end(), t + one_at, t + one_at + one_len);
      }
      P->seq.insert(P->seq.end(), line, line + len);
    }
    n_lines++;
  }
  if (in_record) close_record();
}

// open-addressing set of byte strings kept elsewhere: slot = {hash, id + 1}
struct FlatSet {
  std::vector<uint64_t> h;
  std::vector<uint32_t> id;
  uint64_t mask;
  explicit FlatSet(size_t n) {
    size_t cap = 16;
    while (cap < 2 * n + 2) cap <<= 1;
    h.assign(cap, 0);
    id.assign(cap, 0);
    mask = cap - 1;
  }
  // eq(id) tells whether the stored string `id` equals the probed one; returns the id found, or inserts new_id
  template <typename Eq>
  uint32_t find_or_insert(uint64_t hash, uint32_t new_id, Eq eq, bool* inserted) {
    for (uint64_t i = hash & mask;; i = (i + 1) & mask) {
      if (id[i] == 0) { h[i] = hash; id[i] = new_id + 1; *inserted = true; return new_id; }
      if (h[i] == hash && eq(id[i] - 1)) { *inserted = false; return id[i] - 1; }
    }
  }
  void prefetch(uint64_t hash) const { __builtin_prefetch(&h[hash & mask]); __builtin_prefetch(&id[hash & mask]); }
};

int parse(const uint8_t* t, size_t n, rp_reads* R) {
  const bool dbg = getenv("RP_DEBUG_INGEST") != nullptr;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double t0 = now();
  // pass 1, parallel: the text is cut at header lines into one piece per thread; every piece yields its
  // records (header, joined + trimmed sequence, hash of the sequence)
  const unsigned nt = host_threads(n, 1 << 20);
  std::vector<size_t> cut(nt + 1, n);
  cut[0] = 0;
  for (unsigned i = 1; i < nt; i++) {
    // the first line start at or after i*n/nt that begins with '>'
    size_t pos = std::max(cut[i - 1], n / nt * i);
    size_t found = n;
    while (pos < n) {
      const uint8_t* g = (const uint8_t*)memchr(t + pos, '>', n - pos);
      if (!g) break;
      const size_t q = (size_t)(g - t);
      if (q == 0 || t[q - 1] == '\n' || t[q - 1] == '\r') { found = q; break; }
      pos = q + 1;
    }
    cut[i] = found;
  }
  std::vector<Piece> pieces(nt);
  {
    std::vector<std::thread> th;
    for (unsigned i = 1; i < nt; i++)
      if (cut[i] < cut[i + 1]) th.emplace_back(parse_piece, t, cut[i], cut[i + 1], &pieces[i]);
    parse_piece(t, cut[0], cut[1], &pieces[0]);
    for (auto& x : th) x.join();
  }
  const double t1 = now();
  if (pieces[0].data_before_header) return set_error(RP_E_IO, "FASTA: sequence data before the first '>' header");
  size_t nrec = 0, hdr_bytes = 0, seq_bytes = 0;
  bool any_gap = false;
  std::vector<size_t> rec0(nt + 1, 0), hdr0(nt + 1, 0), seq0(nt + 1, 0);
  for (unsigned i = 0; i < nt; i++) {
    nrec += pieces[i].hdr_len.size(); hdr_bytes += pieces[i].hdr.size(); seq_bytes += pieces[i].seq.size();
    rec0[i + 1] = nrec; hdr0[i + 1] = hdr_bytes; seq0[i + 1] = seq_bytes;
    any_gap |= pieces[i].any_gap;
  }
  if (nrec == 0) return set_error(RP_E_IO, "No valid fasta sequences were found");  // FASTAPointer.checkSize :238-241
  if (nrec >= 0xFFFFFFFFull) return set_error(RP_E_UNSUPPORTED, "more than 2^32-2 records");
  // gather the pieces (parallel): headers into R; the sequences stay where the pieces parsed them
  std::vector<const uint8_t*> sptr(nrec);
  std::vector<uint32_t> slen(nrec);
  std::vector<uint64_t> hashes(nrec);
  if (!R->hdr.resize(hdr_bytes)) return set_error(RP_E_NOMEM, "out of memory for %zu B of headers", hdr_bytes);
  R->hdr_off.resize(nrec + 1);
  auto run_parallel = [&](unsigned n_jobs, auto&& job) {
    std::vector<std::thread> th;
    for (unsigned i = 1; i < n_jobs; i++) th.emplace_back(job, i);
    job(0u);
    for (auto& x : th) x.join();
  };
  run_parallel(nt, [&](unsigned i) {
    const Piece& P = pieces[i];
    if (!P.hdr.empty()) memcpy(R->hdr.data() + hdr0[i], P.hdr.data(), P.hdr.size());
    size_t ho = hdr0[i];
    for (size_t j = 0; j < P.hdr_len.size(); j++) {
      const size_t r = rec0[i] + j;
      const uint64_t at = P.seq_at[j];
      R->hdr_off[r] = ho; slen[r] = P.seq_len[j]; hashes[r] = P.hash[j];
      sptr[r] = (at & kInJoined) ? P.seq.data() + (at & ~kInJoined) : t + at;
      ho += P.hdr_len[j];
    }
  });
  R->hdr_off[nrec] = hdr_bytes;
  const double t2 = now();
  // pass 2: exact-sequence uniques (what is placed), numbered in order of first appearance.  The hash space
  // is split over the threads: thread p finds, among the records whose hash falls to it, the first record of
  // every distinct sequence; ids then come from a prefix count over the records that are such a first.
  std::vector<uint32_t> first_of(nrec);
  const unsigned np = nt;
  // which thread a record's hash falls to (a byte per record: every thread reads all of them), and how many fall to each
  std::vector<uint8_t> bucket(nrec);
  std::vector<size_t> mine_part((size_t)nt * np, 0);
  run_parallel(nt, [&](unsigned i) {
    size_t* cnt = &mine_part[(size_t)i * np];
    for (size_t r = nrec * i / nt; r < nrec * (i + 1) / nt; r++) {
      const unsigned p = (unsigned)(((hashes[r] >> 40) * np) >> 24);  // 24 hash bits scaled to [0, np)
      bucket[r] = (uint8_t)p;
      cnt[p]++;
    }
  });
  run_parallel(np, [&](unsigned p) {
    size_t mine = 0;
    for (unsigned i = 0; i < nt; i++) mine += mine_part[(size_t)i * np + p];
    FlatSet set(mine);
    for (size_t r = 0; r < nrec; r++) {
      if (bucket[r] != p) continue;
      const uint8_t* sp = sptr[r];
      const uint32_t sn = slen[r];
      bool ins;
      first_of[r] = set.find_or_insert(hashes[r], (uint32_t)r, [&](uint32_t o) {
        return slen[o] == sn && (sn == 0 || memcmp(sptr[o], sp, sn) == 0);
      }, &ins);
    }
  });
  // ids in order of first appearance: count the firsts (and their bytes) per range of records, scan, then number
  R->unique_of.resize(nrec);
  R->group_of.resize(nrec);
  std::vector<size_t> n_first(nt + 1, 0), b_first(nt + 1, 0);
  run_parallel(nt, [&](unsigned i) {
    size_t c = 0, bytes = 0;
    for (size_t r = nrec * i / nt; r < nrec * (i + 1) / nt; r++)
      if (first_of[r] == r) { c++; bytes += slen[r]; }
    n_first[i + 1] = c; b_first[i + 1] = bytes;
  });
  for (unsigned i = 0; i < nt; i++) { n_first[i + 1] += n_first[i]; b_first[i + 1] += b_first[i]; }
  const size_t n_unique = n_first[nt];
  std::vector<uint32_t> first_rec(n_unique);  // unique id -> its first record
  R->seq_off.resize(n_unique + 1);
  R->seq_off[n_unique] = b_first[nt];
  run_parallel(nt, [&](unsigned i) {
    size_t u = n_first[i], at = b_first[i];
    for (size_t r = nrec * i / nt; r < nrec * (i + 1) / nt; r++)
      if (first_of[r] == r) {
        R->unique_of[r] = (uint32_t)u;
        first_rec[u] = (uint32_t)r;
        R->seq_off[u] = at;
        u++; at += slen[r];
      }
  });
  if (!R->seq.resize(R->seq_off.back())) return set_error(RP_E_NOMEM, "out of memory for %zu B of sequences", (size_t)R->seq_off.back());
  run_parallel(nt, [&](unsigned i) {
    for (size_t r = nrec * i / nt; r < nrec * (i + 1) / nt; r++)
      if (first_of[r] != r) R->unique_of[r] = R->unique_of[first_of[r]];
    for (size_t u = n_unique * i / nt; u < n_unique * (i + 1) / nt; u++)
      if (slen[first_rec[u]]) memcpy(R->seq.data() + R->seq_off[u], sptr[first_rec[u]], slen[first_rec[u]]);
  });
  pieces.clear();
  if (dbg) fprintf(stderr, "ingest: parse %.3f s (%u threads), gather %.3f s, unique %.3f s\n", t1 - t0, nt, t2 - t1, now() - t2);
  // gap-stripped groups (the reference's checksum key: fasta.getSequence(true) removes '-', Fasta.java:35-39,
  // PlacementProcess.java:593).  Without any '-' in the file a group is a unique sequence.
  if (!any_gap) {
    R->group_of = R->unique_of;
    R->n_groups = (uint32_t)n_unique;
    return RP_OK;
  }
  std::vector<uint32_t> group_of_unique(n_unique);
  {
    // stripped keys, per unique sequence, in order of first appearance (= the order groups are numbered in)
    std::vector<uint8_t> stripped;
    std::vector<uint64_t> st_off(1, 0);
    stripped.reserve(R->seq.size());
    FlatSet groups(n_unique);
    for (size_t u = 0; u < n_unique; u++) {
      const size_t b = stripped.size();
      for (uint64_t i = R->seq_off[u]; i < R->seq_off[u + 1]; i++)
        if (R->seq[i] != '-') stripped.push_back(R->seq[i]);
      const size_t gn = stripped.size() - b;
      bool ins;
      const uint32_t g = groups.find_or_insert(hash_bytes(stripped.data() + b, gn), R->n_groups, [&](uint32_t o) {
        return st_off[o + 1] - st_off[o] == gn && (gn == 0 || memcmp(stripped.data() + st_off[o], stripped.data() + b, gn) == 0);
      }, &ins);
      if (ins) { st_off.push_back(stripped.size()); R->n_groups++; }
      else stripped.resize(b);
      group_of_unique[u] = g;
    }
  }
  for (size_t r = 0; r < nrec; r++) R->group_of[r] = group_of_unique[R->unique_of[r]];
  return RP_OK;
}

// ---- Java number formatting (json-simple prints Float / Double with toString()) ---------------------
// Double.toString / Float.toString layout: shortest digits that round-trip; plain decimal for
// 1e-3 <= |x| < 1e7 (at least one digit after the point), otherwise d.dddE[-]n.
template <typename T>
void java_number(std::string& out, T v) {
  if (isnan(v) || isinf(v)) { out += "null"; return; }  // JSONValue.toJSONString: NaN / Infinity -> null
  if (v == 0) { out += signbit(v) ? "-0.0" : "0.0"; return; }
  char buf[48], o[64];
  auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);  // shortest round-trip digits
  // buf = [-]d[.ddd]e[+-]xx  ->  digits, decimal exponent
  const char* p = buf;
  char* w = o;
  if (*p == '-') { *w++ = '-'; p++; }
  char digits[24];
  int nd = 0;
  for (; *p != 'e'; p++)
    if (*p != '.') digits[nd++] = *p;
  p++;
  const bool eneg = *p == '-';
  p++;
  int exp10 = 0;
  for (; p < res.ptr; p++) exp10 = exp10 * 10 + (*p - '0');
  if (eneg) exp10 = -exp10;
  if (exp10 >= -3 && exp10 < 7) {
    if (exp10 >= 0) {
      for (int i = 0; i <= exp10; i++) *w++ = i < nd ? digits[i] : '0';
      *w++ = '.';
      if (nd > exp10 + 1) { memcpy(w, digits + exp10 + 1, nd - exp10 - 1); w += nd - exp10 - 1; }
      else *w++ = '0';
    } else {
      *w++ = '0'; *w++ = '.';
      for (int i = 0; i < -exp10 - 1; i++) *w++ = '0';
      memcpy(w, digits, nd); w += nd;
    }
  } else {
    *w++ = digits[0];
    *w++ = '.';
    if (nd > 1) { memcpy(w, digits + 1, nd - 1); w += nd - 1; }
    else *w++ = '0';
    *w++ = 'E';
    w = std::to_chars(w, o + sizeof o, exp10).ptr;
  }
  out.append(o, (size_t)(w - o));
}
inline void append_int(std::string& out, int v) {
  char b[16];
  out.append(b, (size_t)(std::to_chars(b, b + sizeof b, v).ptr - b));
}

void json_string(std::string& out, const uint8_t* p, size_t n) {  // JSONValue.escape
  out += '"';
  for (size_t i = 0; i < n; i++) {
    size_t j = i;  // run of characters that need no escape
    while (j < n && p[j] >= 0x20 && p[j] != 0x7F && p[j] != '"' && p[j] != '\\' && p[j] != '/') j++;
    if (j > i) { out.append((const char*)p + i, j - i); i = j; if (i == n) break; }
    const uint8_t c = p[i];
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\b': out += "\\b"; break;
      case '\f': out += "\\f"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      case '/': out += "\\/"; break;
      default:
        if (c < 0x20 || c == 0x7F) { char b[8]; snprintf(b, sizeof b, "\\u%04X", c); out += b; }
        else out += (char)c;
    }
  }
  out += '"';
}

}  // namespace

extern "C" {

int rp_reads_from_memory(const uint8_t* text, uint64_t n_bytes, rp_reads** out) {
  if (!out || (!text && n_bytes)) return set_error(RP_E_INVALID, "NULL argument");
  *out = nullptr;
  rp_reads* R = new rp_reads();
  const int rc = parse(text, (size_t)n_bytes, R);
  if (rc) { delete R; return rc; }
  *out = R;
  return RP_OK;
}

int rp_reads_load_fasta(const char* path, rp_reads** out) {
  if (!path || !out) return set_error(RP_E_INVALID, "NULL argument");
  *out = nullptr;
  const int fd = open(path, O_RDONLY);
  if (fd < 0) return set_error(RP_E_IO, "cannot open %s: %s", path, strerror(errno));
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); return set_error(RP_E_IO, "cannot stat %s", path); }
  if (st.st_size == 0) { close(fd); return set_error(RP_E_IO, "No valid fasta sequences were found"); }
  void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return set_error(RP_E_IO, "mmap of %s failed", path);
  const int rc = rp_reads_from_memory((const uint8_t*)m, (uint64_t)st.st_size, out);
  munmap(m, (size_t)st.st_size);
  return rc;
}

void rp_reads_free(rp_reads* r) { delete r; }

int rp_reads_describe(const rp_reads* r, uint64_t* n_records, uint64_t* n_unique, uint64_t* n_groups) {
  if (!r) return set_error(RP_E_INVALID, "reads is NULL");
  if (n_records) *n_records = r->hdr_off.size() - 1;
  if (n_unique) *n_unique = r->seq_off.size() - 1;
  if (n_groups) *n_groups = r->n_groups;
  return RP_OK;
}

int rp_reads_unique(const rp_reads* r, const uint8_t** seq, const uint64_t** seq_off) {
  if (!r) return set_error(RP_E_INVALID, "reads is NULL");
  if (seq) *seq = r->seq.data();
  if (seq_off) *seq_off = r->seq_off.data();
  return RP_OK;
}

int rp_reads_records(const rp_reads* r, const uint8_t** hdr, const uint64_t** hdr_off, const uint32_t** unique_of,
                     const uint32_t** group_of) {
  if (!r) return set_error(RP_E_INVALID, "reads is NULL");
  if (hdr) *hdr = r->hdr.data();
  if (hdr_off) *hdr_off = r->hdr_off.data();
  if (unique_of) *unique_of = r->unique_of.data();
  if (group_of) *group_of = r->group_of.data();
  return RP_OK;
}

int rp_jplace_write(const char* path, const rp_reads* r, int32_t keep_at_most, const int32_t* n_rows,
                    const uint16_t* node, const float* score, const double* lwr, const int32_t* status,
                    const int32_t* edge_id, const float* branch_len, int32_t n_nodes, const char* tree_newick,
                    const char* invocation, int32_t guppy_compat, const char* not_placed_path,
                    uint64_t* n_placements) {
  if (!path || !r || !n_rows || !node || !score || !lwr || !status || !edge_id || !branch_len)
    return set_error(RP_E_INVALID, "NULL argument");
  const size_t nrec = r->hdr_off.size() - 1;
  const int K = keep_at_most;
  FILE* f = fopen(path, "wb");
  if (!f) return set_error(RP_E_IO, "cannot open %s for writing", path);
  FILE* fnp = nullptr;
  if (not_placed_path && !(fnp = fopen(not_placed_path, "wb"))) {
    fclose(f);
    return set_error(RP_E_IO, "cannot open %s for writing", not_placed_path);
  }
  // The reference walks the records in file order (PlacementProcess.java:568).  A record whose gap-stripped
  // sequence is already REGISTERED only adds [name-up-to-first-space, 1] to that placement's "nm" (:596-624);
  // a placement is registered only once it produced rows (:1047 sits inside the nsBound block, after the
  // unplaced `continue` :797-806), so duplicates of an unplaced sequence are placed again on their own.
  const bool dbg = getenv("RP_DEBUG_INGEST") != nullptr;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double t0 = now();
  double t_fmt = 0, t_wr = 0;
  std::vector<int64_t> placement_of_group(r->n_groups, -1);
  constexpr uint32_t kNil = 0xFFFFFFFFu;
  std::vector<uint32_t> first_record, dup_head, dup_tail;  // per placement; its duplicates form a list through dup_next
  std::vector<uint32_t> dup_next(nrec, kNil);
  for (size_t i = 0; i < nrec; i++) {
    const uint32_t g = r->group_of[i], u = r->unique_of[i];
    if (placement_of_group[g] >= 0) {
      const size_t p = (size_t)placement_of_group[g];
      if (dup_head[p] == kNil) dup_head[p] = (uint32_t)i;
      else dup_next[dup_tail[p]] = (uint32_t)i;
      dup_tail[p] = (uint32_t)i;
      continue;
    }
    if (status[u] == RP_STATUS_BAD_CHAR || status[u] == RP_STATUS_TOO_SHORT || status[u] == RP_STATUS_TOO_LONG) {
      // the reference stops here (System.exit / exception); the caller decides -- the writer refuses
      fclose(f);
      if (fnp) fclose(fnp);
      return set_error(RP_E_INVALID, "record %zu cannot be processed by the reference (status %d)", i, status[u]);
    }
    if (status[u] == RP_STATUS_UNPLACED) {
      if (fnp) { fwrite(r->hdr.data() + r->hdr_off[i], 1, r->hdr_off[i + 1] - r->hdr_off[i], fnp); fputc('\n', fnp); }
      continue;
    }
    if (n_rows[u] <= 0) continue;  // below nsBound: nothing written, nothing registered (:974)
    placement_of_group[g] = (int64_t)first_record.size();
    first_record.push_back((uint32_t)i);
    dup_head.push_back(kNil);
    dup_tail.push_back(kNil);
  }
  const size_t n_pl = first_record.size();
  const double t1 = now();
  // what a row says about its node -- edge number and distal length -- is printed once per node, not once per row
  // (a third of the numbers of the document: Float.toString is ~50 ns a piece)
  std::vector<std::string> edge_txt((size_t)std::max(0, n_nodes)), distal_txt((size_t)std::max(0, n_nodes));
  {
    const unsigned ntn = host_threads((size_t)std::max(0, n_nodes), 4096);
    auto job = [&](unsigned t) {
      for (size_t x = (size_t)n_nodes * t / ntn; x < (size_t)n_nodes * (t + 1) / ntn; x++) {
        append_int(edge_txt[x], edge_id[x]);
        java_number(distal_txt[x], branch_len[x] / 2.0f);  // getBranchLengthToAncestor()/2 (:1015, :1022)
      }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < ntn; t++) th.emplace_back(job, t);
    job(0);
    for (auto& x : th) x.join();
  }
  // one placement object, appended to o; false if a row names a node the tree does not have
  auto format_placement = [&](std::string& o, size_t p) -> bool {
    const uint32_t i = first_record[p], u = r->unique_of[i];
    o += p ? ",\n{\n\t\"p\":\n\t[" : "\n{\n\t\"p\":\n\t[";
    for (int k = 0; k < n_rows[u]; k++) {
      const size_t at = (size_t)u * K + k;
      const unsigned x = node[at];
      if ((int)x >= n_nodes) return false;
      if (k) o += ",\n\t";
      o += '[';
      if (guppy_compat) {  // distal_length, edge_num, like_weight_ratio, likelihood, pendant_length (:1005-1016)
        o += distal_txt[x]; o += ','; o += edge_txt[x]; o += ',';
        java_number(o, lwr[at]); o += ','; java_number(o, score[at]); o += ",0.0";
      } else {             // edge_num, likelihood, like_weight_ratio, distal_length, pendant_length (:1017-1024)
        o += edge_txt[x]; o += ','; java_number(o, score[at]); o += ',';
        java_number(o, lwr[at]); o += ','; o += distal_txt[x]; o += ",0.0";
      }
      o += ']';
    }
    o += "],\n\t\"nm\":\n\t[[";
    json_string(o, r->hdr.data() + r->hdr_off[i], r->hdr_off[i + 1] - r->hdr_off[i]);  // full header (:1041)
    o += ",1]";
    for (uint32_t d = dup_head[p]; d != kNil; d = dup_next[d]) {
      const uint8_t* h = r->hdr.data() + r->hdr_off[d];
      size_t n = r->hdr_off[d + 1] - r->hdr_off[d];
      const void* sp = memchr(h, ' ', n);  // header up to the first space (:598-602)
      if (sp) n = (const uint8_t*)sp - h;
      o += ",\n\t[";
      json_string(o, h, n);
      o += ",1]";
    }
    o += "]\n}";
    return true;
  };
  std::string o;
  o.reserve(1 << 20);
  o += "{\n\"tree\":";
  if (tree_newick) json_string(o, (const uint8_t*)tree_newick, strlen(tree_newick));
  else o += "null";
  o += ",\n\"placements\":\n[";
  bool ok = fwrite(o.data(), 1, o.size(), f) == o.size();
  o.clear();
  // the placements are formatted in blocks, one block per thread and round, and written in order
  constexpr size_t kBlock = 8192;
  const unsigned nt = host_threads(n_pl, kBlock);
  std::vector<std::string> buf(nt);
  std::vector<char> good(nt, 1);
  for (size_t base = 0; base < n_pl && ok; base += (size_t)nt * kBlock) {
    auto job = [&](unsigned t) {
      std::string& b = buf[t];
      b.clear();
      const size_t lo = base + (size_t)t * kBlock, hi = std::min(n_pl, lo + kBlock);
      for (size_t p = lo; p < hi && good[t]; p++) good[t] = format_placement(b, p);
    };
    const double ta = now();
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; t++)
      if (base + (size_t)t * kBlock < n_pl) th.emplace_back(job, t);
    job(0);
    for (auto& x : th) x.join();
    const double tb = now();
    t_fmt += tb - ta;
    for (unsigned t = 0; t < nt && ok; t++) {
      if (base + (size_t)t * kBlock >= n_pl) break;
      ok = good[t] && fwrite(buf[t].data(), 1, buf[t].size(), f) == buf[t].size();
    }
    t_wr += now() - tb;
  }
  if (dbg) fprintf(stderr, "jplace: plan %.3f s, format %.3f s (%u threads), write %.3f s\n", t1 - t0, t_fmt, nt, t_wr);
  if (!ok) {
    fclose(f);
    if (fnp) fclose(fnp);
    return set_error(RP_E_INVALID, "a result row names node >= n_nodes, or the disk is full");
  }
  o += "\n],\n\"version\":3,\n\"metadata\":{\"invocation\":";
  json_string(o, (const uint8_t*)(invocation ? invocation : ""), invocation ? strlen(invocation) : 0);
  o += guppy_compat ? "},\n\"fields\":[\"distal_length\",\"edge_num\",\"like_weight_ratio\",\"likelihood\",\"pendant_length\"]\n}\n"
                    : "},\n\"fields\":[\"edge_num\",\"likelihood\",\"like_weight_ratio\",\"distal_length\",\"pendant_length\"]\n}\n";
  ok = fwrite(o.data(), 1, o.size(), f) == o.size();
  ok = (fclose(f) == 0) && ok;
  if (fnp) ok = (fclose(fnp) == 0) && ok;
  if (n_placements) *n_placements = n_pl;
  return ok ? RP_OK : set_error(RP_E_IO, "short write to %s", path);
}

// inputs/PHYMLWrapper.java:206-229 (and the RAxML-ng / PAML twins): the per-site preparation of the posteriors
int rp_pp_prepare(const float* probs, const uint8_t* state_of_column, int32_t n_nodes, int32_t n_sites, int32_t n_states,
                  float site_pp_threshold, int32_t as_log10, float* pp_out, uint8_t* states_out) {
  if (!probs || !state_of_column || !pp_out || !states_out) return set_error(RP_E_INVALID, "NULL argument");
  if (n_nodes < 0 || n_sites < 0 || n_states < 1 || n_states > 32) return set_error(RP_E_INVALID, "bad shape");
  const size_t n_rows = (size_t)n_nodes * (size_t)n_sites;
  const unsigned nt = host_threads(n_rows, 1 << 14);
  auto job = [&](unsigned t) {
    float p[32];
    uint8_t st[32];
    for (size_t r = n_rows * t / nt; r < n_rows * (t + 1) / nt; r++) {
      const float* in = probs + r * n_states;
      for (int i = 0; i < n_states; i++) {
        float v = in[i];
        if (v < site_pp_threshold) v = site_pp_threshold;       // :218-219
        if (as_log10) v = (float)log10((double)v);              // :220-221  (float)Math.log10(sp.proba)
        p[i] = v;
        st[i] = state_of_column[i];
      }
      // Collections.sort with SiteProba.compareTo: stable, descending; a stable insertion sort gives the same
      // order for any input the comparator orders consistently ((a - b) < 0.0 / > 0.0 in float)
      for (int i = 1; i < n_states; i++) {
        const float v = p[i];
        const uint8_t s8 = st[i];
        int j = i - 1;
        while (j >= 0 && (p[j] - v) < 0.0f) { p[j + 1] = p[j]; st[j + 1] = st[j]; j--; }
        p[j + 1] = v;
        st[j + 1] = s8;
      }
      memcpy(pp_out + r * n_states, p, sizeof(float) * n_states);
      memcpy(states_out + r * n_states, st, n_states);
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++) th.emplace_back(job, t);
  job(0);
  for (auto& x : th) x.join();
  return RP_OK;
}

// alignement/Alignment.java:231-260
int rp_gap_intervals(const uint8_t* chars, int32_t n_rows, int32_t n_cols, uint64_t* gap_off, int32_t* gap_len, uint64_t cap,
                     uint64_t* n_len) {
  if (!chars || !gap_off || !n_len || n_rows < 0 || n_cols < 0) return set_error(RP_E_INVALID, "bad argument");
  std::vector<std::vector<int32_t>> iv((size_t)n_cols);  // gapIntervals[col]: empty = null
  for (int i = 0; i < n_rows; i++) {
    const uint8_t* row = chars + (size_t)i * n_cols;
    int first = -1;
    uint8_t prev = 'n';
    for (int j = 0; j < n_cols; j++) {
      const uint8_t c = row[j];
      if (c == '-') {
        if (prev != '-' && first == -1) first = j;                       // :240-245
      } else if (first != -1) {
        const int32_t len = j - first;                                   // :249
        std::vector<int32_t>& l = iv[(size_t)first];
        if (std::find(l.begin(), l.end(), len) == l.end()) l.push_back(len);  // :250-252
        first = -1;
      }
      prev = c;
    }
  }
  uint64_t n = 0;
  for (int j = 0; j < n_cols; j++) { gap_off[j] = n; n += iv[(size_t)j].size(); }
  gap_off[n_cols] = n;
  *n_len = n;
  if (!gap_len) return RP_OK;
  if (cap < n) return set_error(RP_E_INVALID, "gap_len holds %llu entries, %llu needed", (unsigned long long)cap, (unsigned long long)n);
  for (int j = 0; j < n_cols; j++)
    if (!iv[(size_t)j].empty()) memcpy(gap_len + gap_off[j], iv[(size_t)j].data(), iv[(size_t)j].size() * sizeof(int32_t));
  return RP_OK;
}

// test hook: Float.toString / Double.toString as the writer prints them
int rp_java_number(double v, int32_t as_float, char* out, int32_t cap) {
  std::string s;
  if (as_float) java_number(s, (float)v);
  else java_number(s, v);
  if ((int)s.size() + 1 > cap) return set_error(RP_E_INVALID, "buffer too small");
  memcpy(out, s.c_str(), s.size() + 1);
  return RP_OK;
}

}  // extern "C"
