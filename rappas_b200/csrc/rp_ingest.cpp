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
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <charconv>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rappas_b200.h"

namespace rp {
int set_error(int code, const char* fmt, ...);
}
using rp::set_error;

struct rp_reads {
  // every FASTA record, in file order
  std::vector<uint8_t> hdr;        // headers (first line without '>'), concatenated
  std::vector<uint64_t> hdr_off;   // [n_records + 1]
  std::vector<uint32_t> unique_of; // record -> index of its exact sequence among the unique ones
  std::vector<uint32_t> group_of;  // record -> id of its gap-stripped sequence (the reference's MD5 key)
  // distinct exact sequences (gaps kept: Main_PLACEMENT_v07.java:195), in order of first appearance
  std::vector<uint8_t> seq;
  std::vector<uint64_t> seq_off;   // [n_unique + 1]
  uint32_t n_groups = 0;
};

namespace {

struct Slice {
  const uint8_t* p;
  size_t n;
  bool operator==(const Slice& o) const { return n == o.n && (n == 0 || memcmp(p, o.p, n) == 0); }
};
struct SliceHash {
  size_t operator()(const Slice& s) const {  // FNV-1a 64; equality is checked on the bytes, so this only buckets
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < s.n; i++) { h ^= s.p[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};

// java.io.BufferedReader.readLine: a line ends at \n, \r or \r\n
inline size_t next_line(const uint8_t* t, size_t n, size_t pos, size_t* line_end) {
  size_t e = pos;
  while (e < n && t[e] != '\n' && t[e] != '\r') e++;
  *line_end = e;
  if (e < n && t[e] == '\r' && e + 1 < n && t[e + 1] == '\n') return e + 2;
  return e < n ? e + 1 : n;
}

int parse(const uint8_t* t, size_t n, rp_reads* R) {
  // pass 1: records -> (header, joined + trimmed sequence) in a scratch arena
  std::vector<uint8_t> all_seq;
  std::vector<uint64_t> all_off{0};
  R->hdr_off.assign(1, 0);
  bool in_record = false;
  size_t pos = 0;
  auto close_record = [&]() {
    // String.trim(): strip code points <= U+0020 from both ends of the joined sequence (FASTAPointer.java:143-145)
    size_t b = all_off.back(), e = all_seq.size();
    while (e > b && all_seq[e - 1] <= 0x20) e--;
    size_t s = b;
    while (s < e && all_seq[s] <= 0x20) s++;
    if (s > b) memmove(&all_seq[b], &all_seq[s], e - s);
    all_seq.resize(b + (e - s));
    all_off.push_back(all_seq.size());
  };
  while (pos < n) {
    size_t le;
    const size_t nxt = next_line(t, n, pos, &le);
    const uint8_t* line = t + pos;
    const size_t len = le - pos;
    pos = nxt;
    if (len == 0 || line[0] == '#') continue;  // empty and '#' lines are skipped (FASTAPointer.java:82-87)
    if (line[0] == '>') {
      if (in_record) close_record();
      R->hdr.insert(R->hdr.end(), line + 1, line + len);
      R->hdr_off.push_back(R->hdr.size());
      in_record = true;
      continue;
    }
    if (!in_record) return set_error(RP_E_IO, "FASTA: sequence data before the first '>' header");
    all_seq.insert(all_seq.end(), line, line + len);
  }
  if (in_record) close_record();
  const size_t nrec = R->hdr_off.size() - 1;
  if (nrec == 0) return set_error(RP_E_IO, "No valid fasta sequences were found");  // FASTAPointer.checkSize :238-241
  if (nrec >= 0xFFFFFFFFull) return set_error(RP_E_UNSUPPORTED, "more than 2^32-2 records");
  // pass 2: exact-sequence uniques (what is placed) and gap-stripped groups (the reference's checksum key)
  R->unique_of.resize(nrec);
  R->group_of.resize(nrec);
  R->seq_off.assign(1, 0);
  R->seq.reserve(all_seq.size());
  std::unordered_map<Slice, uint32_t, SliceHash> uniq, groups;
  uniq.reserve(nrec * 2);
  groups.reserve(nrec * 2);
  std::vector<std::vector<uint8_t>> stripped_store;  // keeps the gap-stripped keys that differ from their sequence alive
  for (size_t r = 0; r < nrec; r++) {
    Slice s{all_seq.data() + all_off[r], (size_t)(all_off[r + 1] - all_off[r])};
    auto it = uniq.find(s);
    if (it == uniq.end()) {
      // note: slices must point into storage that never moves -> R->seq was reserved for the worst case
      const size_t b = R->seq.size();
      R->seq.insert(R->seq.end(), s.p, s.p + s.n);
      R->seq_off.push_back(R->seq.size());
      const uint32_t id = (uint32_t)(R->seq_off.size() - 2);
      uniq.emplace(Slice{R->seq.data() + b, s.n}, id);
      R->unique_of[r] = id;
    } else {
      R->unique_of[r] = it->second;
    }
    // fasta.getSequence(true): '-' removed (Fasta.java:35-39), PlacementProcess.java:593
    Slice g = s;
    if (memchr(s.p, '-', s.n)) {
      std::vector<uint8_t> st;
      st.reserve(s.n);
      for (size_t i = 0; i < s.n; i++)
        if (s.p[i] != '-') st.push_back(s.p[i]);
      stripped_store.push_back(std::move(st));
      g = Slice{stripped_store.back().data(), stripped_store.back().size()};
    } else {
      // point into R->seq (stable) rather than the scratch arena
      const uint32_t u = R->unique_of[r];
      g = Slice{R->seq.data() + R->seq_off[u], s.n};
    }
    auto gt = groups.find(g);
    if (gt == groups.end()) {
      groups.emplace(g, R->n_groups);
      R->group_of[r] = R->n_groups++;
    } else {
      R->group_of[r] = gt->second;
    }
  }
  return RP_OK;
}

// ---- Java number formatting (json-simple prints Float / Double with toString()) ---------------------
// Double.toString / Float.toString layout: shortest digits that round-trip; plain decimal for
// 1e-3 <= |x| < 1e7 (at least one digit after the point), otherwise d.dddE[-]n.
template <typename T>
void java_number(std::string& out, T v) {
  if (isnan(v) || isinf(v)) { out += "null"; return; }  // JSONValue.toJSONString: NaN / Infinity -> null
  if (v == 0) { out += signbit(v) ? "-0.0" : "0.0"; return; }
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);  // shortest round-trip digits
  std::string s(buf, res.ptr);
  // s = [-]d[.ddd]e[+-]xx
  const bool neg = s[0] == '-';
  const size_t ms = neg ? 1 : 0, ep = s.find('e');
  std::string digits;
  for (size_t i = ms; i < ep; i++)
    if (s[i] != '.') digits += s[i];
  const int exp10 = atoi(s.c_str() + ep + 1);
  if (neg) out += '-';
  if (exp10 >= -3 && exp10 < 7) {
    if (exp10 >= 0) {
      for (int i = 0; i <= exp10; i++) out += i < (int)digits.size() ? digits[i] : '0';
      out += '.';
      if ((int)digits.size() > exp10 + 1) out.append(digits, exp10 + 1, std::string::npos);
      else out += '0';
    } else {
      out += "0.";
      for (int i = 0; i < -exp10 - 1; i++) out += '0';
      out += digits;
    }
  } else {
    out += digits[0];
    out += '.';
    if (digits.size() > 1) out.append(digits, 1, std::string::npos);
    else out += '0';
    out += 'E';
    out += std::to_string(exp10);
  }
}

void json_string(std::string& out, const uint8_t* p, size_t n) {  // JSONValue.escape
  out += '"';
  for (size_t i = 0; i < n; i++) {
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
  std::vector<int64_t> placement_of_group(r->n_groups, -1);
  struct Placement { uint32_t first_record; std::vector<uint32_t> dups; };
  std::vector<Placement> placements;
  for (size_t i = 0; i < nrec; i++) {
    const uint32_t g = r->group_of[i], u = r->unique_of[i];
    if (placement_of_group[g] >= 0) { placements[(size_t)placement_of_group[g]].dups.push_back((uint32_t)i); continue; }
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
    placement_of_group[g] = (int64_t)placements.size();
    placements.push_back({(uint32_t)i, {}});
  }
  std::string o;
  o.reserve(1 << 20);
  o += "{\n\"tree\":";
  if (tree_newick) json_string(o, (const uint8_t*)tree_newick, strlen(tree_newick));
  else o += "null";
  o += ",\n\"placements\":\n[";
  bool ok = true;
  for (size_t p = 0; p < placements.size() && ok; p++) {
    const uint32_t i = placements[p].first_record, u = r->unique_of[i];
    o += p ? ",\n{\n\t\"p\":\n\t[" : "\n{\n\t\"p\":\n\t[";
    for (int k = 0; k < n_rows[u]; k++) {
      const size_t at = (size_t)u * K + k;
      const unsigned x = node[at];
      if ((int)x >= n_nodes) { ok = false; break; }
      if (k) o += ",\n\t";
      o += '[';
      const float distal = branch_len[x] / 2.0f;  // getBranchLengthToAncestor()/2 (:1015, :1022)
      if (guppy_compat) {  // distal_length, edge_num, like_weight_ratio, likelihood, pendant_length (:1005-1016)
        java_number(o, distal); o += ','; o += std::to_string(edge_id[x]); o += ',';
        java_number(o, lwr[at]); o += ','; java_number(o, score[at]); o += ",0.0";
      } else {             // edge_num, likelihood, like_weight_ratio, distal_length, pendant_length (:1017-1024)
        o += std::to_string(edge_id[x]); o += ','; java_number(o, score[at]); o += ',';
        java_number(o, lwr[at]); o += ','; java_number(o, distal); o += ",0.0";
      }
      o += ']';
    }
    o += "],\n\t\"nm\":\n\t[[";
    json_string(o, r->hdr.data() + r->hdr_off[i], r->hdr_off[i + 1] - r->hdr_off[i]);  // full header (:1041)
    o += ",1]";
    for (uint32_t d : placements[p].dups) {
      const uint8_t* h = r->hdr.data() + r->hdr_off[d];
      size_t n = r->hdr_off[d + 1] - r->hdr_off[d];
      const void* sp = memchr(h, ' ', n);  // header up to the first space (:598-602)
      if (sp) n = (const uint8_t*)sp - h;
      o += ",\n\t[";
      json_string(o, h, n);
      o += ",1]";
    }
    o += "]\n}";
    if (o.size() > (1 << 20)) { ok = fwrite(o.data(), 1, o.size(), f) == o.size(); o.clear(); }
  }
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
  if (n_placements) *n_placements = placements.size();
  return ok ? RP_OK : set_error(RP_E_IO, "short write to %s", path);
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
