// rp_db.cu -- host side of the device-resident phylo-kmer DB: alphabet tables, threshold,
// flat CSR -> (open-addressing table + posting blocks), upload, .rgdb file I/O, error plumbing.
//
// Replaces, for the placement path only, the in-JVM CustomHash_v4_FastUtil81
// (core/hash/CustomHash_v4_FastUtil81.java:36: Object2ObjectOpenCustomHashMap<byte[],Char2FloatOpenHashMap>)
// and the part of SessionNext_v2.load (main_v2/SessionNext_v2.java:158-207) the hot path consumes.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include <algorithm>
#include <thread>

#include "rp_common.h"

namespace rp {

static thread_local char t_err[512];
std::atomic<uint64_t> g_kernel_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof t_err, fmt, ap);
  va_end(ap);
  return code;
}

int alphabet_bits(int alphabet) { return alphabet == RP_ALPHA_NUCL ? 2 : 5; }
int alphabet_states(int alphabet) { return alphabet == RP_ALPHA_NUCL ? 4 : 20; }

// AmbigSequenceKnife ctor (core/algos/AmbigSequenceKnife.java:95):
//   maxAmbigPerMer=(int)Math.floor(Math.pow(k, 1.0/s.getNonAmbiguousStatesCount()))
int max_ambig_per_mer(int alphabet, int k) {
  return (int)floor(pow((double)k, 1.0 / (double)alphabet_states(alphabet)));
}

// Character classes.  isAmbiguous() is consulted before stateToByte() (AmbigSequenceKnife.java:106,123),
// so an ambiguity letter wins over a state letter of the same spelling.
void build_alphabet_tables(int alphabet, AlphabetTables* t) {
  memset(t->cls, kClsBad, sizeof t->cls);
  memset(t->alt_n, 0, sizeof t->alt_n);
  memset(t->alt_states, 0, sizeof t->alt_states);
  auto both_cases = [&](char upper, uint8_t v) {
    t->cls[(unsigned char)upper] = v;
    t->cls[(unsigned char)(upper + 32)] = v;
  };
  int next_set = 0;
  auto add_set = [&](const char* letters_cased, std::initializer_list<int> states) {
    int id = next_set++;
    t->alt_n[id] = (uint8_t)states.size();
    int i = 0;
    for (int s : states) t->alt_states[id][i++] = (uint8_t)s;
    for (const char* p = letters_cased; *p; ++p) t->cls[(unsigned char)*p] = (uint8_t)(kClsAmb | id);
  };
  if (alphabet == RP_ALPHA_NUCL) {
    // DNAStatesShifted.charToByte, core/DNAStatesShifted.java:182-209
    enum { A = 0, T = 1, C = 2, G = 3 };
    both_cases('A', A); both_cases('T', T); both_cases('U', T); both_cases('C', C); both_cases('G', G);
    // IUPAC alternatives in the literal array order of core/DNAStatesShifted.java:62-96
    add_set("Rr", {A, G}); add_set("Yy", {C, T}); add_set("Ss", {C, G}); add_set("Ww", {A, T});
    add_set("Kk", {G, T}); add_set("Mm", {A, C});
    add_set("Bb", {C, G, T}); add_set("Dd", {A, G, T}); add_set("Hh", {A, C, T}); add_set("Vv", {A, C, G});
    add_set("Nn", {A, C, G, T});
    // '.' and '-' are registered as new byte[4] and never filled (:57-58): four times state 0
    add_set(".-", {0, 0, 0, 0});
  } else {
    // AAStates, core/AAStates.java:23-34, 74-93
    const char* order = "RHKDESTNQCGPAILMFWYV";
    for (int i = 0; i < 20; i++) both_cases(order[i], (uint8_t)i);
    if (alphabet == RP_ALPHA_AMINO_UO) {  // :118-123
      both_cases('U', 9);
      both_cases('O', 14);
    }
    // :97-107
    add_set("-*!Xx", {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19});
    add_set("Bb", {3, 7});
    add_set("Zz", {4, 8});
    add_set("Jj", {13, 14});
  }
}

static int log2_u64(uint64_t x) { int l = 0; while ((1ull << l) < x) l++; return l; }

DbView make_db_view(const rp_db* db, const DeviceCtx* dc) {
  DbView v;
  memset(&v, 0, sizeof v);
  v.n_parts = (int)dc->parts.size();
  v.table_parts = db->table_replicated ? 1 : v.n_parts;
  for (int i = 0; i < v.n_parts; i++) {
    const Partition& pt = db->parts[dc->parts[i]];
    v.table[i] = pt.d_table;
    v.blocks[i] = pt.d_blocks;
    v.bucket_shift[i] = 32 - log2_u64(pt.n_buckets);
  }
  if (db->table_replicated) {  // every partition's device holds the whole table: probe the local copy
    const Partition& pt = db->parts[dc->local_part];
    v.table[0] = pt.d_table;
    v.bucket_shift[0] = 32 - log2_u64(pt.n_buckets);
  }
  // direct-address table of a replicated small-key-space DB (this device's copy)
  v.direct = (!db->partitioned && v.n_parts == 1) ? db->parts[dc->parts[0]].d_direct : nullptr;
  v.alphabet = db->desc.alphabet;
  v.k = db->desc.k;
  v.bits = alphabet_bits(db->desc.alphabet);
  v.n_nodes = db->desc.n_nodes;
  v.max_amb = max_ambig_per_mer(db->desc.alphabet, db->desc.k);
  v.T = db->desc.thr_log10;
  v.Tlin = db->desc.thr_lin;
  return v;
}

static int check_desc(const rp_db_desc* d) {
  if (!d) return set_error(RP_E_INVALID, "desc is NULL");
  if (d->alphabet < 0 || d->alphabet > 2) return set_error(RP_E_INVALID, "alphabet %d not in {0,1,2}", d->alphabet);
  int kmax = d->alphabet == RP_ALPHA_NUCL ? 31 : 12;
  if (d->k < 2 || d->k > kmax) return set_error(RP_E_INVALID, "k=%d out of range [2,%d]", d->k, kmax);
  if (d->n_nodes < 1 || d->n_nodes > 65535)
    return set_error(RP_E_INVALID, "n_nodes=%d out of range [1,65535] (node ids are Java chars)", d->n_nodes);
  return RP_OK;
}

// ---- host build of table + blocks (multi-threaded over key ranges) ------------------------------
struct HostImage {
  std::vector<uint64_t> table;   // 2 u64 per slot, 2 slots per bucket
  std::vector<uint64_t> direct;  // meta per planar key (kEmptyKey = absent): whole-DB images of small nucleotide key spaces
  uint8_t* blocks = nullptr;     // malloc'd, block_bytes
  uint64_t n_buckets = 0, block_bytes = 0, max_block_bytes = 0;
  ~HostImage() { free(blocks); }
};

// Serial cuckoo insertion of the keys the greedy parallel pass could not place: random-walk
// eviction between a key's two buckets.  At load <= 0.5 with 2x2 buckets this is a handful of keys.
static bool cuckoo_insert(std::vector<uint64_t>& tab, int shift, uint64_t key, uint64_t meta, uint64_t* rng) {
  for (int kick = 0; kick < 2000; kick++) {
    const KeyHash m = hash_key(key);
    const uint32_t b[2] = {bucket1(m, shift), bucket2(m, shift)};
    for (int c = 0; c < 2; c++)
      for (int s = 0; s < kBucketSlots; s++) {
        uint64_t* slot = &tab[2 * ((uint64_t)b[c] * kBucketSlots + s)];
        if (slot[0] == kEmptyKey) { slot[0] = key; slot[1] = meta; return true; }
      }
    *rng = *rng * 6364136223846793005ull + 1442695040888963407ull;
    const uint32_t pick = (uint32_t)(*rng >> 33);
    uint64_t* victim = &tab[2 * ((uint64_t)b[pick & 1] * kBucketSlots + ((pick >> 1) & 1))];
    std::swap(key, victim[0]);
    std::swap(meta, victim[1]);
  }
  return false;
}

// Builds the image of the keys sel[0..n_sel) (all keys if sel == nullptr) as partition `part`.
// owners != nullptr (with sel == nullptr): the table holds ALL keys, key i pointing into the posting blocks
// of partition owners[i] (offsets counted per partition, in key order, so every builder agrees on them),
// and only the blocks of partition `part` are materialised -- the layout with a replicated table.
static int build_image(const rp_db_desc* d, const uint64_t* keys, const uint64_t* offsets, const uint16_t* post_node,
                       const float* post_score, const uint64_t* sel, uint64_t n_sel, int part, HostImage* img,
                       const uint8_t* owners = nullptr) {
  const uint64_t nk = sel ? n_sel : d->n_keys;
  auto key_at = [&](uint64_t j) { return sel ? sel[j] : j; };
  auto owner_at = [&](uint64_t j) { return owners ? (int)owners[key_at(j)] : part; };
  if (offsets && d->n_keys && offsets[d->n_keys] != d->n_postings)
    return set_error(RP_E_INVALID, "offsets[n_keys]=%llu != n_postings=%llu", (unsigned long long)offsets[d->n_keys],
                     (unsigned long long)d->n_postings);
  uint64_t nb = kMinBuckets;
  while (nb < nk) nb <<= 1;  // 2 slots per bucket -> load factor in (0.25, 0.5]
  if (nb > (1ull << 31)) return set_error(RP_E_UNSUPPORTED, "n_keys=%llu: more than 2^31 buckets", (unsigned long long)nk);
  const int shift = 32 - log2_u64(nb);
  img->n_buckets = nb;
  const uint64_t n_slots = nb * kBucketSlots;
  img->table.assign(2 * n_slots, 0);
  for (uint64_t i = 0; i < n_slots; i++) img->table[2 * i] = kEmptyKey;
  // block offsets (32 B units) of key j inside its owner partition's blocks
  std::vector<uint64_t> boff(nk + 1, 0);
  uint64_t run[kMaxParts] = {0};
  uint64_t max_bb = 0;
  for (uint64_t j = 0; j < nk; j++) {
    const uint64_t i = key_at(j);
    if (offsets[i + 1] < offsets[i]) return set_error(RP_E_INVALID, "offsets not monotone at key %llu", (unsigned long long)i);
    uint64_t P = offsets[i + 1] - offsets[i];
    if (P > 65535) return set_error(RP_E_INVALID, "key %llu has %llu postings (> 65535)", (unsigned long long)i, (unsigned long long)P);
    uint64_t& r = run[owners ? owner_at(j) : 0];
    boff[j] = r;
    r += block_bytes_for(P) / kBlockAlign;
    if (r > kMetaOffMask) return set_error(RP_E_INVALID, "posting blocks exceed 2^37 * 32 B");
    max_bb = std::max(max_bb, block_bytes_for(P));
  }
  img->block_bytes = run[owners ? part : 0] * kBlockAlign;
  img->max_block_bytes = max_bb;
  img->blocks = (uint8_t*)calloc(img->block_bytes ? img->block_bytes : 32, 1);
  if (!img->blocks) return set_error(RP_E_NOMEM, "cannot allocate %llu B for posting blocks", (unsigned long long)img->block_bytes);

  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (nk < 4096) nt = 1;
  std::atomic<int> err{0};
  const int n_nodes = d->n_nodes;
  const int bits = alphabet_bits(d->alphabet), k = d->k;
  // Direct-address table beside the cuckoo table when the whole key space is small (nucleotide k <= 12: 4^k x 8 B
  // <= 128 MB): one 8 B load per window, no hashing, no key compare.  Whole-DB images only.
  const bool want_direct = !sel && !owners && d->alphabet == RP_ALPHA_NUCL && 2 * k <= 24 && !getenv("RP_NO_DIRECT");
  if (want_direct) img->direct.assign((size_t)1 << (2 * k), kEmptyKey);
  const uint64_t code_limit = (bits * k >= 64) ? ~0ull : (1ull << (bits * k));
  std::vector<std::vector<std::pair<uint64_t, uint64_t>>> leftover(nt);
  auto pack_range = [&](unsigned tid, uint64_t k0, uint64_t k1) {
    std::vector<std::pair<uint16_t, float>> tmp;
    for (uint64_t j = k0; j < k1 && !err.load(std::memory_order_relaxed); j++) {
      const uint64_t i = key_at(j);
      const uint64_t lo = offsets[i], P = offsets[i + 1] - offsets[i];
      if (keys[i] == kEmptyKey || keys[i] >= code_limit) { err = 4; return; }
      tmp.resize(P);
      bool sorted = true;
      for (uint64_t p = 0; p < P; p++) {
        tmp[p] = {post_node[lo + p], post_score[lo + p]};
        if (post_node[lo + p] >= n_nodes) { err = 1; return; }
        if (p && tmp[p].first <= tmp[p - 1].first) sorted = false;
      }
      if (!sorted) {
        // node order inside a key only decides L order (tie-breaks); ascending ids make the
        // shared-memory accumulation bank-conflict free for contiguous runs
        std::sort(tmp.begin(), tmp.end(), [](auto& a, auto& b) { return a.first < b.first; });
        for (uint64_t p = 1; p < P; p++)
          if (tmp[p].first == tmp[p - 1].first) { err = 2; return; }  // one value per (k-mer,node): CustomHash_v4:76-89
      }
      const int own = owner_at(j);
      uint8_t* blk = img->blocks + boff[j] * kBlockAlign;
      for (uint64_t base = 0; base < P && own == part; base += kSubBlock) {
        uint64_t m = std::min<uint64_t>(kSubBlock, P - base);
        float* sc = (float*)(blk + (base / kSubBlock) * kSubBlockBytes);
        uint16_t* nd = (uint16_t*)((uint8_t*)sc + 4 * m);
        for (uint64_t q = 0; q < m; q++) { sc[q] = tmp[base + q].second; nd[q] = tmp[base + q].first; }
      }
      // greedy 2-choice placement (lock-free: CAS on the key word, then publish meta)
      const uint64_t key = planar_from_code(keys[i], bits, k);
      // node range of the list in sixteenths of the padded tree (tmp is sorted by node here)
      const uint64_t n_pad = (uint64_t)padded_nodes(n_nodes);
      const uint64_t qmin = P ? (uint64_t)tmp[0].first * 16 / n_pad : 0, qmax = P ? (uint64_t)tmp[P - 1].first * 16 / n_pad : 0;
      const uint64_t meta = ((uint64_t)own << kMetaPartShift) | (qmax << kMetaQmaxShift) | (qmin << kMetaQminShift) |
                            (boff[j] << 16) | P;
      if (want_direct) img->direct[key] = meta;  // (a duplicate key is caught by the cuckoo placement below)
      const KeyHash m32 = hash_key(key);
      const uint32_t bk[2] = {bucket1(m32, shift), bucket2(m32, shift)};
      bool placed = false;
      for (int c = 0; c < 2 && !placed; c++)
        for (int s = 0; s < kBucketSlots && !placed; s++) {
          uint64_t* slot = &img->table[2 * ((uint64_t)bk[c] * kBucketSlots + s)];
          uint64_t expect = kEmptyKey;
          if (__atomic_compare_exchange_n(slot, &expect, key, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
            slot[1] = meta;
            placed = true;
          } else if (expect == key) { err = 3; return; }
        }
      if (!placed) leftover[tid].push_back({key, meta});
    }
  };
  if (nt == 1) {
    pack_range(0, 0, nk);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back(pack_range, t, nk * t / nt, nk * (t + 1) / nt);
    for (auto& x : th) x.join();
  }
  if (!err.load()) {
    uint64_t rng = 0x243F6A8885A308D3ull;
    for (auto& lv : leftover)
      for (auto& kv : lv) {
        // a key stays inside its own two buckets for ever, so a duplicate is visible there
        const KeyHash m32 = hash_key(kv.first);
        const uint32_t bk[2] = {bucket1(m32, shift), bucket2(m32, shift)};
        for (int c = 0; c < 2; c++)
          for (int s = 0; s < kBucketSlots; s++)
            if (img->table[2 * ((uint64_t)bk[c] * kBucketSlots + s)] == kv.first) err = 3;
        if (err.load()) break;
        if (!cuckoo_insert(img->table, shift, kv.first, kv.second, &rng)) { err = 5; break; }
      }
  }
  switch (err.load()) {
    case 1: return set_error(RP_E_INVALID, "posting node id >= n_nodes");
    case 2: return set_error(RP_E_INVALID, "a key lists the same node twice");
    case 3: return set_error(RP_E_INVALID, "duplicate key");
    case 4: return set_error(RP_E_INVALID, "key out of range for this alphabet and k (or the reserved 0xFFFFFFFFFFFFFFFF)");
    case 5: return set_error(RP_E_INVALID, "cuckoo placement failed: a set of keys shares both candidate buckets (2 x 2 slots); "
                              "with independent 32-bit bucket hashes this needs > 4 keys colliding in 64 hash bits");
    default: break;
  }
  return RP_OK;
}

static void free_device_ctx(DeviceCtx* dc) {
  if (!dc) return;
  if (dc->device >= 0 && cudaSetDevice(dc->device) == cudaSuccess) {
    std::vector<StreamCtx*> all = {&dc->sc[0], &dc->sc[1]};
    for (auto* ds : dc->dev_slots) {
      if (ds->user) cudaStreamSynchronize(ds->user);  // (a destroyed caller stream only returns an error here)
      cudaGetLastError();
      all.push_back(&ds->sc);
    }
    for (StreamCtx* sp : all) {
      StreamCtx& s = *sp;
      if (s.stream) cudaStreamSynchronize(s.stream);
      cudaFree(s.d_counter); cudaFree(s.d_amb_S); cudaFree(s.d_amb_C);
      cudaFree(s.d_seq); cudaFree(s.d_off); cudaFree(s.d_n_rows); cudaFree(s.d_node); cudaFree(s.d_score);
      cudaFree(s.d_lwr); cudaFree(s.d_counts); cudaFree(s.d_status);
      if (s.h_in) cudaFreeHost(s.h_in);
      if (s.h_out) cudaFreeHost(s.h_out);
      if (s.ev_k0) cudaEventDestroy(s.ev_k0);
      if (s.ev_k1) cudaEventDestroy(s.ev_k1);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (auto* ds : dc->dev_slots) { if (ds->last) cudaEventDestroy(ds->last); delete ds; }
  }
  delete dc;
}

}  // namespace rp

using namespace rp;

extern "C" {

const char* rp_last_error(void) { return t_err; }
const char* rp_version(void) { return "rappas_b200 0.1 (sm_100a, abi 1)"; }
uint64_t rp_kernel_launch_count(void) { return g_kernel_launches.load(); }

int rp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// Main_DBBUILD_3.java:165-166 (float/double sequence kept literally)
void rp_threshold(float omega, int32_t alphabet, int32_t k, float* thr_lin, float* thr_log10) {
  float ratio = omega / (float)alphabet_states(alphabet);
  float lin = (float)pow(0.0 + (double)ratio, (double)k);
  float lg = (float)log10((double)lin);
  if (thr_lin) *thr_lin = lin;
  if (thr_log10) *thr_log10 = lg;
}

uint64_t rp_pack_kmer(int32_t alphabet, const uint8_t* states, int32_t k) {
  uint64_t code = 0;
  const int bits = alphabet_bits(alphabet);
  for (int i = 0; i < k; i++) code |= (uint64_t)states[i] << (bits * i);
  return code;
}

int rp_db_load(const rp_db_desc* desc, const uint64_t* keys, const uint64_t* offsets, const uint16_t* post_node,
               const float* post_score, const int32_t* devices, int32_t n_devices, int32_t partitioned,
               rp_db** out) {
  int rc = check_desc(desc);
  if (rc) return rc;
  if (!out) return set_error(RP_E_INVALID, "out is NULL");
  *out = nullptr;
  if (n_devices < 1 || !devices) return set_error(RP_E_INVALID, "need at least one device");
  if (desc->n_keys && (!keys || !offsets)) return set_error(RP_E_INVALID, "keys/offsets are NULL");
  if (desc->n_postings && (!post_node || !post_score)) return set_error(RP_E_INVALID, "posting arrays are NULL");
  int ndev_avail = rp_device_count();
  if (ndev_avail == 0)
    return set_error(RP_E_CUDA, "no CUDA device visible: librappas_b200 has no CPU fallback");
  for (int i = 0; i < n_devices; i++)
    if (devices[i] < 0 || devices[i] >= ndev_avail) return set_error(RP_E_INVALID, "device %d not present", devices[i]);

  static const uint64_t zero_off[1] = {0};
  if (!desc->n_keys) offsets = zero_off;
  // distinct execution devices; with partitioned != 0 every entry of devices[] is one partition (an
  // entry may repeat a device: several partitions in one HBM, which is how a 1-GPU box tests the mode)
  std::vector<int> exec;
  for (int i = 0; i < n_devices; i++)
    if (std::find(exec.begin(), exec.end(), devices[i]) == exec.end()) exec.push_back(devices[i]);
  const int n_parts = partitioned ? n_devices : 1;
  if (n_parts > kMaxParts) return set_error(RP_E_UNSUPPORTED, "at most %d partitions", kMaxParts);

  if (partitioned < 0 || partitioned > 2) return set_error(RP_E_INVALID, "partitioned must be 0, 1 or 2");
  rp_db* db = new rp_db();
  db->desc = *desc;
  db->partitioned = partitioned ? 1 : 0;
  db->table_replicated = partitioned == 2;
  build_alphabet_tables(desc->alphabet, &db->alpha);
  auto fail = [&](int code) { rp_db_free(db); return code; };
  auto upload = [&](const HostImage& img, int device) -> int {
    Partition pt;
    pt.device = device;
    pt.n_buckets = img.n_buckets;
    pt.block_bytes = img.block_bytes;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pt.d_table, img.n_buckets * 32);
    // +512 B: idle lanes of the last block's last chunk may address (never load) past its end
    if (e == cudaSuccess) e = cudaMalloc((void**)&pt.d_blocks, img.block_bytes + 512);
    if (e == cudaSuccess) e = cudaMemcpy(pt.d_table, img.table.data(), img.n_buckets * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && img.block_bytes) e = cudaMemcpy(pt.d_blocks, img.blocks, img.block_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && !img.direct.empty()) {
      e = cudaMalloc((void**)&pt.d_direct, img.direct.size() * 8);
      if (e == cudaSuccess) e = cudaMemcpy(pt.d_direct, img.direct.data(), img.direct.size() * 8, cudaMemcpyHostToDevice);
    }
    db->parts.push_back(pt);  // owned (and freed) by the db even on failure
    if (e != cudaSuccess) {
      const int code = (e == cudaErrorMemoryAllocation) ? RP_E_NOMEM : RP_E_CUDA;
      set_error(code, "device %d: %s while uploading the DB (%llu B table + %llu B blocks)", device,
                cudaGetErrorString(e), (unsigned long long)(img.n_buckets * 32), (unsigned long long)img.block_bytes);
      cudaGetLastError();
      return code;
    }
    return RP_OK;
  };

  if (!partitioned) {
    HostImage img;
    rc = build_image(desc, keys, offsets, post_node, post_score, nullptr, 0, 0, &img);
    if (rc) return fail(rc);
    db->n_buckets = img.n_buckets;
    db->block_bytes = img.block_bytes;
    db->max_block_bytes = img.max_block_bytes;
    for (int d : exec)
      if ((rc = upload(img, d))) return fail(rc);
  } else {
    // keys go to partition owner_of(mix(planar key)): the kernel recomputes the same owner from the key
    const int bits = alphabet_bits(desc->alphabet);
    std::vector<std::vector<uint64_t>> sel(n_parts);
    std::vector<uint8_t> owners(desc->n_keys);
    for (uint64_t i = 0; i < desc->n_keys; i++) {
      owners[i] = (uint8_t)owner_of(hash_key(planar_from_code(keys[i], bits, desc->k)), n_parts);
      sel[owners[i]].push_back(i);
    }
    for (int p = 0; p < n_parts; p++) {
      HostImage img;
      rc = partitioned == 2 ? build_image(desc, keys, offsets, post_node, post_score, nullptr, 0, p, &img, owners.data())
                            : build_image(desc, keys, offsets, post_node, post_score, sel[p].data(), sel[p].size(), p, &img);
      if (rc) return fail(rc);
      db->n_buckets = std::max(db->n_buckets, img.n_buckets);
      db->block_bytes += img.block_bytes;
      db->max_block_bytes = std::max(db->max_block_bytes, img.max_block_bytes);
      if ((rc = upload(img, devices[p]))) return fail(rc);
    }
    // every executing device must reach every partition: peer-map the HBM of the others (NVLink)
    for (int a : exec)
      for (int b : exec) {
        if (a == b) continue;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, a, b);
        if (!can) { set_error(RP_E_UNSUPPORTED, "device %d cannot map the memory of device %d (no peer access)", a, b); return fail(RP_E_UNSUPPORTED); }
        cudaSetDevice(a);
        cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
          set_error(RP_E_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", a, b, cudaGetErrorString(e));
          return fail(RP_E_CUDA);
        }
        cudaGetLastError();
      }
  }
  for (size_t i = 0; i < exec.size(); i++) {
    DeviceCtx* dc = new DeviceCtx();
    db->dev.push_back(dc);
    dc->device = exec[i];
    if (partitioned) for (int p = 0; p < n_parts; p++) dc->parts.push_back(p);
    else dc->parts.push_back((int)i);
    if (partitioned)  // a partition resident on this device (its table is the one this device probes when replicated)
      for (int p = n_parts - 1; p >= 0; p--)
        if (devices[p] == exec[i]) dc->local_part = p;
    cudaDeviceProp prop;
    cudaError_t e = cudaSetDevice(dc->device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dc->device);
    if (e != cudaSuccess) { set_error(RP_E_CUDA, "device %d: %s", dc->device, cudaGetErrorString(e)); return fail(RP_E_CUDA); }
    dc->sm_count = prop.multiProcessorCount;
    dc->smem_optin = prop.sharedMemPerBlockOptin;
    if ((rc = compute_geometry(db, dc))) return fail(rc);
  }
  *out = db;
  return RP_OK;
}

// ---- hash-partitioned DB across PROCESSES (one process per GPU, the bench / torchrun deployment) ------
// Rank p builds and uploads partition p only, exports a 152-byte blob (two cudaIpcMemHandle_t + sizes),
// the ranks exchange the blobs through whatever they have (torch.distributed all_gather), and every rank
// attaches the others' partitions: the kernels then reach them exactly like the in-process peer pointers.
struct PartBlob {
  cudaIpcMemHandle_t table, blocks;
  uint64_t n_buckets, block_bytes, n_keys;
};
static_assert(sizeof(PartBlob) == RP_PART_BLOB_BYTES, "blob layout");

int rp_db_load_partition(const rp_db_desc* desc, const uint64_t* keys, const uint64_t* offsets,
                         const uint16_t* post_node, const float* post_score, int32_t device, int32_t part,
                         int32_t n_parts, int32_t replicate_table, uint8_t* blob_out, rp_db** out) {
  int rc = check_desc(desc);
  if (rc) return rc;
  if (!out || !blob_out) return set_error(RP_E_INVALID, "NULL argument");
  *out = nullptr;
  if (n_parts < 1 || n_parts > kMaxParts || part < 0 || part >= n_parts) return set_error(RP_E_INVALID, "bad partition index");
  if (rp_device_count() == 0) return set_error(RP_E_CUDA, "no CUDA device visible: librappas_b200 has no CPU fallback");
  static const uint64_t zero_off[1] = {0};
  if (!desc->n_keys) offsets = zero_off;
  const int bits = alphabet_bits(desc->alphabet);
  std::vector<uint64_t> sel;
  std::vector<uint8_t> owners(desc->n_keys);
  for (uint64_t i = 0; i < desc->n_keys; i++) {
    owners[i] = (uint8_t)owner_of(hash_key(planar_from_code(keys[i], bits, desc->k)), n_parts);
    if ((int)owners[i] == part) sel.push_back(i);
  }
  HostImage img;
  rc = replicate_table ? build_image(desc, keys, offsets, post_node, post_score, nullptr, 0, part, &img, owners.data())
                       : build_image(desc, keys, offsets, post_node, post_score, sel.data(), sel.size(), part, &img);
  if (rc) return rc;
  rp_db* db = new rp_db();
  db->desc = *desc;
  db->partitioned = 2;  // 2 = waiting for rp_db_attach_partitions
  db->table_replicated = replicate_table != 0;
  build_alphabet_tables(desc->alphabet, &db->alpha);
  db->parts.resize(n_parts);
  Partition& pt = db->parts[part];
  pt.device = device;
  pt.n_buckets = img.n_buckets;
  pt.block_bytes = img.block_bytes;
  pt.n_keys = replicate_table ? desc->n_keys : sel.size();
  db->max_block_bytes = img.max_block_bytes;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&pt.d_table, img.n_buckets * 32);
  if (e == cudaSuccess) e = cudaMalloc((void**)&pt.d_blocks, img.block_bytes + 512);
  if (e == cudaSuccess) e = cudaMemcpy(pt.d_table, img.table.data(), img.n_buckets * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && img.block_bytes) e = cudaMemcpy(pt.d_blocks, img.blocks, img.block_bytes, cudaMemcpyHostToDevice);
  PartBlob blob;
  memset(&blob, 0, sizeof blob);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&blob.table, pt.d_table);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&blob.blocks, pt.d_blocks);
  if (e != cudaSuccess) {
    set_error(e == cudaErrorMemoryAllocation ? RP_E_NOMEM : RP_E_CUDA, "device %d: %s while uploading partition %d", device,
              cudaGetErrorString(e), part);
    cudaGetLastError();
    rp_db_free(db);
    return e == cudaErrorMemoryAllocation ? RP_E_NOMEM : RP_E_CUDA;
  }
  blob.n_buckets = img.n_buckets;
  blob.block_bytes = img.block_bytes;
  blob.n_keys = sel.size();
  memcpy(blob_out, &blob, sizeof blob);
  DeviceCtx* dc = new DeviceCtx();
  dc->device = device;
  dc->local_part = part;
  db->dev.push_back(dc);
  *out = db;
  return RP_OK;
}

int rp_db_partition_blob(rp_db* db, uint8_t* blob_out) {
  if (!db || !blob_out) return set_error(RP_E_INVALID, "NULL argument");
  if (db->dev.size() != 1) return set_error(RP_E_INVALID, "not a single-partition handle");
  const int part = db->dev[0]->local_part;
  if (part < 0 || part >= (int)db->parts.size() || !db->parts[part].d_table || db->parts[part].ipc)
    return set_error(RP_E_INVALID, "the handle holds no partition of its own");
  Partition& pt = db->parts[part];
  RP_CUDA_TRY(cudaSetDevice(pt.device));
  PartBlob blob;
  memset(&blob, 0, sizeof blob);
  RP_CUDA_TRY(cudaIpcGetMemHandle(&blob.table, pt.d_table));
  RP_CUDA_TRY(cudaIpcGetMemHandle(&blob.blocks, pt.d_blocks));
  blob.n_buckets = pt.n_buckets;
  blob.block_bytes = pt.block_bytes;
  blob.n_keys = db->desc.n_keys;
  memcpy(blob_out, &blob, sizeof blob);
  return RP_OK;
}

int rp_db_attach_partitions(rp_db* db, const uint8_t* blobs, int32_t n_parts) {
  if (!db || !blobs) return set_error(RP_E_INVALID, "NULL argument");
  if (db->partitioned != 2 || (int)db->parts.size() != n_parts || db->dev.size() != 1)
    return set_error(RP_E_INVALID, "db was not created by rp_db_load_partition with %d partitions", n_parts);
  DeviceCtx* dc = db->dev[0];
  RP_CUDA_TRY(cudaSetDevice(dc->device));
  db->block_bytes = 0;
  db->n_buckets = 0;
  for (int p = 0; p < n_parts; p++) {
    PartBlob blob;
    memcpy(&blob, blobs + (size_t)p * sizeof blob, sizeof blob);
    Partition& pt = db->parts[p];
    if (pt.d_table == nullptr) {  // not mine: map the owner's allocations into this process
      pt.device = dc->device;
      pt.ipc = true;
      pt.n_buckets = blob.n_buckets;
      pt.block_bytes = blob.block_bytes;
      RP_CUDA_TRY(cudaIpcOpenMemHandle((void**)&pt.d_table, blob.table, cudaIpcMemLazyEnablePeerAccess));
      RP_CUDA_TRY(cudaIpcOpenMemHandle((void**)&pt.d_blocks, blob.blocks, cudaIpcMemLazyEnablePeerAccess));
    }
    db->block_bytes += pt.block_bytes;
    db->n_buckets = std::max(db->n_buckets, pt.n_buckets);
    dc->parts.push_back(p);
  }
  cudaDeviceProp prop;
  RP_CUDA_TRY(cudaGetDeviceProperties(&prop, dc->device));
  dc->sm_count = prop.multiProcessorCount;
  dc->smem_optin = prop.sharedMemPerBlockOptin;
  db->partitioned = 1;
  return compute_geometry(db, dc);
}

// owner partition of each ABI k-mer code (host helper; the sharding tests check it against a restatement)
int rp_partition_of_keys(int32_t alphabet, int32_t k, const uint64_t* keys, uint64_t n_keys, int32_t n_parts, int32_t* out) {
  if ((!keys || !out) && n_keys) return set_error(RP_E_INVALID, "NULL argument");
  if (n_parts < 1 || n_parts > kMaxParts) return set_error(RP_E_INVALID, "n_parts out of range");
  const int bits = alphabet_bits(alphabet);
  for (uint64_t i = 0; i < n_keys; i++) out[i] = (int32_t)owner_of(hash_key(planar_from_code(keys[i], bits, k)), n_parts);
  return RP_OK;
}

void rp_db_free(rp_db* db) {
  if (!db) return;
  for (auto* dc : db->dev) free_device_ctx(dc);
  for (auto& pt : db->parts) {
    if (pt.device < 0 || cudaSetDevice(pt.device) != cudaSuccess) continue;
    if (pt.ipc) { cudaIpcCloseMemHandle(pt.d_table); cudaIpcCloseMemHandle(pt.d_blocks); }
    else { cudaFree(pt.d_table); cudaFree(pt.d_blocks); cudaFree(pt.d_direct); }
  }
  delete db;
}

int rp_db_describe(const rp_db* db, rp_db_desc* out) {
  if (!db || !out) return set_error(RP_E_INVALID, "NULL argument");
  *out = db->desc;
  return RP_OK;
}

int rp_db_device_bytes(const rp_db* db, uint64_t* table_bytes, uint64_t* block_bytes) {
  if (!db) return set_error(RP_E_INVALID, "db is NULL");
  // replicated: one replica; partitioned: the sum over the partitions
  uint64_t tb = 0;
  for (auto& pt : db->parts) tb += pt.n_buckets * 32;
  if (!db->partitioned && !db->parts.empty()) tb = db->parts[0].n_buckets * 32;
  if (table_bytes) *table_bytes = tb;
  if (block_bytes) *block_bytes = db->block_bytes;
  return RP_OK;
}

double rp_last_kernel_ms(const rp_db* db) { return db ? db->last_kernel_ms.load() : 0.0; }

// ---- .rgdb : flat little-endian export of the phylo-kmer DB (written by the Java exporter, or by
// rp_db_save_file).  Layout (all sections 64 B aligned):
//   [0,64)   header: char magic[8]="RGDB\0\0\0\1"; i32 alphabet,k,n_nodes; f32 thr_log10,thr_lin; i32 0;
//                    u64 n_keys, n_postings; u64 off_keys, off_offsets, off_nodes  (off_scores follows)
//   keys u64[n_keys] | offsets u64[n_keys+1] | post_node u16[n_postings] | post_score f32[n_postings]
struct RgdbHeader {
  char magic[8];
  int32_t alphabet, k, n_nodes;
  float thr_log10, thr_lin;
  int32_t zero;
  uint64_t n_keys, n_postings;
  uint64_t off_keys, off_offsets, off_nodes, off_scores;
};
static_assert(sizeof(RgdbHeader) == 80, "header layout");
static const char kMagic[8] = {'R', 'G', 'D', 'B', 0, 0, 0, 1};
static uint64_t align64(uint64_t x) { return (x + 63) & ~63ull; }

int rp_db_save_file(const char* path, const rp_db_desc* d, const uint64_t* keys, const uint64_t* offsets,
                    const uint16_t* post_node, const float* post_score) {
  int rc = check_desc(d);
  if (rc) return rc;
  RgdbHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, kMagic, 8);
  h.alphabet = d->alphabet; h.k = d->k; h.n_nodes = d->n_nodes; h.thr_log10 = d->thr_log10; h.thr_lin = d->thr_lin;
  h.n_keys = d->n_keys; h.n_postings = d->n_postings;
  h.off_keys = align64(sizeof h);
  h.off_offsets = align64(h.off_keys + 8 * d->n_keys);
  h.off_nodes = align64(h.off_offsets + 8 * (d->n_keys + 1));
  h.off_scores = align64(h.off_nodes + 2 * d->n_postings);
  FILE* f = fopen(path, "wb");
  if (!f) return set_error(RP_E_IO, "cannot open %s for writing", path);
  auto put = [&](uint64_t off, const void* p, uint64_t n) -> bool {
    if (fseek(f, (long)off, SEEK_SET) != 0) return false;
    return n == 0 || fwrite(p, 1, n, f) == n;
  };
  static const uint64_t zero_off[1] = {0};
  bool ok = put(0, &h, sizeof h) && put(h.off_keys, keys, 8 * d->n_keys) &&
            put(h.off_offsets, d->n_keys ? offsets : zero_off, 8 * (d->n_keys + 1)) &&
            put(h.off_nodes, post_node, 2 * d->n_postings) && put(h.off_scores, post_score, 4 * d->n_postings);
  ok = (fclose(f) == 0) && ok;
  if (!ok) return set_error(RP_E_IO, "short write to %s", path);
  return RP_OK;
}

int rp_db_load_file(const char* path, const int32_t* devices, int32_t n_devices, int32_t partitioned, rp_db** out) {
  if (!path || !out) return set_error(RP_E_INVALID, "NULL argument");
  int fd = open(path, O_RDONLY);
  if (fd < 0) return set_error(RP_E_IO, "cannot open %s", path);
  struct stat st;
  if (fstat(fd, &st) != 0 || (uint64_t)st.st_size < sizeof(RgdbHeader)) {
    close(fd);
    return set_error(RP_E_IO, "%s: too short for an .rgdb header", path);
  }
  void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) return set_error(RP_E_IO, "mmap of %s failed", path);
  const uint8_t* base = (const uint8_t*)m;
  RgdbHeader h;
  memcpy(&h, base, sizeof h);
  int rc = RP_OK;
  if (memcmp(h.magic, kMagic, 8) != 0) rc = set_error(RP_E_IO, "%s: bad magic (not an .rgdb v1 file)", path);
  uint64_t need = h.off_scores + 4 * h.n_postings;
  if (!rc && (h.off_keys < sizeof h || h.off_offsets < h.off_keys + 8 * h.n_keys ||
              h.off_nodes < h.off_offsets + 8 * (h.n_keys + 1) || h.off_scores < h.off_nodes + 2 * h.n_postings ||
              need > (uint64_t)st.st_size))
    rc = set_error(RP_E_IO, "%s: section table inconsistent with file size", path);
  if (!rc) {
    rp_db_desc d;
    memset(&d, 0, sizeof d);
    d.alphabet = h.alphabet; d.k = h.k; d.n_nodes = h.n_nodes; d.thr_log10 = h.thr_log10; d.thr_lin = h.thr_lin;
    d.n_keys = h.n_keys; d.n_postings = h.n_postings;
    rc = rp_db_load(&d, (const uint64_t*)(base + h.off_keys), (const uint64_t*)(base + h.off_offsets),
                    (const uint16_t*)(base + h.off_nodes), (const float*)(base + h.off_scores), devices, n_devices,
                    partitioned, out);
  }
  munmap(m, (size_t)st.st_size);
  return rc;
}

}  // extern "C"
