/*
 * rappas_oracle.c -- CPU ORACLE of the RAPPAS placement hot path.  TEST INFRASTRUCTURE ONLY
 * (see rappas_oracle.h).  PARITY UNPINNED: no reference golden vectors exist and the Java
 * reference cannot run in this environment; pinned by hand-derived known answers only.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/src).  Java numeric semantics are kept explicit:
 *   - `float op float` is an IEEE binary32 operation (compile with -ffp-contract=off so
 *     gcc never fuses a*b+c, and without -ffast-math);
 *   - `f32 += f64` is  f32 = (float)((double)f32 + f64)  (JLS 15.26.2);
 *   - `int * float` converts the int to float first (JLS 5.6.2).
 * Math.pow / Math.log10 are taken from libm; Java allows 1 ulp on both, so the last bit of
 * an ambiguity-path score or of an LWR may differ from a JVM's.
 */
#define _GNU_SOURCE
#include "rappas_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[256];
const char* rpo_last_error(void) { return g_err; }
static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}

/* ------------------------------------------------------------------ alphabets ---- */

static int is_amino(int32_t alphabet) { return alphabet == RP_ALPHA_AMINO || alphabet == RP_ALPHA_AMINO_UO; }
static int n_states(int32_t alphabet) { return is_amino(alphabet) ? 20 : 4; }

/* DNAStatesShifted.charToByte, core/DNAStatesShifted.java:182-209 : A=0 T/U=1 C=2 G=3 */
static int nucl_state(int c) {
  switch (c) {
    case 'a': case 'A': return 0;
    case 't': case 'T': case 'u': case 'U': return 1;
    case 'c': case 'C': return 2;
    case 'g': case 'G': return 3;
    default: return -1;
  }
}

/* DNAStatesShifted ctor, core/DNAStatesShifted.java:45-96.  '.' and '-' are registered with
 * `new byte[4]` and never filled, i.e. four times state 0 (:57-58). */
static int nucl_alts(int c, uint8_t* out) {
  enum { A = 0, T = 1, C = 2, G = 3 };
  switch (c) {
    case 'R': case 'r': out[0] = A; out[1] = G; return 2;
    case 'Y': case 'y': out[0] = C; out[1] = T; return 2;
    case 'S': case 's': out[0] = C; out[1] = G; return 2;
    case 'W': case 'w': out[0] = A; out[1] = T; return 2;
    case 'K': case 'k': out[0] = G; out[1] = T; return 2;
    case 'M': case 'm': out[0] = A; out[1] = C; return 2;
    case 'B': case 'b': out[0] = C; out[1] = G; out[2] = T; return 3;
    case 'D': case 'd': out[0] = A; out[1] = G; out[2] = T; return 3;
    case 'H': case 'h': out[0] = A; out[1] = C; out[2] = T; return 3;
    case 'V': case 'v': out[0] = A; out[1] = C; out[2] = G; return 3;
    case 'N': case 'n': out[0] = A; out[1] = C; out[2] = G; out[3] = T; return 4;
    case '.': case '-': out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 0; return 4;
    default: return 0;
  }
}

/* AAStates ctor, core/AAStates.java:23-34, 74-93 (+ :118-123 with convertUO) */
static int amino_state(int c, int convert_uo) {
  static const char order[] = "RHKDESTNQCGPAILMFWYV";
  if (c >= 'a' && c <= 'z') c -= 32;
  if (convert_uo) {
    if (c == 'U') return 9;   /* U -> C */
    if (c == 'O') return 14;  /* O -> L */
  }
  for (int i = 0; i < 20; i++)
    if (order[i] == c) return i;
  return -1;
}

/* AAStates ctor, core/AAStates.java:97-112 */
static int amino_alts(int c, uint8_t* out) {
  switch (c) {
    case '-': case '*': case '!': case 'X': case 'x':
      for (int i = 0; i < 20; i++) out[i] = (uint8_t)i;
      return 20;
    case 'B': case 'b': out[0] = 3;  out[1] = 7;  return 2; /* D,N */
    case 'Z': case 'z': out[0] = 4;  out[1] = 8;  return 2; /* E,Q */
    case 'J': case 'j': out[0] = 13; out[1] = 14; return 2; /* I,L */
    default: return 0;
  }
}

int32_t rpo_ambiguity_equivalence(int32_t alphabet, int32_t c, uint8_t* out) {
  if (c < 0 || c > 255) return 0;
  return is_amino(alphabet) ? amino_alts(c, out) : nucl_alts(c, out);
}

/* isAmbiguous is tested BEFORE stateToByte (AmbigSequenceKnife.java:106, 123) */
int32_t rpo_char_class(int32_t alphabet, int32_t c) {
  uint8_t tmp[20];
  if (c < 0 || c > 255) return -2;
  if (rpo_ambiguity_equivalence(alphabet, c, tmp) > 0) return -1;
  int s = is_amino(alphabet) ? amino_state(c, alphabet == RP_ALPHA_AMINO_UO) : nucl_state(c);
  return s < 0 ? -2 : s;
}

/* AmbigSequenceKnife ctor, core/algos/AmbigSequenceKnife.java:95 */
int32_t rpo_max_ambig_per_mer(int32_t alphabet, int32_t k) {
  return (int32_t)floor(pow((double)k, 1.0 / (double)n_states(alphabet)));
}

/* Main_DBBUILD_3.java:165-166:
 *   float PPStarThreshold=(float)Math.pow((0.0+omega/s.getNonAmbiguousStatesCount()),k);
 *   float PPStarThresholdAsLog=(float)Math.log10(PPStarThreshold);                       */
void rpo_threshold(float omega, int32_t alphabet, int32_t k, float* thr_lin, float* thr_log10) {
  float ratio = omega / (float)n_states(alphabet); /* float / int -> float */
  float lin = (float)pow(0.0 + (double)ratio, (double)k);
  float lg = (float)log10((double)lin);
  if (thr_lin) *thr_lin = lin;
  if (thr_log10) *thr_log10 = lg;
}

/* nucl: DNAStatesShifted.compressMer, core/DNAStatesShifted.java:115-143 -- base i goes to
 * bits 2*(i%4) of byte i/4; read little-endian this is sum b_i * 4^i.
 * amino: AAStates.compressMer is the identity (:195-197); two byte[k] are equal iff the
 * 5-bit packing sum b_i * 32^i is equal (HashStrategy.java:26-29). */
uint64_t rpo_pack_kmer(int32_t alphabet, const uint8_t* states, int32_t k) {
  uint64_t code = 0;
  int bits = is_amino(alphabet) ? 5 : 2;
  for (int i = 0; i < k; i++) code |= (uint64_t)states[i] << (bits * i);
  return code;
}

/* ------------------------------------------------------------------ DB ----------- */

struct rpo_db {
  rp_db_desc desc;
  uint64_t* keys;
  uint64_t* offsets;
  uint16_t* post_node;
  float* post_score;
  /* open addressing: slot = key index + 1, 0 = empty */
  uint64_t* slots;
  uint64_t mask;
};

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

static int check_desc(const rp_db_desc* d) {
  if (!d) return fail(RP_E_INVALID, "desc is NULL");
  if (d->alphabet < 0 || d->alphabet > 2) return fail(RP_E_INVALID, "bad alphabet");
  int amino = is_amino(d->alphabet);
  if (d->k < 2 || d->k > (amino ? 12 : 31)) return fail(RP_E_INVALID, "k out of range");
  if (d->n_nodes < 1 || d->n_nodes > 65535) return fail(RP_E_INVALID, "n_nodes out of range");
  return RP_OK;
}

int rpo_db_load(const rp_db_desc* desc, const uint64_t* keys, const uint64_t* offsets,
                const uint16_t* post_node, const float* post_score, rpo_db** out) {
  int rc = check_desc(desc);
  if (rc) return rc;
  if (!out) return fail(RP_E_INVALID, "out is NULL");
  rpo_db* db = (rpo_db*)calloc(1, sizeof *db);
  if (!db) return fail(RP_E_NOMEM, "calloc");
  db->desc = *desc;
  uint64_t nk = desc->n_keys, np = desc->n_postings;
  db->keys = (uint64_t*)malloc((nk ? nk : 1) * sizeof(uint64_t));
  db->offsets = (uint64_t*)malloc((nk + 1) * sizeof(uint64_t));
  db->post_node = (uint16_t*)malloc((np ? np : 1) * sizeof(uint16_t));
  db->post_score = (float*)malloc((np ? np : 1) * sizeof(float));
  uint64_t cap = 16;
  while (cap < 2 * nk) cap <<= 1;
  db->slots = (uint64_t*)calloc(cap, sizeof(uint64_t));
  db->mask = cap - 1;
  if (!db->keys || !db->offsets || !db->post_node || !db->post_score || !db->slots) {
    rpo_db_free(db);
    return fail(RP_E_NOMEM, "malloc");
  }
  if (nk) memcpy(db->keys, keys, nk * sizeof(uint64_t));
  if (nk) memcpy(db->offsets, offsets, (nk + 1) * sizeof(uint64_t)); else db->offsets[0] = 0;
  if (np) memcpy(db->post_node, post_node, np * sizeof(uint16_t));
  if (np) memcpy(db->post_score, post_score, np * sizeof(float));
  if (db->offsets[nk] != np) { rpo_db_free(db); return fail(RP_E_INVALID, "offsets[n_keys] != n_postings"); }
  for (uint64_t p = 0; p < np; p++)
    if (db->post_node[p] >= (uint32_t)desc->n_nodes) { rpo_db_free(db); return fail(RP_E_INVALID, "node id >= n_nodes"); }
  for (uint64_t i = 0; i < nk; i++) {
    uint64_t h = mix64(db->keys[i]) & db->mask;
    while (db->slots[h]) {
      if (db->keys[db->slots[h] - 1] == db->keys[i]) { rpo_db_free(db); return fail(RP_E_INVALID, "duplicate key"); }
      h = (h + 1) & db->mask;
    }
    db->slots[h] = i + 1;
  }
  *out = db;
  return RP_OK;
}

void rpo_db_free(rpo_db* db) {
  if (!db) return;
  free(db->keys); free(db->offsets); free(db->post_node); free(db->post_score); free(db->slots);
  free(db);
}

/* CustomHash_v4_FastUtil81.getPairsOfTopPosition2, core/hash/CustomHash_v4_FastUtil81.java:146-153;
 * key equality = Arrays.equals on the packed bytes (HashStrategy.java:26-29) = integer equality here.
 * Returns key index or -1. */
static int64_t db_lookup(const rpo_db* db, uint64_t code) {
  uint64_t h = mix64(code) & db->mask;
  while (db->slots[h]) {
    uint64_t i = db->slots[h] - 1;
    if (db->keys[i] == code) return (int64_t)i;
    h = (h + 1) & db->mask;
  }
  return -1;
}

/* ------------------------------------------------------------------ knife -------- */

/* Per-read tables of AmbigSequenceKnife.initTables, core/algos/AmbigSequenceKnife.java:98-130 */
typedef struct {
  int32_t cap;
  int8_t* state;     /* sequence[]: state byte, -1 = ambiguous                        */
  int32_t* ambcnt;   /* ambiguityCountPerMer[]                                        */
} knife_t;

static int knife_reserve(knife_t* kn, int64_t len) {
  if (len <= kn->cap) return 0;
  int64_t cap = len + 64;
  free(kn->state); free(kn->ambcnt);
  kn->state = (int8_t*)malloc((size_t)cap);
  kn->ambcnt = (int32_t*)malloc((size_t)cap * sizeof(int32_t));
  kn->cap = (int32_t)cap;
  return (kn->state && kn->ambcnt) ? 0 : -1;
}

/* returns RP_STATUS_PLACED(0) when windows can be enumerated, else TOO_SHORT / BAD_CHAR */
static int knife_init(knife_t* kn, const rp_db_desc* d, const uint8_t* s, int64_t len) {
  int k = d->k;
  memset(kn->ambcnt, 0, (size_t)(len > 0 ? len : 0) * sizeof(int32_t));
  for (int64_t i = 0; i < len; i++) {
    int cls = rpo_char_class(d->alphabet, s[i]);
    if (cls == -1) { /* :106-119 */
      for (int64_t j = i - k + 1; j < i + 1; j++)
        if (j > -1 && j < len) kn->ambcnt[j]++;
      kn->state[i] = -1;
    } else if (cls == -2) { /* :122-128 -> System.exit(1) in the reference */
      return RP_STATUS_BAD_CHAR;
    } else {
      kn->state[i] = (int8_t)cls;
    }
  }
  /* SAMPLING_LINEAR: merOrder = new int[len-k+1], :144-150 ; negative size throws */
  if (len - k + 1 < 0) return RP_STATUS_TOO_SHORT;
  return RP_STATUS_PLACED;
}

/* AmbigSequenceKnife.getNextByteWord alternatives, :235-260.  For window j with the
 * ambiguous offsets o_1<o_2<..: n = prod |A_m| ; alternative t takes A_m[t mod |A_m|] at o_m
 * (the inner loop restarts j=0 every |A_m| words: `jump` mod |A_m|).
 * Writes n codes to out_codes (cap 4^2=16 nucl / 20 amino at maxAmbig<=2) and returns n. */
#define RPO_MAX_ALT 64
static int window_alternatives(const knife_t* kn, const rp_db_desc* d, const uint8_t* s, int64_t j,
                               uint64_t* out_codes) {
  int k = d->k;
  int n = 1;
  uint8_t alts[20];
  for (int i = 0; i < k; i++)
    if (kn->state[j + i] == -1) n *= rpo_ambiguity_equivalence(d->alphabet, s[j + i], alts);
  if (n > RPO_MAX_ALT) return -1;
  uint8_t word[32];
  for (int t = 0; t < n; t++) {
    for (int i = 0; i < k; i++) {
      if (kn->state[j + i] != -1) {
        word[i] = (uint8_t)kn->state[j + i];
      } else {
        int na = rpo_ambiguity_equivalence(d->alphabet, s[j + i], alts);
        word[i] = alts[t % na];
      }
    }
    out_codes[t] = rpo_pack_kmer(d->alphabet, word, k);
  }
  return n;
}

/* ------------------------------------------------------------------ scoring ------ */

typedef struct {
  int32_t node;
  float score;
} score_t; /* PlacementProcess.Score, PlacementProcess.java:1247-1262 */

/* java.lang.Float.compare: total order, -0 < +0, NaN (canonical) greatest */
static int32_t float_to_int_bits(float f) {
  if (f != f) return 0x7fc00000;
  int32_t b; memcpy(&b, &f, 4); return b;
}
static int float_compare(float a, float b) {
  if (a < b) return -1;
  if (a > b) return 1;
  int32_t ia = float_to_int_bits(a), ib = float_to_int_bits(b);
  return ia == ib ? 0 : (ia < ib ? -1 : 1);
}

/* java.util.PriorityQueue (JDK 8) on an array: offer = siftUpComparable, poll = siftDownComparable */
static void pq_offer(score_t* q, int* size, score_t x) {
  int k = (*size)++;
  while (k > 0) {
    int parent = (k - 1) >> 1;
    if (float_compare(x.score, q[parent].score) >= 0) break;
    q[k] = q[parent];
    k = parent;
  }
  q[k] = x;
}
static void pq_poll(score_t* q, int* size) {
  int s = --(*size);
  score_t x = q[s];
  if (s != 0) {
    int k = 0, half = s >> 1;
    while (k < half) {
      int child = 2 * k + 1, right = child + 1;
      score_t c = q[child];
      if (right < s && float_compare(c.score, q[right].score) > 0) c = q[child = right];
      if (float_compare(x.score, c.score) <= 0) break;
      q[k] = c;
      k = child;
    }
    q[k] = x;
  }
}

typedef struct {
  int32_t n_nodes;
  float* S;       /* PlacementProcess.java:496 */
  int32_t* C;     /* :495 */
  int32_t* L;     /* :493 */
  int32_t nL;
  float* S_amb;   /* :1130 / :1186 (allocated per call in the reference; zero between uses here) */
  int32_t* C_amb; /* :1131 / :1187 */
  int32_t* L_amb; /* :1132 / :1189 */
  knife_t kn;
} work_t;

static int work_init(work_t* w, int32_t n_nodes) {
  memset(w, 0, sizeof *w);
  w->n_nodes = n_nodes;
  w->S = (float*)calloc((size_t)n_nodes, sizeof(float));
  w->C = (int32_t*)calloc((size_t)n_nodes, sizeof(int32_t));
  w->L = (int32_t*)malloc((size_t)n_nodes * sizeof(int32_t));
  w->S_amb = (float*)calloc((size_t)n_nodes, sizeof(float));
  w->C_amb = (int32_t*)calloc((size_t)n_nodes, sizeof(int32_t));
  w->L_amb = (int32_t*)malloc((size_t)n_nodes * sizeof(int32_t));
  return (w->S && w->C && w->L && w->S_amb && w->C_amb && w->L_amb) ? 0 : -1;
}
static void work_free(work_t* w) {
  free(w->S); free(w->C); free(w->L); free(w->S_amb); free(w->C_amb); free(w->L_amb);
  free(w->kn.state); free(w->kn.ambcnt);
}

/* treatAmbiguitiesWithMean, PlacementProcess.java:1129-1174 ; ...WithMax, :1185-1236 */
static void treat_ambiguities(const rpo_db* db, work_t* w, const uint64_t* codes, int W_size, int Q,
                              int with_max) {
  const float T = db->desc.thr_log10, Tlin = db->desc.thr_lin;
  int nLamb = 0;
  for (int i = 0; i < W_size; i++) {
    int64_t ki = db_lookup(db, codes[i]);
    if (ki < 0) continue; /* :1145-1147 */
    for (uint64_t p = db->offsets[ki]; p < db->offsets[ki + 1]; p++) {
      int x = db->post_node[p];
      float v = db->post_score[p];
      if (with_max) { /* :1208-1218 */
        if (w->C_amb[x] == 0) { w->L_amb[nLamb++] = x; w->S_amb[x] = v; }
        w->C_amb[x] += 1;
        if (v > w->S_amb[x]) w->S_amb[x] = v;
      } else { /* :1149-1156 ; S_amb[x]+=Math.pow(10,v): f32 += f64 */
        if (w->C_amb[x] == 0) w->L_amb[nLamb++] = x;
        w->C_amb[x] += 1;
        w->S_amb[x] = (float)((double)w->S_amb[x] + pow(10.0, (double)v));
      }
    }
  }
  for (int i = 0; i < nLamb; i++) { /* :1161-1172 / :1223-1234 */
    int x = w->L_amb[i];
    if (w->C[x] == 0) {
      w->L[w->nL++] = x;
      w->S[x] = (float)Q * T; /* S[x]=Q*PPStarThresholdAsLog10 : int*float */
    }
    w->C[x] += 1;
    if (with_max) {
      float d = w->S_amb[x] - T;
      w->S[x] = w->S[x] + d;
    } else {
      /* float avgProba=(S_amb[x] + (W_size-C_amb[x])*PPStarThreshold) / W_size;   all f32 */
      float pad = (float)(W_size - w->C_amb[x]) * Tlin;
      float num = w->S_amb[x] + pad;
      float avg = num / (float)W_size;
      /* S[x]+=Math.log10(avgProba)-PPStarThresholdAsLog10;   f32 += f64 */
      w->S[x] = (float)((double)w->S[x] + (log10((double)avg) - (double)T));
    }
    w->C_amb[x] = 0;
    w->S_amb[x] = 0.0f; /* the reference allocates fresh zeroed arrays per call */
  }
}

/* The per-read body of processQueries, PlacementProcess.java:645-764: fills S, C, L.
 * counts[4] = RP_CNT_*.  Returns status (PLACED here only means "windows enumerated"). */
static int score_read(const rpo_db* db, const rp_place_cfg* cfg, work_t* w, const uint8_t* s, int64_t len,
                      int32_t* counts) {
  const rp_db_desc* d = &db->desc;
  const float T = d->thr_log10;
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  w->nL = 0;
  if (knife_reserve(&w->kn, len) != 0) return -1;
  int st = knife_init(&w->kn, d, s, len);
  if (st != RP_STATUS_PLACED) return st;
  int k = d->k;
  int Q = (int)(len - k + 1); /* sk.getMerCount() */
  int max_amb = rpo_max_ambig_per_mer(d->alphabet, k);
  uint64_t alt_codes[RPO_MAX_ALT];
  for (int j = 0; j < Q; j++) { /* merOrder[j] = j, AmbigSequenceKnife.java:144-150 */
    if (w->kn.ambcnt[j] < 1) { /* plain window, :224-226 ; PlacementProcess.java:698-735 */
      uint64_t code = rpo_pack_kmer(d->alphabet, (const uint8_t*)(w->kn.state + j), k);
      int64_t ki = db_lookup(db, code);
      counts[RP_CNT_WINDOWS]++;
      if (ki < 0) continue; /* :713-716 */
      counts[RP_CNT_MATCHED]++;
      for (uint64_t p = db->offsets[ki]; p < db->offsets[ki + 1]; p++) {
        int x = db->post_node[p];
        if (w->C[x] == 0) { /* :726-729 */
          w->L[w->nL++] = x;
          float qt = (float)Q * T;
          w->S[x] = w->S[x] + qt;
        }
        w->C[x] += 1; /* :731 */
        float dv = db->post_score[p] - T; /* :733 */
        w->S[x] = w->S[x] + dv;
      }
    } else if (w->kn.ambcnt[j] > max_amb) { /* AmbigSequenceKnife.java:230-232 ; PlacementProcess.java:691-696 */
      counts[RP_CNT_WINDOWS]++;
      counts[RP_CNT_SKIPPED]++;
    } else if (!cfg->treat_amb) { /* :745-749 */
      counts[RP_CNT_WINDOWS]++;
      counts[RP_CNT_SKIPPED]++;
    } else { /* :738-744 */
      counts[RP_CNT_AMBIG]++;
      int n = window_alternatives(&w->kn, d, s, j, alt_codes);
      if (n < 0) return -1;
      /* queryKmerMatchingDB is passed by value: matches here are NOT counted (:741-743, :1159) */
      treat_ambiguities(db, w, alt_codes, n, Q, cfg->amb_with_max);
      counts[RP_CNT_WINDOWS]++;
    }
  }
  return RP_STATUS_PLACED;
}

/* fillBestScoreList (:396-451) + the row loop of processQueries (:974-1000).
 * Writes up to K rows best-first; returns the number of rows. */
static int select_rows(const rp_place_cfg* cfg, const work_t* w, uint16_t* out_node, float* out_score,
                       double* out_lwr) {
  const int K = cfg->keep_at_most;
  score_t heap[RP_MAX_KEEP + 2];
  score_t B[RP_MAX_KEEP];
  int nb = w->nL < K ? w->nL : K; /* :828-832 */
  int hs = 0;
  for (int i = 0; i < w->nL; i++) { /* :402-409 */
    score_t e = {w->L[i], w->S[w->L[i]]};
    pq_offer(heap, &hs, e);
    if (hs > nb) pq_poll(heap, &hs);
  }
  float lowest = 0.0f, best = -3.4028234663852886e38f; /* -Float.MAX_VALUE */
  double sum = 0.0;
  for (int i = 0; i < hs; i++) { /* :415-428, heap array order */
    sum += pow(10.0, (double)heap[i].score);
    if (heap[i].score < lowest) lowest = heap[i].score;
    if (heap[i].score > best) best = heap[i].score;
  }
  for (int i = 0; i < K; i++) { B[i].node = -1; B[i].score = -INFINITY; } /* :499-501, :1071-1073 */
  for (int i = 0; i < hs; i++) B[i] = heap[i]; /* :431-435 */
  for (int i = 1; i < K; i++) { /* Arrays.sort(Object[]): stable, ascending by compareTo */
    score_t x = B[i];
    int j = i - 1;
    while (j >= 0 && float_compare(B[j].score, x.score) > 0) { B[j + 1] = B[j]; j--; }
    B[j + 1] = x;
  }
  float shift = (-308.0f >= lowest) ? best : 0.0f; /* computeWeightRatioShift :384-390 */
  if (shift != 0.0f) { /* :442-448 ; f32 subtraction then widened */
    sum = 0.0;
    for (int ii = K - nb; ii < K; ii++) {
      float df = B[ii].score - shift;
      sum += pow(10.0, (double)df);
    }
  }
  int rows = 0;
  if (B[K - 1].score >= cfg->ns_bound) { /* :974 */
    float best2 = B[K - 1].score;      /* :978 */
    float lowest2 = B[K - nb].score;   /* :979 */
    float shift2 = (-308.0f >= lowest2) ? best2 : 0.0f; /* :980 */
    double best_ratio = -1;
    for (int i = K - 1; i > K - nb - 1; i--) { /* :984 */
      /* computeWeightRatio(Score, double shift, double sum): (double)(s.score - shift) with
       * shift already a double => f64 subtraction (:392-393) */
      double lwr = pow(10.0, (double)B[i].score - (double)shift2) / sum;
      if (i == K - 1) best_ratio = lwr;
      if (i < K - 1 && lwr < (best_ratio * (double)cfg->keep_factor)) break; /* :998 */
      out_node[rows] = (uint16_t)B[i].node;
      out_score[rows] = B[i].score;
      out_lwr[rows] = lwr;
      rows++;
    }
  }
  return rows;
}

static void reset_read(work_t* w) { /* :1067-1075 */
  for (int i = 0; i < w->nL; i++) { w->S[w->L[i]] = 0.0f; w->C[w->L[i]] = 0; }
  w->nL = 0;
}

static int check_cfg(const rp_place_cfg* cfg) {
  if (!cfg) return fail(RP_E_INVALID, "cfg is NULL");
  if (cfg->keep_at_most < 1 || cfg->keep_at_most > RP_MAX_KEEP) return fail(RP_E_INVALID, "keep_at_most out of range");
  return RP_OK;
}

static int place_range(rpo_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off,
                       int64_t r0, int64_t r1, int32_t* out_n_rows, uint16_t* out_node, float* out_score,
                       double* out_lwr, int32_t* out_counts, int32_t* out_status) {
  work_t w;
  if (work_init(&w, db->desc.n_nodes) != 0) { work_free(&w); return RP_E_NOMEM; }
  const int K = cfg->keep_at_most;
  int rc = RP_OK;
  for (int64_t r = r0; r < r1; r++) {
    int32_t counts[4];
    const uint8_t* s = seq + seq_off[r];
    int64_t len = (int64_t)(seq_off[r + 1] - seq_off[r]);
    for (int i = 0; i < K; i++) {
      out_node[r * K + i] = 0xFFFF; out_score[r * K + i] = -INFINITY; out_lwr[r * K + i] = 0.0;
    }
    int st = score_read(db, cfg, &w, s, len, counts);
    if (st < 0) { rc = RP_E_NOMEM; break; }
    int rows = 0;
    if (st == RP_STATUS_PLACED) {
      if (w.nL < 1) st = RP_STATUS_UNPLACED; /* :797-806 */
      else rows = select_rows(cfg, &w, out_node + r * K, out_score + r * K, out_lwr + r * K);
    } else {
      counts[0] = counts[1] = counts[2] = counts[3] = 0;
    }
    reset_read(&w);
    out_n_rows[r] = rows;
    out_status[r] = st;
    if (out_counts) memcpy(out_counts + 4 * r, counts, sizeof counts);
  }
  work_free(&w);
  return rc;
}

int rpo_place_batch(rpo_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off,
                    int64_t n_reads, int32_t* out_n_rows, uint16_t* out_node, float* out_score,
                    double* out_lwr, int32_t* out_counts, int32_t* out_status) {
  if (!db) return fail(RP_E_INVALID, "db is NULL");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  rc = place_range(db, cfg, seq, seq_off, 0, n_reads, out_n_rows, out_node, out_score, out_lwr, out_counts,
                   out_status);
  return rc ? fail(rc, "out of memory") : RP_OK;
}

typedef struct {
  rpo_db* db; const rp_place_cfg* cfg; const uint8_t* seq; const uint64_t* seq_off; int64_t r0, r1;
  int32_t* out_n_rows; uint16_t* out_node; float* out_score; double* out_lwr; int32_t* out_counts;
  int32_t* out_status; int rc;
} mt_arg;

static void* mt_main(void* p) {
  mt_arg* a = (mt_arg*)p;
  a->rc = place_range(a->db, a->cfg, a->seq, a->seq_off, a->r0, a->r1, a->out_n_rows, a->out_node, a->out_score,
                      a->out_lwr, a->out_counts, a->out_status);
  return NULL;
}

int rpo_place_batch_mt(rpo_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off,
                       int64_t n_reads, int32_t* out_n_rows, uint16_t* out_node, float* out_score,
                       double* out_lwr, int32_t* out_counts, int32_t* out_status, int32_t n_threads) {
  if (!db) return fail(RP_E_INVALID, "db is NULL");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 1024) n_threads = 1024;
  if ((int64_t)n_threads > n_reads) n_threads = (int32_t)(n_reads > 0 ? n_reads : 1);
  pthread_t* th = (pthread_t*)malloc((size_t)n_threads * sizeof *th);
  mt_arg* args = (mt_arg*)malloc((size_t)n_threads * sizeof *args);
  if (!th || !args) { free(th); free(args); return fail(RP_E_NOMEM, "malloc"); }
  for (int t = 0; t < n_threads; t++) {
    mt_arg a = {db, cfg, seq, seq_off, n_reads * t / n_threads, n_reads * (t + 1) / n_threads,
                out_n_rows, out_node, out_score, out_lwr, out_counts, out_status, 0};
    args[t] = a;
    if (pthread_create(&th[t], NULL, mt_main, &args[t]) != 0) { mt_main(&args[t]); th[t] = 0; }
  }
  for (int t = 0; t < n_threads; t++) {
    if (th[t]) pthread_join(th[t], NULL);
    if (args[t].rc) rc = args[t].rc;
  }
  free(th); free(args);
  return rc ? fail(rc, "worker failed") : RP_OK;
}

int rpo_node_scores(rpo_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off,
                    int64_t n_reads, float* out_scores, int32_t* out_hitcount) {
  if (!db) return fail(RP_E_INVALID, "db is NULL");
  int rc = check_cfg(cfg);
  if (rc) return rc;
  work_t w;
  if (work_init(&w, db->desc.n_nodes) != 0) { work_free(&w); return fail(RP_E_NOMEM, "malloc"); }
  const int64_t N = db->desc.n_nodes;
  for (int64_t r = 0; r < n_reads; r++) {
    int32_t counts[4];
    for (int64_t x = 0; x < N; x++) {
      out_scores[r * N + x] = NAN;
      if (out_hitcount) out_hitcount[r * N + x] = 0;
    }
    int st = score_read(db, cfg, &w, seq + seq_off[r], (int64_t)(seq_off[r + 1] - seq_off[r]), counts);
    if (st < 0) { work_free(&w); return fail(RP_E_NOMEM, "malloc"); }
    for (int i = 0; i < w.nL; i++) {
      int x = w.L[i];
      out_scores[r * N + x] = w.S[x];
      if (out_hitcount) out_hitcount[r * N + x] = w.C[x];
    }
    reset_read(&w);
  }
  work_free(&w);
  return RP_OK;
}

int rpo_extract_kmers(rpo_db* db, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                      const uint64_t* win_off, uint64_t* out_code, uint8_t* out_kind, int32_t* out_nalt,
                      int32_t* out_hits, int32_t* out_status) {
  if (!db) return fail(RP_E_INVALID, "db is NULL");
  const rp_db_desc* d = &db->desc;
  knife_t kn = {0, NULL, NULL};
  int k = d->k, max_amb = rpo_max_ambig_per_mer(d->alphabet, k);
  uint64_t alt_codes[RPO_MAX_ALT];
  for (int64_t r = 0; r < n_reads; r++) {
    const uint8_t* s = seq + seq_off[r];
    int64_t len = (int64_t)(seq_off[r + 1] - seq_off[r]);
    int64_t Q = len - k + 1 > 0 ? len - k + 1 : 0;
    if ((int64_t)(win_off[r + 1] - win_off[r]) != Q) { free(kn.state); free(kn.ambcnt); return fail(RP_E_INVALID, "win_off mismatch"); }
    if (knife_reserve(&kn, len) != 0) { free(kn.state); free(kn.ambcnt); return fail(RP_E_NOMEM, "malloc"); }
    int st = knife_init(&kn, d, s, len);
    if (out_status) out_status[r] = st;
    for (int64_t j = 0; j < Q; j++) {
      uint64_t o = win_off[r] + (uint64_t)j;
      out_code[o] = ~0ULL; out_kind[o] = RP_WIN_SKIPPED; out_nalt[o] = 0; out_hits[o] = -1;
      if (st != RP_STATUS_PLACED) continue;
      if (kn.ambcnt[j] < 1) {
        uint64_t code = rpo_pack_kmer(d->alphabet, (const uint8_t*)(kn.state + j), k);
        int64_t ki = db_lookup(db, code);
        out_code[o] = code; out_kind[o] = RP_WIN_PLAIN; out_nalt[o] = 1;
        out_hits[o] = ki < 0 ? -1 : (int32_t)(db->offsets[ki + 1] - db->offsets[ki]);
      } else if (kn.ambcnt[j] <= max_amb) {
        int n = window_alternatives(&kn, d, s, j, alt_codes);
        if (n < 0) { free(kn.state); free(kn.ambcnt); return fail(RP_E_INVALID, "too many alternatives"); }
        int any = 0; int32_t tot = 0;
        for (int t = 0; t < n; t++) {
          int64_t ki = db_lookup(db, alt_codes[t]);
          if (ki >= 0) { any = 1; tot += (int32_t)(db->offsets[ki + 1] - db->offsets[ki]); }
        }
        out_code[o] = alt_codes[0]; out_kind[o] = RP_WIN_AMBIG; out_nalt[o] = n; out_hits[o] = any ? tot : -1;
      }
    }
  }
  free(kn.state); free(kn.ambcnt);
  return RP_OK;
}
