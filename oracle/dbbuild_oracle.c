/*
 * dbbuild_oracle.c -- CPU ORACLE of the RAPPAS phylo-k-mer generation.  TEST INFRASTRUCTURE ONLY
 * (same rules as rappas_oracle.h: only tests/ and bench legs labelled as CPU baselines may load it).
 *
 * PARITY UNPINNED: restated from a reading of the Java sources; the reference cannot run here.
 *
 * Follows, statement by statement:
 *   driver        main_v2/Main_DBBUILD_3.java:648-716  (for node, for pos in [0, len-k+2), for j: exploreWords(pos, j);
 *                                                        ONE WordExplorer_v3 per (node, pos): its fields survive the j loop)
 *   explorer      core/algos/WordExplorer_v3.java:98-199 (recursive; kept recursive here on purpose)
 *   registration  core/hash/CustomHash_v4_FastUtil81.java:73-90 (addTuple: the maximum per (k-mer, node))
 *   word -> code  core/DNAStatesShifted.java:115-143 (compressMer, = sum b_i 4^i), amino: sum b_i 32^i (rp_pack_kmer)
 *
 * Arithmetic: currentLogSum is a Java float; getPP returns double (PProbasSorted.java:45), so
 * `currentLogSum += pp` is (float)((double)sum + (double)pp) and the same for -=.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rappas_b200.h"

typedef struct {
  uint64_t key;
  uint16_t node;
  float score;
} Tuple;

typedef struct {
  const rp_dbbuild_desc* d;
  const float* pp;          /* [n_nodes][n_sites][n_states] */
  const uint8_t* states;
  const uint64_t* gap_off;  /* may be NULL */
  const int32_t* gap_len;
  /* WordExplorer_v3 fields (:33-52) */
  int node_id, ref_position, current_k, bound_reaching_k, idx_of_first_jump;
  int bound_reached;
  float current_log_sum;
  uint8_t word[64];
  uint16_t original_id;
  /* output */
  Tuple* out;
  uint64_t n_out, cap;
  int oom;
} Explorer;

static void emit(Explorer* e) {
  if (e->n_out == e->cap) {
    uint64_t nc = e->cap ? e->cap * 2 : 1024;
    Tuple* t = (Tuple*)realloc(e->out, nc * sizeof(Tuple));
    if (!t) { e->oom = 1; return; }
    e->out = t;
    e->cap = nc;
  }
  uint64_t code = 0;
  const int bits = e->d->alphabet == RP_ALPHA_NUCL ? 2 : 5;
  for (int i = 0; i < e->d->k; i++) code |= (uint64_t)e->word[i] << (bits * i);
  Tuple* t = &e->out[e->n_out++];
  t->key = code;
  t->node = e->original_id;
  t->score = e->current_log_sum;
}

/* WordExplorer_v3.exploreWords(int i, int j), :98-199 */
static void explore_words(Explorer* e, int i, int j) {
  const rp_dbbuild_desc* d = e->d;
  if (i > d->n_sites - 1) return;                                   /* :109-111 */
  if (e->current_k == 0) e->idx_of_first_jump = -1;                 /* :113-115 */
  const size_t at = ((size_t)e->node_id * d->n_sites + i) * d->n_states + j;
  e->word[e->current_k] = e->states[at];                            /* :117 */
  const double p = (double)e->pp[at];
  e->current_log_sum = (float)((double)e->current_log_sum + p);     /* :119 */
  e->bound_reached = e->current_log_sum < d->thr_log10;             /* :120 */
  if (e->bound_reached) e->bound_reaching_k = e->current_k;         /* :121-123 */
  if (e->current_k == d->k - 1) {                                   /* :126 */
    if (!e->bound_reached) emit(e);                                 /* :128-138 */
    e->current_log_sum = (float)((double)e->current_log_sum - p);   /* :141 */
    return;
  }
  for (int j2 = 0; j2 < d->n_states; j2++) {                        /* :147 */
    if (e->bound_reached && e->bound_reaching_k == e->current_k + 1) break;   /* :148-150 */
    e->current_k++;
    explore_words(e, i + 1, j2);                                    /* :155-157 */
    e->current_k--;
    if (d->gap_jumps && e->gap_off && i < d->n_sites - 1) {         /* :161 */
      const uint64_t g0 = e->gap_off[i + 1], g1 = e->gap_off[i + 2];
      if (g1 > g0) {                                                /* gapIntervals[i+1] != null, :163 */
        if (d->gap_jumps == 1) {                                    /* !limitTo1Jump, :165-171 */
          for (uint64_t g = g0; g < g1; g++) {
            e->current_k++;
            explore_words(e, (i + 1) + e->gap_len[g], j2);
            e->current_k--;
          }
        } else if (e->idx_of_first_jump == -1) {                    /* :174-184 */
          e->idx_of_first_jump = i;
          for (uint64_t g = g0; g < g1; g++) {
            e->current_k++;
            explore_words(e, (i + 1) + e->gap_len[g], j2);
            e->current_k--;
          }
        }
      }
    }
  }
  e->current_log_sum = (float)((double)e->current_log_sum - p);     /* :198 */
}

static int cmp_tuple(const void* a, const void* b) {
  const Tuple* x = (const Tuple*)a;
  const Tuple* y = (const Tuple*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  if (x->node != y->node) return x->node < y->node ? -1 : 1;
  return 0;
}

/* Runs the whole build; returns malloc'd CSR arrays (caller frees with rpo_dbbuild_free_arrays).
 * raw_out != NULL also returns the tuples in the reference's addTuple order (for order-sensitive checks). */
int rpo_dbbuild(const rp_dbbuild_desc* d, const float* pp, const uint8_t* states, const uint16_t* original_id,
                const uint64_t* gap_off, const int32_t* gap_len, uint64_t* n_keys, uint64_t* n_postings,
                uint64_t* n_tuples, uint64_t** keys, uint64_t** offsets, uint16_t** post_node, float** post_score) {
  if (!d || !pp || !states || !original_id || d->k < 1 || d->k > 64) return 1;
  Explorer e;
  memset(&e, 0, sizeof e);
  e.d = d; e.pp = pp; e.states = states; e.gap_off = gap_off; e.gap_len = gap_len;
  for (int node = 0; node < d->n_nodes; node++) {                   /* Main_DBBUILD_3.java:648 */
    for (int pos = 0; pos < d->n_sites - d->k + 2; pos++) {         /* :692 */
      /* new WordExplorer_v3(...), :700-706 / WordExplorer_v3.java:70-91 */
      e.node_id = node; e.ref_position = pos; e.current_k = 0; e.current_log_sum = 0.0f;
      e.bound_reached = 0; e.bound_reaching_k = -1; e.idx_of_first_jump = -1;
      e.original_id = original_id[node];
      memset(e.word, 0, sizeof e.word);
      for (int j = 0; j < d->n_states; j++) explore_words(&e, pos, j);   /* :712-714 */
      if (e.oom) { free(e.out); return 3; }
    }
  }
  *n_tuples = e.n_out;
  /* addTuple (:73-90): per (k-mer, node) the maximum survives -- putIfAbsent, then replaced only if PPStar > old */
  qsort(e.out, e.n_out, sizeof(Tuple), cmp_tuple);
  uint64_t nk = 0, np = 0;
  for (uint64_t t = 0; t < e.n_out; t++) {
    if (t == 0 || e.out[t].key != e.out[t - 1].key) nk++;
    if (t == 0 || cmp_tuple(&e.out[t], &e.out[t - 1]) != 0) np++;
  }
  *keys = (uint64_t*)malloc((nk + 1) * sizeof(uint64_t));
  *offsets = (uint64_t*)malloc((nk + 1) * sizeof(uint64_t));
  *post_node = (uint16_t*)malloc((np + 1) * sizeof(uint16_t));
  *post_score = (float*)malloc((np + 1) * sizeof(float));
  uint64_t ik = 0, ip = 0;
  for (uint64_t t = 0; t < e.n_out; t++) {
    const int new_key = t == 0 || e.out[t].key != e.out[t - 1].key;
    const int new_post = t == 0 || cmp_tuple(&e.out[t], &e.out[t - 1]) != 0;
    if (new_key) { (*keys)[ik] = e.out[t].key; (*offsets)[ik] = ip; ik++; }
    if (new_post) { (*post_node)[ip] = e.out[t].node; (*post_score)[ip] = e.out[t].score; ip++; }
    else if (e.out[t].score > (*post_score)[ip - 1]) (*post_score)[ip - 1] = e.out[t].score;
  }
  (*offsets)[nk] = np;
  *n_keys = nk;
  *n_postings = np;
  free(e.out);
  return 0;
}

void rpo_dbbuild_free_arrays(uint64_t* keys, uint64_t* offsets, uint16_t* post_node, float* post_score) {
  free(keys); free(offsets); free(post_node); free(post_score);
}
