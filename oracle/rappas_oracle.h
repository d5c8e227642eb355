/*
 * rappas_oracle.h -- CPU ORACLE of the RAPPAS placement hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker / the CPU baseline.  The product
 * (rappas_b200/) never links, imports or falls back to it.
 *
 * PARITY UNPINNED: the reference (phylo42/RAPPAS v1.22, Java) ships no tests, golden
 * vectors or sample data and cannot be compiled or run here (no JDK,
 * lib/fastutil-8.2.2.jar absent).  This file restates the algorithm from a reading of the
 * Java sources; it is pinned only by hand-derived known answers (tests/test_oracle_known_answers.py)
 * and by an independent pure-Python restatement (oracle/oracle_py.py) written from the
 * same sources.
 *
 * The rpo_* functions mirror the rp_* C ABI of include/rappas_b200.h one to one.
 */
#ifndef RAPPAS_ORACLE_H
#define RAPPAS_ORACLE_H

#include <stdint.h>
#include "../include/rappas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rpo_db rpo_db;

void     rpo_threshold(float omega, int32_t alphabet, int32_t k, float* thr_lin, float* thr_log10);
uint64_t rpo_pack_kmer(int32_t alphabet, const uint8_t* states, int32_t k);
int32_t  rpo_max_ambig_per_mer(int32_t alphabet, int32_t k);
/* char -> class: 0..19 state byte; -1 ambiguous; -2 unsupported */
int32_t  rpo_char_class(int32_t alphabet, int32_t c);
/* alternatives of an ambiguous char in the reference's array order; returns count (0 if not ambiguous) */
int32_t  rpo_ambiguity_equivalence(int32_t alphabet, int32_t c, uint8_t* out_states /*[20]*/);

int  rpo_db_load(const rp_db_desc* desc, const uint64_t* keys, const uint64_t* offsets,
                 const uint16_t* post_node, const float* post_score, rpo_db** out);
void rpo_db_free(rpo_db* db);

int  rpo_place_batch(rpo_db* db, const rp_place_cfg* cfg,
                     const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                     int32_t* out_n_rows, uint16_t* out_node, float* out_score, double* out_lwr,
                     int32_t* out_counts, int32_t* out_status);
/* the same, reads sharded over n_threads pthreads (reads are independent: S,C,L are reset
 * per read, PlacementProcess.java:1067-1075).  The reference itself is single-threaded. */
int  rpo_place_batch_mt(rpo_db* db, const rp_place_cfg* cfg,
                        const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                        int32_t* out_n_rows, uint16_t* out_node, float* out_score, double* out_lwr,
                        int32_t* out_counts, int32_t* out_status, int32_t n_threads);

int  rpo_extract_kmers(rpo_db* db, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                       const uint64_t* win_off, uint64_t* out_code, uint8_t* out_kind,
                       int32_t* out_nalt, int32_t* out_hits, int32_t* out_status);

/* full S[] of every read after its last window: out_scores[n_reads][n_nodes], NaN where C[x]==0;
 * out_hitcount[n_reads][n_nodes] = C[x] (may be NULL) */
int  rpo_node_scores(rpo_db* db, const rp_place_cfg* cfg,
                     const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                     float* out_scores, int32_t* out_hitcount);

const char* rpo_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
