/*
 * rappas_b200.h -- C ABI of the B200-native RAPPAS placement engine.
 *
 * This header is the drop-in boundary.  The reference (phylo42/RAPPAS, Java) has no
 * FFI of its own; the seam is the call
 *     Main_PLACEMENT_v07.java:255-257  ->  PlacementProcess.processQueries
 *     (src/core/algos/PlacementProcess.java:471-483)
 * and every entry point below names the reference code it replaces.  The ABI is plain
 * C: flat pointers + sizes, POD structs passed by pointer, no callbacks, so that it
 * binds from JNI, Panama FFM (java.lang.foreign), ctypes and cgo alike.  INTEGRATION.md
 * shows the Java-side binding.
 *
 * Two libraries export symbols declared here:
 *   librappas_b200.so   (rappas_b200/csrc, CUDA sm_100a)  -> every rp_* symbol
 *   librappas_oracle.so (oracle/, plain C, TEST ONLY)      -> the rpo_* mirror (oracle/rappas_oracle.h)
 *
 * Threading: every call is re-entrant per rp_db handle; rp_last_error() is thread-local.
 * Errors: non-zero return (RP_E_*) + rp_last_error(); the library never aborts the
 * process (the reference calls System.exit(1) from inside the loop,
 * AmbigSequenceKnife.java:124-128 -- here the read gets RP_STATUS_BAD_CHAR instead).
 */
#ifndef RAPPAS_B200_H
#define RAPPAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define RP_ABI_VERSION 1

/* return codes */
#define RP_OK            0
#define RP_E_INVALID     1   /* bad argument / unsupported parameter combination      */
#define RP_E_CUDA        2   /* CUDA runtime error (message in rp_last_error)          */
#define RP_E_NOMEM       3   /* host or device allocation failed                       */
#define RP_E_IO          4   /* file could not be read / written / parsed              */
#define RP_E_UNSUPPORTED 5   /* feature compiled out or not available on this box      */

/* alphabets (SessionNext_v2.states: DNAStatesShifted | AAStates) */
#define RP_ALPHA_NUCL      0 /* A=0 T/U=1 C=2 G=3, DNAStatesShifted.java:182-209       */
#define RP_ALPHA_AMINO     1 /* RHKDESTNQCGPAILMFWYV = 0..19, AAStates.java:23-34      */
#define RP_ALPHA_AMINO_UO  2 /* amino, DB built with --convertUO (U->C, O->L), AAStates.java:118-123 */

/* per-read status (out_status) */
#define RP_STATUS_PLACED    0 /* >=1 row emitted (or rows suppressed only by ns_bound)  */
#define RP_STATUS_UNPLACED  1 /* no k-mer hit: PlacementProcess.java:797-806            */
#define RP_STATUS_TOO_SHORT 2 /* len < k-1; the reference throws NegativeArraySizeException (AmbigSequenceKnife.java:145) */
#define RP_STATUS_BAD_CHAR  3 /* unsupported character; the reference exits(1) (AmbigSequenceKnife.java:124-128) */
#define RP_STATUS_TOO_LONG  4 /* read longer than 2^30 characters: not placed (CUDA library limit)       */

/* per-window kind (rp_extract_kmers) */
#define RP_WIN_PLAIN   0     /* no ambiguity: one k-mer                                */
#define RP_WIN_AMBIG   1     /* 1..maxAmbigPerMer ambiguities: n alternatives          */
#define RP_WIN_SKIPPED 2     /* > maxAmbigPerMer ambiguities: AmbigSequenceKnife.java:230-232 */

/* out_counts columns, PlacementProcess.java:646-649, 790 */
#define RP_CNT_WINDOWS  0    /* queryKmerCount = Q                                     */
#define RP_CNT_MATCHED  1    /* queryKmerMatchingDB (plain windows only) */
#define RP_CNT_AMBIG    2    /* ambiguousMerTreated                                    */
#define RP_CNT_SKIPPED  3    /* skippedKmer                                            */

typedef struct rp_db rp_db; /* opaque: owns the device-resident DB on every selected GPU */

/* What the hot path consumes from SessionNext_v2 (SessionNext_v2.java:43-66, 163-195) */
typedef struct rp_db_desc {
  int32_t  alphabet;    /* RP_ALPHA_*                                                   */
  int32_t  k;           /* session.k ; nucl 2..31, amino 2..12                          */
  int32_t  n_nodes;     /* session.originalTree.getNodeCount(), 1..65535 (char keys)    */
  float    thr_log10;   /* session.PPStarThresholdAsLog10  (T)                          */
  float    thr_lin;     /* session.PPStarThreshold         (T_lin)                      */
  int32_t  reserved0;
  uint64_t n_keys;      /* distinct k-mers                                              */
  uint64_t n_postings;  /* total (node, score) pairs                                    */
} rp_db_desc;

/* Placement flags, ArgumentsParser_v2.java:86-91 + PlacementProcess ctor :73-77 */
typedef struct rp_place_cfg {
  int32_t keep_at_most; /* --keep-at-most, default 7, 1..RP_MAX_KEEP                    */
  float   keep_factor;  /* --keep-factor, default 0.01f                                 */
  int32_t treat_amb;    /* !--noamb, default 1                                          */
  int32_t amb_with_max; /* --ambwithmax, default 0                                      */
  float   ns_bound;     /* --nsbound / calibrationNormScore, default -INFINITY          */
  int32_t reserved0;
} rp_place_cfg;

#define RP_MAX_KEEP 32

/* Threshold exactly as Main_DBBUILD_3.java:165-166 computes it (float/double sequence). */
void rp_threshold(float omega, int32_t alphabet, int32_t k, float* thr_lin, float* thr_log10);

/* k-mer code of k state bytes.  nucl: DNAStatesShifted.compressMer (DNAStatesShifted.java:115-143)
 * read as a little-endian integer, code = sum b_i * 4^i; amino: sum b_i * 32^i (the
 * reference keys on the raw byte[k], AAStates.java:195-197; equality is the same). */
uint64_t rp_pack_kmer(int32_t alphabet, const uint8_t* states, int32_t k);

/* ---- DB: replaces the in-JVM CustomHash_v4_FastUtil81 (CustomHash_v4_FastUtil81.java:36)
 * by a GPU-resident open-addressing table + posting blocks.  Input is the flat CSR export
 * of hash.getHash() (walk pattern: SessionNext_v2.java:250-261):
 *   keys[n_keys]        k-mer codes (rp_pack_kmer), distinct, any order
 *   offsets[n_keys+1]   CSR offsets into the posting arrays
 *   post_node/score     postings of key i at [offsets[i], offsets[i+1]), in
 *                       char2FloatEntrySet() iteration order; node ids distinct within a key
 * devices[n_devices]    CUDA ordinals; the DB is replicated on each (partitioned = 0).
 * partitioned = 1       every entry of devices[] (<= 8, an ordinal may repeat) holds ONE hash partition
 *                       of the keys and their postings (DBs > one GPU's HBM); the distinct devices
 *                       peer-map each other's memory and each places its slice of the reads, probing
 *                       and gathering from the owner's HBM over NVLink.  Same rows as replicated.
 * partitioned = 2       only the posting blocks are partitioned (the bulk of a DB); every partition's device
 *                       also holds the WHOLE table, so probes are local loads and only the posting gathers
 *                       (bulk copies, overlapped) cross NVLink.  For DBs whose table fits beside 1/n of the
 *                       postings.  Same rows again.
 */
int  rp_db_load(const rp_db_desc* desc, const uint64_t* keys, const uint64_t* offsets,
                const uint16_t* post_node, const float* post_score,
                const int32_t* devices, int32_t n_devices, int32_t partitioned, rp_db** out);
/* Same, from an .rgdb file written by the Java exporter (format: DESIGN.md). */
int  rp_db_load_file(const char* path, const int32_t* devices, int32_t n_devices,
                     int32_t partitioned, rp_db** out);
int  rp_db_save_file(const char* path, const rp_db_desc* desc, const uint64_t* keys,
                     const uint64_t* offsets, const uint16_t* post_node, const float* post_score);
/* Hash-partitioned DB with ONE PROCESS PER GPU (the torchrun deployment): rank `part` builds and uploads
 * only its partition on `device` and gets a RP_PART_BLOB_BYTES blob (CUDA IPC handles + sizes); the ranks
 * exchange the blobs (e.g. torch.distributed all_gather) and each calls rp_db_attach_partitions with all
 * n_parts blobs in partition order, which maps the other GPUs' partitions into this process.  After
 * that the handle behaves like rp_db_load(partitioned = 1) restricted to this rank's device
 * (replicate_table != 0: like partitioned = 2, this rank building the whole table and its own blocks). */
#define RP_PART_BLOB_BYTES 152
int  rp_db_load_partition(const rp_db_desc* desc, const uint64_t* keys, const uint64_t* offsets,
                          const uint16_t* post_node, const float* post_score, int32_t device, int32_t part,
                          int32_t n_parts, int32_t replicate_table, uint8_t* blob_out, rp_db** out);
int  rp_db_attach_partitions(rp_db* db, const uint8_t* blobs, int32_t n_parts);
/* owner partition (0..n_parts-1) of each k-mer code, as rp_db_load / the kernels compute it */
int  rp_partition_of_keys(int32_t alphabet, int32_t k, const uint64_t* keys, uint64_t n_keys, int32_t n_parts,
                          int32_t* out);
void rp_db_free(rp_db* db);
int  rp_db_describe(const rp_db* db, rp_db_desc* out);
/* bytes of HBM the DB occupies on one device (table + posting blocks) */
int  rp_db_device_bytes(const rp_db* db, uint64_t* table_bytes, uint64_t* block_bytes);

/* ---- synthetic DB generated ON THE DEVICE, one partition per call ("a synthetic DB too large
 * for one GPU ... each GPU generates its own partition on device from the seed with a counter-based RNG keyed by
 * code").  Every property of a key is a pure function of (seed, code) -- rappas_b200/csrc/rp_synth.h, restated
 * on the host in rappas_b200/synth_hash.py -- so the keys a read sample probes can be regenerated for the CPU
 * oracle while the whole DB exists nowhere but in the HBM of its owners.  Nucleotide, k <= 16.
 *   occupancy        fraction of the 4^k codes that are keys (CustomHash_v4_FastUtil81.java:49: ">= 75 %")
 *   plen_table       [65536] inverse CDF of the posting-list length, values in [1, n_nodes]
 * n_parts == 1: a whole DB, ready for rp_place_batch.  n_parts > 1: partition `part` of the keys
 * (owner = rp_partition_of_keys); complete it with rp_db_partition_blob + rp_db_attach_partitions (peer-memory
 * form) or hand it to rp_xchg_create (exchange form).  desc->n_keys / n_postings are outputs (rp_db_describe). */
int  rp_db_synth_partition(const rp_db_desc* desc, uint64_t seed, double occupancy, const uint16_t* plen_table,
                           int32_t device, int32_t part, int32_t n_parts, rp_db** out);
/* CUDA-IPC blob of the partition this handle holds (what rp_db_load_partition returns in blob_out) */
int  rp_db_partition_blob(rp_db* db, uint8_t* blob_out);

/* ---- EXCHANGE FORM of a hash-partitioned DB (north_star: "NCCL all-to-all of k-mer probes over NVLink, only for
 * DBs exceeding one GPU's HBM").  One rank per GPU (one process each); rank p holds partition p and
 * places its own reads: the keys of their windows travel to the owners (all-to-all), the owners answer with the
 * posting lists of the hits (all-to-all back, pipelined over sub-batches), and the home GPU runs the same fused
 * placement kernel over what it received -- rows are bit-identical to the replicated DB.  rp_xchg_place is a
 * COLLECTIVE: every rank calls it, with its own reads (possibly none).
 *   rp_xchg_unique_id   rank 0 creates the NCCL id; the host program hands it to the other ranks (any transport)
 *   rp_xchg_create      `partition` = this rank's rp_db_load_partition / rp_db_synth_partition handle (kept by the
 *                       caller, must outlive the rp_xchg); NCCL is loaded with dlopen("libnccl.so.2") (RP_NCCL_LIB)
 *   rp_xchg_create_local  all `world` ranks inside ONE process on one GPU (collectives = device copies): how a
 *                       1-GPU box tests the form; rp_xchg_place then takes `world` batches at once
 *   rp_xchg_place       n_local = 1 (NCCL) or world (local); arrays of n_local pointers, each as in rp_place_batch
 *                       (seq_off[l][0] == 0; out_counts may be NULL or hold NULL entries)
 * Transport of the posting blocks (read when the rp_xchg is created): by default the owners write them into the
 * homes' receive buffers over peer memory (CUDA IPC between the ranks' processes: all ranks on one node with peer
 * access, which is what NVSwitch gives); RP_XCHG_PUSH=0 = ncclSend / ncclRecv instead.  Tuning knobs, all optional:
 * RP_XCHG_PROBES (probes per pipelined sub-batch, 16 M), RP_XCHG_RESERVE_SMS (SMs the placement kernel leaves free,
 * 8), RP_XCHG_ONE_STREAM=1, RP_XCHG_DEBUG=1 (phase times and the device time line of rank 0 on stderr). */
#define RP_XCHG_ID_BYTES 128
typedef struct rp_xchg rp_xchg;
int  rp_xchg_unique_id(uint8_t* id_out);
int  rp_xchg_create(rp_db* partition, int32_t rank, int32_t world, const uint8_t* id, rp_xchg** out);
int  rp_xchg_create_local(rp_db** partitions, int32_t world, rp_xchg** out);
int  rp_xchg_place(rp_xchg* x, const rp_place_cfg* cfg, int32_t n_local, const uint8_t* const* seq,
                   const uint64_t* const* seq_off, const int64_t* n_reads, int32_t* const* out_n_rows,
                   uint16_t* const* out_node, float* const* out_score, double* const* out_lwr,
                   int32_t* const* out_counts, int32_t* const* out_status);
/* of the last rp_xchg_place: device time from "reads resident" to "rows ready" (max over local ranks), probes sent by
 * the local ranks, posting bytes they received, and -- as OWNERS -- the probes that hit and the postings they listed */
int  rp_xchg_stats(const rp_xchg* x, double* device_ms, uint64_t* probes, uint64_t* payload_bytes, uint64_t* hits,
                   uint64_t* postings);
void rp_xchg_free(rp_xchg* x);
/* The host-side bookkeeping of one rp_xchg_place (pure host code, no GPU: exported so that the N > 1 logic can be
 * tested on a CPU).  probes[(w * n_sub + j) * world + o] = probes of rank w's sub-batch j owned by o;
 * units[(o * world + p) * n_sub + j] = 32 B units owner o sends home p for sub-batch j (NULL: keys only).  Outputs for
 * `rank` (any may be NULL): key_send_off / key_recv_off [world + 1], seg_first [world * n_sub + 1],
 * home_first [world * n_sub], pay_* [n_sub * world] (32 B units), caps = {send_cap, recv_cap, recv_total}. */
int  rp_xchg_plan(int32_t world, int32_t n_sub, int32_t rank, int32_t direct_local, const uint64_t* probes,
                  const uint64_t* units, uint64_t* key_send_off, uint64_t* key_recv_off, uint64_t* seg_first,
                  uint64_t* home_first, uint64_t* pay_send_off, uint64_t* pay_send_cnt, uint64_t* pay_recv_off,
                  uint64_t* pay_recv_cnt, uint64_t* caps);

/* ---- placement of one batch of reads: replaces the per-read body of
 * PlacementProcess.processQueries (PlacementProcess.java:645-838) and the LWR / keep-factor
 * loop (PlacementProcess.java:974-1000).
 *   seq, seq_off[n_reads+1]   concatenated ASCII reads exactly as Fasta.getSequence(false)
 *                             returns them (gaps kept, Main_PLACEMENT_v07.java:195)
 * outputs (caller-allocated, host memory):
 *   out_n_rows[n]             rows emitted for the read (0 if unplaced / below ns_bound)
 *   out_node/score/lwr        [n][keep_at_most], best first; unused slots: 0xFFFF / -inf / 0
 *   out_counts[n][4]          RP_CNT_* (may be NULL)
 *   out_status[n]             RP_STATUS_*
 * Reads are sharded over the DB's devices in contiguous slices; no inter-GPU traffic
 * when the DB is replicated.
 */
int  rp_place_batch(rp_db* db, const rp_place_cfg* cfg,
                    const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                    int32_t* out_n_rows, uint16_t* out_node, float* out_score, double* out_lwr,
                    int32_t* out_counts, int32_t* out_status);

/* Host memory for rp_place_batch.  The call pipelines H2D / kernel / D2H over chunks of the batch; that needs
 * page-locked host memory.  Buffers from rp_host_alloc (or registered with rp_host_register: e.g. the address of a
 * direct ByteBuffer / a MemorySegment) are copied from and to directly.  Any other (pageable) buffer is detected
 * (cudaPointerGetAttributes) and staged through the library's own pinned ring by the calling thread and a few
 * helpers -- correct, and overlapped with the GPU, but it costs the host one extra pass over the bytes. */
int  rp_host_alloc(void** out, uint64_t bytes);
void rp_host_free(void* p);
int  rp_host_register(void* p, uint64_t bytes);
int  rp_host_unregister(void* p);

/* Same work with every buffer already resident on devices[device_index] and the kernels
 * enqueued on `stream` (a cudaStream_t; NULL = default stream).  Asynchronous: returns
 * after launch.  This is what bench.py times as the kernel-only number. */
int  rp_place_batch_device(rp_db* db, int32_t device_index, const rp_place_cfg* cfg,
                           const uint8_t* d_seq, const uint64_t* d_seq_off, int64_t n_reads,
                           int32_t* d_out_n_rows, uint16_t* d_out_node, float* d_out_score,
                           double* d_out_lwr, int32_t* d_out_counts, int32_t* d_out_status,
                           void* stream);

/* ---- parity diagnostics for k-mer extraction + lookup (K1 + K2):
 * AmbigSequenceKnife.initTables/getNextByteWord (AmbigSequenceKnife.java:98-174, 209-272),
 * DNAStatesShifted.compressMer, CustomHash_v4_FastUtil81.getPairsOfTopPosition2 (:146-153).
 * win_off[n_reads+1] = prefix sum of max(len-k+1, 0).  Per window:
 *   out_code    code of the plain window / of alternative 0 of an ambiguous window; ~0 if skipped
 *   out_kind    RP_WIN_*
 *   out_nalt    1, n (= product of alternative counts), or 0 if skipped
 *   out_hits    postings found: plain -> len of the posting list or -1 on a miss;
 *               ambiguous -> sum over alternatives found, -1 if none matched; skipped -> -1
 * Windows of reads with status != 0 are filled with (~0, SKIPPED, 0, -1).
 */
int  rp_extract_kmers(rp_db* db, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                      const uint64_t* win_off, uint64_t* out_code, uint8_t* out_kind,
                      int32_t* out_nalt, int32_t* out_hits, int32_t* out_status);

/* The same two stages AS THE PLACEMENT KERNEL'S PRODUCER COMPUTES THEM (rp_extract_kmers is a separate, simpler
 * kernel): one placement launch that also records, per window, the k-mer code it built from the ballots and what
 * its probe found.  out_code: code of a plain window, ~0 otherwise; out_hits: postings found (-1 = not in the DB),
 * -2 = skipped window, -3 = ambiguous window (treated by the alternatives path), 0xFCFCFCFC = not visited (the
 * read was cut short by an unsupported character).  Diagnostic: device 0, one launch, no pipelining. */
int  rp_place_windows(rp_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off, int64_t n_reads,
                      const uint64_t* win_off, uint64_t* out_code, int32_t* out_hits);

/* ---- parity diagnostic for the scoring stage (K3): the full per-node vector S[] of every
 * read after its last window (PlacementProcess.java:719-735, 1161-1172, 1223-1234).
 *   out_scores[n_reads][n_nodes]   NaN where the node was never touched (C[x]==0)
 *   out_hitcount                   C[x]; the CUDA library does not keep C[] (only first-touch
 *                                  matters to S) and requires NULL here; the oracle fills it.
 */
int  rp_node_scores(rp_db* db, const rp_place_cfg* cfg, const uint8_t* seq, const uint64_t* seq_off,
                    int64_t n_reads, float* out_scores, int32_t* out_hitcount);

/* ---- host side either end of the hot path (plain C++, no GPU work) -----------
 * rp_reads = a parsed query file:
 *   records   every FASTA record in file order: header = first line without '>', sequence = the other
 *             lines joined and trimmed, gaps kept; empty and '#' lines skipped
 *             (FASTAPointer.nextSequenceAsFastaObject, inputs/FASTAPointer.java:67-149)
 *   unique    the distinct exact sequences in order of first appearance: what is sent to
 *             rp_place_batch (a sequence is placed once however often it occurs)
 *   group_of  id of the record's sequence with '-' removed = the reference's duplicate key
 *             (MD5 of getSequence(true), core/algos/PlacementProcess.java:591-595)
 */
typedef struct rp_reads rp_reads;
int  rp_reads_load_fasta(const char* path, rp_reads** out);
int  rp_reads_from_memory(const uint8_t* text, uint64_t n_bytes, rp_reads** out);
void rp_reads_free(rp_reads* r);
int  rp_reads_describe(const rp_reads* r, uint64_t* n_records, uint64_t* n_unique, uint64_t* n_groups);
/* pointers stay valid until rp_reads_free; seq_off has n_unique+1 entries, hdr_off n_records+1 */
int  rp_reads_unique(const rp_reads* r, const uint8_t** seq, const uint64_t** seq_off);
int  rp_reads_records(const rp_reads* r, const uint8_t** hdr, const uint64_t** hdr_off,
                      const uint32_t** unique_of, const uint32_t** group_of);

/* .jplace writer: the row assembly of PlacementProcess.java:974-1047 (column order :1005-1024, "nm" of
 * duplicates :596-624, registration of a placement only once it has rows :1047) and the file shell of
 * Main_PLACEMENT_v07.java:270-315, without json-simple and the seven whole-document regex passes.
 * Results are those of rp_place_batch over rp_reads_unique (row-major [n_unique][keep_at_most]).
 *   edge_id / branch_len [n_nodes]   PhyloNode.getJplaceEdgeId() / getBranchLengthToAncestor() per node id
 *   not_placed_path                  optional: headers of reads without any k-mer hit (:797-806)
 * Numbers are printed as Java's Float.toString / Double.toString would (what json-simple emits);
 * the document equals the reference's as JSON (key order and whitespace are not reproduced).
 * Fails if a record has status BAD_CHAR / TOO_SHORT / TOO_LONG: the reference aborts on those. */
int  rp_jplace_write(const char* path, const rp_reads* r, int32_t keep_at_most, const int32_t* n_rows,
                     const uint16_t* node, const float* score, const double* lwr, const int32_t* status,
                     const int32_t* edge_id, const float* branch_len, int32_t n_nodes, const char* tree_newick,
                     const char* invocation, int32_t guppy_compat, const char* not_placed_path,
                     uint64_t* n_placements);
/* Float.toString (as_float != 0) / Double.toString of v, as the writer prints numbers (test hook) */
int  rp_java_number(double v, int32_t as_float, char* out, int32_t cap);

/* ---- DB build: the phylo-k-mer generation of Main_DBBUILD_3.java:648-750.
 * For every tested ancestral node and every alignment position, WordExplorer_v3.exploreWords
 * (core/algos/WordExplorer_v3.java:98-199) walks the k-mers whose summed log10 posterior stays above the
 * threshold, states in descending-probability order, and registers (k-mer, node, log10 PP*) with
 * CustomHash_v4_FastUtil81.addTuple (core/hash/CustomHash_v4_FastUtil81.java:73-90: the maximum per
 * (k-mer, node)).  One GPU thread runs one (node, position) explorer, in the reference's own visiting
 * order -- the running sum is an f32 that is added to and subtracted from along the walk, so the order is
 * part of the result; the tuples are then sorted and max-reduced on the device.  The output is the CSR
 * form rp_db_load takes (keys ascending, a key's postings by ascending node id).
 *   pp, states [n_nodes][n_sites][n_states]   PProbasSorted (core/PProbasSorted.java:19-20): log10 posteriors
 *                                             in descending order per site, and the state of each
 *   original_id [n_nodes]                     extendedTree.getFakeToOriginalId(nodeMapping.get(nodeId))
 *   gap_off [n_sites + 1], gap_len            Alignment.getGapIntervals(): the jump lengths registered at a
 *                                             site (CSR; an empty range = null); NULL = no gap jumps
 *   gap_jumps                                 0 off, 1 every jump combination, 2 at most one jump per k-mer
 *                                             (doGapJumps / limitTo1Jump)                                   */
typedef struct {
  int32_t alphabet;   /* RP_ALPHA_* */
  int32_t k;
  int32_t n_nodes;    /* ancestral nodes tested */
  int32_t n_sites;    /* alignment columns */
  int32_t n_states;   /* 4 or 20 */
  float   thr_log10;  /* PPStarThresholdAsLog10 */
  int32_t gap_jumps;
  int32_t reserved0;
} rp_dbbuild_desc;
typedef struct rp_dbbuild rp_dbbuild;
/* Raw ancestral-reconstruction posteriors -> the PProbasSorted arrays rp_dbbuild_run takes.  Per (node, site):
 * clamp at site_pp_threshold (Float.MIN_VALUE in Main_DBBUILD_3.java:164), log10 in double narrowed to float
 * when as_log10 != 0, then a stable sort, highest first (SiteProba.compareTo, core/SiteProba.java:26-34, under
 * Collections.sort: equal values keep the column order) -- inputs/PHYMLWrapper.java:206-229 and the same block
 * of the RAxML-ng / PAML wrappers.  probs [n_nodes][n_sites][n_states] in the column order of the AR file,
 * state_of_column[n_states] = the state byte of each column.  Host code (multi-threaded), no GPU needed. */
int  rp_pp_prepare(const float* probs, const uint8_t* state_of_column, int32_t n_nodes, int32_t n_sites,
                   int32_t n_states, float site_pp_threshold, int32_t as_log10, float* pp_out, uint8_t* states_out);
/* Alignment.updateGapIntervals (alignement/Alignment.java:231-260): for every column at which a run of '-' starts
 * in some row and ends before the end of that row, the distinct run lengths in order of first appearance (row
 * by row) -- the jump lengths rp_dbbuild_run explores, in the order it explores them.  chars [n_rows][n_cols];
 * gap_off [n_cols + 1]; gap_len may be NULL to query *n_len first (RP_E_INVALID if cap is too small). */
int  rp_gap_intervals(const uint8_t* chars, int32_t n_rows, int32_t n_cols, uint64_t* gap_off, int32_t* gap_len,
                      uint64_t cap, uint64_t* n_len);
int  rp_dbbuild_run(const rp_dbbuild_desc* desc, const float* pp, const uint8_t* states, const uint16_t* original_id,
                    const uint64_t* gap_off, const int32_t* gap_len, int32_t device, rp_dbbuild** out);
/* arrays owned by the handle; n_tuples = addTuple calls (the reference's "Tuples explored"), kernel_ms = device time */
int  rp_dbbuild_result(const rp_dbbuild* b, uint64_t* n_keys, uint64_t* n_postings, uint64_t* n_tuples,
                       const uint64_t** keys, const uint64_t** offsets, const uint16_t** post_node,
                       const float** post_score, double* kernel_ms);
void rp_dbbuild_free(rp_dbbuild* b);

/* ---- introspection used by bench.py / tests */
int  rp_device_count(void);
/* number of kernels this library launched since load (all threads); bench.py's gpu_launches */
uint64_t rp_kernel_launch_count(void);
/* device-event time (ms) of the placement kernel of the last rp_place_batch call, max over devices */
double rp_last_kernel_ms(const rp_db* db);
const char* rp_version(void);
const char* rp_last_error(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RAPPAS_B200_H */
