/* gpusim_b200.h — C ABI of the B200-native brute-force Tanimoto scan + top-k engine.
 *
 * This is the drop-in boundary for gpusimilarity's one hot path: everything that
 * `gpusim::FingerprintDB` (reference fingerprintdb_cuda.h:53-147) asks of CUDA goes through
 * these entry points.  Plain pointers and sizes only; no Qt, Thrust or torch types.
 * Strings (SMILES / ids, the dbkey gate) never cross this boundary: the engine speaks
 * GLOBAL ROW NUMBERS and the C++ adapter (include/gpusim/fingerprintdb_cuda.h) or the
 * Python mirror (gpusimilarity_b200/fingerprintdb.py) maps rows back to strings.
 *
 * Result order is canonical: score descending, then global row ascending — what the
 * reference's stable per-chunk sort produces (fingerprintdb_cuda.cu:245,280-282).
 * Scores are bit-identical to `float(common) / float(total - common)` with an IEEE
 * round-to-nearest divide (fingerprintdb_cuda.cu:100-101); `score >= cutoff ? score : 0`
 * (:102) makes 0/0 rows score 0.
 *
 * Every function returns GSB_OK (0) or an error code; gsb_last_error() gives the message
 * (thread-local).  There is NO CPU fallback behind the GPU entry points: without a usable
 * CUDA device they fail with GSB_ERR_CUDA.
 */
#ifndef GPUSIM_B200_H
#define GPUSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSB_OK 0
#define GSB_ERR_INVALID 1   /* bad argument (std::invalid_argument in the adapter)        */
#define GSB_ERR_CUDA 2      /* CUDA runtime / device failure (reference: thrust::system_error) */
#define GSB_ERR_NOMEM 3     /* no device with enough memory (reference .cu:65-66)          */
#define GSB_ERR_STATE 4     /* call out of order, e.g. search before upload                */
#define GSB_ERR_CORRUPT 5   /* row count does not match the data (reference .cu:153-156)   */
#define GSB_ERR_IO 6        /* .fsim file could not be read / wrong version (gpusim.cpp:186-189) */

#define GSB_MAX_WORDS 128   /* widest fingerprint: 4096 bits */

/* similarity metrics (gsb_db_set_metric); every kernel shares the scan, only the epilogue differs */
#define GSB_METRIC_TANIMOTO 0 /* c / (pq + pd - c)   reference fingerprintdb_cuda.cu:89-103 (default) */
#define GSB_METRIC_DICE 1     /* 2c / (pq + pd)                                                      */
#define GSB_METRIC_TVERSKY 2  /* c / (alpha (pq - c) + beta (pd - c) + c), f32, every op rounded      */

/* how gsb_db_search_batch serves a batch (gsb_db_batch_mode) */
#define GSB_BATCH_LOOPED 0    /* one single-query scan per query                                      */
#define GSB_BATCH_POPC 1      /* up to 256 queries per pass over the database (POPC kernel)           */
#define GSB_BATCH_SLICED 2    /* up to 1024 queries per pass (bit-sliced kernel)                      */
#define GSB_BATCH_TENSOR 3    /* 128 queries per pass on the tensor cores (tcgen05), 1024 per call    */

/* *n of a device-side result when the launch failed without killing the context (a grid barrier
 * or peer flag timed out): the host-buffer entry points turn it into GSB_ERR_CUDA */
#define GSB_COUNT_ERROR 0xffffffffu

typedef struct gsb_db gsb_db;

/* ---- library / devices -------------------------------------------------------------- */
const char* gsb_last_error(void);
const char* gsb_version(void);
/* reference get_gpu_count(), fingerprintdb_cuda.cu:41-52 */
int gsb_device_count(void);
/* reference get_gpu_free_memory(), fingerprintdb_cuda.cu:33-39 (0 on error) */
uint64_t gsb_device_free_bytes(int device);
/* reference get_available_gpu_memory(), fingerprintdb_cuda.cu:401-413 */
uint64_t gsb_available_device_bytes(void);
/* reference get_next_gpu(), fingerprintdb_cuda.cu:54-68: round-robin over devices, first one
 * with more than required_bytes free (checks the device it returns — the reference checks a
 * different one); GSB_ERR_NOMEM if none. */
int gsb_next_device(uint64_t required_bytes, int* device);
/* HBM footprint of `rows` fingerprints of fp_bits bits uploaded with fold_factor: rows are padded to
 * a power-of-two word count and carry a 2-byte popcount each (DESIGN.md "HBM layout").  What the
 * fold-factor policy of the server sizes against (the reference compares raw bytes, gpusim.cpp:131-141). */
uint64_t gsb_layout_bytes(int fp_bits, uint64_t rows, unsigned fold_factor);
/* Last resort after a sticky CUDA error: cudaDeviceReset on every visible device.  Device state of
 * EVERY gsb_db dies with it; each database must be put up again with gsb_db_upload (its host rows are
 * untouched) before the next search.  gsb_server_recover does both for a server's databases. */
int gsb_devices_reset(void);

/* ---- database life cycle ------------------------------------------------------------ */
/* reference FingerprintDB::FingerprintDB + FingerprintDBStorage ctor, .cu:117-166.
 * chunk i holds chunk_bytes[i] bytes of packed rows (fp_bits/8 bytes each, whole rows).
 * The bytes are copied (the reference copies them too, .cu:123-125).  fp_bits % 32 == 0.
 * GSB_ERR_CORRUPT when the chunks do not add up to fp_count rows. */
int gsb_db_create(const void* const* chunk_ptrs, const uint64_t* chunk_bytes, int n_chunks,
                  int fp_bits, uint64_t fp_count, gsb_db** out);
/* A synthetic shard generated directly in device memory (benchmarks at sizes no host file
 * can hold): rows [row_base, row_base + n_rows) of the counter-based database described in
 * DESIGN.md ("Synthetic data"; host twin: oracle/oracle.py synth_rows).  Already uploaded. */
int gsb_db_create_synthetic(int device, int fp_bits, uint64_t n_rows, uint64_t row_base,
                            uint64_t seed, uint32_t plant_period, gsb_db** out);
/* The same rows split into contiguous, equal shards over `devices` (one process driving several
 * GPUs, the reference's own mode, .cu:176-183); a device may be listed more than once. */
int gsb_db_create_synthetic_sharded(const int* devices, int n_devices, int fp_bits, uint64_t n_rows,
                                    uint64_t row_base, uint64_t seed, uint32_t plant_period, gsb_db** out);
/* reference FingerprintDB::copyToGPU, .cu:168-195.  Rows are split into contiguous, equal
 * shards over `devices` (NULL / 0 = every visible device that is needed).  fold_factor is
 * bumped to the next divisor of the word count (.cu:170-173); with fold_factor > 1 the
 * folded rows are what is uploaded and searched (re-scored with the full rows kept on the host). */
int gsb_db_upload(gsb_db* db, const int* devices, int n_devices, unsigned fold_factor);
void gsb_db_destroy(gsb_db* db);
/* Similarity metric of every later search of this database (GSB_METRIC_*; alpha / beta are the
 * Tversky weights, ignored otherwise).  The reference scores Tanimoto only; SURVEY §8 f4. */
int gsb_db_set_metric(gsb_db* db, int metric, float alpha, float beta);

uint64_t gsb_db_count(const gsb_db* db);          /* FingerprintDB::count(), .h:74            */
int gsb_db_fp_bits(const gsb_db* db);             /* getFingerprintBitcount(), .h:124-127      */
uint64_t gsb_db_data_bytes(const gsb_db* db);     /* getFingerprintDataSize(), .h:123          */
unsigned gsb_db_fold_factor(const gsb_db* db);    /* effective factor after upload             */
int gsb_db_shard_count(const gsb_db* db);
/* reference FingerprintDB::getFingerprint, .cu:212-226 (without its chunk-boundary
 * off-by-one).  out_words has fp_bits/32 entries. */
int gsb_db_get_fingerprint(const gsb_db* db, uint64_t row, int32_t* out_words);

/* ---- search ------------------------------------------------------------------------- */
/* reference FingerprintDB::search, .cu:341-381 (+ search_storage :228-339).  HOST buffers in
 * and out; one call = query upload, one fused scan+select launch per shard, merge, results
 * back.  out_rows/out_scores hold k entries; *out_n = min(k, survivors) are written.
 * *out_approx = rows with score >= cutoff when cutoff > 0, else the row count (:265-277). */
int gsb_db_search(const gsb_db* db, const int32_t* query_words, int n_words, uint32_t k,
                  float cutoff, uint32_t* out_rows, float* out_scores, uint32_t* out_n,
                  uint64_t* out_approx);
/* The same search split in two, so that a caller can keep several queries in flight (up to 4 per
 * database): _async queues the launch(es) and returns a ticket at once, _wait blocks until that
 * query's results are on the host and hands them out.  The query travels as a kernel parameter and
 * the last CTA of the launch stores the results straight into mapped pinned host memory, so there
 * is no cudaMemcpy and no stream synchronize on the path; consecutive queries overlap on the device
 * (programmatic dependent launch: the scan of query i+1 starts while the last CTA of query i is
 * still sorting).  gsb_db_search is _async followed by _wait. */
int gsb_db_search_async(const gsb_db* db, const int32_t* query_words, int n_words, uint32_t k,
                        float cutoff, uint64_t* ticket);
int gsb_db_search_wait(const gsb_db* db, uint64_t ticket, uint32_t* out_rows, float* out_scores,
                       uint32_t* out_n, uint64_t* out_approx);
/* New (the reference serves one query per request, gpusim.cpp:407-414): n_queries queries
 * over the same database; query q's results land at out_rows + q*k etc.  Identical results to
 * n_queries calls of gsb_db_search.  With the default layout, rows of at most 1024 bits, no fold
 * and k <= 512 the queries share ONE pass over the database per group: 1024 queries per pass
 * with the bit-sliced kernel (6 or more queries; gsb_sliced.cuh), 256 with the POPC kernel.
 * Folded databases take the same path as long as k * F * floor(log2 2F) <= 512 candidates per
 * query (second stage on the device).  Otherwise the queries are searched one after the other;
 * gsb_db_batch_mode says beforehand which of the three it will be (and a batch of 8 or more
 * queries that ends up looping says so once on stderr).
 * GSB_BATCH_KERNEL=0/2/3 forces looping / the POPC kernel / the bit-sliced kernel. */
int gsb_db_search_batch(const gsb_db* db, const int32_t* query_words, int n_words,
                        int n_queries, uint32_t k, float cutoff, uint32_t* out_rows,
                        float* out_scores, uint32_t* out_n, uint64_t* out_approx);
/* *mode = GSB_BATCH_* that gsb_db_search_batch would use for this k / batch size / cutoff,
 * *queries_per_pass (may be NULL) the queries sharing one pass over the database. */
int gsb_db_batch_mode(const gsb_db* db, uint32_t k, int n_queries, float cutoff, int* mode,
                      uint32_t* queries_per_pass);
/* reference FingerprintDB::search_cpu, fingerprintdb_cuda.cpp:20-54: host threads, CPU
 * scores (no cutoff, 0/0 = NaN), first k in stable score-descending order.  A separate
 * entry point of the reference API — never used as a fallback by gsb_db_search.
 * Searches every chunk and clamps k to the row count (the reference does neither). */
int gsb_db_search_cpu(const gsb_db* db, const int32_t* query_words, int n_words, uint32_t k,
                      uint32_t* out_rows, float* out_scores, uint32_t* out_n);

/* ---- device-resident entry points (one process per GPU; torch.distributed plumbing) ---- */
/* Packed candidate: (float bits of score << 32) | (0xFFFFFFFF - global row).  Larger key =
 * better in canonical order, so one unsigned compare orders candidates. */
typedef uint64_t gsb_key;

/* All searches of one gsb_db share its per-shard workspace.  Launches on ONE stream are ordered by
 * the stream (and overlap head-to-tail through programmatic dependent launch); a launch that
 * arrives on ANOTHER stream than the work still in flight is made to wait for it with an event, so
 * callers need no synchronisation of their own.  A stream handed to these entry points must stay
 * valid until the next search of the database (or its destruction).
 *
 * Asynchronous search of this process's single shard on `stream` (a cudaStream_t; NULL = the
 * legacy default stream).  d_query: fp_bits/32 words in device memory.  d_out_keys: k keys,
 * sorted best first, unused tail zero-filled; d_out_n: entries written; d_out_survivors: this
 * shard's approximate count.  All three live in device memory.  One launch, no host sync. */
int gsb_db_search_device(const gsb_db* db, void* stream, const int32_t* d_query, uint32_t k,
                         float cutoff, gsb_key* d_out_keys, uint32_t* d_out_n,
                         uint64_t* d_out_survivors);
/* Merge n_lists sorted candidate lists (list i = d_keys + i*list_stride, d_counts[i] valid
 * entries; d_counts may be NULL = every list is full with zero keys as padding) into the global
 * top-k.  Used after the all-gather of per-shard candidates.  Asynchronous on `stream`. */
int gsb_merge_device(int device, void* stream, const gsb_key* d_keys, const uint32_t* d_counts,
                     int n_lists, uint32_t list_stride, uint32_t k, uint32_t* d_out_rows,
                     float* d_out_scores, uint32_t* d_out_n);

/* Multi-query scan of this process's shard: n_queries queries in device memory
 * ([n_queries][fp_bits/32] words) are scored against every row in ONE pass over the database.
 * d_out_keys [n_queries][k] (best first, zero padded), d_out_n [n_queries], d_out_survivors
 * [n_queries].  Needs the default layout, fingerprints of at most 1024 bits and k <= 512.
 * n_queries is limited to gsb_db_batch_max_queries() per call. */
int gsb_db_search_batch_device(const gsb_db* db, void* stream, const int32_t* d_queries, int n_queries,
                               uint32_t k, float cutoff, gsb_key* d_out_keys, uint32_t* d_out_n,
                               uint64_t* d_out_survivors);
/* Queries one gsb_db_search_batch_device call accepts for this k / batch size / cutoff: 1024
 * where the bit-sliced kernel applies (see gsb_db_search_batch), else 256.  The database must
 * have been uploaded (GSB_ERR_STATE otherwise: the device layout is chosen there). */
int gsb_db_batch_max_queries(const gsb_db* db, uint32_t k, int n_queries, float cutoff, uint32_t* out_max);
/* Merge of all-gathered per-rank batch records (per rank: [n_queries][k] keys, [n_queries]
 * survivors, [n_queries] counts, u64 each): one CTA per query. */
int gsb_merge_batch_device(int device, void* stream, const gsb_key* d_records, int n_ranks, int n_queries,
                           uint32_t k, uint32_t* d_out_rows, float* d_out_scores, uint32_t* d_out_n,
                           uint64_t* d_out_approx);

/* Fused scan + cross-GPU exchange + merge in ONE launch per rank: the last CTA of every rank's
 * scan stores its shard's candidate record straight into every rank's exchange buffer through
 * NVLink peer-mapped memory, raises an arrival flag there, waits for all ranks' flags in its own
 * buffer and merges.  Every rank ends with the global top-k (rows, scores, count, approximate
 * count) in its own device buffers; no NCCL call and no second launch on the query path.
 * peer_base[r] is rank r's exchange buffer as mapped into THIS process (e.g. torch symmetric
 * memory `buffer_ptrs`), gsb_exchange_bytes() large and zero-filled before the first query.
 * seq is the query number: 1, 2, 3, ... identical on all ranks. */
typedef struct gsb_exchange {
    uint64_t peer_base[16];
    uint32_t rank, world;
    uint64_t seq;
} gsb_exchange;
int gsb_exchange_bytes(uint32_t world, uint32_t k, uint64_t* bytes);
int gsb_db_search_device_fused(const gsb_db* db, void* stream, const int32_t* d_query, uint32_t k,
                               float cutoff, const gsb_exchange* xchg, uint32_t* d_out_rows,
                               float* d_out_scores, uint32_t* d_out_n, uint64_t* d_out_approx);

/* The general form of the two device entry points above.  The query comes from host memory
 * (h_query: copied into the launch parameters, no H2D copy) or device memory (d_query; pass
 * GSB_QUERY_STABLE when nothing queued on the stream after the previous search writes it, which
 * lets the launch overlap that search's tail — without the flag the kernel first waits for the
 * work before it).  xchg == NULL: shard-local search (sink.keys / n / approx).  xchg != NULL: fused
 * cross-GPU search (sink.rows / scores / n / approx).  Every sink pointer must be device
 * accessible: device memory, or pinned host memory — then the results land on the host without any
 * copy, and `done` (optional) is set to done_value after everything else is visible system-wide:
 * poll it with gsb_wait_word instead of synchronising the stream. */
#define GSB_QUERY_STABLE 1u
typedef struct gsb_sink {
    gsb_key* keys;
    uint32_t* rows;
    float* scores;
    uint32_t* n;       /* entries written, or GSB_COUNT_ERROR */
    uint64_t* approx;
    uint64_t* done;
    uint64_t done_value;
} gsb_sink;
int gsb_db_search_enqueue(const gsb_db* db, void* stream, const int32_t* h_query, const int32_t* d_query,
                          uint32_t flags, uint32_t k, float cutoff, const gsb_exchange* xchg,
                          const gsb_sink* sink);
/* Spin until *word == value (a completion word in pinned host memory); timeout_us = 0 waits for
 * ever.  GSB_ERR_CUDA on timeout. */
int gsb_wait_word(const uint64_t* word, uint64_t value, uint64_t timeout_us);

/* ---- .fsim ingest without Qt (reference GPUSimServer::extractData, gpusim.cpp:173-253, and the
 * Decompress*Runnable helpers :48-85): big-endian QDataStream framing, qUncompress = zlib after a
 * 4-byte length, chunks inflated in parallel.  Strings stay valid until gsb_fsim_close. ---- */
typedef struct gsb_fsim gsb_fsim;
int gsb_fsim_open(const char* path, gsb_fsim** out);  /* GSB_ERR_IO: unreadable / wrong version */
void gsb_fsim_close(gsb_fsim* f);
const char* gsb_fsim_last_error(void);
const char* gsb_fsim_dbkey(const gsb_fsim* f);
int gsb_fsim_fp_bits(const gsb_fsim* f);
uint64_t gsb_fsim_fp_count(const gsb_fsim* f);
int gsb_fsim_chunk_count(const gsb_fsim* f);
const void* gsb_fsim_chunk_data(const gsb_fsim* f, int chunk);
uint64_t gsb_fsim_chunk_bytes(const gsb_fsim* f, int chunk);
uint64_t gsb_fsim_string_count(const gsb_fsim* f, int which);          /* which: 0 SMILES, 1 ids */
const char* gsb_fsim_string(const gsb_fsim* f, int which, uint64_t index);
/* gsb_db_create over the file's fingerprint chunks WITHOUT copying them: the database adopts the
 * inflated buffers (afterwards gsb_fsim_chunk_count() is 0; the strings stay with the file). */
int gsb_fsim_create_db(gsb_fsim* f, gsb_db** out);

/* ---- serving (reference GPUSimServer, gpusim.h:23-95 / gpusim.cpp:87-461), without Qt ----
 * Loads the .fsim files (database name = file base name, gpusim.cpp:114-116), applies the
 * fold-factor policy (:121-151; gpu_bitcount = 0 means automatic; GSB_ERR_INVALID "GPU bitset not
 * sufficiently small to fit on GPU" like the reference's std::invalid_argument), uploads when a
 * GPU is used (:159-163).  use_gpu = 0 is --cpu_only (main.cpp:21-28,64-65). */
typedef struct gsb_server gsb_server;
int gsb_server_create(const char* const* fsim_paths, int n_paths, int gpu_bitcount, int use_gpu,
                      gsb_server** out);
void gsb_server_destroy(gsb_server* srv);
/* After a search failed with GSB_ERR_CUDA: upload every database again, resetting the devices first if
 * the context itself is gone.  The serve loop calls it by itself. */
int gsb_server_recover(gsb_server* srv);
const char* gsb_server_last_error(void);
void gsb_server_set_use_gpu(gsb_server* srv, int use_gpu);   /* setUseGPU, gpusim.h:82          */
int gsb_server_using_gpu(const gsb_server* srv);             /* usingGPU, gpusim.cpp:168-171    */
unsigned gsb_server_fold_factor(const gsb_server* srv);
int gsb_server_database_count(const gsb_server* srv);
int gsb_server_get_fingerprint(const gsb_server* srv, const char* dbname, uint64_t row,
                               int32_t* out_words);          /* getFingerprint, gpusim.cpp:456-459 */
/* incomingSearchRequest (gpusim.cpp:376-454) on a byte buffer: parse the reference's request,
 * run searchDatabases (:306-374: per-database search, merge, SMILES de-duplication with ids joined
 * by ";:;"), serialise the reference's response.  *response is malloc'ed: gsb_server_free. */
int gsb_server_handle_request(gsb_server* srv, const void* request, uint64_t request_bytes,
                              void** response, uint64_t* response_bytes);
void gsb_server_free(void* p);
/* New: n requests that ask the same thing (databases + keys, result count, cutoff, query width)
 * answered from ONE pass over each database (gsb_db_search_batch).  The event loop does this by
 * itself for requests that arrive together.  responses[i] is malloc'ed: gsb_server_free. */
int gsb_server_handle_batch(gsb_server* srv, const void* const* requests, const uint64_t* request_bytes,
                            int n, void** responses, uint64_t* response_bytes);
/* setupSocket (gpusim.cpp:255-274): unix socket, default "/tmp/gpusimilarity" (what QLocalServer
 * name "gpusimilarity" resolves to); then the one-request-at-a-time event loop. */
int gsb_server_listen(gsb_server* srv, const char* socket_path);
int gsb_server_serve(gsb_server* srv, uint64_t max_requests);
void gsb_server_stop(gsb_server* srv);

/* ---- folding (reference calculation_functors.cpp:22-41, fingerprintdb_cuda.cpp:56-69) ---- */
int gsb_fold_fingerprint(const int32_t* words, int n_words, int factor, int32_t* out_words);

/* ---- introspection for benchmarks ------------------------------------------------------ */
typedef struct gsb_scan_info {
    int device;
    int grid;                 /* persistent CTAs launched per query                         */
    int block;                /* threads per CTA                                            */
    int stages;               /* TMA ring depth                                             */
    uint32_t tile_rows;       /* rows per TMA bulk load                                     */
    uint32_t tile_bytes;      /* bytes per TMA bulk load                                    */
    uint32_t smem_bytes;      /* dynamic shared memory per CTA                              */
    uint32_t cand_capacity;   /* per-CTA candidate buffer entries                           */
    uint64_t shard_rows;
    uint64_t db_bytes_per_query;  /* bytes one query's scan reads from HBM (layout bytes)     */
} gsb_scan_info;
/* Geometry gsb_db_search would use for shard `shard` and this k. */
int gsb_db_scan_info(const gsb_db* db, int shard, uint32_t k, gsb_scan_info* out);
/* Device self-test: the kernel's own correctly-rounded division against __fdiv_rn for every
 * (common, union) pair a fingerprint of up to 8192 bits can produce; *mismatches must be 0. */
int gsb_selftest_division(int device, uint64_t* mismatches);
/* Kernels launched by this library since load (all threads). */
uint64_t gsb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GPUSIM_B200_H */
