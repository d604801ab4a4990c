/* CPU oracle (plain C) for the brute-force Tanimoto scan + top-k path.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline ("port") leg through oracle/oracle_c.py.  The product never links this.
 *
 * Pinned by tests/test_oracle.py against the reference's own known answers and against the
 * reference sources compiled verbatim (oracle/_ref), see oracle/oracle.py.
 *
 * Follows:
 *   scoring            reference calculation_functors.cpp:6-20  (TanimotoFunctorCPU)
 *   cutoff / zeroing   reference fingerprintdb_cuda.cu:100-102  (TanimotoFunctor)
 *   survivors, order   reference fingerprintdb_cuda.cu:263-290, 366-380
 *   fold               reference calculation_functors.cpp:22-41
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float score;
    uint32_t row;
} cand_t;

/* calculation_functors.cpp:6-20, one row. */
static inline float score_row(const int32_t* q, const int32_t* d, int words)
{
    int total = 0, common = 0;
    for (int i = 0; i < words; i++) {
        const int32_t fp1 = q[i], fp2 = d[i];
        total += __builtin_popcount((unsigned) fp1) + __builtin_popcount((unsigned) fp2);
        common += __builtin_popcount((unsigned) (fp1 & fp2));
    }
    return (float) common / (float) (total - common);
}

/* canonical order: score descending, then row ascending. */
static inline int better(cand_t a, cand_t b)
{
    return a.score > b.score || (a.score == b.score && a.row < b.row);
}

/* bounded "keep the k best" container: binary heap whose root is the WORST kept entry */
typedef struct {
    cand_t* h;
    uint32_t n, cap;
} heap_t;

static void heap_push(heap_t* hp, cand_t c)
{
    cand_t* h = hp->h;
    if (hp->n < hp->cap) {
        uint32_t i = hp->n++;
        while (i > 0) {
            uint32_t p = (i - 1) / 2;
            if (!better(h[p], c))
                break;
            h[i] = h[p];
            i = p;
        }
        h[i] = c;
        return;
    }
    if (hp->cap == 0 || !better(c, h[0]))
        return;
    uint32_t i = 0;
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, w = i;
        cand_t worst = c;
        if (l < hp->n && better(worst, h[l])) { w = l; worst = h[l]; }
        if (r < hp->n && better(worst, h[r])) { w = r; worst = h[r]; }
        if (w == i)
            break;
        h[i] = h[w];
        i = w;
    }
    h[i] = c;
}

static int cmp_canonical(const void* a, const void* b)
{
    const cand_t x = *(const cand_t*) a, y = *(const cand_t*) b;
    return better(x, y) ? -1 : (better(y, x) ? 1 : 0);
}

typedef struct {
    const int32_t* q;
    const int32_t* db;
    int words;
    uint64_t lo, hi, row_base;
    float cutoff;
    float* scores_out; /* optional: raw CPU scores */
    heap_t heap;
    uint64_t survivors;
} job_t;

static void* scan_job(void* arg)
{
    job_t* j = (job_t*) arg;
    const int drop_zero = j->cutoff > 0.0f; /* .cu:265 */
    for (uint64_t r = j->lo; r < j->hi; r++) {
        float s = score_row(j->q, j->db + r * (uint64_t) j->words, j->words);
        if (j->scores_out)
            j->scores_out[r] = s;
        if (!j->heap.cap && !drop_zero)
            continue;
        s = (s >= j->cutoff) ? s : 0.0f; /* .cu:102; NaN -> 0 */
        if (drop_zero && s == 0.0f)
            continue;
        j->survivors++;
        cand_t c = {s, (uint32_t) (r + j->row_base)};
        heap_push(&j->heap, c);
    }
    return NULL;
}

static int run_jobs(const int32_t* q, int words, const int32_t* db, uint64_t n_rows,
                    uint64_t row_base, uint32_t k, float cutoff, float* scores_out,
                    int n_threads, cand_t** merged, uint32_t* n_merged, uint64_t* survivors)
{
    if (n_threads < 1)
        n_threads = 1;
    job_t* jobs = (job_t*) calloc(n_threads, sizeof(job_t));
    pthread_t* th = (pthread_t*) calloc(n_threads, sizeof(pthread_t));
    const uint64_t per = (n_rows + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        job_t* j = &jobs[t];
        j->q = q; j->db = db; j->words = words; j->row_base = row_base;
        j->lo = (uint64_t) t * per < n_rows ? (uint64_t) t * per : n_rows;
        j->hi = j->lo + per < n_rows ? j->lo + per : n_rows;
        j->cutoff = cutoff; j->scores_out = scores_out;
        j->heap.cap = k; j->heap.n = 0;
        j->heap.h = k ? (cand_t*) malloc((size_t) k * sizeof(cand_t)) : NULL;
        pthread_create(&th[t], NULL, scan_job, j);
    }
    uint64_t surv = 0;
    uint32_t total = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        surv += jobs[t].survivors;
        total += jobs[t].heap.n;
    }
    cand_t* all = (cand_t*) malloc((size_t) (total ? total : 1) * sizeof(cand_t));
    uint32_t o = 0;
    for (int t = 0; t < n_threads; t++) {
        memcpy(all + o, jobs[t].heap.h, (size_t) jobs[t].heap.n * sizeof(cand_t));
        o += jobs[t].heap.n;
        free(jobs[t].heap.h);
    }
    qsort(all, total, sizeof(cand_t), cmp_canonical);
    *merged = all; *n_merged = total; *survivors = surv;
    free(jobs); free(th);
    return 0;
}

/* Raw CPU scores for every row (TanimotoFunctorCPU semantics: no cutoff, 0/0 = NaN). */
void oracle_score(const int32_t* query, int words, const int32_t* db, uint64_t n_rows,
                  float* out_scores, int n_threads)
{
    cand_t* m; uint32_t nm; uint64_t s;
    run_jobs(query, words, db, n_rows, 0, 0, -1.0f, out_scores, n_threads, &m, &nm, &s);
    free(m);
}

/* FingerprintDB::search semantics (unfolded): returns min(k, survivors) results in canonical
 * order with global row ids, and the approximate (survivor) count. */
void oracle_search(const int32_t* query, int words, const int32_t* db, uint64_t n_rows,
                   uint64_t row_base, uint32_t k, float cutoff, uint32_t* out_rows,
                   float* out_scores, uint32_t* out_n, uint64_t* out_approx, int n_threads)
{
    cand_t* m; uint32_t nm; uint64_t surv;
    run_jobs(query, words, db, n_rows, row_base, k, cutoff, NULL, n_threads, &m, &nm, &surv);
    const uint32_t n = nm < k ? nm : k;
    for (uint32_t i = 0; i < n; i++) {
        out_rows[i] = m[i].row;
        out_scores[i] = m[i].score;
    }
    *out_n = n;
    *out_approx = cutoff > 0.0f ? surv : n_rows; /* .cu:272-277, 367-369 */
    free(m);
}

/* FoldFingerprintFunctorCPU (calculation_functors.cpp:22-41): OR of the `factor` contiguous
 * segments of words/factor words. */
void oracle_fold(const int32_t* fp, int words, int factor, int32_t* out)
{
    const int new_words = words / factor;
    memset(out, 0, (size_t) new_words * sizeof(int32_t));
    for (int w = 0; w < words; w++)
        out[w % new_words] |= fp[w];
}

/* ---- synthetic database generator: C twin of oracle.py synth_rows (and of the device generator
 * in gpusimilarity_b200/csrc/gsb_kernels.cuh), threaded so tests can build 10 M-row inputs. ---- */
static inline uint64_t mix64(uint64_t x)
{
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint32_t hash32(uint64_t seed, uint64_t row, uint64_t word, uint64_t salt)
{
    return (uint32_t) (mix64(row * 0x9E3779B97F4A7C15ull + word * 0xD1B54A32D192ED03ull +
                             salt * 0x8CB92BA72F3D8DD7ull + seed) >> 32);
}
static inline uint32_t random_word(uint64_t seed, uint64_t row, uint32_t w)
{
    uint32_t x = hash32(seed, row, w, 0);
    for (uint32_t salt = 1; salt < 5; salt++)
        x &= hash32(seed, row, w, salt);
    return x;
}
#define SYNTH_TEMPLATE_ROW 0xFFFFFFFFull
#define SYNTH_MAX_FLIPS 24

typedef struct {
    uint64_t seed, lo, hi, row_base;
    int words;
    uint32_t plant_period;
    uint32_t* out;
} synth_job_t;

static void* synth_job(void* arg)
{
    synth_job_t* j = (synth_job_t*) arg;
    for (uint64_t r = j->lo; r < j->hi; r++) {
        const uint64_t grow = j->row_base + r;
        uint32_t* row = j->out + r * (uint64_t) j->words;
        int planted = j->plant_period > 0 && (hash32(j->seed, grow, 0, 7) % j->plant_period) == 0;
        if (!planted) {
            for (int w = 0; w < j->words; w++)
                row[w] = random_word(j->seed, grow, (uint32_t) w);
        } else {
            for (int w = 0; w < j->words; w++)
                row[w] = random_word(j->seed, SYNTH_TEMPLATE_ROW, (uint32_t) w);
            const uint32_t nflip = 1 + hash32(j->seed, grow, 1, 7) % SYNTH_MAX_FLIPS;
            for (uint32_t f = 0; f < nflip; f++) {
                const uint32_t pos = hash32(j->seed, grow, 2 + f, 7) % (uint32_t) (j->words * 32);
                row[pos >> 5] ^= 1u << (pos & 31);
            }
        }
    }
    return NULL;
}

void oracle_synth_rows(uint64_t seed, uint64_t n_rows, uint64_t row_base, int words,
                       uint32_t plant_period, uint32_t* out, int n_threads)
{
    if (n_threads < 1)
        n_threads = 1;
    synth_job_t* jobs = (synth_job_t*) calloc(n_threads, sizeof(synth_job_t));
    pthread_t* th = (pthread_t*) calloc(n_threads, sizeof(pthread_t));
    const uint64_t per = (n_rows + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        synth_job_t* j = &jobs[t];
        j->seed = seed; j->row_base = row_base; j->words = words; j->plant_period = plant_period;
        j->out = out;
        j->lo = (uint64_t) t * per < n_rows ? (uint64_t) t * per : n_rows;
        j->hi = j->lo + per < n_rows ? j->lo + per : n_rows;
        pthread_create(&th[t], NULL, synth_job, j);
    }
    for (int t = 0; t < n_threads; t++)
        pthread_join(th[t], NULL);
    free(jobs); free(th);
}

/* ---- streamed search over the synthetic database: rows [row_base, row_base + n_rows) are generated
 * on the fly (never materialised), scored and selected exactly like oracle_search.  Lets the tests
 * check a 1 B-row GPU search (BASELINE configs[2]) bit for bit instead of by invariants. ---- */
typedef struct {
    const int32_t* q;
    uint64_t seed, lo, hi, row_base;
    int words;
    uint32_t plant_period;
    float cutoff;
    heap_t heap;
    uint64_t survivors;
} stream_job_t;

static void* stream_job(void* arg)
{
    stream_job_t* j = (stream_job_t*) arg;
    const int drop_zero = j->cutoff > 0.0f;
    uint32_t row[128], tmpl[128];
    for (int w = 0; w < j->words; w++)
        tmpl[w] = random_word(j->seed, SYNTH_TEMPLATE_ROW, (uint32_t) w);
    for (uint64_t r = j->lo; r < j->hi; r++) {
        const uint64_t grow = j->row_base + r;
        int planted = j->plant_period > 0 && (hash32(j->seed, grow, 0, 7) % j->plant_period) == 0;
        if (!planted) {
            for (int w = 0; w < j->words; w++)
                row[w] = random_word(j->seed, grow, (uint32_t) w);
        } else {
            memcpy(row, tmpl, (size_t) j->words * 4);
            const uint32_t nflip = 1 + hash32(j->seed, grow, 1, 7) % SYNTH_MAX_FLIPS;
            for (uint32_t f = 0; f < nflip; f++) {
                const uint32_t pos = hash32(j->seed, grow, 2 + f, 7) % (uint32_t) (j->words * 32);
                row[pos >> 5] ^= 1u << (pos & 31);
            }
        }
        float s = score_row(j->q, (const int32_t*) row, j->words);
        s = (s >= j->cutoff) ? s : 0.0f;
        if (drop_zero && s == 0.0f)
            continue;
        j->survivors++;
        cand_t c = {s, (uint32_t) grow};
        heap_push(&j->heap, c);
    }
    return NULL;
}

void oracle_stream_search(const int32_t* query, int words, uint64_t seed, uint32_t plant_period,
                          uint64_t n_rows, uint64_t row_base, uint32_t k, float cutoff,
                          uint32_t* out_rows, float* out_scores, uint32_t* out_n,
                          uint64_t* out_approx, int n_threads)
{
    if (n_threads < 1)
        n_threads = 1;
    stream_job_t* jobs = (stream_job_t*) calloc(n_threads, sizeof(stream_job_t));
    pthread_t* th = (pthread_t*) calloc(n_threads, sizeof(pthread_t));
    const uint64_t per = (n_rows + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        stream_job_t* j = &jobs[t];
        j->q = query; j->seed = seed; j->row_base = row_base; j->words = words;
        j->plant_period = plant_period; j->cutoff = cutoff;
        j->lo = (uint64_t) t * per < n_rows ? (uint64_t) t * per : n_rows;
        j->hi = j->lo + per < n_rows ? j->lo + per : n_rows;
        j->heap.cap = k; j->heap.n = 0;
        j->heap.h = k ? (cand_t*) malloc((size_t) k * sizeof(cand_t)) : NULL;
        pthread_create(&th[t], NULL, stream_job, j);
    }
    uint64_t surv = 0;
    uint32_t total = 0;
    for (int t = 0; t < n_threads; t++) {
        pthread_join(th[t], NULL);
        surv += jobs[t].survivors;
        total += jobs[t].heap.n;
    }
    cand_t* all = (cand_t*) malloc((size_t) (total ? total : 1) * sizeof(cand_t));
    uint32_t o = 0;
    for (int t = 0; t < n_threads; t++) {
        memcpy(all + o, jobs[t].heap.h, (size_t) jobs[t].heap.n * sizeof(cand_t));
        o += jobs[t].heap.n;
        free(jobs[t].heap.h);
    }
    qsort(all, total, sizeof(cand_t), cmp_canonical);
    const uint32_t n = total < k ? total : k;
    for (uint32_t i = 0; i < n; i++) {
        out_rows[i] = all[i].row;
        out_scores[i] = all[i].score;
    }
    *out_n = n;
    *out_approx = cutoff > 0.0f ? surv : n_rows;
    free(all); free(jobs); free(th);
}

/* ---- the same for a batch of queries in ONE pass over the generated rows (row generation is the
 * expensive part): out arrays are [nq][k]; bench.py checks BASELINE configs[4] (1024 queries,
 * top-100) result lists and the single-query top-1000 against it at full size. ---- */
typedef struct {
    const int32_t* qs;
    int nq;
    uint64_t seed, lo, hi, row_base;
    int words;
    uint32_t plant_period;
    float cutoff;
    heap_t* heaps;      /* [nq] */
    uint64_t* survivors; /* [nq] */
} mstream_job_t;

/* common[q] = popcount(query q AND row) for every query.  Cloned for CPUs with AVX-512 VPOPCNTDQ
 * (the compiler vectorises the popcount reduction there); plain POPCNT elsewhere. */
__attribute__((target_clones("arch=icelake-server", "default")))
static void common_counts(const int32_t* qs, int nq, int words, const uint32_t* row, int* out)
{
    if ((words & 1) == 0) { /* popcount is additive: two 32-bit words at a time */
        const int w64 = words / 2;
        uint64_t r[64];
        memcpy(r, row, (size_t) words * 4);
        for (int q = 0; q < nq; q++) {
            uint64_t a[64];
            memcpy(a, qs + (size_t) q * words, (size_t) words * 4);
            long long c = 0;
            for (int w = 0; w < w64; w++)
                c += __builtin_popcountll(a[w] & r[w]);
            out[q] = (int) c;
        }
    } else {
        for (int q = 0; q < nq; q++) {
            const uint32_t* qw = (const uint32_t*) qs + (size_t) q * words;
            int c = 0;
            for (int w = 0; w < words; w++)
                c += __builtin_popcount(qw[w] & row[w]);
            out[q] = c;
        }
    }
}

static void* mstream_job(void* arg)
{
    mstream_job_t* j = (mstream_job_t*) arg;
    const int drop_zero = j->cutoff > 0.0f;
    const int words = j->words;
    uint32_t row[128], tmpl[128];
    int* popq = (int*) malloc((size_t) j->nq * sizeof(int));
    int* commons = (int*) malloc((size_t) j->nq * sizeof(int));
    for (int q = 0; q < j->nq; q++) {
        popq[q] = 0;
        for (int w = 0; w < words; w++)
            popq[q] += __builtin_popcount((unsigned) j->qs[(size_t) q * words + w]);
    }
    for (int w = 0; w < words; w++)
        tmpl[w] = random_word(j->seed, SYNTH_TEMPLATE_ROW, (uint32_t) w);
    for (uint64_t r = j->lo; r < j->hi; r++) {
        const uint64_t grow = j->row_base + r;
        int planted = j->plant_period > 0 && (hash32(j->seed, grow, 0, 7) % j->plant_period) == 0;
        if (!planted) {
            for (int w = 0; w < words; w++)
                row[w] = random_word(j->seed, grow, (uint32_t) w);
        } else {
            memcpy(row, tmpl, (size_t) words * 4);
            const uint32_t nflip = 1 + hash32(j->seed, grow, 1, 7) % SYNTH_MAX_FLIPS;
            for (uint32_t f = 0; f < nflip; f++) {
                const uint32_t pos = hash32(j->seed, grow, 2 + f, 7) % (uint32_t) (words * 32);
                row[pos >> 5] ^= 1u << (pos & 31);
            }
        }
        int popd = 0;
        for (int w = 0; w < words; w++)
            popd += __builtin_popcount(row[w]);
        common_counts(j->qs, j->nq, words, row, commons);
        for (int q = 0; q < j->nq; q++) {
            const int common = commons[q];
            /* Cheap, conservative skip (no cutoff only): with the heap full, a row whose quotient is
             * more than 0.1 % below the worst kept score cannot round up to it. */
            heap_t* hp = &j->heaps[q];
            if (!drop_zero && hp->n == hp->cap && hp->cap > 0 &&
                (float) common < hp->h[0].score * (float) (popq[q] + popd - common) * 0.999f)
                continue;
            /* calculation_functors.cpp:18 with total = popq + popd */
            float s = (float) common / (float) (popq[q] + popd - common);
            s = (s >= j->cutoff) ? s : 0.0f;
            if (drop_zero && s == 0.0f)
                continue;
            j->survivors[q]++;
            cand_t c = {s, (uint32_t) grow};
            heap_push(&j->heaps[q], c);
        }
    }
    free(popq);
    free(commons);
    return NULL;
}

void oracle_stream_search_multi(const int32_t* queries, int nq, int words, uint64_t seed,
                                uint32_t plant_period, uint64_t n_rows, uint64_t row_base, uint32_t k,
                                float cutoff, uint32_t* out_rows, float* out_scores, uint32_t* out_n,
                                uint64_t* out_approx, int n_threads)
{
    if (n_threads < 1)
        n_threads = 1;
    mstream_job_t* jobs = (mstream_job_t*) calloc(n_threads, sizeof(mstream_job_t));
    pthread_t* th = (pthread_t*) calloc(n_threads, sizeof(pthread_t));
    const uint64_t per = (n_rows + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        mstream_job_t* j = &jobs[t];
        j->qs = queries; j->nq = nq; j->seed = seed; j->row_base = row_base; j->words = words;
        j->plant_period = plant_period; j->cutoff = cutoff;
        j->lo = (uint64_t) t * per < n_rows ? (uint64_t) t * per : n_rows;
        j->hi = j->lo + per < n_rows ? j->lo + per : n_rows;
        j->heaps = (heap_t*) calloc(nq, sizeof(heap_t));
        j->survivors = (uint64_t*) calloc(nq, sizeof(uint64_t));
        for (int q = 0; q < nq; q++) {
            j->heaps[q].cap = k;
            j->heaps[q].h = k ? (cand_t*) malloc((size_t) k * sizeof(cand_t)) : NULL;
        }
        pthread_create(&th[t], NULL, mstream_job, j);
    }
    for (int t = 0; t < n_threads; t++)
        pthread_join(th[t], NULL);
    cand_t* all = (cand_t*) malloc(((size_t) n_threads * k + 1) * sizeof(cand_t));
    for (int q = 0; q < nq; q++) {
        uint32_t total = 0;
        uint64_t surv = 0;
        for (int t = 0; t < n_threads; t++) {
            memcpy(all + total, jobs[t].heaps[q].h, (size_t) jobs[t].heaps[q].n * sizeof(cand_t));
            total += jobs[t].heaps[q].n;
            surv += jobs[t].survivors[q];
        }
        qsort(all, total, sizeof(cand_t), cmp_canonical);
        const uint32_t n = total < k ? total : k;
        for (uint32_t i = 0; i < n; i++) {
            out_rows[(size_t) q * k + i] = all[i].row;
            out_scores[(size_t) q * k + i] = all[i].score;
        }
        out_n[q] = n;
        out_approx[q] = cutoff > 0.0f ? surv : n_rows;
    }
    free(all);
    for (int t = 0; t < n_threads; t++) {
        for (int q = 0; q < nq; q++)
            free(jobs[t].heaps[q].h);
        free(jobs[t].heaps); free(jobs[t].survivors);
    }
    free(jobs); free(th);
}
