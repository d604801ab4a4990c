// C driver around the reference's own FingerprintDB, compiled verbatim from /root/reference by
// oracle/Makefile into oracle/_ref/libgpusim_ref.so.  TEST INFRASTRUCTURE: used by tests/,
// tests/golden/make_golden.py and bench.py's reference / cpu_baseline legs only.
//
// The reference hands results back as (smiles char*, id char*) pairs.  To recover row numbers,
// every row's "smiles" and "id" pointers point into one tag array, so pointer - base == row.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "calculation_functors.h"
#include "fingerprintdb_cuda.h"

using gpusim::Fingerprint;
using gpusim::FingerprintDB;

namespace
{
struct RefDB {
    std::vector<char> tags;
    std::unique_ptr<FingerprintDB> db;
    std::string key;
};
thread_local std::string g_err;
} // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_gpu_count() { return static_cast<int>(gpusim::get_gpu_count()); }

void* ref_db_create(int fp_bitcount, int fp_count, const char* dbkey,
                    const void* const* chunk_ptrs, const uint64_t* chunk_bytes,
                    int n_chunks)
{
    try {
        auto* h = new RefDB;
        h->key = dbkey;
        h->tags.assign(static_cast<size_t>(fp_count) + 1, 0);
        std::vector<std::vector<char>> data(n_chunks);
        for (int c = 0; c < n_chunks; c++) {
            const char* p = static_cast<const char*>(chunk_ptrs[c]);
            data[c].assign(p, p + chunk_bytes[c]);
        }
        std::vector<char*> smiles(fp_count), ids(fp_count);
        for (int i = 0; i < fp_count; i++) {
            smiles[i] = h->tags.data() + i;
            ids[i] = h->tags.data() + i;
        }
        h->db.reset(new FingerprintDB(fp_bitcount, fp_count, QString(dbkey), data,
                                      smiles, ids));
        return h;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void ref_db_destroy(void* handle) { delete static_cast<RefDB*>(handle); }

int ref_db_copy_to_gpu(void* handle, unsigned fold_factor)
{
    try {
        static_cast<RefDB*>(handle)->db->copyToGPU(fold_factor);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// Returns the number of results, or -1.  calc: 0 = FingerprintDB::search (CUDA),
// 1 = FingerprintDB::search_cpu.  approx is passed through untouched when the reference
// does not assign it (search_cpu, key mismatch).
int ref_db_search(void* handle, const int* query, int n_words, const char* dbkey,
                  unsigned k, float cutoff, int calc, uint32_t* out_rows,
                  float* out_scores, unsigned long* approx)
{
    try {
        auto* h = static_cast<RefDB*>(handle);
        Fingerprint q(query, query + n_words);
        std::vector<char*> smiles, ids;
        std::vector<float> scores;
        if (calc == 0)
            h->db->search(q, QString(dbkey), k, cutoff, smiles, ids, scores, *approx);
        else
            h->db->search_cpu(q, QString(dbkey), k, cutoff, smiles, ids, scores, *approx);
        for (size_t i = 0; i < smiles.size(); i++) {
            out_rows[i] = static_cast<uint32_t>(smiles[i] - h->tags.data());
            out_scores[i] = scores[i];
        }
        return static_cast<int>(smiles.size());
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

int ref_db_get_fingerprint(void* handle, unsigned row, int* out_words, int n_words)
{
    auto fp = static_cast<RefDB*>(handle)->db->getFingerprint(row);
    for (int i = 0; i < n_words && i < static_cast<int>(fp.size()); i++)
        out_words[i] = fp[i];
    return static_cast<int>(fp.size());
}

void ref_bubble_sort(int* indices, float* scores, int n, int number_required)
{
    std::vector<int> vi(indices, indices + n);
    std::vector<float> vs(scores, scores + n);
    gpusim::top_results_bubble_sort(vi, vs, number_required);
    std::memcpy(indices, vi.data(), n * sizeof(int));
    std::memcpy(scores, vs.data(), n * sizeof(float));
}

void ref_fold(const int* fp, int n_words, int factor, int* out)
{
    std::vector<int> unfolded(fp, fp + n_words);
    std::vector<int> folded(n_words / factor, 0);
    gpusim::FoldFingerprintFunctorCPU(factor, n_words, unfolded, folded)(0);
    std::memcpy(out, folded.data(), folded.size() * sizeof(int));
}

// TanimotoFunctorCPU over rows [0, n_rows) of one <= 2^26-row block on n_threads threads
// (what QtConcurrent::blockingMap does in search_cpu, fingerprintdb_cuda.cpp:42-44).
void ref_score_cpu(const int* query, int n_words, const int* db, int n_rows, float* out,
                   int n_threads)
{
    Fingerprint q(query, query + n_words);
    std::vector<int> dbv;   // the functor wants std::vector storage: alias without copying
    std::vector<float> outv;
    // Point the functor's public raw-pointer members at the caller's buffers.
    gpusim::TanimotoFunctorCPU f(q, n_words, dbv, outv);
    f.m_dbdata = db;
    f.m_output = out;
    if (n_threads < 1)
        n_threads = 1;
    std::vector<std::thread> pool;
    const int per = (n_rows + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; t++) {
        const int lo = std::min(n_rows, t * per), hi = std::min(n_rows, lo + per);
        if (lo >= hi)
            break;
        pool.emplace_back([&f, lo, hi]() {
            for (int i = lo; i < hi; i++)
                f(i);
        });
    }
    for (auto& th : pool)
        th.join();
}

} // extern "C"
