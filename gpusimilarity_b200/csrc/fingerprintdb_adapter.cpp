// gpusim::FingerprintDB (include/gpusim/fingerprintdb_cuda.h) over the C ABI.  Host-only C++;
// builds against real Qt5 (QObject/QString) or, where Qt is not installed, against the
// header-only stand-ins used for compile checks (oracle/qt_shims).
#include "gpusim/fingerprintdb_cuda.h"

#include <cstdint>
#include <stdexcept>
#include <string>

#include "gpusim/calculation_functors.h"
#include "gpusim_b200.h"

namespace gpusim
{

namespace
{
[[noreturn]] void raise(int rc)
{
    const std::string msg = gsb_last_error();
    if (rc == GSB_ERR_INVALID)
        throw std::invalid_argument(msg);
    throw std::runtime_error(msg); // what the reference throws for OOM / corruption (.cu:65,154)
}
} // namespace

unsigned int get_gpu_count() { return static_cast<unsigned int>(gsb_device_count()); }

unsigned int get_next_gpu(size_t required_memory)
{
    int dev = 0;
    const int rc = gsb_next_device(required_memory, &dev);
    if (rc != GSB_OK)
        raise(rc);
    return static_cast<unsigned int>(dev);
}

size_t get_available_gpu_memory() { return gsb_available_device_bytes(); }

FingerprintDB::FingerprintDB(int fp_bitcount, int fp_count, const QString& dbkey,
                             std::vector<std::vector<char>>& data, std::vector<char*>& smiles_vector,
                             std::vector<char*>& ids_vector)
    : m_dbkey(dbkey)
{
    m_fp_intsize = fp_bitcount / (sizeof(int) * 8);
    m_total_count = fp_count;
    std::vector<const void*> ptrs;
    std::vector<uint64_t> sizes;
    for (auto& chunk : data) {
        ptrs.push_back(chunk.data());
        sizes.push_back(chunk.size());
    }
    const int rc = gsb_db_create(ptrs.data(), sizes.data(), static_cast<int>(ptrs.size()), fp_bitcount,
                                 static_cast<uint64_t>(fp_count), &m_db);
    if (rc != GSB_OK)
        raise(rc);
    m_total_data_size = static_cast<size_t>(m_total_count) * static_cast<size_t>(m_fp_intsize) * sizeof(int);
    m_smiles.swap(smiles_vector); // reference .cu:164-165
    m_ids.swap(ids_vector);
}

FingerprintDB::~FingerprintDB() { gsb_db_destroy(m_db); }

void FingerprintDB::copyToGPU(unsigned int fold_factor)
{
    const int rc = gsb_db_upload(m_db, nullptr, 0, fold_factor);
    if (rc != GSB_OK)
        raise(rc);
}

Fingerprint FingerprintDB::getFingerprint(unsigned int index) const
{
    Fingerprint out(m_fp_intsize);
    const int rc = gsb_db_get_fingerprint(m_db, index, out.data());
    if (rc != GSB_OK)
        raise(rc);
    return out;
}

void FingerprintDB::search(const Fingerprint& query, const QString& dbkey, unsigned int max_return_count,
                           float similarity_cutoff, std::vector<char*>& results_smiles,
                           std::vector<char*>& results_ids, std::vector<float>& results_scores,
                           unsigned long& approximate_result_count) const
{
    if (dbkey != m_dbkey) // reference .cu:349-352: silent empty result
        return;
    std::vector<uint32_t> rows(max_return_count ? max_return_count : 1);
    std::vector<float> scores(rows.size());
    uint32_t n = 0;
    uint64_t approx = 0;
    const int rc = gsb_db_search(m_db, query.data(), static_cast<int>(query.size()), max_return_count,
                                 similarity_cutoff, rows.data(), scores.data(), &n, &approx);
    if (rc != GSB_OK)
        raise(rc);
    approximate_result_count = approx;
    for (uint32_t i = 0; i < n; i++) {
        results_scores.push_back(scores[i]);
        results_smiles.push_back(m_smiles[rows[i]]);
        results_ids.push_back(m_ids[rows[i]]);
    }
}

void FingerprintDB::search_cpu(const Fingerprint& query, const QString& dbkey, unsigned int max_return_count,
                               float /*similarity_cutoff*/, std::vector<char*>& results_smiles,
                               std::vector<char*>& results_ids, std::vector<float>& results_scores,
                               unsigned long& /*approximate_result_count*/) const
{
    if (dbkey != m_dbkey) // reference fingerprintdb_cuda.cpp:28-31
        return;
    std::vector<uint32_t> rows(max_return_count ? max_return_count : 1);
    std::vector<float> scores(rows.size());
    uint32_t n = 0;
    const int rc = gsb_db_search_cpu(m_db, query.data(), static_cast<int>(query.size()), max_return_count,
                                     rows.data(), scores.data(), &n);
    if (rc != GSB_OK)
        raise(rc);
    for (uint32_t i = 0; i < n; i++) {
        results_smiles.push_back(m_smiles[rows[i]]);
        results_ids.push_back(m_ids[rows[i]]);
        results_scores.push_back(scores[i]);
    }
}

void top_results_bubble_sort(std::vector<int>& indices, std::vector<float>& scores, int number_required)
{
    // partial bubble sort: after pass i the i-th best has floated to position i; the strict
    // comparison keeps equal scores in their original (ascending index) order
    const int count = static_cast<int>(indices.size());
    for (int pass = 0; pass < number_required; pass++) {
        for (int pos = count - 1; pos > pass; pos--) {
            if (scores[pos] > scores[pos - 1]) {
                std::swap(indices[pos], indices[pos - 1]);
                std::swap(scores[pos], scores[pos - 1]);
            }
        }
    }
}

void TanimotoFunctorCPU::operator()(const int& fp_index) const
{
    const int* row = m_dbdata + static_cast<size_t>(m_fp_intsize) * fp_index;
    int common = 0, total = 0;
    for (int w = 0; w < m_fp_intsize; w++) {
        const unsigned q = static_cast<unsigned>(m_ref_fp[w]), d = static_cast<unsigned>(row[w]);
        common += __builtin_popcount(q & d);
        total += __builtin_popcount(q) + __builtin_popcount(d);
    }
    m_output[fp_index] = static_cast<float>(common) / static_cast<float>(total - common);
}

void FoldFingerprintFunctorCPU::operator()(const int& fp_index) const
{
    const int* in = m_unfolded + static_cast<size_t>(fp_index) * m_unfolded_fp_intsize;
    int* out = m_folded + static_cast<size_t>(fp_index) * m_folded_fp_intsize;
    // accumulates into the output like the reference (callers pass zeroed storage)
    for (int w = 0; w < m_unfolded_fp_intsize; w++)
        out[w % m_folded_fp_intsize] |= in[w];
}

} // namespace gpusim
