// Drop-in replacement for the reference's fingerprintdb_cuda.h (:22-160): the same
// gpusim::FingerprintDB surface that gpusim.cpp and test/test_gpusim.cpp compile against,
// implemented over the C ABI of libgpusim_b200.so (include/gpusim_b200.h) in
// gpusimilarity_b200/csrc/fingerprintdb_adapter.cpp.  CUDA types never appear here; the
// engine returns global row numbers and this class maps them to the SMILES / id strings.
//
// Differences from the reference, all documented in SURVEY App. D ("do NOT replicate"):
// search_cpu searches every chunk and clamps k to the row count; getFingerprint has no
// chunk-boundary off-by-one; get_next_gpu checks the device it returns.
#ifndef FINGERPRINTDB_CUDA
#define FINGERPRINTDB_CUDA

#include <memory>
#include <utility>
#include <vector>

#include <QObject>
#include <QString>

#include "types.h"

struct gsb_db;

namespace gpusim
{

class FingerprintDB;

typedef std::pair<char*, char*> ResultData;
typedef std::pair<float, ResultData> SortableResult;

unsigned int get_gpu_count();                      // reference .h:32, .cu:41-52
unsigned int get_next_gpu(size_t required_memory); // reference .h:33, .cu:54-68 (throws std::runtime_error)

class FingerprintDB : public QObject
{
  public:
    // reference .h:58-61 / .cu:133-166: copies the fingerprint chunks, TAKES (swaps out) the
    // smiles / ids vectors, throws std::runtime_error when fp_count does not match the data.
    FingerprintDB(int fp_bitcount, int fp_count, const QString& dbkey,
                  std::vector<std::vector<char>>& data, std::vector<char*>& smiles_vector,
                  std::vector<char*>& ids_vector);
    ~FingerprintDB() override;
    FingerprintDB(const FingerprintDB&) = delete;
    FingerprintDB& operator=(const FingerprintDB&) = delete;

    // reference .h:71 / .cu:168-195
    void copyToGPU(unsigned int fold_factor);

    unsigned int count() const { return m_total_count; } // .h:74

    // reference .h:93 / .cu:212-226
    Fingerprint getFingerprint(unsigned int index) const;

    // reference .h:106-111 / .cu:341-381: results are appended; approximate_result_count is
    // assigned, and left untouched (with nothing appended) on a dbkey mismatch.
    void search(const Fingerprint& query, const QString& dbkey, unsigned int max_return_count,
                float similarity_cutoff, std::vector<char*>& results_smiles,
                std::vector<char*>& results_ids, std::vector<float>& results_scores,
                unsigned long& approximate_result_count) const;

    // reference .h:113-118 / fingerprintdb_cuda.cpp:20-54: cutoff ignored, count not written.
    void search_cpu(const Fingerprint& query, const QString& dbkey, unsigned int max_return_count,
                    float similarity_cutoff, std::vector<char*>& results_smiles,
                    std::vector<char*>& results_ids, std::vector<float>& results_scores,
                    unsigned long& approximate_result_count) const;

    char* getSmiles(int index) const { return m_smiles[index]; } // .h:120
    char* getID(int index) const { return m_ids[index]; }        // .h:121

    size_t getFingerprintDataSize() const { return m_total_data_size; } // .h:123
    int getFingerprintBitcount() const { return m_fp_intsize * sizeof(int) * 8; } // .h:124-127

  protected:
    gsb_db* m_db = nullptr;
    int m_total_count = 0, m_fp_intsize = 0;
    size_t m_total_data_size = 0;
    std::vector<char*> m_smiles;
    std::vector<char*> m_ids;
    QString m_dbkey;
};

size_t get_available_gpu_memory(); // reference .h:149, .cu:401-413

// reference .h:155-156, fingerprintdb_cuda.cpp:92-103
void top_results_bubble_sort(std::vector<int>& indices, std::vector<float>& scores, int number_required);

} // namespace gpusim

#endif
