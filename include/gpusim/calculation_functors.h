// Same functor classes as the reference's calculation_functors.h:11-45 (used directly by
// test/test_gpusim.cpp:154,163 and by search_cpu), written from the algorithm description.
#pragma once

#include <vector>

#include "types.h"

namespace gpusim
{

// reference calculation_functors.cpp:6-20: output[i] = float(common) / float(total - common)
struct TanimotoFunctorCPU {
    const int* m_ref_fp;
    const int m_fp_intsize;
    const int* m_dbdata;
    float* m_output;

    TanimotoFunctorCPU(const Fingerprint& ref_fp, int fp_intsize, const std::vector<int>& dbdata,
                       std::vector<float>& output)
        : m_ref_fp(ref_fp.data()), m_fp_intsize(fp_intsize), m_dbdata(dbdata.data()),
          m_output(output.data())
    {
    }
    void operator()(const int& fp_index) const;
};

// reference calculation_functors.cpp:22-41: folded row = OR of the `factor` contiguous segments
class FoldFingerprintFunctorCPU
{
    const int m_unfolded_fp_intsize;
    const int m_folded_fp_intsize;
    const int* m_unfolded;
    int* m_folded;

  public:
    FoldFingerprintFunctorCPU(const int factor, const int fp_intsize, const std::vector<int>& unfolded,
                              std::vector<int>& folded)
        : m_unfolded_fp_intsize(fp_intsize), m_folded_fp_intsize(fp_intsize / factor),
          m_unfolded(unfolded.data()), m_folded(folded.data())
    {
    }
    void operator()(const int& fp_index) const;
};

} // namespace gpusim
