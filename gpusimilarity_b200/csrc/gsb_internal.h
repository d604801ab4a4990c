// Internal C++ seam between the translation units of libgpusim_b200.so (not part of the C ABI).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

struct gsb_db;

// A host buffer of fingerprint rows.  With a CUDA device present it is PINNED (page-locked, mapped
// and portable) from birth, so that
//   * the .fsim reader inflates a chunk straight into it (reference DecompressAssignFPRunnable,
//     gpusim.cpp:48-63, inflates into a QByteArray and copies that into a std::vector),
//   * the upload DMAs out of it without a staging copy (reference: pageable cudaMemcpy, .cu:181-182),
//   * the second stage of a folded search reads candidate rows in place from the device.
// Falls back to plain memory when pinning fails or no device exists (search_cpu only needs that).
class GsbHostBuf
{
  public:
    GsbHostBuf() = default;
    ~GsbHostBuf() { release(); }
    GsbHostBuf(GsbHostBuf&& o) noexcept : m_p(o.m_p), m_n(o.m_n), m_pinned(o.m_pinned)
    {
        o.m_p = nullptr;
        o.m_n = 0;
        o.m_pinned = false;
    }
    GsbHostBuf& operator=(GsbHostBuf&& o) noexcept
    {
        if (this != &o) {
            release();
            m_p = o.m_p, m_n = o.m_n, m_pinned = o.m_pinned;
            o.m_p = nullptr, o.m_n = 0, o.m_pinned = false;
        }
        return *this;
    }
    GsbHostBuf(const GsbHostBuf&) = delete;
    GsbHostBuf& operator=(const GsbHostBuf&) = delete;
    // false when the memory cannot be had at all
    bool allocate(size_t n);
    void release();
    uint8_t* data() { return m_p; }
    const uint8_t* data() const { return m_p; }
    size_t size() const { return m_n; }
    bool empty() const { return m_n == 0; }
    bool pinned() const { return m_pinned; }
    void shrink(size_t n) { m_n = n < m_n ? n : m_n; }

  private:
    uint8_t* m_p = nullptr;
    size_t m_n = 0;
    bool m_pinned = false;
};

// gsb_db_create without the copy: the database takes ownership of the chunk buffers (whole rows
// each).  Used by the .fsim reader so that a 128 GB database is not held twice in host memory.
int gsb_db_create_adopt(std::vector<GsbHostBuf>&& chunks, int fp_bits, uint64_t fp_count, gsb_db** out);
