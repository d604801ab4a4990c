// Pure integer helpers of the bit-sliced multi-query scan (gsb_sliced.cuh).  Host and device:
// tests/cpp/test_sliced_math.cpp emulates a warp with them on the CPU (layout, swizzle, transpose,
// carry-save counting, bit-sliced compare) against plain popcounts.
//
// The idea: the scores of one query against 32 rows need common = popc(q & d) per row (reference
// TanimotoFunctor, fingerprintdb_cuda.cu:97-98).  With a 32-row batch stored bit-transposed —
// word T[pos] holds bit `pos` of each of the 32 rows — common for all 32 rows at once is the
// column sum of the words T[pos] over the SET bits of q only (~40 of 1024 for Morgan
// fingerprints), accumulated with carry-save adders into bit-sliced counters: a handful of LOP3
// per set bit and 32 rows instead of 32 AND + 32 POPC per row.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define GSB_HD __host__ __device__ __forceinline__
#else
#define GSB_HD inline
#endif
#ifdef __CUDA_ARCH__
#define GSB_UNROLL _Pragma("unroll")
#else
#define GSB_UNROLL
#endif

namespace gsb
{

constexpr uint32_t kSlicedTileBatches = 32;   // 32-row batches per tile: lane l of a warp works on batch l
// Shared memory per batch for rows of `words` 32-bit words: words*32 + 1 transposed words (the last
// one is always zero: list padding) + up to 124 bytes of lane skew, in whole 128-byte lines (so that
// every region starts in bank 0); never smaller than the raw batch the TMA copy delivers.
GSB_HD constexpr uint32_t sliced_region_bytes(uint32_t words)
{
    return ((words * 32u + 1u) * 4u + 124u + 127u) / 128u * 128u;
}
// Word index (inside a lane's transposed batch) of the word that is always zero, and the list
// entry that points at it (padding of the lists to whole groups).
GSB_HD constexpr uint32_t sliced_zero_index(uint32_t words)
{
    return words * 32u;
}
GSB_HD constexpr uint16_t sliced_zero_entry(uint32_t words)
{
    return static_cast<uint16_t>(words * 32u * 4u);
}
constexpr uint32_t kSlicedRegionBytes = sliced_region_bytes(32); // 1024-bit rows: 4224
static_assert(kSlicedRegionBytes == 4224 && sliced_region_bytes(16) == 2176 && sliced_region_bytes(8) == 1152 &&
                  sliced_region_bytes(4) == 640,
              "regions hold the raw batch (32 rows + 64 bytes of popcounts) and the skewed transposed batch");
constexpr uint32_t kSlicedGroup = 8;          // list entries consumed per carry-save round

// Word index of bit position `pos` (0 .. row bits - 1) inside a lane's transposed batch.  The rotation by
// the word column (pos >> 5) makes the transposition's stores conflict free (lane = column writes
// 32 words that land in 32 banks); lane l's batch starts 4*l bytes into its region, so the loads
// of one position by the 32 lanes of a warp hit 32 different banks as well.
GSB_HD uint32_t sliced_word_index(uint32_t pos)
{
    return (pos & ~31u) | ((pos + (pos >> 5)) & 31u);
}
// List entry: byte offset of the position's word from the lane's batch base.
GSB_HD uint16_t sliced_entry(uint32_t pos)
{
    return static_cast<uint16_t>(sliced_word_index(pos) * 4u);
}
// Byte offset of lane l's transposed batch inside the tile buffer.
GSB_HD uint32_t sliced_lane_base(uint32_t l, uint32_t words = 32)
{
    return l * sliced_region_bytes(words) + 4u * l;
}

// Byte permute (PRMT): result byte i = byte (sel >> 4i) & 7 of the pair {y, x} (x = bytes 0-3).
GSB_HD uint32_t byte_perm(uint32_t x, uint32_t y, uint32_t sel)
{
#ifdef __CUDA_ARCH__
    return __byte_perm(x, y, sel);
#else
    const uint64_t pair = (static_cast<uint64_t>(y) << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
        r |= static_cast<uint32_t>((pair >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}
template <int J, uint32_t M> GSB_HD void transpose32_stage(uint32_t (&a)[32])
{
GSB_UNROLL
    for (int k0 = 0; k0 < 32; k0 += 2 * J) {
GSB_UNROLL
        for (int i = 0; i < J; i++) {
            const int k = k0 + i;
            if (J == 16) { // whole 16-bit halves and bytes move with one byte permute per word
                const uint32_t lo = byte_perm(a[k], a[k + J], 0x5410), hi = byte_perm(a[k], a[k + J], 0x7632);
                a[k] = lo;
                a[k + J] = hi;
            } else if (J == 8) {
                const uint32_t lo = byte_perm(a[k], a[k + J], 0x6240), hi = byte_perm(a[k], a[k + J], 0x7351);
                a[k] = lo;
                a[k + J] = hi;
            } else {
                const uint32_t t = ((a[k] >> J) ^ a[k + J]) & M;
                a[k + J] ^= t;
                a[k] ^= t << J;
            }
        }
    }
}
// 32x32 bit-matrix transpose in registers: afterwards bit r of a[b] is bit b of the old a[r].
GSB_HD void transpose32(uint32_t (&a)[32])
{
    transpose32_stage<16, 0x0000ffffu>(a);
    transpose32_stage<8, 0x00ff00ffu>(a);
    transpose32_stage<4, 0x0f0f0f0fu>(a);
    transpose32_stage<2, 0x33333333u>(a);
    transpose32_stage<1, 0x55555555u>(a);
}

// Carry-save adder over 32 independent bit columns: a + b + c = 2*h + l.
GSB_HD void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c)
{
    const uint32_t u = a ^ b;
    h = (a & b) | (u & c);
    l = u ^ c;
}

// Running bit-sliced counters of one lane: column r of (ones, twos, fours, hi[0..NP)) is the number
// of words added so far that had bit r set.
template <int NP> struct SlicedCount {
    static constexpr int kPlanes = 3 + NP;
    uint32_t ones = 0, twos = 0, fours = 0;
    uint32_t hi[NP];
    GSB_HD SlicedCount()
    {
        for (int i = 0; i < NP; i++)
            hi[i] = 0;
    }
    // add eight words (Harley-Seal round: 7 carry-save adders, then the carry out ripples into hi)
    GSB_HD void add8(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4, uint32_t x5, uint32_t x6,
                     uint32_t x7)
    {
        uint32_t ta, tb, fa, fb, eights;
        csa(ta, ones, ones, x0, x1);
        csa(tb, ones, ones, x2, x3);
        csa(fa, twos, twos, ta, tb);
        csa(ta, ones, ones, x4, x5);
        csa(tb, ones, ones, x6, x7);
        csa(fb, twos, twos, ta, tb);
        csa(eights, fours, fours, fa, fb);
        uint32_t carry = eights;
GSB_UNROLL
        for (int i = 0; i < NP; i++) {
            const uint32_t t = hi[i] & carry;
            hi[i] ^= carry;
            carry = t;
        }
    }
    // add sixteen words: two rounds whose carries meet in one more adder, so that only one carry
    // (weight 16) ripples into the upper planes — 36 logic ops instead of 44
    GSB_HD void add16(const uint32_t (&x)[16])
    {
        uint32_t ta, tb, fa, fb, ea, eb, sixteens;
        csa(ta, ones, ones, x[0], x[1]);
        csa(tb, ones, ones, x[2], x[3]);
        csa(fa, twos, twos, ta, tb);
        csa(ta, ones, ones, x[4], x[5]);
        csa(tb, ones, ones, x[6], x[7]);
        csa(fb, twos, twos, ta, tb);
        csa(ea, fours, fours, fa, fb);
        csa(ta, ones, ones, x[8], x[9]);
        csa(tb, ones, ones, x[10], x[11]);
        csa(fa, twos, twos, ta, tb);
        csa(ta, ones, ones, x[12], x[13]);
        csa(tb, ones, ones, x[14], x[15]);
        csa(fb, twos, twos, ta, tb);
        csa(eb, fours, fours, fa, fb);
        csa(sixteens, hi[0], hi[0], ea, eb);
        uint32_t carry = sixteens;
GSB_UNROLL
        for (int i = 1; i < NP; i++) {
            const uint32_t t = hi[i] & carry;
            hi[i] ^= carry;
            carry = t;
        }
    }
    GSB_HD uint32_t plane(int p) const
    {
        return p == 0 ? ones : (p == 1 ? twos : (p == 2 ? fours : hi[p - 3]));
    }
    // columns whose count is >= m (m < 2^(3+NP)): bit-sliced compare against a constant
    GSB_HD uint32_t at_least(uint32_t m) const
    {
        uint32_t gt = 0, eq = ~0u;
GSB_UNROLL
        for (int p = 3 + NP - 1; p >= 0; p--) {
            const uint32_t c = plane(p);
            if ((m >> p) & 1u) {
                eq &= c;
            } else {
                gt |= eq & c;
                eq &= ~c;
            }
        }
        return gt | eq;
    }
    // the same with a different m in every lane and no branches: borrow chain of (count - m),
    // one three-input logic op per plane; a column is >= m iff no borrow comes out of the top
    GSB_HD uint32_t at_least_lane(uint32_t m) const
    {
        uint32_t borrow = 0;
GSB_UNROLL
        for (int p = 0; p < 3 + NP; p++) {
            const uint32_t c = plane(p);
            const uint32_t mb = static_cast<uint32_t>(static_cast<int32_t>(m << (31 - p)) >> 31); // bit p of m, everywhere
            borrow = (~c & mb) | (~(c ^ mb) & borrow);
        }
        return ~borrow;
    }
    // count of column r
    GSB_HD uint32_t column(uint32_t r) const
    {
        uint32_t c = 0;
GSB_UNROLL
        for (int p = 0; p < 3 + NP; p++)
            c |= ((plane(p) >> r) & 1u) << p;
        return c;
    }
};

// Smallest common-bit count that can still reach score `ts` for a query of `pq` set bits: the
// union is at least pq, so score <= div(common, pq); rows below the returned count are skipped
// without being scored.  `div` is the kernel's correctly rounded division (monotone in common).
// Returns pq + 1 (or more) when nothing can pass.
template <class Div> GSB_HD uint32_t sliced_filter_min(float ts, uint32_t pq, Div div)
{
    if (!(ts > 0.0f))
        return 0;
    if (pq == 0)
        return 1; // every score is 0/0 -> 0 (fingerprintdb_cuda.cu:102), below any positive ts
    const float est = ts * static_cast<float>(pq);
    uint32_t m = est >= 2.0f ? static_cast<uint32_t>(est) - 1u : 0u;
    if (m > pq)
        m = pq;
    while (m <= pq && !(div(m, pq) >= ts))
        m++;
    return m;
}

// Tighter, per-batch bound: every row of a batch has at least pd_min set bits, the union is
// pq + pd - common, so score >= ts needs common >= ts/(1+ts) * (pq + pd_min).  sliced_tq() is that
// factor rounded DOWN (by far more than the rounding of the score and of this arithmetic), and
// sliced_lane_min() the resulting count, never above the exact bound.
GSB_HD float sliced_tq(float ts)
{
    if (!(ts > 0.0f))
        return 0.0f;
    return ts / (1.0f + ts) * (1.0f - 1.0f / 1048576.0f);
}
GSB_HD uint32_t sliced_lane_min(float tq, float pq_plus_pdmin)
{
    const float x = tq * pq_plus_pdmin - 0.001f; // the product is off by < 2^-13 for sums up to 2048
#ifdef __CUDA_ARCH__
    return __float2uint_ru(x); // ceil; saturating: negative values and NaN give 0
#else
    if (!(x > 0.0f))
        return 0u;
    const uint32_t f = static_cast<uint32_t>(x);
    return static_cast<float>(f) < x ? f + 1u : f; // ceil
#endif
}

// Global per-query score histogram (threshold sharing between CTAs): monotone bucket of a score's
// float bits with 7 mantissa bits kept; everything below 2^-6 shares bucket 0, 1.0 is bucket 769.
constexpr uint32_t kSlicedHistBuckets = 1024;
constexpr uint32_t kSlicedHistShift = 16;
constexpr uint32_t kSlicedHistBase = (0x3c800000u >> kSlicedHistShift) - 1u;
GSB_HD uint32_t sliced_bucket(uint32_t score_bits)
{
    const uint32_t b = score_bits >> kSlicedHistShift;
    return b > kSlicedHistBase ? (b - kSlicedHistBase < kSlicedHistBuckets ? b - kSlicedHistBase : kSlicedHistBuckets - 1u) : 0u;
}
// smallest score bits that fall into bucket b (b >= 1)
GSB_HD uint32_t sliced_bucket_floor_bits(uint32_t b)
{
    return (b + kSlicedHistBase) << kSlicedHistShift;
}

} // namespace gsb
