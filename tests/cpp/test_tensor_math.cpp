// CPU emulation of the operands of the tensor-core multi-query scan (gsb_tensor.cuh) with the
// shared helpers of gsb_tensor_math.h: the expanders' slab stores (B operand), the query staging
// (A operand, tensor-memory columns) and the way tcgen05.mma reads both (canonical K-major layout
// without swizzle: core matrix = 8 rows x 16 bytes, stride byte offset between 8-row groups,
// leading byte offset between the two 16-byte chunks of a 32-byte K-step), against plain
// popcounts; bank-conflict freedom of the stores; the epilogue filter against the exact bound.
// Exit code 0 = all good.
#include "../../gpusimilarity_b200/csrc/gsb_sliced_math.h"
#include "../../gpusimilarity_b200/csrc/gsb_tensor_math.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <set>
#include <vector>

using namespace gsb;

static int g_fail = 0;
#define CHECK(c)                                                                                 \
    do {                                                                                         \
        if (!(c)) {                                                                              \
            if (g_fail++ < 20)                                                                   \
                std::fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, #c);                 \
        }                                                                                        \
    } while (0)

static float as_float(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

int main()
{
    std::mt19937 rng(12345);
    for (double density : {0.03, 0.2, 0.5, 1.0}) {
        std::bernoulli_distribution bit(density);
        std::vector<uint32_t> rows(kTcTileRows * 32), queries(kTcQueries * 32);
        for (auto& w : rows) {
            w = 0;
            for (int b = 0; b < 32; b++)
                w |= static_cast<uint32_t>(bit(rng)) << b;
        }
        for (auto& w : queries) {
            w = 0;
            for (int b = 0; b < 32; b++)
                w |= static_cast<uint32_t>(bit(rng)) << b;
        }
        // ---- B operand: the four slabs as the expander warps write them (warp e: rows 16e..16e+15, lane = row word)
        std::vector<std::vector<uint8_t>> slab(kTcSlabs, std::vector<uint8_t>(kTcSlabBytes, 0xee));
        for (uint32_t g = 0; g < kTcSlabs; g++)
            for (uint32_t e = 0; e < 8; e++)
                for (uint32_t i = 0; i < 16; i++) {
                    std::set<uint32_t> banks[2]; // an 8-byte store is served per half warp
                    for (uint32_t lane = 0; lane < 32; lane++) {
                        const uint32_t r = e * 16 + i, w = rows[r * 32 + lane];
                        const uint32_t off = tc_slab_offset(e * 16, lane) + (i >> 3) * kTcSbo + (i & 7) * 16;
                        CHECK(off == tc_slab_offset(r, lane));
                        CHECK(off + 8 <= kTcSlabBytes);
                        const uint32_t v0 = tc_row_word(w, 2 * g), v1 = tc_row_word(w, 2 * g + 1);
                        std::memcpy(&slab[g][off], &v0, 4);
                        std::memcpy(&slab[g][off + 4], &v1, 4);
                        banks[lane / 16].insert((off / 4) % 32);
                        banks[lane / 16].insert((off / 4 + 1) % 32);
                    }
                    CHECK(banks[0].size() == 32 && banks[1].size() == 32);
                }
        // ---- A operand: 256 32-bit columns per query lane
        std::vector<uint32_t> a_cols(kTcQueries * 256);
        for (uint32_t q = 0; q < kTcQueries; q++)
            for (uint32_t ks = 0; ks < 32; ks++)
                for (uint32_t u = 0; u < 8; u++) {
                    uint32_t iw, plane;
                    tc_a_source(ks, u, &iw, &plane);
                    a_cols[q * 256 + ks * 8 + u] = tc_query_word(queries[q * 32 + iw], plane);
                }
        // ---- the MMA: K-step ks = 8 g + s reads 32 bytes of K per row / query
        for (uint32_t q = 0; q < kTcQueries; q += 7)
            for (uint32_t r = 0; r < kTcTileRows; r++) {
                int64_t d = 0;
                for (uint32_t g = 0; g < kTcSlabs; g++)
                    for (uint32_t s = 0; s < kTcSlabSteps; s++)
                        for (uint32_t kb = 0; kb < 32; kb++) {
                            // B: descriptor start = slab + 2 s LBO; chunk kb / 16 at + LBO; row group at + SBO
                            const uint32_t boff = (2 * s + kb / 16) * kTcLbo + (r / 8) * kTcSbo + (r % 8) * 16 + kb % 16;
                            const uint32_t b = slab[g][boff];
                            const uint32_t a = (a_cols[q * 256 + (g * 8 + s) * 8 + kb / 4] >> (8 * (kb % 4))) & 0xffu;
                            d += static_cast<int64_t>(a) * b;
                        }
                uint32_t common = 0;
                for (uint32_t w = 0; w < 32; w++)
                    common += __builtin_popcount(rows[r * 32 + w] & queries[q * 32 + w]);
                CHECK(d == 128 * static_cast<int64_t>(common));
            }
    }
    // ---- the filter never rejects a pair the exact comparison accepts
    auto div = [](uint32_t c, uint32_t u) { return u ? static_cast<float>(c) / static_cast<float>(u) : 0.0f; };
    uint64_t accepted = 0, exact_ok = 0;
    for (float ts : {0.0f, 1e-6f, 0.013f, 0.1f, 0.25f, 0.3333333f, 0.5f, 0.77f, 0.999f, 1.0f}) {
        const float tq = sliced_tq(ts), slope = tc_filter_slope(tq);
        for (uint32_t pq : {0u, 1u, 17u, 48u, 100u, 512u, 1024u}) {
            const float thr = tc_filter_threshold(tq, pq);
            for (uint32_t pd = 0; pd <= 1024; pd += (pd < 80 ? 1 : 13))
                for (uint32_t c = 0; c <= (pq < pd ? pq : pd); c++) {
                    const float t = std::fma(slope, static_cast<float>(pd), as_float(kTcMagicBits | (128u * c)));
                    const bool pass = t >= thr;
                    const bool exact = div(c, pq + pd - c) >= ts;
                    if (exact) {
                        exact_ok++;
                        CHECK(pass);
                    }
                    accepted += pass;
                }
        }
    }
    // and it is tight: at a threshold of 0.5 it lets through about what the exact test lets through
    {
        const float ts = 0.5f, tq = sliced_tq(ts), slope = tc_filter_slope(tq);
        uint64_t pass_n = 0, exact_n = 0;
        for (uint32_t pq = 20; pq < 200; pq += 7)
            for (uint32_t pd = 20; pd < 200; pd += 3)
                for (uint32_t c = 0; c <= (pq < pd ? pq : pd); c++) {
                    pass_n += std::fma(slope, static_cast<float>(pd), as_float(kTcMagicBits | (128u * c))) >= tc_filter_threshold(tq, pq);
                    exact_n += div(c, pq + pd - c) >= ts;
                }
        CHECK(pass_n >= exact_n && pass_n <= exact_n + exact_n / 50 + 400);
        std::printf("filter at 0.5: %llu pass, %llu exact\n", (unsigned long long) pass_n, (unsigned long long) exact_n);
    }
    std::printf("%s (%llu accepted, %llu exact)\n", g_fail ? "FAILED" : "ok", (unsigned long long) accepted,
                (unsigned long long) exact_ok);
    return g_fail ? 1 : 0;
}
