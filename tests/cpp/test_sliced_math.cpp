// CPU emulation of one warp of the bit-sliced multi-query scan (gsb_sliced.cuh) with the shared
// integer helpers: tile layout, transposition, bank-conflict freedom, carry-save counting,
// bit-sliced compare and the filter bound, against plain popcounts.  Exit code 0 = all good.
#include "../../gpusimilarity_b200/csrc/gsb_sliced_math.h"

#include <cstdio>
#include <cstdlib>
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

static uint32_t rd32(const std::vector<uint8_t>& b, size_t off)
{
    uint32_t v;
    std::memcpy(&v, &b[off], 4);
    return v;
}
static void wr32(std::vector<uint8_t>& b, size_t off, uint32_t v)
{
    std::memcpy(&b[off], &v, 4);
}

template <int NP>
static void run_queries(const std::vector<uint8_t>& tile, const std::vector<uint32_t>& rows, std::mt19937& rng,
                        int n_queries, double density, int W = 32)
{
    std::bernoulli_distribution bit(density);
    for (int qi = 0; qi < n_queries; qi++) {
        uint32_t q[32];
        for (int w = 0; w < W; w++) {
            q[w] = 0;
            for (int b = 0; b < 32; b++)
                q[w] |= static_cast<uint32_t>(bit(rng)) << b;
        }
        // list of entries, padded to a multiple of kSlicedGroup with the zero position
        std::vector<uint16_t> list;
        for (int w = 0; w < W; w++)
            for (int b = 0; b < 32; b++)
                if ((q[w] >> b) & 1u)
                    list.push_back(sliced_entry(w * 32 + b));
        const uint32_t pq = static_cast<uint32_t>(list.size());
        while (list.size() % kSlicedGroup)
            list.push_back(sliced_zero_entry(W));
        if (pq >= (1u << (3 + NP)))
            continue; // the caller picks NP from the list length
        // every entry is read by the 32 lanes at once: 32 different banks
        for (uint16_t e : list) {
            std::set<uint32_t> banks;
            for (uint32_t l = 0; l < 32; l++)
                banks.insert(((sliced_lane_base(l, W) + e) / 4) % 32);
            CHECK(banks.size() == 32);
        }
        const uint32_t m = static_cast<uint32_t>(rng() % (pq + 2));
        for (uint32_t l = 0; l < 32; l++) {
            SlicedCount<NP> cnt;
            const size_t base = sliced_lane_base(l, W);
            size_t g = 0;
            for (; g + 16 <= list.size() && (qi & 1); g += 16) { // odd queries: pairs of groups through add16
                uint32_t x[16];
                for (int i = 0; i < 16; i++)
                    x[i] = rd32(tile, base + list[g + i]);
                cnt.add16(x);
            }
            for (; g < list.size(); g += 8) {
                uint32_t x[8];
                for (int i = 0; i < 8; i++)
                    x[i] = rd32(tile, base + list[g + i]);
                cnt.add8(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
            }
            uint32_t want_ge = 0;
            for (uint32_t r = 0; r < 32; r++) {
                uint32_t common = 0;
                for (int w = 0; w < W; w++)
                    common += __builtin_popcount(q[w] & rows[(l * 32 + r) * W + w]);
                CHECK(cnt.column(r) == common);
                if (common >= m)
                    want_ge |= 1u << r;
            }
            if (m < (1u << (3 + NP))) {
                CHECK(cnt.at_least(m) == want_ge);
                CHECK(cnt.at_least_lane(m) == want_ge);
            }
        }
    }
}

// One tile of rows of W words: raw batches as the TMA copies leave them, in-place transposition
// as the kernel does it (lane = word column; gangs of 32 / W batches per warp), then queries.
static void check_width(int W, std::mt19937& rng)
{
    const uint32_t region_bytes = sliced_region_bytes(W), zero_pos = sliced_zero_index(W);
    std::vector<uint8_t> tile(kSlicedTileBatches * region_bytes, 0xAB);
    std::vector<uint32_t> rows(32 * 32 * W);
    std::bernoulli_distribution bit(1.0 / 12);
    for (auto& w : rows) {
        w = 0;
        for (int b = 0; b < 32; b++)
            w |= static_cast<uint32_t>(bit(rng)) << b;
    }
    CHECK(region_bytes % 128 == 0 && region_bytes >= 32u * W * 4 + 64);
    for (uint32_t l = 0; l < 32; l++)
        for (uint32_t r = 0; r < 32; r++) {
            uint32_t pc = 0;
            for (int w = 0; w < W; w++) {
                wr32(tile, l * region_bytes + (r * W + w) * 4, rows[(l * 32 + r) * W + w]);
                pc += __builtin_popcount(rows[(l * 32 + r) * W + w]);
            }
            const uint16_t p16 = static_cast<uint16_t>(pc);
            std::memcpy(&tile[l * region_bytes + 32 * W * 4 + r * 2], &p16, 2);
        }
    for (uint32_t b = 0; b < 32; b++) {
        const size_t region = b * region_bytes;
        std::vector<std::vector<uint32_t>> x(W, std::vector<uint32_t>(32)); // [column][register]
        for (int col = 0; col < W; col++)
            for (uint32_t r = 0; r < 32; r++)
                x[col][r] = rd32(tile, region + (r * W + col) * 4);
        // (all reads happen before any write: __syncwarp in the kernel)
        for (int col = 0; col < W; col++) {
            uint32_t regs[32];
            for (int i = 0; i < 32; i++)
                regs[i] = x[col][i];
            transpose32(regs);
            for (int i = 0; i < 32; i++)
                x[col][i] = regs[i];
        }
        const size_t tbase = sliced_lane_base(b, W);
        CHECK(tbase >= region && tbase + (zero_pos + 1) * 4 <= region + region_bytes);
        for (uint32_t bb = 0; bb < 32; bb++) {
            std::set<uint32_t> banks;
            for (int col = 0; col < W; col++) {
                const uint32_t idx = sliced_word_index(col * 32 + bb);
                CHECK(idx == col * 32 + ((bb + col) & 31));
                banks.insert(((tbase / 4) + idx) % 32);
                wr32(tile, tbase + idx * 4, x[col][bb]);
            }
            CHECK(banks.size() == static_cast<size_t>(W)); // conflict-free stores within a batch
        }
        wr32(tile, tbase + zero_pos * 4, 0u);
    }
    CHECK(sliced_zero_entry(W) == zero_pos * 4);
    // the word index is a bijection on the row's bit positions (the zero word sits right after them)
    {
        std::set<uint32_t> seen;
        for (uint32_t pos = 0; pos < zero_pos; pos++)
            seen.insert(sliced_word_index(pos));
        CHECK(seen.size() == zero_pos && *seen.rbegin() == zero_pos - 1);
    }
    // sparse queries on the small counter, dense ones on the large counter
    run_queries<4>(tile, rows, rng, 40, 1.0 / 30, W);
    run_queries<4>(tile, rows, rng, 10, W >= 16 ? 1.0 / 10 : 1.0 / 4, W);
    run_queries<8>(tile, rows, rng, 10, 0.5, W);
    run_queries<8>(tile, rows, rng, 3, 1.0, W);
    run_queries<8>(tile, rows, rng, 3, 0.0, W);
}

int main()
{
    std::mt19937 rng(12345);
    // ---- transpose32
    {
        uint32_t a[32], b[32];
        for (int i = 0; i < 32; i++)
            a[i] = b[i] = rng();
        transpose32(b);
        for (int r = 0; r < 32; r++)
            for (int c = 0; c < 32; c++)
                CHECK(((b[c] >> r) & 1u) == ((a[r] >> c) & 1u));
    }
    // ---- tile layout, transposition and counting for every supported row width
    for (int W : {32, 16, 8, 4})
        check_width(W, rng);
    // ---- filter bound: smallest count whose best-case score reaches ts
    {
        auto div = [](uint32_t c, uint32_t u) { return u ? static_cast<float>(c) / static_cast<float>(u) : 0.0f / 0.0f; };
        for (uint32_t pq = 0; pq <= 1024; pq += (pq < 80 ? 1 : 37)) {
            for (uint32_t u = 1; u <= 2048; u += (u < 100 ? 1 : 53))
                for (uint32_t c = 0; c <= u; c += (u < 100 ? 1 : 7)) {
                    const float ts = static_cast<float>(c) / static_cast<float>(u);
                    uint32_t want = 0;
                    if (ts > 0.0f) {
                        if (pq == 0)
                            want = 1;
                        else
                            while (want <= pq && !(div(want, pq) >= ts))
                                want++;
                    }
                    CHECK(sliced_filter_min(ts, pq, div) == want);
                }
            CHECK(sliced_filter_min(0.0f, pq, div) == 0);
            const float nan = 0.0f / 0.0f;
            CHECK(sliced_filter_min(nan, pq, div) == 0);
        }
    }
    // ---- per-batch bound: a row that reaches ts is never below sliced_lane_min, for any pd_min <= pd,
    // including the equality case ts == the row's own score; and the bound is not sloppy
    {
        uint64_t loose = 0, total = 0;
        for (uint32_t pq = 1; pq <= 1024; pq += (pq < 70 ? 1 : 61))
            for (uint32_t pd = 0; pd <= 1024; pd += (pd < 70 ? 1 : 67))
                for (uint32_t c = 0; c <= (pq < pd ? pq : pd); c += (c < 40 ? 1 : 13)) {
                    const uint32_t u = pq + pd - c;
                    if (u == 0)
                        continue;
                    const float ts = static_cast<float>(c) / static_cast<float>(u);
                    const float tq = sliced_tq(ts);
                    for (uint32_t pdmin = pd > 3 ? pd - 3 : 0; pdmin <= pd; pdmin++) {
                        const uint32_t m = sliced_lane_min(tq, static_cast<float>(pq + pdmin));
                        CHECK(m <= c);
                        if (pdmin == pd) {
                            total++;
                            loose += (m + 1 < c);
                        }
                    }
                    // a slightly lower threshold keeps the row, a slightly higher one may drop it
                    CHECK(sliced_lane_min(sliced_tq(ts * 0.999f), static_cast<float>(pq + pd)) <= c);
                }
        CHECK(loose == 0); // with the true popcount the bound is within one count of exact
        CHECK(sliced_lane_min(sliced_tq(0.0f), 100.0f) == 0 && sliced_lane_min(sliced_tq(-1.0f), 100.0f) == 0);
        (void) total;
    }
    // ---- histogram buckets: monotone in the score, floor of a bucket maps back into it
    {
        uint32_t prev = 0;
        for (uint32_t u = 1; u <= 2048; u += 3)
            for (uint32_t c = 0; c <= u; c += 1 + u / 64) {
                const float a = static_cast<float>(c) / static_cast<float>(u);
                uint32_t bits;
                std::memcpy(&bits, &a, 4);
                const uint32_t b = sliced_bucket(bits);
                CHECK(b < kSlicedHistBuckets);
                if (b >= 1) {
                    CHECK(sliced_bucket_floor_bits(b) <= bits && sliced_bucket(sliced_bucket_floor_bits(b)) == b);
                    CHECK(sliced_bucket(sliced_bucket_floor_bits(b) - 1) == b - 1);
                }
                (void) prev;
            }
        const float one = 1.0f, tiny = 0.015f;
        uint32_t b1, b2;
        std::memcpy(&b1, &one, 4);
        std::memcpy(&b2, &tiny, 4);
        CHECK(sliced_bucket(b1) == 769 && sliced_bucket(b2) == 0 && sliced_bucket(0) == 0);
        CHECK(sliced_bucket(0x7fc00000u) == kSlicedHistBuckets - 1); // never out of range
    }
    if (g_fail) {
        std::fprintf(stderr, "%d check(s) failed\n", g_fail);
        return 1;
    }
    std::printf("sliced math ok\n");
    return 0;
}
