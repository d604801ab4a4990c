// Internal C++ seam between the translation units of libgpusim_b200.so (not part of the C ABI).
#pragma once

#include <cstdint>
#include <vector>

struct gsb_db;

// gsb_db_create without the copy: the database takes ownership of the chunk buffers (whole rows
// each).  Used by the .fsim reader so that a 128 GB database is not held twice in host memory.
int gsb_db_create_adopt(std::vector<std::vector<uint8_t>>&& chunks, int fp_bits, uint64_t fp_count, gsb_db** out);
