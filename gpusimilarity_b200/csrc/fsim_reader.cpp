// Qt-free reader for gpusimilarity's .fsim database files — the native equivalent of
// GPUSimServer::extractData + Decompress*Runnable (reference gpusim.cpp:48-85, 173-253).
//
// On-disk format (big-endian QDataStream; see gpusimilarity_b200/fsim.py for the full table):
//   i32 version(==3), cstr dbkey, i32 fp_bitcount, i32 fp_count,
//   i32 n; n x QByteArray(qCompress(raw fingerprints))
//   i32 n; n x QByteArray(qCompress(cstr...))   SMILES
//   i32 n; n x QByteArray(qCompress(cstr...))   ids
// qCompress = u32 BE uncompressed length + zlib stream.  Chunks are inflated in parallel (the
// reference uses a QThreadPool, gpusim.cpp:202-236).
#include "../../include/gpusim_b200.h"
#include "gsb_internal.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace
{
thread_local std::string g_fsim_err;

struct Blob {
    std::vector<uint8_t> compressed;
    GsbHostBuf raw;            // fingerprint chunks: pinned when a CUDA device is present, the upload DMAs straight out of it
    std::vector<uint8_t> text; // SMILES / id chunks: plain memory
    bool is_fp = false;
    std::string error;
};

struct Cursor {
    const std::vector<uint8_t>& buf;
    size_t off = 0;
    bool ok = true;
    explicit Cursor(const std::vector<uint8_t>& b) : buf(b) {}
    uint32_t u32()
    {
        if (off + 4 > buf.size()) {
            ok = false;
            return 0;
        }
        const uint8_t* p = buf.data() + off;
        off += 4;
        return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | uint32_t(p[3]);
    }
    // QByteArray / char*: u32 length (0xFFFFFFFF = null) + bytes
    bool bytes(const uint8_t** p, size_t* n)
    {
        uint32_t len = u32();
        if (!ok)
            return false;
        if (len == 0xFFFFFFFFu)
            len = 0;
        if (off + len > buf.size()) {
            ok = false;
            return false;
        }
        *p = buf.data() + off;
        *n = len;
        off += len;
        return true;
    }
};

void inflate_blob(Blob* b)
{
    if (b->compressed.size() < 4) {
        b->raw.release();
        b->text.clear();
        return;
    }
    const uint8_t* p = b->compressed.data();
    const uLongf expected = (uLongf(p[0]) << 24) | (uLongf(p[1]) << 16) | (uLongf(p[2]) << 8) | uLongf(p[3]);
    uint8_t* dst;
    if (b->is_fp) { // inflate straight into the (pinned) buffer the upload will DMA from
        if (!b->raw.allocate(expected)) {
            b->error = "out of host memory";
            return;
        }
        dst = b->raw.data();
    } else {
        b->text.resize(expected);
        dst = b->text.data();
    }
    uLongf got = expected;
    const int rc = uncompress(dst, &got, p + 4, static_cast<uLong>(b->compressed.size() - 4));
    if (rc != Z_OK || got != expected)
        b->error = "qUncompress failed (zlib rc " + std::to_string(rc) + ")";
    std::vector<uint8_t>().swap(b->compressed);
}

// sequence of cstr (u32 length incl. NUL + bytes) -> offsets into the raw buffer
void split_strings(const std::vector<uint8_t>& raw, std::vector<const char*>* out, std::string* err)
{
    Cursor c(raw);
    while (c.off < raw.size()) {
        const uint8_t* p;
        size_t n;
        if (!c.bytes(&p, &n)) {
            *err = "truncated string block";
            return;
        }
        if (n && p[n - 1] != 0) { // every entry is handed out as a C string: it must end inside its bytes
            *err = "string entry without terminating NUL";
            return;
        }
        out->push_back(n ? reinterpret_cast<const char*>(p) : "");
    }
}
} // namespace

struct gsb_fsim {
    std::string dbkey;
    int fp_bits = 0;
    uint64_t fp_count = 0;
    std::vector<Blob> fp, smi, ids;
    std::vector<const char*> smiles_ptrs, id_ptrs;
};

extern "C" {

const char* gsb_fsim_last_error(void) { return g_fsim_err.c_str(); }

int gsb_fsim_open(const char* path, gsb_fsim** out)
{
    try {
    if (!path || !out) {
        g_fsim_err = "null argument";
        return GSB_ERR_INVALID;
    }
    std::ifstream in(path, std::ios::binary | std::ios::ate);
    if (!in) {
        g_fsim_err = std::string("cannot open ") + path;
        return GSB_ERR_IO;
    }
    std::vector<uint8_t> file(static_cast<size_t>(in.tellg()));
    in.seekg(0);
    in.read(reinterpret_cast<char*>(file.data()), static_cast<std::streamsize>(file.size()));
    Cursor c(file);
    const int32_t version = static_cast<int32_t>(c.u32());
    if (!c.ok || version != 3) { // reference gpusim.cpp:186-189
        g_fsim_err = "Database version incompatible with this GPUSim version";
        return GSB_ERR_IO;
    }
    std::unique_ptr<gsb_fsim> f(new gsb_fsim);
    const uint8_t* p;
    size_t n;
    if (!c.bytes(&p, &n)) {
        g_fsim_err = "truncated header";
        return GSB_ERR_IO;
    }
    f->dbkey.assign(reinterpret_cast<const char*>(p), n && p[n - 1] == 0 ? n - 1 : n);
    f->fp_bits = static_cast<int32_t>(c.u32());
    f->fp_count = static_cast<uint64_t>(static_cast<int32_t>(c.u32()));
    for (std::vector<Blob>* group : {&f->fp, &f->smi, &f->ids}) {
        const int32_t count = static_cast<int32_t>(c.u32());
        if (!c.ok || count < 0) {
            g_fsim_err = "truncated chunk table";
            return GSB_ERR_IO;
        }
        group->resize(count);
        for (Blob& b : *group) {
            b.is_fp = group == &f->fp;
            if (!c.bytes(&p, &n)) {
                g_fsim_err = "truncated chunk";
                return GSB_ERR_IO;
            }
            b.compressed.assign(p, p + n);
        }
    }
    std::vector<uint8_t>().swap(file);
    // inflate every chunk on its own thread (bounded by the core count)
    std::vector<Blob*> all;
    for (std::vector<Blob>* group : {&f->fp, &f->smi, &f->ids})
        for (Blob& b : *group)
            all.push_back(&b);
    const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), all.size()));
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; t++)
        pool.emplace_back([&all, t, nt]() {
            for (size_t i = t; i < all.size(); i += nt)
                inflate_blob(all[i]);
        });
    for (auto& th : pool)
        th.join();
    for (Blob* b : all)
        if (!b->error.empty()) {
            g_fsim_err = b->error;
            return GSB_ERR_IO;
        }
    std::string err;
    for (Blob& b : f->smi)
        split_strings(b.text, &f->smiles_ptrs, &err);
    for (Blob& b : f->ids)
        split_strings(b.text, &f->id_ptrs, &err);
    if (!err.empty()) {
        g_fsim_err = err;
        return GSB_ERR_IO;
    }
    // a row without its SMILES / id would crash the server at query time: fail at load instead
    if (f->smiles_ptrs.size() != f->fp_count || f->id_ptrs.size() != f->fp_count) {
        g_fsim_err = "SMILES / id counts (" + std::to_string(f->smiles_ptrs.size()) + " / " +
                     std::to_string(f->id_ptrs.size()) + ") do not match the fingerprint count " +
                     std::to_string(f->fp_count) + ", potential database corruption.";
        return GSB_ERR_CORRUPT;
    }
    *out = f.release();
    return GSB_OK;
    } catch (const std::bad_alloc&) {
        g_fsim_err = "out of host memory";
        return GSB_ERR_NOMEM;
    } catch (const std::exception& e) {
        g_fsim_err = std::string("exception: ") + e.what();
        return GSB_ERR_IO;
    }
}

void gsb_fsim_close(gsb_fsim* f) { delete f; }
const char* gsb_fsim_dbkey(const gsb_fsim* f) { return f->dbkey.c_str(); }
int gsb_fsim_fp_bits(const gsb_fsim* f) { return f->fp_bits; }
uint64_t gsb_fsim_fp_count(const gsb_fsim* f) { return f->fp_count; }
int gsb_fsim_chunk_count(const gsb_fsim* f) { return static_cast<int>(f->fp.size()); }
const void* gsb_fsim_chunk_data(const gsb_fsim* f, int i) { return f->fp[i].raw.data(); }
uint64_t gsb_fsim_chunk_bytes(const gsb_fsim* f, int i) { return f->fp[i].raw.size(); }
uint64_t gsb_fsim_string_count(const gsb_fsim* f, int which)
{
    return which == 0 ? f->smiles_ptrs.size() : f->id_ptrs.size();
}
const char* gsb_fsim_string(const gsb_fsim* f, int which, uint64_t index)
{
    const auto& v = which == 0 ? f->smiles_ptrs : f->id_ptrs;
    return index < v.size() ? v[index] : nullptr;
}

int gsb_fsim_create_db(gsb_fsim* f, gsb_db** out)
{
    // the database adopts the inflated fingerprint chunks (no second copy of a 100 GB database in
    // host memory); the file object keeps the SMILES / id strings
    std::vector<GsbHostBuf> chunks;
    for (Blob& b : f->fp)
        chunks.push_back(std::move(b.raw));
    f->fp.clear();
    return gsb_db_create_adopt(std::move(chunks), f->fp_bits, f->fp_count, out);
}

} // extern "C"
