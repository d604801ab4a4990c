// Qt-free equivalent of the reference's GPUSimServer (gpusim.h:23-95, gpusim.cpp:87-461): load
// .fsim databases, pick the fold factor, answer similarity searches over the local socket
// "gpusimilarity" with the reference's QDataStream wire format, merge several databases and
// de-duplicate identical SMILES.  Host-side only; all scoring goes through the C ABI.
//
// Wire format (big-endian QDataStream, default stream version => `float` travels as an 8-byte
// double):
//   request : i32 n_db; n_db x (cstr dbname, cstr dbkey); i32 request_num; i32 results_requested;
//             f64 cutoff; QByteArray fingerprint                      (gpusim.cpp:384-414)
//   response: i32 request_num; i32 n; u64 approximate_count;
//             n x cstr smiles; n x cstr ids; n x f64 score            (gpusim.cpp:431-453)
#include "../../include/gpusim_b200.h"

#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{
thread_local std::string g_srv_err;
thread_local int g_last_rc = GSB_OK; // code of the latest failed search (the serve loop recovers from GSB_ERR_CUDA)

struct Database {
    std::string name, key;
    gsb_fsim* file = nullptr; // owns the SMILES / id strings
    gsb_db* db = nullptr;
    ~Database()
    {
        gsb_db_destroy(db);
        gsb_fsim_close(file);
    }
};

struct Result {
    float score;
    const char* smiles;
    const char* id;
};

// ---- QDataStream primitives -------------------------------------------------------------
struct Reader {
    const uint8_t* p;
    size_t n, off = 0;
    bool ok = true;
    bool need(size_t k)
    {
        if (off + k > n)
            ok = false;
        return ok;
    }
    uint32_t u32()
    {
        if (!need(4))
            return 0;
        const uint8_t* q = p + off;
        off += 4;
        return (uint32_t(q[0]) << 24) | (uint32_t(q[1]) << 16) | (uint32_t(q[2]) << 8) | uint32_t(q[3]);
    }
    int32_t i32() { return static_cast<int32_t>(u32()); }
    double f64()
    {
        if (!need(8))
            return 0;
        uint64_t v = 0;
        for (int i = 0; i < 8; i++)
            v = (v << 8) | p[off + i];
        off += 8;
        double d;
        std::memcpy(&d, &v, 8);
        return d;
    }
    std::string bytes()
    {
        uint32_t len = u32();
        if (len == 0xFFFFFFFFu)
            len = 0;
        if (!need(len))
            return std::string();
        std::string s(reinterpret_cast<const char*>(p + off), len);
        off += len;
        return s;
    }
    std::string cstr()
    {
        std::string s = bytes();
        if (!s.empty() && s.back() == '\0')
            s.pop_back();
        return s;
    }
};

struct Writer {
    std::vector<uint8_t> buf;
    void u32(uint32_t v)
    {
        for (int s = 24; s >= 0; s -= 8)
            buf.push_back(static_cast<uint8_t>(v >> s));
    }
    void u64(uint64_t v)
    {
        for (int s = 56; s >= 0; s -= 8)
            buf.push_back(static_cast<uint8_t>(v >> s));
    }
    void f64(double d)
    {
        uint64_t v;
        std::memcpy(&v, &d, 8);
        u64(v);
    }
    void cstr(const char* s)
    {
        const size_t len = std::strlen(s) + 1;
        u32(static_cast<uint32_t>(len));
        buf.insert(buf.end(), s, s + len);
    }
};

std::string base_name(const std::string& path)
{
    // QFileInfo::baseName(): file name up to (not including) the first '.'
    const size_t slash = path.find_last_of('/');
    const std::string file = slash == std::string::npos ? path : path.substr(slash + 1);
    return file.substr(0, file.find('.'));
}
} // namespace

struct gsb_server {
    std::map<std::string, std::unique_ptr<Database>> dbs; // QHash in the reference; order is irrelevant there
    bool use_gpu = true;
    unsigned fold_factor = 1;
    int listen_fd = -1;
    std::string socket_path;
    std::atomic<bool> stop{false};

    bool using_gpu() const { return use_gpu && gsb_device_count() != 0; } // gpusim.cpp:168-171

    // GPUSimServer::similaritySearch (gpusim.cpp:276-293) on one database, results as pointers
    int similarity_search(Database& d, const int32_t* q, int n_words, const std::string& key, unsigned k,
                          float cutoff, bool gpu, std::vector<Result>* out, uint64_t* approx)
    {
        *approx = 0;
        if (key != d.key) // fingerprintdb_cuda.cu:349-352: silently empty
            return GSB_OK;
        // the request's result count comes straight from the socket: never size buffers by it alone
        k = static_cast<unsigned>(std::min<uint64_t>(k, gsb_db_count(d.db)));
        std::vector<uint32_t> rows(k ? k : 1);
        std::vector<float> scores(rows.size());
        uint32_t n = 0;
        int rc;
        if (gpu)
            rc = gsb_db_search(d.db, q, n_words, k, cutoff, rows.data(), scores.data(), &n, approx);
        else
            rc = gsb_db_search_cpu(d.db, q, n_words, k, rows.data(), scores.data(), &n);
        if (rc != GSB_OK) {
            g_srv_err = gsb_last_error();
            g_last_rc = rc;
            return rc;
        }
        for (uint32_t i = 0; i < n; i++) {
            const char* smi = gsb_fsim_string(d.file, 0, rows[i]);
            const char* id = gsb_fsim_string(d.file, 1, rows[i]);
            out->push_back({scores[i], smi ? smi : "", id ? id : ""});
        }
        return GSB_OK;
    }

    // GPUSimServer::searchDatabases (gpusim.cpp:306-374)
    int search_databases(const int32_t* q, int n_words, int results_requested, float cutoff,
                         const std::vector<std::pair<std::string, std::string>>& name_key,
                         std::vector<std::string>* smiles, std::vector<std::string>* ids, std::vector<float>* scores,
                         uint64_t* approx_total)
    {
        std::map<std::string, std::string> ordered(name_key.begin(), name_key.end()); // std::map<QString,QString>
        std::vector<Result> all;
        for (const auto& nk : ordered) {
            auto it = dbs.find(nk.first);
            if (it == dbs.end()) // "Unknown database requested": skipped (:322-325)
                continue;
            uint64_t approx = 0;
            const int rc = similarity_search(*it->second, q, n_words, nk.second,
                                             static_cast<unsigned>(std::max(results_requested, 0)), cutoff, using_gpu(),
                                             &all, &approx);
            if (rc != GSB_OK)
                return rc;
            *approx_total += approx; // :332
        }
        merge_and_dedup(all, results_requested, smiles, ids, scores);
        return GSB_OK;
    }

    // the merge half of searchDatabases (gpusim.cpp:339-373)
    static void merge_and_dedup(std::vector<Result>& all, int results_requested, std::vector<std::string>* smiles,
                                std::vector<std::string>* ids, std::vector<float>* scores)
    {
        // :339-340 sort + reverse = descending score; ties keep (database, rank) order here, the
        // reference breaks them by pointer value
        std::stable_sort(all.begin(), all.end(), [](const Result& a, const Result& b) { return a.score > b.score; });
        std::map<std::string, std::string> smiles_to_ids; // :342-359
        for (const Result& r : all) {
            auto it = smiles_to_ids.find(r.smiles);
            if (it != smiles_to_ids.end())
                it->second += std::string(";:;") + r.id;
            else
                smiles_to_ids[r.smiles] = r.id;
            if (smiles_to_ids.size() >= static_cast<size_t>(std::max(results_requested, 0)))
                break;
        }
        std::set<std::string> written; // :361-373
        int count = 0;
        for (const Result& r : all) {
            if (count >= results_requested)
                break;
            if (!written.insert(r.smiles).second)
                continue;
            scores->push_back(r.score);
            smiles->push_back(r.smiles);
            ids->push_back(smiles_to_ids[r.smiles]);
            count++;
        }
    }

    struct Request {
        std::vector<std::pair<std::string, std::string>> name_key;
        int request_num = 0, results_requested = 0;
        float cutoff = 0;
        std::vector<int32_t> query;
    };

    static bool parse_request(const uint8_t* data, size_t len, Request* r)
    {
        Reader rd{data, len};
        const int n_db = rd.i32();
        if (n_db < 0 || n_db > 4096) // every entry takes at least 8 bytes of the request
            return false;
        for (int i = 0; i < n_db && rd.ok; i++) {
            std::string name = rd.cstr();
            std::string key = rd.cstr();
            r->name_key.emplace_back(name, key);
        }
        r->request_num = rd.i32();
        r->results_requested = rd.i32();
        r->cutoff = static_cast<float>(rd.f64());
        const std::string fp = rd.bytes();
        if (!rd.ok)
            return false;
        r->query.resize(fp.size() / 4);
        std::memcpy(r->query.data(), fp.data(), r->query.size() * 4);
        return true;
    }

    static void write_response(const Request& r, const std::vector<std::string>& smiles,
                               const std::vector<std::string>& ids, const std::vector<float>& scores, uint64_t approx,
                               std::vector<uint8_t>* response)
    {
        Writer w;
        w.u32(static_cast<uint32_t>(r.request_num));
        w.u32(static_cast<uint32_t>(smiles.size()));
        w.u64(approx);
        for (const auto& s : smiles)
            w.cstr(s.c_str());
        for (const auto& s : ids)
            w.cstr(s.c_str());
        for (float s : scores)
            w.f64(static_cast<double>(s));
        response->swap(w.buf);
    }

    // GPUSimServer::incomingSearchRequest (gpusim.cpp:376-454) without the socket
    int handle_request(const uint8_t* data, size_t len, std::vector<uint8_t>* response)
    {
        Request r;
        if (!parse_request(data, len, &r)) {
            g_srv_err = "truncated request";
            return GSB_ERR_INVALID;
        }
        std::vector<std::string> smiles, ids;
        std::vector<float> scores;
        uint64_t approx = 0;
        const int rc = search_databases(r.query.data(), static_cast<int>(r.query.size()), r.results_requested, r.cutoff,
                                        r.name_key, &smiles, &ids, &scores, &approx);
        if (rc != GSB_OK)
            return rc;
        write_response(r, smiles, ids, scores, approx, response);
        return GSB_OK;
    }

    // Requests that arrived together and ask the same thing of the same databases (same names and
    // keys, result count, cutoff, query width) are answered from ONE gsb_db_search_batch call per
    // database, i.e. one pass over the database for all of them.  The reference serves strictly
    // one query at a time (python/gpusim_server.py:32,100-121 holds a mutex around the socket).
    int handle_batch(const std::vector<const Request*>& group, std::vector<std::vector<uint8_t>>* responses)
    {
        const Request& first = *group[0];
        const size_t nq = group.size();
        const int n_words = static_cast<int>(first.query.size());
        const unsigned k_req = static_cast<unsigned>(std::max(first.results_requested, 0));
        std::map<std::string, std::string> ordered(first.name_key.begin(), first.name_key.end());
        std::vector<std::vector<Result>> all(nq);
        std::vector<uint64_t> approx_total(nq, 0);
        std::vector<int32_t> queries(nq * n_words);
        for (size_t q = 0; q < nq; q++)
            std::memcpy(queries.data() + q * n_words, group[q]->query.data(), n_words * 4);
        for (const auto& nk : ordered) {
            auto it = dbs.find(nk.first);
            if (it == dbs.end() || nk.second != it->second->key)
                continue;
            Database& d = *it->second;
            const unsigned k = static_cast<unsigned>(std::min<uint64_t>(k_req, gsb_db_count(d.db)));
            std::vector<uint32_t> rows(nq * std::max(k, 1u)), cnt(nq);
            std::vector<float> scores(rows.size());
            std::vector<uint64_t> approx(nq);
            const int rc = gsb_db_search_batch(d.db, queries.data(), n_words, static_cast<int>(nq), k, first.cutoff,
                                               rows.data(), scores.data(), cnt.data(), approx.data());
            if (rc != GSB_OK) {
                g_srv_err = gsb_last_error();
                g_last_rc = rc;
                return rc;
            }
            for (size_t q = 0; q < nq; q++) {
                approx_total[q] += approx[q];
                for (uint32_t i = 0; i < cnt[q]; i++) {
                    const char* smi = gsb_fsim_string(d.file, 0, rows[q * k + i]);
                    const char* id = gsb_fsim_string(d.file, 1, rows[q * k + i]);
                    all[q].push_back({scores[q * k + i], smi ? smi : "", id ? id : ""});
                }
            }
        }
        responses->resize(nq);
        for (size_t q = 0; q < nq; q++) {
            std::vector<std::string> smiles, ids;
            std::vector<float> scores;
            merge_and_dedup(all[q], first.results_requested, &smiles, &ids, &scores);
            write_response(*group[q], smiles, ids, scores, approx_total[q], &(*responses)[q]);
        }
        return GSB_OK;
    }
};

// No exception may cross the C ABI (a request asking for 2^31 results must not end in
// std::terminate inside the hosting process).
#define SRV_TRY try {
#define SRV_CATCH                                                                                \
    }                                                                                            \
    catch (const std::bad_alloc&)                                                                \
    {                                                                                            \
        g_srv_err = "out of host memory";                                                        \
        return GSB_ERR_NOMEM;                                                                    \
    }                                                                                            \
    catch (const std::exception& e)                                                              \
    {                                                                                            \
        g_srv_err = std::string("exception: ") + e.what();                                       \
        return GSB_ERR_INVALID;                                                                  \
    }                                                                                            \
    catch (...)                                                                                  \
    {                                                                                            \
        g_srv_err = "unknown exception";                                                         \
        return GSB_ERR_INVALID;                                                                  \
    }

extern "C" {

const char* gsb_server_last_error(void) { return g_srv_err.c_str(); }

// Failure containment for the daemon: after a search came back with GSB_ERR_CUDA the databases are
// put up again — first on the context as it is (the kernels report barrier / peer timeouts through
// an error word and leave the context usable), and if that fails too after a reset of every
// device.  The host rows and strings are untouched, so nothing has to be read from disk again.
int gsb_server_recover(gsb_server* srv)
{
    SRV_TRY
    if (!srv)
        return GSB_ERR_INVALID;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (attempt == 1 && gsb_devices_reset() != GSB_OK)
            break;
        bool ok = true;
        for (auto& kv : srv->dbs)
            if (gsb_db_upload(kv.second->db, nullptr, 0, srv->fold_factor) != GSB_OK) {
                g_srv_err = gsb_last_error();
                ok = false;
                break;
            }
        if (ok)
            return GSB_OK;
    }
    return GSB_ERR_CUDA;
    SRV_CATCH
}

int gsb_server_create(const char* const* fsim_paths, int n_paths, int gpu_bitcount, int use_gpu, gsb_server** out)
{
    SRV_TRY
    if (!out || n_paths < 0 || (n_paths > 0 && !fsim_paths)) {
        g_srv_err = "null argument";
        return GSB_ERR_INVALID;
    }
    std::unique_ptr<gsb_server> srv(new gsb_server);
    srv->use_gpu = use_gpu != 0;
    uint64_t total_db_memory = 0, max_compounds = 0;
    int max_bitcount = 0;
    for (int i = 0; i < n_paths; i++) { // gpusim.cpp:97-117
        std::unique_ptr<Database> d(new Database);
        int rc = gsb_fsim_open(fsim_paths[i], &d->file);
        if (rc != GSB_OK) {
            g_srv_err = gsb_fsim_last_error();
            return rc;
        }
        rc = gsb_fsim_create_db(d->file, &d->db);
        if (rc != GSB_OK) {
            g_srv_err = gsb_last_error();
            return rc;
        }
        d->name = base_name(fsim_paths[i]);
        d->key = gsb_fsim_dbkey(d->file);
        total_db_memory += gsb_db_data_bytes(d->db);
        max_compounds = std::max<uint64_t>(max_compounds, gsb_db_count(d->db));
        max_bitcount = std::max(max_bitcount, gsb_db_fp_bits(d->db));
        srv->dbs[d->name] = std::move(d);
    }
    // fold-factor policy, gpusim.cpp:131-151
    unsigned fold_factor = 1;
    if (gsb_device_count() > 0) {
        uint64_t gpu_memory = gsb_available_device_bytes();
        const uint64_t reserve = sizeof(int) * max_compounds;
        gpu_memory = gpu_memory > reserve ? gpu_memory - reserve : 0;
        if (total_db_memory > gpu_memory && gpu_memory > 0)
            fold_factor = static_cast<unsigned>(std::ceil(static_cast<float>(total_db_memory) /
                                                          static_cast<float>(gpu_memory)));
        // The reference sizes against raw bytes; what has to fit is the device layout (rows padded to
        // a power-of-two width, 2 B popcount per row) plus the upload staging and the multi-query
        // workspaces: fold further until that fits, instead of failing the upload with NOMEM.
        const uint64_t headroom = (3ull << 30) * static_cast<uint64_t>(gsb_device_count());
        auto footprint = [&srv](unsigned f) {
            uint64_t sum = 0;
            for (auto& kv : srv->dbs)
                sum += gsb_layout_bytes(gsb_db_fp_bits(kv.second->db), gsb_db_count(kv.second->db), f);
            return sum;
        };
        while (gpu_memory > headroom && footprint(fold_factor) > gpu_memory - headroom &&
               fold_factor < static_cast<unsigned>(max_bitcount / 32))
            fold_factor++;
    }
    if (gpu_bitcount > 0) {
        const unsigned arg_fold_factor = static_cast<unsigned>(max_bitcount / gpu_bitcount);
        if (arg_fold_factor < fold_factor) {
            g_srv_err = "GPU bitset not sufficiently small to fit on GPU";
            return GSB_ERR_INVALID;
        }
        fold_factor = std::max(1u, arg_fold_factor);
    }
    srv->fold_factor = fold_factor;
    if (srv->using_gpu()) { // gpusim.cpp:159-163
        for (auto& kv : srv->dbs) {
            const int rc = gsb_db_upload(kv.second->db, nullptr, 0, fold_factor);
            if (rc != GSB_OK) {
                g_srv_err = gsb_last_error();
                return rc;
            }
        }
    }
    *out = srv.release();
    return GSB_OK;
    SRV_CATCH
}

void gsb_server_destroy(gsb_server* srv)
{
    if (!srv)
        return;
    if (srv->listen_fd >= 0) {
        close(srv->listen_fd);
        unlink(srv->socket_path.c_str());
    }
    delete srv;
}

void gsb_server_set_use_gpu(gsb_server* srv, int use_gpu) { srv->use_gpu = use_gpu != 0; } // gpusim.h:82
int gsb_server_using_gpu(const gsb_server* srv) { return srv->using_gpu() ? 1 : 0; }
unsigned gsb_server_fold_factor(const gsb_server* srv) { return srv->fold_factor; }
int gsb_server_database_count(const gsb_server* srv) { return static_cast<int>(srv->dbs.size()); }

// GPUSimServer::getFingerprint (gpusim.cpp:456-459)
int gsb_server_get_fingerprint(const gsb_server* srv, const char* dbname, uint64_t row, int32_t* out_words)
{
    auto it = srv->dbs.find(dbname);
    if (it == srv->dbs.end()) {
        g_srv_err = "unknown database";
        return GSB_ERR_INVALID;
    }
    const int rc = gsb_db_get_fingerprint(it->second->db, row, out_words);
    if (rc != GSB_OK)
        g_srv_err = gsb_last_error();
    return rc;
}

// One request in the reference wire format -> one response.  *response is malloc'ed; free with
// gsb_server_free.
int gsb_server_handle_request(gsb_server* srv, const void* request, uint64_t request_bytes, void** response,
                              uint64_t* response_bytes)
{
    SRV_TRY
    std::vector<uint8_t> out;
    const int rc = srv->handle_request(static_cast<const uint8_t*>(request), request_bytes, &out);
    if (rc != GSB_OK)
        return rc;
    *response = std::malloc(out.size() ? out.size() : 1);
    std::memcpy(*response, out.data(), out.size());
    *response_bytes = out.size();
    return GSB_OK;
    SRV_CATCH
}

void gsb_server_free(void* p) { std::free(p); }

// n requests that ask the same thing (same databases and keys, result count, cutoff, query width),
// answered from one gsb_db_search_batch call per database.  responses[i] is malloc'ed.
int gsb_server_handle_batch(gsb_server* srv, const void* const* requests, const uint64_t* request_bytes, int n,
                            void** responses, uint64_t* response_bytes)
{
    SRV_TRY
    if (n <= 0 || !requests || !request_bytes || !responses || !response_bytes) {
        g_srv_err = "null argument";
        return GSB_ERR_INVALID;
    }
    std::vector<gsb_server::Request> parsed(n);
    std::vector<const gsb_server::Request*> group;
    for (int i = 0; i < n; i++) {
        if (!gsb_server::parse_request(static_cast<const uint8_t*>(requests[i]), request_bytes[i], &parsed[i])) {
            g_srv_err = "truncated request";
            return GSB_ERR_INVALID;
        }
        const auto &x = parsed[0], &y = parsed[i];
        if (x.name_key != y.name_key || x.results_requested != y.results_requested || x.cutoff != y.cutoff ||
            x.query.size() != y.query.size()) {
            g_srv_err = "requests of one batch must ask the same thing";
            return GSB_ERR_INVALID;
        }
        group.push_back(&parsed[i]);
    }
    if (!srv->using_gpu()) {
        g_srv_err = "batched search needs the GPU path";
        return GSB_ERR_STATE;
    }
    std::vector<std::vector<uint8_t>> out;
    const int rc = srv->handle_batch(group, &out);
    if (rc != GSB_OK)
        return rc;
    for (int i = 0; i < n; i++) {
        responses[i] = std::malloc(out[i].size() ? out[i].size() : 1);
        std::memcpy(responses[i], out[i].data(), out[i].size());
        response_bytes[i] = out[i].size();
    }
    return GSB_OK;
    SRV_CATCH
}

// GPUSimServer::setupSocket (gpusim.cpp:255-274): listen on <dir>/<name> ("/tmp/gpusimilarity"),
// removing a stale socket file once.
int gsb_server_listen(gsb_server* srv, const char* socket_path)
{
    srv->socket_path = socket_path && *socket_path ? socket_path : "/tmp/gpusimilarity";
    sockaddr_un addr{};
    addr.sun_family = AF_UNIX;
    std::snprintf(addr.sun_path, sizeof(addr.sun_path), "%s", srv->socket_path.c_str());
    for (int attempt = 0; attempt < 2; attempt++) {
        const int fd = socket(AF_UNIX, SOCK_STREAM, 0);
        if (fd < 0)
            break;
        if (bind(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) == 0 && listen(fd, 128) == 0) { // (QLocalServer's own backlog is 50)
            srv->listen_fd = fd;
            return GSB_OK;
        }
        close(fd);
        unlink(srv->socket_path.c_str());
    }
    g_srv_err = "Server start failed on " + srv->socket_path;
    return GSB_ERR_IO;
}

void gsb_server_stop(gsb_server* srv) { srv->stop = true; }

// Event loop: one request in flight at a time, like the reference's Qt main thread.  Returns
// after gsb_server_stop() or after max_requests (0 = unlimited) were answered.
int gsb_server_serve(gsb_server* srv, uint64_t max_requests)
{
    SRV_TRY
    if (srv->listen_fd < 0) {
        g_srv_err = "not listening";
        return GSB_ERR_STATE;
    }
    std::vector<pollfd> fds{{srv->listen_fd, POLLIN, 0}};
    std::map<int, std::vector<uint8_t>> pending;
    uint64_t served = 0;
    // MSG_NOSIGNAL: a client that went away while its search ran must not kill the daemon (and the
    // databases it holds) with SIGPIPE; the connection is dropped when its read side reports EOF
    auto send_all = [](int fd, const std::vector<uint8_t>& buf) {
        size_t sent = 0;
        while (sent < buf.size()) {
            const ssize_t w = send(fd, buf.data() + sent, buf.size() - sent, MSG_NOSIGNAL);
            if (w <= 0)
                break;
            sent += static_cast<size_t>(w);
        }
    };
    // a failed search still gets a well-formed (empty) response, so that the client does not sit
    // out its 30 s timeout; the reason goes to stderr
    auto send_empty = [&send_all](int fd, const gsb_server::Request& r, const char* why) {
        std::fprintf(stderr, "[gpusimserver] request %d failed: %s\n", r.request_num, why);
        std::vector<uint8_t> response;
        gsb_server::write_response(r, {}, {}, {}, 0, &response);
        send_all(fd, response);
    };
    while (!srv->stop && (max_requests == 0 || served < max_requests)) {
        if (poll(fds.data(), fds.size(), 100) <= 0)
            continue;
        if (fds[0].revents & POLLIN) {
            const int c = accept(srv->listen_fd, nullptr, nullptr);
            if (c >= 0)
                fds.push_back({c, POLLIN, 0});
        }
        // read what is there; every connection with a complete request joins this round
        std::vector<std::pair<int, gsb_server::Request>> ready;
        for (size_t i = 1; i < fds.size();) {
            if (!(fds[i].revents & (POLLIN | POLLHUP | POLLERR))) {
                i++;
                continue;
            }
            uint8_t chunk[65536];
            const ssize_t got = read(fds[i].fd, chunk, sizeof(chunk));
            if (got <= 0) {
                close(fds[i].fd);
                pending.erase(fds[i].fd);
                fds.erase(fds.begin() + i);
                continue;
            }
            auto& buf = pending[fds[i].fd];
            buf.insert(buf.end(), chunk, chunk + got);
            gsb_server::Request r;
            if (gsb_server::parse_request(buf.data(), buf.size(), &r)) {
                ready.emplace_back(fds[i].fd, std::move(r));
                buf.clear();
            }
            i++;
        }
        // group identical "shapes" and answer each group from one batched search
        std::vector<bool> taken(ready.size(), false);
        for (size_t a = 0; a < ready.size(); a++) {
            if (taken[a])
                continue;
            std::vector<size_t> members{a};
            for (size_t b = a + 1; b < ready.size(); b++) {
                const auto &x = ready[a].second, &y = ready[b].second;
                if (!taken[b] && x.name_key == y.name_key && x.results_requested == y.results_requested &&
                    x.cutoff == y.cutoff && x.query.size() == y.query.size())
                    members.push_back(b);
            }
            for (size_t m : members)
                taken[m] = true;
            if (members.size() >= 2 && srv->using_gpu()) {
                std::vector<const gsb_server::Request*> group;
                for (size_t m : members)
                    group.push_back(&ready[m].second);
                std::vector<std::vector<uint8_t>> responses;
                if (srv->handle_batch(group, &responses) == GSB_OK)
                    for (size_t g = 0; g < members.size(); g++)
                        send_all(ready[members[g]].first, responses[g]);
                else {
                    for (size_t m : members)
                        send_empty(ready[m].first, ready[m].second, g_srv_err.c_str());
                    if (g_last_rc == GSB_ERR_CUDA)
                        gsb_server_recover(srv);
                }
            } else {
                for (size_t m : members) {
                    const auto& r = ready[m].second;
                    std::vector<std::string> smiles, ids;
                    std::vector<float> scores;
                    uint64_t approx = 0;
                    std::vector<uint8_t> response;
                    if (srv->search_databases(r.query.data(), static_cast<int>(r.query.size()), r.results_requested,
                                              r.cutoff, r.name_key, &smiles, &ids, &scores, &approx) == GSB_OK) {
                        gsb_server::write_response(r, smiles, ids, scores, approx, &response);
                        send_all(ready[m].first, response);
                    } else {
                        send_empty(ready[m].first, r, g_srv_err.c_str());
                        if (g_last_rc == GSB_ERR_CUDA)
                            gsb_server_recover(srv);
                    }
                }
            }
            served += members.size();
        }
    }
    return GSB_OK;
    SRV_CATCH
}

} // extern "C"
