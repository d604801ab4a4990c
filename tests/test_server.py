"""GPUSimServer equivalent: multi-database merge + SMILES de-duplication (reference
gpusim.cpp:306-374), the socket wire protocol (gpusim.cpp:376-454) and the fold-factor policy.
On a machine without a GPU the server answers through search_cpu exactly as the reference does
(usingGPU() false, gpusim.cpp:168-171,327-331); the gpu-marked tests run the same through CUDA."""
import os
import shutil
import struct
import subprocess
import threading
import time

import numpy as np
import pytest

import gpusimilarity_b200 as gsb
from conftest import GOLDEN, ROOT
from gpusimilarity_b200.fsim import write_fsim
from gpusimilarity_b200.server import (GPUSimServer, decode_response, encode_request, search_over_socket)
from oracle import oracle as O


@pytest.fixture()
def two_dbs(tmp_path):
    a, b = str(tmp_path / "small.fsim"), str(tmp_path / "small_copy.fsim")
    shutil.copyfile(os.path.join(GOLDEN, "small.fsim"), a)   # reference test/CMakeLists.txt:11-12
    shutil.copyfile(os.path.join(GOLDEN, "small.fsim"), b)
    return a, b


def _search_multiple(server, small_db, golden):
    t = golden["reference_tests"]["TestSearchMultiple"]
    fp = server.getFingerprint(t["query_row"], "small")
    assert np.array_equal(fp, small_db[t["query_row"]])
    smiles, ids, scores, approx = server.searchDatabases(fp, t["k"], 0.0, {"small": "pass", "small_copy": "pass"})
    assert len(smiles) == t["k"]                                  # test_gpusim.cpp:96
    assert ids[0].decode() == t["top_id"]                         # test_gpusim.cpp:98
    assert len(set(smiles)) == len(smiles) and scores == sorted(scores, reverse=True)
    return smiles, ids, scores, approx


def test_search_multiple_cpu(two_dbs, small_db, small_fsim, golden):
    """reference TestSearchMultiple (test/test_gpusim.cpp:71-99) through the CPU path."""
    server = GPUSimServer(list(two_dbs), use_gpu=False)
    assert not server.usingGPU()
    smiles, ids, scores, approx = _search_multiple(server, small_db, golden)
    # same answer as the oracle's restatement of searchDatabases
    rows, sc = O.search_cpu(small_db[3], small_db, 10)
    one = ([small_fsim.smiles[r] for r in rows], [small_fsim.ids[r] for r in rows], list(sc))
    w_smiles, w_ids, w_scores = O.search_databases([one, one], 10)
    assert smiles == w_smiles and ids == w_ids and np.allclose(scores, w_scores, rtol=0, atol=0)
    # unknown database names are skipped, wrong keys give nothing (gpusim.cpp:322-325, .cu:349-352)
    s2, i2, f2, _ = server.searchDatabases(small_db[3], 10, 0.0, {"small": "pass", "nope": "x"})
    assert len(s2) == 10 and i2[0] == b"ZINC00000022"
    s3, _, _, _ = server.searchDatabases(small_db[3], 10, 0.0, {"small": "wrong"})
    assert s3 == []


def test_wire_format_round_trip(two_dbs, small_db):
    server = GPUSimServer([two_dbs[0]], use_gpu=False)
    req = encode_request({"small": "pass"}, 123456, 5, 0.0, small_db[0])
    # layout check against the reference writer (gpusim_search.py:36-47)
    assert req[:4] == struct.pack(">i", 1) and req[4:8] == struct.pack(">I", 6) and req[8:14] == b"small\0"
    resp = server.handleRequest(req)
    num, approx, smiles, ids, scores = decode_response(resp)
    assert num == 123456 and len(smiles) == 5 and ids[0] == b"ZINC00000007" and scores[0] == 1.0
    assert len(resp) == 16 + sum(4 + len(s) + 1 for s in smiles + ids) + 8 * 5
    with pytest.raises(gsb.GsbError):
        server.handleRequest(req[:20])


def test_socket_server_thread(two_dbs, small_db, tmp_path):
    server = GPUSimServer(list(two_dbs), use_gpu=False)
    path = str(tmp_path / "gpusimilarity.sock")
    server.listen(path)
    th = threading.Thread(target=server.serve, args=(3,), daemon=True)
    th.start()
    for i, row in enumerate((0, 3, 42)):
        resp = search_over_socket(encode_request({"small": "pass", "small_copy": "pass"}, 1000 + i, 7, 0.0,
                                                 small_db[row]), path)
        num, approx, smiles, ids, scores = decode_response(resp)
        assert num == 1000 + i and len(smiles) == 7 and scores[0] == 1.0 and b";:;" in ids[0]
    th.join(timeout=10)
    assert not th.is_alive()


def test_many_clients_at_once(two_dbs, small_db, tmp_path):
    """80 clients connect at the same moment (more than the accept queue held before: a unix socket
    refuses the surplus with EAGAIN): every one gets its own, right answer."""
    n = 80
    server = GPUSimServer([two_dbs[0]], use_gpu=False)
    path = str(tmp_path / "crowd.sock")
    server.listen(path)
    th = threading.Thread(target=server.serve, args=(n,), daemon=True)
    th.start()
    out = [None] * n
    go = threading.Barrier(n)

    def client(i):
        go.wait()
        out[i] = decode_response(search_over_socket(encode_request({"small": "pass"}, 5000 + i, 3, 0.0, small_db[i % 100]), path))

    clients = [threading.Thread(target=client, args=(i,)) for i in range(n)]
    for c in clients:
        c.start()
    for c in clients:
        c.join(timeout=60)
    th.join(timeout=10)
    for i, r in enumerate(out):
        assert r is not None and r[0] == 5000 + i and len(r[2]) == 3 and r[4][0] == 1.0, i


def test_hostile_requests_do_not_hurt_the_daemon(two_dbs, small_db, tmp_path):
    """a result count of 2^31-1 is clamped to the row count (no 16 GB allocation, no
    exception through the C ABI); a request that fails (wrong query width) still gets a well-formed
    empty response; a client that hangs up before its answer does not kill the server (MSG_NOSIGNAL)."""
    import socket as pysocket
    server = GPUSimServer([two_dbs[0]], use_gpu=False)
    num, approx, smiles, ids, scores = decode_response(
        server.handleRequest(encode_request({"small": "pass"}, 9, 2**31 - 1, 0.0, small_db[0])))
    assert num == 9 and len(smiles) == 100                          # every row, not 2^31-1 slots
    with pytest.raises(gsb.GsbError):
        server.handleRequest(struct.pack(">i", 2**30) + b"\0" * 64)  # absurd database count
    path = str(tmp_path / "hostile.sock")
    server.listen(path)
    th = threading.Thread(target=server.serve, args=(4,), daemon=True)
    th.start()
    bad = encode_request({"small": "pass"}, 77, 5, 0.0, small_db[0][:16])    # 512-bit query, 1024-bit database
    num, approx, smiles, ids, scores = decode_response(search_over_socket(bad, path, timeout=10))
    assert num == 77 and smiles == [] and approx == 0
    with pysocket.socket(pysocket.AF_UNIX, pysocket.SOCK_STREAM) as c:      # hang up right after sending
        c.connect(path)
        c.sendall(encode_request({"small": "pass"}, 78, 100, 0.0, small_db[1]))
    for i in range(2):                                                       # the daemon is still there
        num, _, smiles, ids, _ = decode_response(search_over_socket(
            encode_request({"small": "pass"}, 80 + i, 3, 0.0, small_db[0]), path, timeout=10))
        assert num == 80 + i and ids[0] == b"ZINC00000007"
    th.join(timeout=10)
    assert not th.is_alive()


def test_inconsistent_fsim_fails_at_load(tmp_path, small_fsim):
    """a .fsim whose SMILES / id counts do not match its fingerprint count must fail when it
    is opened, not crash the server at query time."""
    path = str(tmp_path / "short.fsim")
    write_fsim(path, small_fsim.fingerprints(), list(small_fsim.smiles)[:-1], list(small_fsim.ids),
               dbkey=small_fsim.dbkey)
    with pytest.raises(gsb.GsbError):
        GPUSimServer([path], use_gpu=False)


def test_server_binary_cli(two_dbs, small_db, tmp_path):
    """The Qt-free gpusimserver: reference command line (main.cpp:21-28), --cpu_only."""
    binary = os.path.join(ROOT, "gpusimilarity_b200", "gpusimserver_b200")
    if not os.path.exists(binary):
        subprocess.run(["make", "-C", ROOT, "adapter"], check=True, capture_output=True)
    path = str(tmp_path / "cli.sock")
    proc = subprocess.Popen([binary, "--cpu_only", "--socket", path, two_dbs[0]], stderr=subprocess.PIPE)
    try:
        for _ in range(100):
            if os.path.exists(path):
                break
            time.sleep(0.05)
        resp = search_over_socket(encode_request({"small": "pass"}, 7, 3, 0.0, small_db[0]), path)
        num, _, smiles, ids, scores = decode_response(resp)
        assert num == 7 and ids[0] == b"ZINC00000007" and len(smiles) == 3
    finally:
        proc.terminate()
        proc.wait(timeout=10)
    assert subprocess.run([binary], capture_output=True).returncode == 1   # "Not enough arguments."


def test_gpu_bitcount_policy(two_dbs):
    with pytest.raises(gsb.GsbError):
        GPUSimServer(["/nonexistent/file.fsim"])
    srv = GPUSimServer([two_dbs[0]], gpu_bitcount=256, use_gpu=False)
    assert srv.foldFactor() == 4                                  # 1024 / 256 (gpusim.cpp:144-151)


@pytest.mark.gpu
def test_search_multiple_gpu(two_dbs, small_db, golden):
    server = GPUSimServer(list(two_dbs))
    assert server.usingGPU()
    smiles, ids, scores, approx = _search_multiple(server, small_db, golden)
    assert approx == 200                                          # += per database (gpusim.cpp:332)
    cpu = GPUSimServer(list(two_dbs), use_gpu=False)
    c_smiles, c_ids, c_scores, _ = cpu.searchDatabases(small_db[3], 10, 0.0, {"small": "pass", "small_copy": "pass"})
    assert smiles == c_smiles and ids == c_ids and scores == c_scores
    # cutoff through the wire (f64 on the wire, f32 in the engine): reference TestSimilarityCutoff counts
    for cutoff, n_res, n_approx in zip((0, 0.1, 0.3, 0.4), (10, 10, 3, 1), (100, 86, 3, 1)):
        s, i, f, a = server.searchDatabases(small_db[0], 10, cutoff, {"small": "pass"})
        assert (len(s), a) == (n_res, n_approx)


@pytest.mark.gpu
def test_batched_requests_equal_single_requests(two_dbs, small_db, tmp_path):
    """Daemon-side batching: requests of one shape go through gsb_db_search_batch (one pass over
    each database) and must produce byte-identical responses; also through the socket loop with
    clients that send at the same time."""
    server = GPUSimServer(list(two_dbs))
    names = {"small": "pass", "small_copy": "pass"}
    reqs = [encode_request(names, 500 + i, 8, 0.05, small_db[row]) for i, row in enumerate((0, 3, 17, 42, 99, 3))]
    single = [server.handleRequest(r) for r in reqs]
    assert server.handleBatch(reqs) == single
    with pytest.raises(gsb.GsbError):
        server.handleBatch([reqs[0], encode_request(names, 1, 9, 0.05, small_db[0])])
    path = str(tmp_path / "batch.sock")
    server.listen(path)
    th = threading.Thread(target=server.serve, args=(len(reqs),), daemon=True)
    th.start()
    out = [None] * len(reqs)

    def client(i):
        out[i] = search_over_socket(reqs[i], path)

    clients = [threading.Thread(target=client, args=(i,)) for i in range(len(reqs))]
    for c in clients:
        c.start()
    for c in clients:
        c.join(timeout=30)
    th.join(timeout=10)
    assert out == single


@pytest.mark.gpu
def test_folded_server(two_dbs, small_db):
    server = GPUSimServer([two_dbs[0]], gpu_bitcount=512)
    assert server.foldFactor() == 2
    smiles, ids, scores, _ = server.searchDatabases(small_db[0], 5, 0.0, {"small": "pass"})
    assert ids[0] == b"ZINC00000007" and scores[0] == 1.0


@pytest.mark.gpu
def test_daemon_recovers_after_device_reset(two_dbs, small_db, tmp_path):
    """the daemon's way back after GSB_ERR_CUDA — gsb_server_recover puts the databases
    up again, after a cudaDeviceReset if need be.  Runs in its own process: the reset takes the whole
    CUDA context of the process with it."""
    script = f"""
import sys, faulthandler, numpy as np
faulthandler.enable()                                             # a crash names the step it happened in
sys.path.insert(0, {ROOT!r})
from gpusimilarity_b200._lib import lib, check
from gpusimilarity_b200.server import GPUSimServer
from gpusimilarity_b200.fsim import read_fsim
db = read_fsim({two_dbs[0]!r}).fingerprints()
srv = GPUSimServer([{two_dbs[0]!r}, {two_dbs[1]!r}])
names = {{"small": "pass", "small_copy": "pass"}}
want = srv.searchDatabases(db[3], 10, 0.0, names)
assert lib().gsb_server_recover(srv._h) == 0                      # plain re-upload
assert srv.searchDatabases(db[3], 10, 0.0, names) == want
check(lib().gsb_devices_reset())                                  # the context is gone ...
assert lib().gsb_server_recover(srv._h) == 0                      # ... and the databases come back
assert srv.searchDatabases(db[3], 10, 0.0, names) == want
assert srv.searchDatabases(db[0], 10, 0.3, {{"small": "pass"}})[3] == 3
print("recovered")
"""
    r = subprocess.run([os.sys.executable, "-c", script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "recovered" in r.stdout, r.stdout + r.stderr
