"""Requests per second through the reference's wire protocol (unix socket, one connection per request as
python/gpusim_search.py does it): the Qt-free daemon (gsb_server_*) in this process, N client threads.
The daemon answers requests of one shape that arrive together from ONE multi-query pass.
usage: daemon_bench.py [rows] [clients] [requests_per_client] [results]"""
import os, sys, tempfile, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gpusimilarity_b200.fsim import write_fsim
from gpusimilarity_b200.server import GPUSimServer, encode_request, decode_response, search_over_socket
from oracle import oracle_c as OC

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
n_clients = int(sys.argv[2]) if len(sys.argv) > 2 else 16
per_client = int(sys.argv[3]) if len(sys.argv) > 3 else 64
results = int(sys.argv[4]) if len(sys.argv) > 4 else 20

tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "bench.fsim")
t0 = time.time()
fps = OC.c_synth_db(2024, rows, 32, 997)
write_fsim(path, fps, [b"C" * (1 + i % 7) + str(i).encode() for i in range(rows)], [b"ID%d" % i for i in range(rows)])
print(f"wrote {rows} rows ({rows * 128 / 1e6:.0f} MB of fingerprints) in {time.time() - t0:.1f} s", flush=True)
t0 = time.time()
server = GPUSimServer([path], use_gpu=os.environ.get("DAEMON_CPU") is None)
print(f"daemon up (load + upload) in {time.time() - t0:.2f} s, fold factor {server.foldFactor()}", flush=True)
sock = os.path.join(tmp, "gpusimilarity.sock")
server.listen(sock)
names = {"bench": "pass"}


def run(n_clients, per_client):
    total = n_clients * per_client
    th = threading.Thread(target=server.serve, args=(total,), daemon=True)
    th.start()
    lat = [[] for _ in range(n_clients)]
    ok = [0] * n_clients

    def client(c):
        crng = np.random.default_rng(1000 + c)
        for i in range(per_client):
            row = int(crng.integers(0, rows))
            req = encode_request(names, c * per_client + i, results, 0.0, fps[row])
            t = time.perf_counter()
            resp = search_over_socket(req, sock)
            lat[c].append(time.perf_counter() - t)
            _, approx, smiles, ids, scores = decode_response(resp)
            ok[c] += int(len(scores) >= 1 and scores[0] == 1.0 and ids[0].startswith(b"ID"))
    t = time.perf_counter()
    cl = [threading.Thread(target=client, args=(c,)) for c in range(n_clients)]
    for c in cl:
        c.start()
    for c in cl:
        c.join()
    wall = time.perf_counter() - t
    th.join(timeout=10)
    flat = sorted(x for l in lat for x in l)
    print(f"clients={n_clients:3d} requests={total:5d}: {total / wall:8.0f} requests/s, latency median {1e3 * flat[len(flat) // 2]:.2f} ms "
          f"p99 {1e3 * flat[int(len(flat) * 0.99) - 1]:.2f} ms, {sum(ok)}/{total} answers right (top hit = the query's own row)", flush=True)


run(1, 200)
for c in (4, n_clients, 4 * n_clients):
    run(c, per_client)
server.close()
