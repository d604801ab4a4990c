"""BASELINE config 5: 1 B synthetic rows sharded over the ranks, a batch of queries, top-100.
Launch with torchrun (one rank per GPU) or plain python for one GPU.
usage: bench_batch_dist.py [--rows R] [--queries Q] [--k K]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from gpusimilarity_b200.dist import ShardedBatchSearcher, shard_range
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000_000)
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--k", type=int, default=100)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
base, n = shard_range(args.rows, rank, world)
db = gsb.FingerprintDB.synthetic(n, device=local, seed=0x5EED5EED, plant_period=250000, row_base=base)
seed_rows = (np.arange(args.queries, dtype=np.uint64) * np.uint64(7919)) % np.uint64(args.rows)
qs = O.synth_rows(0x5EED5EED, seed_rows, 32, 250000)          # copies of database rows: top-1 is the row itself
d_q = torch.from_numpy(np.ascontiguousarray(qs)).to(dev)
s = ShardedBatchSearcher(db, args.k, local, dist, world)
stream = torch.cuda.current_stream()
G = s.max_queries(args.queries, 0.0)
groups = [(q0, min(G, args.queries - q0)) for q0 in range(0, args.queries, G)]
def run_all():
    for q0, nq in groups:
        s.search_device(d_q[q0:].data_ptr(), nq, 0.0, stream)
run_all()                                                            # warm-up (allocations)
torch.cuda.synchronize()
if dist: dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run_all(); b.record(); b.synchronize()
ms = a.elapsed_time(b)
if dist:
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
q0, nq = groups[-1]                                                  # the result buffers hold the last group
top1 = s.out_rows.view(-1, args.k)[:nq, 0].cpu().numpy().astype(np.int64) & 0xffffffff
ok = bool((top1 == seed_rows[q0:q0 + nq].astype(np.int64)).all()) and bool((s.out_scores.view(-1, args.k)[:nq, 0] == 1.0).all().item())
if rank == 0:
    print(json.dumps({"config": f"{args.rows} rows x {args.queries} queries, top-{args.k}, {world} GPU(s)",
                      "batch_ms": ms, "queries_per_s": args.queries / ms * 1e3,
                      "row_query_per_s": args.rows * args.queries / ms * 1e3, "queries_per_pass": G, "top1_is_self_last_group": ok}))
if dist:
    dist.barrier(); dist.destroy_process_group()
