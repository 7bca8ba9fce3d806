"""Full-size multi-GPU configs of BASELINE.json (torchrun, one rank per GPU):
  C4  sharded k-means: 10M x 128 SIFT-like descriptors (mixture of 100k centres), k = 65536, NCCL all-reduce
  C5  sharded flat IP index: 10M x 512 unit vectors, 10k queries, top-100, all_gather + on-device merge
Prints one JSON line on rank 0.  Usage: torchrun --nproc-per-node N profiles/run_c4_c5.py [--iters 3] [--scale 1.0]
"""
import argparse, json, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
args = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
from image_search_engine_b200 import ops
from image_search_engine_b200._lib import METRIC_IP
from image_search_engine_b200.parallel import ShardedIndexFlat, ShardedKmeans

out = {"n_gpus": world}

# ---------------- C4 ----------------
N, d, k = int(10_000_000 * args.scale), 128, int(65536 * args.scale)
n_local = N // world
gc = torch.Generator(device=dev); gc.manual_seed(4)            # same centre table on every rank
centres = torch.randn((100_000, d), generator=gc, device=dev).square_()
centres *= 512.0 / centres.norm(dim=1, keepdim=True)
g = torch.Generator(device=dev); g.manual_seed(40 + rank)
x = centres[torch.randint(0, centres.shape[0], (n_local,), generator=g, device=dev)]
x = (x + 8.0 * torch.randn((n_local, d), generator=g, device=dev)).round_().clamp_(0, 255)
km = ShardedKmeans(d, k, seed=42, niter=args.iters, spherical=True)
warm = torch.ones(1 << 20, device=dev)
for _ in range(3):
    dist.all_reduce(warm)          # NCCL communicator set-up is not part of the k-means time
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
km.train(x)
torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
out["c4"] = {"n": n_local * world, "d": d, "k": k, "iters": args.iters, "s_total": t1 - t0,
             "s_per_iter_incl_setup": (t1 - t0) / args.iters,
             "iter_end_times_s": [s_["time"] for s_ in km.iteration_stats], "obj": [float(o) for o in km.obj],
             "nsplit": [s["nsplit"] for s in km.iteration_stats],
             "phases_ms_rank0": [{kk: round(s_[kk], 3) for kk in ("ms_assign", "ms_accumulate", "ms_allreduce", "ms_finalize_host")}
                                 for s_ in km.iteration_stats],
             "allreduce_bytes_per_iter": 4 * (k * d + k) + 8}
del x, km
torch.cuda.empty_cache()

# ---------------- C5 ----------------
nb, dq, nq, topk = int(10_000_000 * args.scale), 512, 10_000, 100
nb_local = nb // world
g5 = torch.Generator(device=dev); g5.manual_seed(50 + rank)
db = torch.empty((nb_local, dq), device=dev)
for i in range(0, nb_local, 250_000):
    db[i:i + 250_000].normal_(generator=g5)
ops.normalize_l2_(db)
q = db[torch.randint(0, nb_local, (nq,), generator=g5, device=dev)] + 0.05 * torch.randn((nq, dq), generator=g5, device=dev)
dist.broadcast(q, src=0)
ops.normalize_l2_(q)
idx = ShardedIndexFlat(dq, METRIC_IP)
idx.add_local(db)
for _ in range(2):
    D, I = idx.search(q, topk)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    D, I = idx.search(q, topk)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
sorted_ok = bool((D[:, :-1] >= D[:, 1:]).all())
out["c5"] = {"nb": nb_local * world, "d": dq, "nq": nq, "topk": topk, "ms_per_batch": float(ms.item()),
             "qps": nq / float(ms.item()) * 1e3, "sorted": sorted_ok, "top1_score_mean": float(D[:, 0].mean()),
             "search_stats_rank0": ops.search_stats()}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
