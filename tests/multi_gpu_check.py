"""torchrun target (one rank per GPU, NCCL): ShardedKmeans / ShardedIndexFlat on real devices must equal
the single-GPU drop-in.  Launched by tests/test_gpu_multi.py; exit code 0 = parity."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from image_search_engine_b200 import faiss_compat
    from image_search_engine_b200._lib import METRIC_IP, METRIC_L2
    from image_search_engine_b200.parallel import ShardedIndexFlat, ShardedKmeans, shard_bounds

    rng = np.random.default_rng(31)
    n, d, k = 60000, 64, 128
    centers = rng.standard_normal((k, d)).astype(np.float32) * 3
    x = (centers[rng.integers(0, k, n)] + rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)
    x[:300] = x[0]                                       # duplicates -> empty clusters -> split path
    b = shard_bounds(n, world)
    skm = ShardedKmeans(d, k, seed=42, niter=6, spherical=True)
    skm.train(x[b[rank]:b[rank + 1]])
    ref = faiss_compat.Kmeans(d, k, seed=42, niter=6, spherical=True)
    ref.train(x)
    np.testing.assert_allclose(skm.obj, ref.obj, rtol=1e-4)
    np.testing.assert_allclose(skm.centroids, ref.centroids, rtol=2e-3, atol=2e-4)
    assert [s["nsplit"] for s in skm.iteration_stats] == [s["nsplit"] for s in ref.iteration_stats]
    cent = torch.from_numpy(skm.centroids).cuda()
    gathered = [torch.empty_like(cent) for _ in range(world)]
    dist.all_gather(gathered, cent)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks ended with different centroids"

    nb, dq = 50001, 96
    db = rng.standard_normal((nb, dq)).astype(np.float32)
    db[40000] = db[7]                                    # cross-shard exact tie
    q = (db[rng.integers(0, nb, 500)] + 0.02 * rng.standard_normal((500, dq))).astype(np.float32)
    q[0] = db[7]
    bd = shard_bounds(nb, world)
    for metric in (METRIC_IP, METRIC_L2):
        six = ShardedIndexFlat(dq, metric)
        six.add_local(db[bd[rank]:bd[rank + 1]])
        D, I = six.search(q, 10)
        full = faiss_compat.IndexFlatIP(dq) if metric == METRIC_IP else faiss_compat.IndexFlatL2(dq)
        full.add(db)
        Dr, Ir = full.search(q, 10)
        I, D = I.cpu().numpy(), D.cpu().numpy()
        mism = (I != Ir).any(axis=1).mean()
        assert mism <= 0.02, f"metric {metric}: {mism:.3%} rows differ from the unsharded search"
        np.testing.assert_allclose(D, Dr, rtol=1e-5, atol=1e-5)
        assert I[0, 0] == 7 or metric == METRIC_L2 and I[0, 0] in (7, 40000)
        assert (I[0, :2] == [7, 40000]).all()
        # query counts that do not divide by the world size (padded all-to-all slices) and fewer than 20 queries
        # (the exact direct-difference kernel on every shard, like the unsharded index): identical to the flat index
        for nq in (33, 7):
            Ds, Is = six.search(q[:nq], 10)
            Df, If = full.search(q[:nq], 10)
            Is, Ds = Is.cpu().numpy(), Ds.cpu().numpy()
            if nq < 20:
                assert np.array_equal(Is, If), f"metric {metric}, nq {nq}"
                np.testing.assert_allclose(Ds, Df, rtol=1e-6, atol=1e-6)
            else:
                assert (Is != If).any(axis=1).mean() <= 0.1
                np.testing.assert_allclose(Ds, Df, rtol=1e-5, atol=1e-5)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_PARITY_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
