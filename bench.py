#!/usr/bin/env python
"""Headline benchmark of the B200 retrieval core (contract: see the task statement / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-knn]

One "step" = one pass of the hot path over one batch of synthetic input at BASELINE.json configs[1]
(C2): quantise 1,000,000 SIFT-like 128-D descriptors against a k=4096 codebook (fused tcgen05
assign), per-image BoVW histogram over 10,000 images and Okapi tf weighting.  The second half of the
metric (kNN QPS at 1M x 2048, top-10 = configs[2], C3) is measured in the same run and reported under
"knn".  N > 1: one process per GPU (torchrun), every rank quantises its own C2-sized shard of images
(no data-path collective, weak scaling) and holds its own 1M-row shard of the flat index (all_gather
of per-rank top-k lists + on-device merge).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C2 = dict(n_desc=1_000_000, d=128, k=4096, n_img=10_000, per_img=100)
C3 = dict(nb=1_000_000, d=2048, nq=10_000, topk=10)
METRIC = "Mdescriptors/s for k-means assign+histogram"
WORKLOAD = ("C2: 1M SIFT-like 128-D float32 descriptors per GPU, k=4096 codebook, 10k images: quantise (assign) + "
            "BoVW histogram (numpy-compat) + Okapi tf")


def sift_like(rng, n, d):
    g = rng.standard_normal((n, d), dtype=np.float32)
    g *= g
    g *= (512.0 / np.linalg.norm(g, axis=1, keepdims=True)).astype(np.float32)
    np.rint(g, out=g)
    np.minimum(g, 255, out=g)
    return g


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        j = json.loads(p.read_text())
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sust=j.get("bf16_tflops_sustained"),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed regions, read through NVML from a thread of
    this process (pynvml).  Spawning `nvidia-smi -lms` instead was measured to slow host-side CUDA calls
    enough to distort millisecond-scale steps (profiles/r01_findings.md); pause() stops sampling for the
    host/PCIe-bound e2e legs."""

    def __init__(self, gpu_index, period_s=0.02):
        import threading
        self.mode = os.environ.get("ISE_BENCH_SAMPLER", "on")   # debugging aid: on | idle | off
        self.period = period_s
        self.rows = []
        self.active = threading.Event()
        self.quit = threading.Event()
        self.h = None
        self.err = None
        try:
            if self.mode == "off":
                raise RuntimeError("sampler disabled by ISE_BENCH_SAMPLER=off")
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ent = vis.split(",")[gpu_index].strip()
                phys = int(ent) if ent.isdigit() else None
            self.h = (pynvml.nvmlDeviceGetHandleByIndex(phys) if phys is not None
                      else pynvml.nvmlDeviceGetHandleByUUID(ent))
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # pragma: no cover - depends on the box
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self.quit.is_set():
            if self.active.is_set():
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((float(sm), float(pw), int(rs)))
                except Exception as e:  # pragma: no cover
                    self.err = repr(e)
            time.sleep(self.period)

    def start(self):
        if self.mode == "on":
            self.active.set()

    def pause(self):
        self.active.clear()

    def stop(self):
        self.pause()
        self.quit.set()
        if self.h is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r[2] & bit for r in self.rows))
        pmax = max(r[1] for r in self.rows)
        load = [r[0] for r in self.rows if r[1] >= 0.5 * pmax] or [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": self.sm_max, "reasons": reasons,
                "samples": len(self.rows), "power_w_max": pmax, "source": "NVML, 20 ms period, device-timed regions"}


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference loops; also the cpu_baseline leg of the GPU arm)
# ----------------------------------------------------------------------------------------------
def cpu_assign_histogram(n_img_sample, seed=2, reps=1):
    """Times the reference's per-image quantise + np.histogram + Okapi loop on `n_img_sample` images."""
    from oracle import cpu_baseline, faiss_shim
    rng = np.random.default_rng(seed)
    X = sift_like(rng, n_img_sample * C2["per_img"], C2["d"])
    cent = X[rng.choice(X.shape[0], C2["k"], replace=X.shape[0] < C2["k"])].copy()
    faiss_shim.normalize_L2(cent)
    index = cpu_baseline.codebook_index(cent)
    descs = [X[i * C2["per_img"]:(i + 1) * C2["per_img"]] for i in range(n_img_sample)]
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_baseline.assign_histogram_step(index, descs, C2["k"])
        best = min(best, time.perf_counter() - t0)
    return X.shape[0] / best / 1e6, best


def cpu_knn(nb_sample, nq_sample, seed=3):
    from oracle import cpu_baseline
    rng = np.random.default_rng(seed)
    db = np.maximum(rng.standard_normal((nb_sample, C3["d"]), dtype=np.float32), 0)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = db[rng.integers(0, nb_sample, nq_sample)] + 0.05 * rng.standard_normal((nq_sample, C3["d"]), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    t0 = time.perf_counter()
    cpu_baseline.flat_search(db, q.astype(np.float32), C3["topk"], "ip")
    dt = time.perf_counter() - t0
    # flat search cost is linear in nb: QPS at the full 1M-row index = sample QPS * nb_sample / nb
    return nq_sample / dt * (nb_sample / C3["nb"]), dt


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        n = os.cpu_count() or 1
    return int(n)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads()
    n_img = 1000  # bounded sample: 100,000 descriptors per step (1/10 of C2), same generator and codebook size
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_assign_histogram(n_img)
        if i >= args.warmup:
            vals.append((v, dt))
        if sum(d for _, d in vals) > 150:
            break
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    sample = (f"{n_img} of {C2['n_img']} images x {C2['per_img']} descriptors per step (1/10 of C2), "
              f"per-image Faiss-shim search + np.histogram + Okapi, NumPy/OpenBLAS")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mdescriptors/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mdescriptors/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mdescriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_knn:
        q, dt = cpu_knn(200_000, 1000)
        line["knn"] = {"metric": "kNN QPS at 1M x 2048 top-10", "value": q, "unit": "queries/s",
                       "sample": "200k of 1M DB rows x 1000 of 10k queries (%.1f s), scaled linearly in nb" % dt}
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from image_search_engine_b200 import BOVW, FaissKMeans, faiss_compat, ops
    from image_search_engine_b200._lib import METRIC_IP
    from image_search_engine_b200.bag_of_visual_words import PackedDescriptions

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = None
    if world > 1 and not os.environ.get("ISE_BENCH_NO_BIND"):
        from image_search_engine_b200.parallel import bind_to_gpu_numa
        numa_cpus = bind_to_gpu_numa(local_rank)     # before any pinned allocation (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- C2 inputs (per rank: its own shard of 10k images) ----------------
    rng = np.random.default_rng(2 + rank)
    X_host = sift_like(rng, C2["n_desc"], C2["d"])
    offsets = np.arange(0, C2["n_desc"] + 1, C2["per_img"], dtype=np.int64)
    packed = PackedDescriptions(X_host, offsets).pin()
    X_dev = packed.matrix.to(dev)
    off_dev = torch.from_numpy(offsets).to(dev)

    # codebook: 2 Lloyd iterations from the Faiss-style random init (untimed set-up)
    km = FaissKMeans(C2["k"], n_init=1, max_iter=2)
    km.fit(X_dev)
    bovw = BOVW(None, n_clusters=C2["k"])
    bovw.clusterer = km

    from image_search_engine_b200.utils import OkapiTransformer
    okapi = OkapiTransformer()

    # CUDA events around every gemm_select launch made inside the timed regions (roofline numerator)
    kernel_events = []
    _orig_gemm_select = ops.gemm_select

    def _timed_gemm_select(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = _orig_gemm_select(*a, **kw)
        e1.record()
        kernel_events.append((e0, e1))
        return r

    ops.gemm_select = _timed_gemm_select
    packed_dev = PackedDescriptions(X_dev, offsets)

    def device_step():
        # the public device-resident path: prepare planes -> fused assign -> histogram + Okapi tf
        return bovw.histograms_device(packed_dev, okapi=okapi)

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = ops.launches()
    kernel_events.clear()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        device_step()
    t_end.record()
    barrier()
    launches = ops.launches() - l0
    ms_step = max_over_ranks(t_start.elapsed_time(t_end) / args.steps)
    # all gemm_select launches of a step (main launch + any fallback re-run) count towards the kernel time
    kern_ms = float(np.sum([a.elapsed_time(b) for a, b in kernel_events]) / args.steps)
    assign_stats = dict(ops.last_search_stats)
    value = world * C2["n_desc"] / (ms_step * 1e-3) / 1e6

    # ---------------- the step's HBM-bound kernels, timed alone (explains the non-tensor share of the step) ----
    def _time(fn, reps=5):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    words_dev = km.transform_device(X_dev)
    hist_out = torch.empty((C2["n_img"], C2["k"]), dtype=torch.float64, device=dev)
    hbm_kernels = []
    for name, fn, nbytes, traffic in (
            # algorithmic: read every float once, write the one FP16 plane integer-valued descriptors need, + norms;
            # actual traffic: the absmax pass reads the matrix a second time
            ("absmax_f32_kernel + prepare_planes_f32x4_kernel", lambda: ops.prepare_operand(X_dev),
             C2["n_desc"] * (C2["d"] * 6 + 4), C2["n_desc"] * (C2["d"] * 10 + 4)),
            ("histogram_kernel<double> (numpy-compat + Okapi)",
             lambda: ops.bovw_histogram(words_dev, off_dev, C2["k"], okapi=True, out=hist_out),
             C2["n_desc"] * 8 + C2["n_img"] * C2["k"] * 8, C2["n_desc"] * 8 + C2["n_img"] * C2["k"] * 8)):
        ms = _time(fn)
        hbm_kernels.append({"kernel": name, "ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
                            "peak": P["hbm"], "frac": nbytes / (ms * 1e-3) / 1e9 / P["hbm"],
                            "traffic_gbs": traffic / (ms * 1e-3) / 1e9, "frac_traffic": traffic / (ms * 1e-3) / 1e9 / P["hbm"]})
    del hist_out, words_dev

    # ---------------- e2e: host (pinned) descriptors -> host histogram matrix ----------------
    sampler.pause()   # clocks are sampled during the device-timed regions only (see ClockSampler.pause)
    bovw.descriptions = None

    def timed_host(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)

    # (1) the Pipeline-level call: pinned host descriptors -> scipy CSR float64 (histogram + Okapi tf), which is what
    #     the reference's pipeline.transform() returns (utils.py:153-202).  H2D chunks overlap the quantisation; the
    #     CSR arrays are built in HBM and come back in one small copy.
    csr_holder = {}

    def e2e_step_csr():
        csr_holder["m"] = bovw.transform_csr(packed, okapi=okapi, n_chunks=16, copy=False)

    e2e_ms = timed_host(e2e_step_csr)
    e2e_val = world * C2["n_desc"] / (e2e_ms * 1e-3) / 1e6
    h2d = int(X_host.nbytes + offsets.nbytes)
    d2h = int((C2["n_img"] + 1) * 4 + C2["n_desc"] * 12)       # indptr + (int32 index, float64 value) per descriptor slot
    csr_nnz = int(csr_holder["m"].nnz)
    assert csr_holder["m"].shape == (C2["n_img"], C2["k"]) and float(csr_holder["m"].sum()) > 0

    # (2) the same step with the reference's DENSE float64 matrix as the result (BOVW.transform's own format):
    #     H2D | kernels | D2H of 328 MB chunk-pipelined on three streams
    out_pin = torch.empty((C2["n_img"], C2["k"]), dtype=torch.float64, pin_memory=True)

    def e2e_step_dense():
        return bovw.histograms_host(packed, out_pin, okapi=okapi, n_chunks=16)

    e2e_dense_ms = timed_host(e2e_step_dense)
    d2h_dense = int(out_pin.numel() * 8)
    del out_pin

    # ---------------- k-means training iteration time (extra) ----------------
    kmeans_iter_ms = float("inf")
    for _ in range(2):      # best of two 5-iteration fits (wall clock incl. the per-iteration host sync)
        barrier()
        t0 = time.perf_counter()
        km2 = FaissKMeans(C2["k"], n_init=1, max_iter=5)
        km2.fit(X_dev)
        torch.cuda.synchronize()
        kmeans_iter_ms = min(kmeans_iter_ms, (time.perf_counter() - t0) * 1e3 / 5)

    del km2

    # ---------------- C3: flat IP search, 1M x 2048 per rank, 10k queries, top-10 ----------------
    knn = None
    if not args.no_knn:
        g = torch.Generator(device=dev)
        g.manual_seed(3 + rank)
        db = torch.empty((C3["nb"], C3["d"]), dtype=torch.float32, device=dev)
        for i in range(0, C3["nb"], 100_000):
            blk = db[i:i + 100_000]
            blk.normal_(generator=g)
            blk.clamp_(min=0)
        ops.normalize_l2_(db)
        gq = torch.Generator(device=dev)
        gq.manual_seed(1003)  # same queries on every rank
        pick = torch.randint(0, C3["nb"], (C3["nq"],), generator=gq, device=dev)
        q = db[pick] + 0.05 * torch.randn((C3["nq"], C3["d"]), generator=gq, device=dev)
        if world > 1:
            dist.broadcast(q, src=0)
        ops.normalize_l2_(q)
        from image_search_engine_b200.parallel import ShardedIndexFlat
        sidx = ShardedIndexFlat(C3["d"], METRIC_IP)
        sidx.add_local(db)                      # planes + norms built once (index build, untimed)
        torch.cuda.synchronize()

        def knn_step(qd):
            # public sharded-index search: prepare(q) -> gemm_select top-10 -> exact re-score ->
            # [all_gather + on-device merge when world > 1]
            return sidx.search(qd, C3["topk"])

        ksteps = max(2, min(args.steps, 5))
        for _ in range(2):
            knn_step(q)
        barrier()
        sampler.start()
        kernel_events.clear()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(ksteps):
            D, I = knn_step(q)
        e.record()
        barrier()
        knn_ms = max_over_ranks(s.elapsed_time(e) / ksteps)
        # sample pre-pass + coarse main launch + split re-run of unproven rows, all tcgen05 gemm_select launches
        knn_kern_ms = float(np.sum([a.elapsed_time(b) for a, b in kernel_events]) / ksteps)
        knn_stats = dict(ops.last_search_stats)
        sampler.pause()
        # self-check (size independent): every query's best hit is the row it was derived from (rank 0's shard)
        hit = float((I[:, 0] == pick).float().mean().item())
        # e2e: pinned host queries in, host (D, I) out
        q_pin = q.cpu().pin_memory()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            qd = q_pin.to(dev, non_blocking=True)
            D, I = knn_step(qd)
            Dh, Ih = D.cpu(), I.cpu()
        torch.cuda.synchronize()
        knn_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / ksteps)
        flops = 2.0 * C3["nq"] * C3["nb"] * C3["d"]
        knn = {
            "metric": "kNN QPS at 1M x 2048 top-10", "value": C3["nq"] / (knn_ms * 1e-3), "unit": "queries/s",
            "ms_per_step": knn_ms, "steps": ksteps, "nb_per_gpu": C3["nb"], "nb_total": C3["nb"] * world,
            "e2e": {"value": C3["nq"] / (knn_e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(q.numel() * 4), "d2h_bytes_per_step": int(C3["nq"] * C3["topk"] * 12)},
            "roofline": {"bound": "tensor", "achieved": flops / (knn_kern_ms * 1e-3) / 1e12, "peak": P["tf_burst"],
                         "unit": "TFLOP/s", "frac": flops / (knn_kern_ms * 1e-3) / 1e12 / P["tf_burst"],
                         "traffic": 104.08e9,
                         "traffic_source": "ncu r01 final: dram read 103.91 GB + write 0.17 GB for the coarse launch "
                                           "(29.9 GB on another box with the L2 hit rate at 91 % instead of 72 %)",
                         "kernel": "gemm_select_kernel<1,1,IP,32> coarse (+ sample pre-pass, split re-run of unproven "
                                   "rows, topk_merge_kernel)", "kernel_ms": knn_kern_ms, "search": knn_stats,
                         "peak_source": P["src"] + ", bf16 burst"},
            "top1_self_hit": hit,
        }
        del db, sidx

    clocks = sampler.stop()

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v1, dt1 = cpu_assign_histogram(300)
        n_img = int(min(C2["n_img"], max(300, 300 * 15.0 / max(dt1, 1e-3))))
        v, dt = cpu_assign_histogram(n_img)
        cpu = {"value": v, "unit": "Mdescriptors/s", "cores": cpu_threads(), "kind": "port",
               "sample": f"{n_img} of {C2['n_img']} images x {C2['per_img']} descriptors ({dt:.1f} s): per-image "
                         f"Faiss-shim IndexFlatIP.search + np.histogram + Okapi on NumPy/OpenBLAS"}
        if knn is not None:
            qv, qdt = cpu_knn(200_000, 1000)
            cpu["knn_qps"] = qv
            cpu["knn_sample"] = f"200k of 1M DB rows x 1000 of 10k queries ({qdt:.1f} s), scaled linearly in nb"

    if rank == 0:
        flops = 2.0 * C2["k"] * C2["d"] * C2["n_desc"]
        ach = flops / (kern_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "Mdescriptors/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16x2-split (fp32 accumulate)", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "l2": "inputs larger than L2 (512 MB descriptors + 328 MB histogram per step)",
                       "step": "prepare planes + gemm_select(top-1) + bovw_histogram(okapi)"},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "Mdescriptors/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "api": "BOVW.transform_csr(pinned PackedDescriptions, okapi=OkapiTransformer()) "
                    "-> scipy CSR float64 [10k x 4096]", "result_nnz": csr_nnz},
            "e2e_dense": {"value": world * C2["n_desc"] / (e2e_dense_ms * 1e-3) / 1e6, "unit": "Mdescriptors/s",
                          "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_dense, "ms_per_step": e2e_dense_ms,
                          "api": "BOVW.histograms_host(...) -> dense float64 [10k x 4096] (BOVW.transform's format)"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": P["tf_burst"], "unit": "TFLOP/s",
                         "frac": ach / P["tf_burst"], "traffic": 272.09e6,
                         "traffic_source": "ncu r01 final (profiles/r01_final_ncu.md): dram read 258.28 MB + write 13.81 MB per launch", "kernel": "gemm_select_kernel<2,2,IP,1>",
                         "kernel_ms": kern_ms, "search": assign_stats, "kernel_share_of_step": kern_ms / ms_step,
                         "peak_source": P["src"] + ", bf16 burst",
                         # FP32-grade scores need hi*hi + hi*lo(centroids): 2 tcgen05 products per algorithmic FLOP
                         "mma_products": 2, "achieved_mma_tflops": 2 * ach, "frac_mma": 2 * ach / P["tf_burst"]},
            "hbm_kernels": hbm_kernels,
            "cpu_baseline": cpu,
            "host_binding": None if numa_cpus is None else f"rank 0 bound to {len(numa_cpus)} CPUs local to its GPU (NVML)",
            "kmeans_iter_ms": kmeans_iter_ms,
            "knn": knn,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _protect_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line
    on stdout, so everything else is routed to stderr and the JSON goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-knn", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
