"""How long does the one-off generation of the split_clusters RNG stream index take on this host?"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from image_search_engine_b200 import ops
h = np.full(65536, 150, np.float32); h[np.random.default_rng(0).choice(65536, 1700, replace=False)] = 0
t = time.perf_counter(); p, _ = ops.split_plan(h, 10_000_000); t_cold = time.perf_counter() - t
t = time.perf_counter(); p, _ = ops.split_plan(h, 10_000_000); t_warm = time.perf_counter() - t
import os
print(f"cold plan (generates ~112 M draws): {t_cold * 1e3:.1f} ms; warm plan: {t_warm * 1e3:.2f} ms; cpus: {os.cpu_count()}, affinity {len(os.sched_getaffinity(0))}")
