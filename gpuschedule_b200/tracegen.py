"""Synthetic trace generator in the live-path CSV schema.

The reference never committed a trace (its .gitignore excludes data/*.csv), so
every workload in this repo is generated here, deterministically, from a seed.
Schema = the columns the live ingest consumes
(/root/reference/core/jobs/job_generator.py:181-193 and
 /root/reference/core/jobs/jobs_manager.py:233-238):

  type, normalized_time, minutes, gpu_per_container, gpu_utilization_avg,
  gpu_utilization_max, memory_max, memory_avg, used_gpus

plus three optional columns used only by the network-cost model
(/root/reference/core/network/network_service.py:3-39):
  model_name, iterations, ps_count

Distributions follow SURVEY.md section 8(d):
  inter-arrival ~ Exponential(mean 1/rate ticks), cumsum -> floor -> x10000
  used_gpus in {1,2,4,8,16,32} w.p. {.35,.20,.20,.15,.07,.03}
  minutes = clip(LogNormal(4,1), 2, 4000) rounded to 3 dp
  util_avg ~ U(5,90); util_max = min(100, avg + U(1,30))  (3 dp)
  memory_max = randint(512, 16384) MiB as integer bytes; memory_avg a fraction
"""
from __future__ import annotations

import numpy as np

# Iteration-count sample of the reference's distribution-driven generator
# (/root/reference/core/jobs/job_generator.py:24-31) -- values only, used as a
# sampling population for the optional `iterations` column.
_ITER_POP = np.array(
    [1, 1, 1, 1, 1, 1, 109, 126, 133, 138, 141, 143, 144, 147, 157, 168, 175,
     192, 193, 198, 235, 237, 242, 253, 258, 272, 272, 274, 288, 326, 326,
     362, 386, 391, 410, 438, 447, 468, 473, 513, 513, 521, 521, 525, 581,
     606, 607, 775, 775, 789, 822, 864, 864, 892, 903, 949, 1011, 1085, 1360,
     1501, 2178, 2239, 2275, 3304, 3469, 4861], dtype=np.int64)

GPU_CHOICES = np.array([1, 2, 4, 8, 16, 32], dtype=np.int64)
GPU_PROBS = np.array([.35, .20, .20, .15, .07, .03])

BASE_SEED = 20260921


def synth_columns(n_jobs: int, seed: int = 1, rate: float = 0.5,
                  gpu_choices=GPU_CHOICES, gpu_probs=GPU_PROBS,
                  gpu_per_container: int = 1, with_network: bool = False,
                  max_mem_mib: int = 16384):
    """Return a dict of column -> numpy array for an `n_jobs` trace."""
    from .model_factory import model_sizes
    rng = np.random.default_rng(BASE_SEED + int(seed))
    gaps = rng.exponential(1.0 / rate, size=n_jobs)
    arrive = np.floor(np.cumsum(gaps)).astype(np.int64)
    arrive -= arrive[0]
    cols = {}
    cols["type"] = np.array(["noninteractive"] * n_jobs, dtype=object)
    cols["normalized_time"] = arrive * 10000
    cols["minutes"] = np.round(
        np.clip(rng.lognormal(4.0, 1.0, size=n_jobs), 2.0, 4000.0), 3)
    cols["gpu_per_container"] = np.full(n_jobs, int(gpu_per_container), dtype=np.int64)
    avg = np.round(rng.uniform(5.0, 90.0, size=n_jobs), 3)
    mx = np.round(np.minimum(100.0, avg + rng.uniform(1.0, 30.0, size=n_jobs)), 3)
    cols["gpu_utilization_avg"] = avg
    cols["gpu_utilization_max"] = np.maximum(mx, avg)
    mem_mib = rng.integers(512, max_mem_mib, size=n_jobs, endpoint=True)
    cols["memory_max"] = mem_mib.astype(np.int64) * (1 << 20)
    cols["memory_avg"] = np.floor(
        cols["memory_max"] * rng.uniform(0.3, 0.9, size=n_jobs)).astype(np.int64)
    g = rng.choice(np.asarray(gpu_choices), size=n_jobs, p=np.asarray(gpu_probs))
    g = np.maximum(g // gpu_per_container, 1) * gpu_per_container
    cols["used_gpus"] = g.astype(np.int64)
    cols["model"] = np.array(["V100"] * n_jobs, dtype=object)
    if with_network:
        names = np.array(sorted(model_sizes.keys()), dtype=object)
        cols["model_name"] = names[rng.integers(0, len(names), size=n_jobs)]
        cols["iterations"] = _ITER_POP[rng.integers(0, len(_ITER_POP), size=n_jobs)]
        cols["ps_count"] = (cols["used_gpus"] // cols["gpu_per_container"]).astype(np.int64)
    return cols


def synth_frame(n_jobs: int, **kw):
    import pandas as pd
    return pd.DataFrame(synth_columns(n_jobs, **kw))


def write_trace(path: str, n_jobs: int, **kw) -> str:
    synth_frame(n_jobs, **kw).to_csv(path, index=False)
    return path


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="write a synthetic live-schema trace CSV")
    ap.add_argument("out")
    ap.add_argument("--jobs", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rate", type=float, default=0.5)
    ap.add_argument("--network", action="store_true")
    a = ap.parse_args()
    write_trace(a.out, a.jobs, seed=a.seed, rate=a.rate, with_network=a.network)
