"""Throughput of the device normalisation statistics (spp_ring_obs_stats: 2 moment passes + 4 radix-select passes over every
agent's replay ring) against the HBM roofline, next to numpy on the host for one agent.  One JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from spp_rl_b200 import Population

P, OB, AC, S = 148, int(os.environ.get("OB", 11)), int(os.environ.get("AC", 3)), int(os.environ.get("S", 1_000_000))
pop = Population(algo="sac", ob_dim=OB, ac_dim=AC, population=P, buffer_size=S, update_batch_size=16, store_actions=False)
pop.set_norm_stats(-np.ones(OB, np.float32), np.ones(OB, np.float32))
pop.ring_fill_synthetic(seed=3, n=S * 999 // 1000, episode_len=1000)
pop.ring_obs_stats()                                   # warm-up (allocations)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    st, nbytes = pop.ring_obs_stats(return_bytes=True)  # synchronous: returns the [P, ob] results on the host
    ts.append(time.perf_counter() - t0)
dt = min(ts)
L = pop.ring_state(0)[2]
obs = pop.ring_sample_batch(0, np.arange(L, dtype=np.int64))[0].astype(np.float64)
t0 = time.perf_counter()
m, s_, p1, p99 = obs.mean(axis=0), obs.std(axis=0), np.percentile(obs, 1, axis=0), np.percentile(obs, 99, axis=0)
cpu = time.perf_counter() - t0
peak = 6537.3
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"kernel": "spp_ring_obs_stats (update_obs_mean_std for the whole population)", "agents": P, "rows_per_agent": L, "ob": OB,
                  "seconds": dt, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / dt / 1e9, "peak_gbs": peak, "frac": nbytes / dt / 1e9 / peak,
                  "agents_per_s": P / dt, "numpy_one_agent_seconds": cpu, "numpy_agents_per_s_per_core": 1.0 / cpu,
                  "bit_exact_percentiles_agent0": bool(np.array_equal(st["p1"][0], p1) and np.array_equal(st["p99"][0], p99))}))
pop.close()
