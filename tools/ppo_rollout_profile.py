"""The device-resident SPP-PPO rollout alone (ppo_rollout_kernel): E vectorised synthetic environments x T steps, Walker2d shapes.
python tools/ppo_rollout_profile.py [E] [T] [reps] -> one JSON line (ms per launch, transitions/s).  For ncu: -k regex:ppo_rollout."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from spp_rl_b200 import Population
from spp_rl_b200.ppo import PpoPolicy


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    ob, ac = 17, 6
    rng = np.random.RandomState(3)
    pol = PpoPolicy(ob, ac, max_rows=E * T, max_batch_rows=1024, min_max_denormalize=True)
    pol.set_norm_stats((-rng.rand(ob) * 2 - 0.5).astype(np.float32), (rng.rand(ob) * 2 + 0.5).astype(np.float32))
    sd = {}
    for name, o, i in (("fc1", 64, ob), ("fc2", 64, 64), ("fc3", ob, 64)):
        b = 1 / np.sqrt(i)
        sd[name + ".weight"] = rng.uniform(-b, b, (o, i)).astype(np.float32)
        sd[name + ".bias"] = rng.uniform(-b, b, (o,)).astype(np.float32)
    sd["log_scale"] = np.full((ob,), -1.34, np.float32)
    pol.load_state_dict("actor", sd)
    pop = Population(algo="ddpg", ob_dim=ob, ac_dim=ac, population=1, acm_kind="acm", min_max_denormalize=True, update_batch_size=64,
                     buffer_size=4096, store_actions=False)
    pop.set_limits(np.ones(ob, np.float32), np.ones(ac, np.float32))
    pol.rollout_synthetic(pop, E, T, seed=1, reset_envs=True); pol.sync()
    t0 = time.perf_counter()
    for r in range(reps):
        pol.rollout_synthetic(pop, E, T, seed=2 + r)
    pol.sync()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    print(json.dumps({"envs": E, "steps": T, "ms_per_launch": ms, "us_per_step": ms * 1e3 / T, "transitions_per_s": E * T / (ms * 1e-3)}))
    pol.close(); pop.close()


if __name__ == "__main__":
    main()
