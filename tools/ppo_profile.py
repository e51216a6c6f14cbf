"""One single-GPU SPP-PPO iteration at config-4 shapes (Walker2d, E envs x T steps) through PpoPolicy, phase by phase with
CUDA-synchronised wall times: critic fit, advantages, normalisation, actor epochs.  For the launch list run it under
`ncu --metrics gpu__time_duration.sum`."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from spp_rl_b200.ppo import PpoPolicy
from tools.ppo_dp import AC, OB, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--critic-targets", type=int, default=10)
    ap.add_argument("--critic-steps", type=int, default=10)
    a = ap.parse_args()
    E, T = a.envs, a.steps
    N = E * T
    d, w = synth(E, T)
    pol = PpoPolicy(OB, AC, max_rows=N, max_batch_rows=a.batch, min_max_denormalize=True, custom_loss=0.5)
    pol.set_limits(1.0); pol.set_norm_stats(d["mn"], d["mx"])
    for net in ("actor", "critic"):
        pol.load_state_dict(net, {k[len(net) + 1:]: v for k, v in w.items() if k.startswith(net + ".")})
    t = {}

    def timed(name, f):
        torch.cuda.synchronize(); pol.sync(); t0 = time.perf_counter()
        r = f()
        pol.sync(); t[name] = (time.perf_counter() - t0) * 1e3
        return r
    ts, tl = np.arange(E, dtype=np.int64), np.full(E, T, np.int64)
    timed("load_rollout_h2d", lambda: pol.load_rollout(d["obs"], d["nobs"], d["act"], d["logp"], d["rew"], d["done"], d["end"], ts, tl, traj_stride=E))
    timed("critic_fit", lambda: pol.update_critic(a.critic_targets, a.critic_steps))
    timed("advantages", lambda: pol.advantages(want_host=False))
    timed("normalize_adv", lambda: pol.normalize_adv())
    rng = np.random.RandomState(1)
    perms = np.stack([rng.permutation(N) for _ in range(a.epochs)]).astype(np.int64)
    _, epochs, _ = timed("actor_epochs", lambda: pol.update_actor(perms, a.batch, 1e9, a.epochs))
    print(json.dumps({"rows": N, "critic_steps": a.critic_targets * a.critic_steps, "actor_epochs": epochs + 1, "ms": t}))


if __name__ == "__main__":
    main()
