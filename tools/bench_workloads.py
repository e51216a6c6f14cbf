"""Device-resident update throughput of the other BASELINE configs (parity-tested shapes; the headline line is bench.py):
  config 3: SPP-DDPG HalfCheetah shapes (ob 17, ac 6, BasicAcM), population 256 per GPU, B = 256, rings of 100 k
  config 5: SPP-SAC Ant shapes (ob 111, ac 8), population 128 per GPU, B = 256
Prints one JSON line per workload (CUDA-event timed, W warm-up bursts, K timed bursts of G updates per agent)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from spp_rl_b200 import Population, init_state


def flops_per_update(algo, ob, ac, b, acm_kind):
    h = 256
    fa1 = 2 * ob * h
    fa = fa1 + 2 * h * h + 2 * h * (2 * ob if algo == "sac" else ob)
    fc1 = 2 * (ob + ac) * h
    fc = fc1 + 2 * h * h + 2 * h
    fm = 2 * (2 * ob * 64 + 64 * 32 + 32 * ac) if acm_kind == "acm" else 2 * (2 * ob * 100 + 100 * 50 + 2 * ob * 50 + 50 * ac)
    if algo == "sac":
        return b * (2 * fa + 6 * fc + 2 * fm + 2 * (2 * fc - fc1) + 2 * fc + fm + (2 * fa - fa1))
    return b * (2 * fa + 3 * fc + 2 * fm + (2 * fc - fc1) + fc + fm + (2 * fa - fa1))


WORKLOADS = {
    "ddpg_hcheetah": dict(algo="ddpg", ob=17, ac=6, acm_kind="basic", P=256, ring=100_000, gamma=0.95, lr=5e-4, custom_loss=1.0),
    "sac_ant": dict(algo="sac", ob=111, ac=8, acm_kind="acm", P=128, ring=100_000, gamma=0.99, lr=1e-3, custom_loss=0.2),
}


def run(name, w, G, K, W, B=256):
    ob, ac = w["ob"], w["ac"]
    pop = Population(algo=w["algo"], ob_dim=ob, ac_dim=ac, population=w["P"], acm_kind=w["acm_kind"], acm_critic=True, norm_closs=False,
                     min_max_denormalize=True, update_batch_size=B, buffer_size=w["ring"], store_actions=False, gamma=w["gamma"],
                     actor_lr=w["lr"], critic_lr=w["lr"], alpha_lr=w["lr"], custom_loss=w["custom_loss"], alpha=0.2, target_entropy=-float(ac))
    rng = np.random.RandomState(1)
    pop.set_limits(np.ones(ob, np.float32), np.ones(ac, np.float32))
    pop.set_norm_stats((-1 - rng.rand(ob)).astype(np.float32), (1 + rng.rand(ob)).astype(np.float32))
    nets = ["actor", "critic_1", "critic_2", "critic_1_targ", "critic_2_targ", "acm"] if w["algo"] == "sac" else \
        ["actor", "actor_targ", "critic", "critic_targ", "acm"]
    for a in range(w["P"]):
        s0 = init_state(w["algo"], ob, ac, 1000 + a, w["acm_kind"], True)      # fresh nn.Linear-style init per agent (independent seeds)
        for net in nets:
            pop.load_state_dict(net, {k[len(net) + 1:]: v for k, v in s0.items() if k.startswith(net + ".")}, agent=a)
    pop.ring_fill_synthetic(seed=7, n=w["ring"] * 999 // 1000, episode_len=1000)
    stream = torch.cuda.Stream()
    for i in range(W):
        pop.update_ring_device(G, seed=100 + i, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for k in range(K):
            pop.update_ring_device(G, seed=1000 + k, stream=stream.cuda_stream)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ups = w["P"] * G * K / (ms * 1e-3)
    f = flops_per_update(w["algo"], ob, ac, B, w["acm_kind"])
    pop.close()
    return {"workload": name, "algo": w["algo"], "ob": ob, "ac": ac, "acm": w["acm_kind"], "agents_per_gpu": w["P"], "batch": B, "grad_steps": G,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "updates_per_s": ups, "flops_per_update": f, "algorithmic_tflops": f * ups / 1e12,
            "ring_capacity": w["ring"], "data": "synthetic"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grad-steps", type=int, default=50)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    for name, w in WORKLOADS.items():
        if a.only and a.only != name:
            continue
        print(json.dumps(run(name, w, a.grad_steps, a.steps, a.warmup)), flush=True)
