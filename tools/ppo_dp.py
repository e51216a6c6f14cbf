"""SPP-PPO data-parallel check and timing: launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tools/ppo_dp.py [--envs E --steps T]
Environments shard over ranks (trajectories never cross GPUs); every optimiser step all-reduces the gradient vector over
NCCL, the advantage statistics are all-reduced in fp64, and all ranks hold identical weights afterwards.  Rank 0 also runs
the same iteration on ONE GPU over all rows and reports the norm-relative difference of the post-iteration weights.
Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from spp_rl_b200.ppo import PpoPolicy
from spp_rl_b200.sharding import env_shard_rows

OB, AC = 17, 6      # Walker2d shapes


def synth(E, T, seed=3):
    rng = np.random.RandomState(seed)
    N = E * T
    mn, mx = (-rng.rand(OB) * 2 - 0.5).astype(np.float32), (rng.rand(OB) * 2 + 0.5).astype(np.float32)
    obs = (rng.rand(N, OB) * (mx - mn) + mn).astype(np.float32)
    d = dict(mn=mn, mx=mx, obs=obs, nobs=(obs + 0.05 * rng.randn(N, OB)).astype(np.float32), act=(0.5 * rng.randn(N, OB)).astype(np.float32),
             logp=(-10 + rng.randn(N)).astype(np.float32), rew=rng.randn(N).astype(np.float32), done=(rng.rand(N) < 1e-3).astype(np.float32))
    end = d["done"].copy()
    end[(T - 1) * E:] = 1
    d["end"] = end
    w = {}
    for net, out in (("actor", OB), ("critic", 1)):
        for name, o, i in (("fc1", 64, OB), ("fc2", 64, 64), ("fc3", out, 64)):
            b = 1 / np.sqrt(i)
            w[net + "." + name + ".weight"] = rng.uniform(-b, b, (o, i)).astype(np.float32)
            w[net + "." + name + ".bias"] = rng.uniform(-b, b, (o,)).astype(np.float32)
    w["actor.log_scale"] = np.full((OB,), -1.34, np.float32)
    return d, w


def make_policy(device, rows, batch_rows, d, w):
    pol = PpoPolicy(OB, AC, max_rows=rows, max_batch_rows=batch_rows, device=device, min_max_denormalize=True, gamma=0.99, gae_lambda=0.95,
                    custom_loss=0.1, entropy_coef=0.0, actor_lr=3e-4, critic_lr=3e-4)
    pol.set_norm_stats(d["mn"], d["mx"])
    for net in ("actor", "critic"):
        pol.load_state_dict(net, {k[len(net) + 1:]: v for k, v in w.items() if k.startswith(net + ".")})
    return pol


def relnorm(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=512)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--batch", type=int, default=16384, help="global PPO minibatch")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--critic-targets", type=int, default=2)
    ap.add_argument("--critic-steps", type=int, default=5)
    ap.add_argument("--dump", default="", help="npz path: post-iteration weights of the data-parallel and the single-GPU run (rank 0)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    E, T = args.envs, args.steps
    N = E * T
    assert E % world == 0
    El = E // world
    d, w = synth(E, T)
    # local rows: environment slice [rank * El, (rank + 1) * El) of every step, step-major with stride El
    rows = env_shard_rows(E, T, rank, world)
    pol = make_policy(local, El * T, args.batch, d, w)
    pol.comm_init(dist)      # NCCL communicator inside the library: every optimiser step all-reduces on the policy's own stream
    pol.load_rollout(d["obs"][rows], d["nobs"][rows], d["act"][rows], d["logp"][rows], d["rew"][rows], d["done"][rows], d["end"][rows],
                     np.arange(El), np.full(El, T), traj_stride=El, global_rows=N)
    rng = np.random.RandomState(11)
    perms = np.stack([rng.permutation(N) for _ in range(args.epochs)]).astype(np.int64)

    def reset(p):      # the initial weights with a fresh optimiser state
        for net in ("actor", "critic"):
            p.load_state_dict(net, {k[len(net) + 1:]: v for k, v in w.items() if k.startswith(net + ".")})
            p.adam_reset(net)
    # warm-up: one shortened iteration (first-use costs of the kernels and of torch's device ops), then back to the start state
    pol.iteration_dp(perms[:1], args.batch, E, 1e9, 1, 1, rank, world)
    reset(pol)
    torch.cuda.synchronize(); dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    res = pol.iteration_dp(perms, args.batch, E, 1e9, args.critic_targets, args.critic_steps, rank, world)
    pol.sync()
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sd_dp = {net: pol.state_dict(net) for net in ("actor", "critic")}
    # every rank must hold the same weights
    flat = torch.from_numpy(np.concatenate([v.ravel() for net in sd_dp for v in sd_dp[net].values()])).cuda()
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([float(torch.equal(flat, ref))], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    out = None
    if rank == 0:
        one = make_policy(local, N, args.batch, d, w)
        one.load_rollout(d["obs"], d["nobs"], d["act"], d["logp"], d["rew"], d["done"], d["end"], np.arange(E), np.full(E, T), traj_stride=E)
        one.iteration_dp(perms[:1], args.batch, E, 1e9, 1, 1, 0, 1)
        reset(one)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        res1 = one.iteration_dp(perms, args.batch, E, 1e9, args.critic_targets, args.critic_steps, 0, 1)
        one.sync()
        s1.record(); torch.cuda.synchronize()
        worst = 0.0
        dump = {}
        for net in ("actor", "critic"):
            sd1 = one.state_dict(net)
            for k, v in sd1.items():
                e = relnorm(sd_dp[net][k], v)
                worst = max(worst, e)
                dump["dp:%s.%s" % (net, k)] = sd_dp[net][k]
                dump["one:%s.%s" % (net, k)] = v
        one.close()
        if args.dump:
            np.savez(args.dump, critic_loss_dp=res["critic_loss"], critic_loss_one=res1["critic_loss"], **dump)
        info = pol.comm_info()
        out = {"metric": "SPP-PPO iteration (critic fit + GAE + advantage normalisation + clipped-ratio actor epochs), data-parallel over environments",
               "n_gpus": world, "envs": E, "steps": T, "rows": N, "global_minibatch": args.batch, "epochs": res["epochs"],
               "critic_steps": args.critic_targets * args.critic_steps, "allreduces": res["allreduces"], "nccl_version": info["nccl_version"], "p2p": info["p2p"], "p2p_steps": info["p2p_steps"],
               "ms_dp": float(ms.item()), "ms_phases_rank0": res["phases_ms"],
               "transitions_per_s_dp": N / (float(ms.item()) * 1e-3), "ms_single_gpu": s0.elapsed_time(s1), "ms_phases_single_gpu": res1["phases_ms"],
               "dp_vs_single_worst_relnorm": worst, "ranks_bit_identical": bool(same.item() == 1.0),
               "critic_loss_dp": res["critic_loss"], "critic_loss_single_gpu": res1["critic_loss"]}
        print(json.dumps(out), flush=True)
    pol.close()
    dist.destroy_process_group()
    if rank == 0 and (out["dp_vs_single_worst_relnorm"] > 2e-5 or not out["ranks_bit_identical"]):
        sys.exit(1)


if __name__ == "__main__":
    main()
