#!/usr/bin/env python
"""bench.py -- SPP-SAC grad updates/sec (Hopper shape) on B200, per the driver contract.

A "step" is one burst of the hot path: every agent of the population performs G = 50 consecutive
SAC_AcM updates (the reference's grad_steps, train/spp_sac_hopper.py:22), each on a B = 256
minibatch gathered from that agent's device-resident replay ring (capacity 1 M transitions).
  value  updates/s with everything resident in HBM (device sampler + device noise), CUDA-event timed.
  e2e    the same metric through the C-ABI host call spp_update_host() -- the reference's
         update(obs, next_obs, action, reward, done, acm_action) signature -- with pinned HOST
         minibatches copied H2D and the loss table copied D2H inside the timed region.
  roofline      algorithmic FLOPs of the update (SURVEY 8d formula) / kernel time vs the measured peak.
  cpu_baseline  the CPU oracle port (plain-tensor restatement of rltoolkit's update) on the host cores.
`--impl reference` times that CPU port alone (the reference is Python; /root/reference does not
exist on the GPU box, so the oracle port is the travelling stand-in).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SPP-SAC grad updates/sec (Hopper shape) + acm rollout transitions/sec"
OB, AC, B, G = 11, 3, 256, 50
RING = 1_000_000
RING_FILL = 999_000            # 999 episodes x 1000 steps (+ 999 terminal observations) fit the 1 M ring
# Hopper-scale observation statistics (synthetic; same order of magnitude as a trained SPP-SAC Hopper run)
MIN_OBS = [0.73, -0.20, -1.34, -1.43, -0.89, -1.26, -2.92, -4.62, -5.10, -8.25, -10.0]
MAX_OBS = [1.68, 0.18, 0.02, 0.04, 0.90, 4.59, 2.83, 2.58, 3.96, 7.69, 10.0]


def flops_per_update(ob=OB, ac=AC, b=B, hidden=256, hm1=64, hm2=32):
    """B*(2Fa + 6Fc + 2Fm + 2(2Fc - Fc1) + 2Fc + Fm + (2Fa - Fa1)), F = 2*MACs of one forward."""
    fa1 = 2 * ob * hidden
    fa = fa1 + 2 * hidden * hidden + 2 * hidden * 2 * ob
    fc1 = 2 * (ob + ac) * hidden
    fc = fc1 + 2 * hidden * hidden + 2 * hidden
    fm = 2 * (2 * ob * hm1 + hm1 * hm2 + hm2 * ac)
    return b * (2 * fa + 6 * fc + 2 * fm + 2 * (2 * fc - fc1) + 2 * fc + fm + (2 * fa - fa1))


# ------------------------------------------------------------------------------------------- CPU side
# The CPU arm times the reference's OWN SAC_AcM.update (rltoolkit/acm/off_policy/sac_acm.py:89-162) when the reference package is
# available -- oracle/_ref (copied byte for byte by oracle/build_ref.py; it travels to the GPU box) or /root/reference -- through
# the import shims of oracle/ref_import.py; otherwise the oracle's plain-tensor port.  Parallel model = the reference's: one
# single-thread run per host core (mp.Pool in train/spp_sac_hopper.py:115, torch.set_num_threads(1) in rltoolkit/evals.py:22-26).
def _cpu_batch(seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    mn, mx = torch.tensor(MIN_OBS), torch.tensor(MAX_OBS)
    obs = torch.rand(B, OB, generator=g) * (mx - mn) + mn
    nobs = obs + 0.02 * torch.randn(B, OB, generator=g)
    act = torch.randn(B, OB, generator=g)
    rew = torch.randn(B, generator=g)
    done = (torch.rand(B, generator=g) < 1e-3).to(torch.int8)
    aacm = torch.tanh(torch.randn(B, AC, generator=g))
    return obs, nobs, act, rew, done, aacm


def _make_cpu_updater(kind, seed, threads):
    """-> a closure running ONE SAC_AcM.update on a fixed B = 256 Hopper-shaped minibatch (fresh N(0,1) draws inside, as rsample does)."""
    import numpy as np
    import torch

    torch.set_num_threads(threads)
    obs, nobs, act, rew, done, aacm = _cpu_batch(seed)
    if kind == "reference":
        import tempfile

        from oracle.ref_import import import_reference
        rl = import_reference(scratch_dir=tempfile.mkdtemp(prefix="spp_ref_"))
        torch.manual_seed(seed)
        m = rl.SAC_AcM(env_name="Hopper-v2", update_batch_size=B, custom_loss=0.2, acm_critic=True, norm_closs=False,
                       denormalize_actor_out=True, min_max_denormalize=True, acm_pre_train_samples=100, acm_val_buffer_size=None,
                       buffer_size=1000, tensorboard_dir=None, log_dir=None, gamma=0.99, actor_lr=1e-3, critic_lr=1e-3, alpha_lr=1e-3,
                       alpha=0.2, verbose=0)
        m.replay_buffer.min_obs, m.replay_buffer.max_obs = torch.tensor(MIN_OBS), torch.tensor(MAX_OBS)
        return lambda: m.update(obs, nobs, act, rew, done, aacm)
    from oracle import offpolicy as op
    from oracle.norm import NormStats
    from spp_rl_b200.init import init_state

    s = {k: torch.from_numpy(v) for k, v in init_state("sac", OB, AC, seed).items()}
    s["log_alpha"] = torch.tensor(np.log(0.2), dtype=torch.float64)
    st = NormStats(True, torch.tensor(MIN_OBS), torch.tensor(MAX_OBS))
    hp = op.OffPolicyHP(gamma=0.99, custom_loss=0.2, norm_closs=False, acm_critic=True, target_entropy=-float(AC),
                        actor_lim=torch.ones(OB), acm_lim=torch.ones(AC))
    g = torch.Generator().manual_seed(seed + 1)
    state = {"alpha": None}

    def step():
        _, state["alpha"] = op.sac_acm_update(s, hp, st, obs, nobs, None, rew, done, aacm, torch.randn(B, OB, generator=g),
                                              torch.randn(B, OB, generator=g), state["alpha"])
    return step


def _cpu_worker_loop(conn, kind, seed, threads):
    try:
        step = _make_cpu_updater(kind, seed, threads)
        for _ in range(3):
            step()
        conn.send(("ready", 0.0))
        while True:
            n = conn.recv()
            if n <= 0:
                break
            t0 = time.perf_counter()
            for _ in range(n):
                step()
            conn.send(("done", time.perf_counter() - t0))
    except Exception as e:      # surfaced by CpuArm
        conn.send(("error", repr(e)))


def cpu_kind():
    here = os.path.join(ROOT, "oracle", "_ref", "rltoolkit")
    return "reference" if (os.path.isdir(here) or os.path.isdir("/root/reference/rltoolkit/rltoolkit")) else "port"


class CpuArm:
    """`procs` persistent worker processes (forked before any CUDA initialisation), each holding its own model; step(n) runs n
    updates in every worker concurrently and returns the wall time of the slowest."""

    def __init__(self, procs, threads=1, kind=None):
        import multiprocessing as mp
        self.kind = kind or cpu_kind()
        ctx = mp.get_context("fork")
        self.workers = []
        for i in range(procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker_loop, args=(b, self.kind, 50 + i, threads), daemon=True)
            p.start()
            self.workers.append((p, a))
        for _, c in self.workers:
            tag, v = c.recv()
            if tag != "ready":
                self.close()
                raise RuntimeError("CPU worker failed: %s" % (v,))

    def step(self, n):
        t0 = time.perf_counter()
        for _, c in self.workers:
            c.send(n)
        spans = []
        for _, c in self.workers:
            tag, v = c.recv()
            if tag != "done":
                raise RuntimeError("CPU worker failed: %s" % (v,))
            spans.append(v)
        return time.perf_counter() - t0, max(spans)

    def close(self):
        for p, c in self.workers:
            try:
                c.send(0)
            except Exception:
                pass
        for p, _ in self.workers:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()


def cpu_baseline_block(cores):
    """cpu_baseline of the N = 1 line: a bounded sample (about 15-25 s of CPU work) of the same update on the box's host cores."""
    kind = cpu_kind()
    n = 60 if kind == "reference" else 40
    arm = CpuArm(cores, 1, kind)
    try:
        wall, span = arm.step(n)
    finally:
        arm.close()
    out = {"value": cores * n / span, "unit": "updates/s", "cores": cores, "kind": kind, "n_cores_host": os.cpu_count(),
           "per_core": n / span,
           "sample": "%d single-thread processes x %d updates of the same SAC Hopper B=256 update (+3 warm-up each), %s"
                     % (cores, n, "the reference's own SAC_AcM.update from oracle/_ref" if kind == "reference" else "oracle port")}
    one = CpuArm(1, cores, kind)      # plus one run with all intra-op threads
    try:
        _, span1 = one.step(30)
    finally:
        one.close()
    out["all_threads_one_process"] = {"value": 30 / span1, "unit": "updates/s", "threads": cores}
    return out


def reference_ppo_block(transitions=16384, limit_s=300):
    """one perform_iteration of the reference's own PPO_AcM on one host core at a reduced batch, phase by phase
    (tools/ref_ppo_phases.py, its own process: single thread, the gym shim's shape-only Walker2d); reported per transition beside the
    config-4 block.  A reported baseline; never on the product path."""
    return _run_ref_tool("ref_ppo_phases.py", transitions, limit_s)


def reference_rollout_block(transitions=10000, limit_s=180):
    """the body of the reference's frame loop (normalize -> noise_action -> process_action -> add_obs / add_timestep,
    ddpg.py:205-220) for a SAC_AcM at Hopper shapes on one host core with a no-op environment (tools/ref_rollout_step.py): the CPU
    number beside the rollout half of the metric.  A reported baseline; never on the product path."""
    return _run_ref_tool("ref_rollout_step.py", transitions, limit_s)


def _run_ref_tool(script, arg, limit_s):
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", script), str(arg)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, timeout=limit_s, cwd=ROOT)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"unavailable": "tools/%s exited %d: %s" % (script, r.returncode, (r.stderr or "").strip().splitlines()[-1:] or "")}
        return json.loads(lines[-1])
    except Exception as e:      # noqa: BLE001  (a baseline leg must not take the bench line with it)
        return {"unavailable": repr(e)[:200]}


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------- GPU side
HOPPER_PKL = os.path.join(ROOT, "tests", "golden", "models", "hopper_sac_acm_model.pkl")


def build_population(device, P, ring=RING):
    """every agent starts from the reference's trained models/hopper_sac_acm_model.pkl (targets = copies of the
    critics), observation statistics from the same pickle, custom_loss 0.2, acm_critic, min-max denormalisation, gamma 0.99,
    lr 1e-3, tau 0.005, alpha 0.2; the rings are prefilled synthetically (obs ~ U(min_obs, max_obs), episodes of 1000)."""
    import pickle

    import numpy as np
    from spp_rl_b200 import Population

    pop = Population(algo="sac", ob_dim=OB, ac_dim=AC, population=P, device=device, acm_kind="acm", acm_critic=True,
                     norm_closs=False, min_max_denormalize=True, update_batch_size=B, buffer_size=ring,
                     store_actions=False, gamma=0.99, tau=0.005, actor_lr=1e-3, critic_lr=1e-3, alpha_lr=1e-3,
                     custom_loss=0.2, alpha=0.2, target_entropy=-float(AC))
    pop.set_limits(np.ones(OB, np.float32), np.ones(AC, np.float32))
    with open(HOPPER_PKL, "rb") as f:
        d = pickle.load(f)
    pop.set_norm_stats(d["min_obs"].numpy(), d["max_obs"].numpy())
    for net, src in (("actor", "actor"), ("critic_1", "critic_1"), ("critic_2", "critic_2"), ("critic_1_targ", "critic_1"),
                     ("critic_2_targ", "critic_2"), ("acm", "acm")):
        pop.load_state_dict(net, d[src])      # agent = -1: every agent
    pop.ring_fill_synthetic(seed=7, n=ring * 999 // 1000, episode_len=1000)
    return pop


def ppo_cfg4_block(local_rank, rank, world, dist, K, W, E_total=4096, T=2048, batch=65536, epochs=2, critic_targets=10, critic_steps=10,
                   acm_batches=100, acm_ring=110_000):
    """BASELINE config 4 -- SPP-PPO Walker2d shapes, 4096 vectorised synthetic envs x 2048-step rollouts -- as WHOLE iterations of
    PPO_AcM.perform_iteration (on_policy.py:55-86): device rollout -> critic fit (10 x 10 full-batch steps) -> GAE -> advantage
    normalisation -> clipped-ratio actor epochs (65 536-row global minibatches) -> add_buffer -> ACM update batches.  Environments
    shard over ranks (strong scaling: the 4096 environments are fixed); every optimiser step all-reduces the gradient vector with
    NCCL inside the library.  The ACM replicas regress on their own rank's ring (local ACMs, no collective)."""
    import numpy as np
    import torch

    from spp_rl_b200 import Population
    from spp_rl_b200.ppo import PpoPolicy

    ob, ac = 17, 6
    if E_total % world:
        return {"skipped": "environments do not divide over %d ranks" % world}
    El = E_total // world
    N, Nl = E_total * T, El * T
    rng = np.random.RandomState(3)
    mn, mx = (-rng.rand(ob) * 2 - 0.5).astype(np.float32), (rng.rand(ob) * 2 + 0.5).astype(np.float32)
    pol = PpoPolicy(ob, ac, max_rows=Nl, max_batch_rows=batch, device=local_rank, min_max_denormalize=True, norm_closs=True, gamma=0.99,
                    gae_lambda=0.95, custom_loss=0.5, entropy_coef=0.0, actor_lr=3e-4, critic_lr=3e-4)
    pol.set_norm_stats(mn, mx)
    for net, out in (("actor", ob), ("critic", 1)):
        sd = {}
        for name, o, i in (("fc1", 64, ob), ("fc2", 64, 64), ("fc3", out, 64)):
            b = 1 / np.sqrt(i)
            sd[name + ".weight"] = rng.uniform(-b, b, (o, i)).astype(np.float32)
            sd[name + ".bias"] = rng.uniform(-b, b, (o,)).astype(np.float32)
        if net == "actor":
            sd["log_scale"] = np.full((ob,), -1.34, np.float32)
        pol.load_state_dict(net, sd)
    pop = Population(algo="ddpg", ob_dim=ob, ac_dim=ac, population=1, device=local_rank, acm_kind="acm", acm_critic=True, min_max_denormalize=True,
                     update_batch_size=64, acm_batch_size=256, buffer_size=acm_ring, store_actions=False, acm_lr=1e-4)
    pop.set_norm_stats(mn, mx)
    pop.set_limits(np.ones(ob, np.float32), np.ones(ac, np.float32))
    if world > 1:
        pol.comm_init(dist)
    pol.set_reserved_sms(1)      # the ACM burst (one 213 KB CTA on the population's stream) runs beside the policy update
    st = pol._ext_stream()
    gen = torch.Generator(device="cuda")
    phases = {}

    import threading

    def iteration(i, timed):
        """rollout, then TWO independent halves side by side: the policy update on the policy's stream (this thread) and add_buffer +
        the ACM regression batches on the population's stream (a second host thread; ctypes drops the GIL inside the library).  The
        reference runs them one after the other (on_policy.py:55-86), but neither reads what the other writes, so the results are the
        same; both finish before the next rollout, which needs the updated actor AND the updated ACM."""
        t = [time.perf_counter()]
        pol.rollout_synthetic(pop, El, T, max_ep_len=1000, done_prob=0.001, seed=1000 * i + rank, reset_envs=(i == 0))
        pol.set_global_rows(N)
        pol.sync(); t.append(time.perf_counter())
        acm_ms = {}

        def acm_side():
            a0 = time.perf_counter()
            pop.ring_add_rollout_store(0, pol)
            a1 = time.perf_counter()
            pop.acm_update_ring(acm_batches, idx=None, seed=5 + i)
            pop.sync()
            acm_ms["add_buffer"], acm_ms["acm_update"] = (a1 - a0) * 1e3, (time.perf_counter() - a1) * 1e3
        th = threading.Thread(target=acm_side)
        th.start()
        gen.manual_seed(77 + i)      # the same global permutations on every rank
        with torch.cuda.stream(st):
            perms = [torch.randperm(N, device="cuda", generator=gen) for _ in range(epochs)]
        res = pol.iteration_dp(perms, batch, E_total, 1e9, critic_targets, critic_steps, rank, world)
        th.join()
        t.append(time.perf_counter())
        if timed:
            phases["rollout"] = phases.get("rollout", 0.0) + (t[1] - t[0]) * 1e3
            phases["update_and_acm_side_overlapped"] = phases.get("update_and_acm_side_overlapped", 0.0) + (t[2] - t[1]) * 1e3
            for name, v in list(res["phases_ms"].items()) + list(acm_ms.items()):
                phases[name] = phases.get(name, 0.0) + v
        return res

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    for w in range(max(W, 1)):
        iteration(w, False)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        res = iteration(max(W, 1) + k, True)
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item()) * 1e3 / K
    info = pol.comm_info()
    out = {"workload": "SPP-PPO Walker2d shapes (ob 17, ac 6, hidden 64): %d vectorised synthetic envs x %d steps = %d transitions per iteration; "
                       "whole PPO_AcM.perform_iteration on the device" % (E_total, T, N),
           "scaling": "strong", "n_gpus": world, "ms_per_iteration": ms, "transitions_per_s": N / (ms * 1e-3),
           "phases_ms_rank0": {k: v / K for k, v in phases.items()}, "critic_steps": critic_targets * critic_steps, "actor_epochs": res["epochs"],
           "global_minibatch": batch, "allreduces_per_iteration": res["allreduces"], "nccl_version": info["nccl_version"],
           "allreduce_path": ("fused reduce + all-reduce kernel over NVLink peer memory (csrc/ppo_p2p.cu)" if info.get("p2p") else ("NCCL" if world > 1 else "none")),
           "acm_update_batches": acm_batches, "critic_loss": res["critic_loss"], "kl": res["kl"], "adv_split_ms_last": list(getattr(pol, "_adv_split_ms", ())),
           "timing": "host clock between barriers + device synchronisation on both sides (the iteration spans two streams and host index work), max over ranks"}
    pol.close(); pop.close()
    return out


def population_train_block(local_rank, rank, world, dist, name, algo, ob, ac, agents_per_gpu, frames, warm_frames, ring=60_000, envs_per_agent=1):
    """BASELINE configs 3 / 5 through the population-batched train loop (Population.train_synthetic): rollout -> update bursts ->
    ACM batches -> ring statistics, scheduled as ddpg.py:191-237 / ddpg_acm.py:52-85 (update_freq 50, grad_steps 50, random_frames 100,
    iterations of 1000 frames); agents shard over ranks with no collective.  frames/s = env frames of all agents / wall time."""
    import numpy as np
    import torch

    from spp_rl_b200 import Population, init_state

    pop = Population(algo=algo, ob_dim=ob, ac_dim=ac, population=agents_per_gpu, device=local_rank, acm_kind="acm", acm_critic=True,
                     min_max_denormalize=True, update_batch_size=B, acm_batch_size=128, buffer_size=ring, store_actions=False,
                     gamma=0.99, tau=0.005, actor_lr=1e-3, critic_lr=1e-3, custom_loss=0.2)
    pop.set_limits(np.ones(ob, np.float32), np.ones(ac, np.float32))
    pop.set_norm_stats(-np.ones(ob, np.float32), np.ones(ob, np.float32))
    for a in range(agents_per_gpu):
        s0 = init_state(algo, ob, ac, 1000 + rank * agents_per_gpu + a, "acm", True)
        for net in (("actor", "critic_1", "critic_2", "acm") if algo == "sac" else ("actor", "critic", "acm")):
            pop.load_state_dict(net, {k[len(net) + 1:]: v for k, v in s0.items() if k.startswith(net + ".")}, agent=a)
    pop.sync_targets()
    kw = dict(envs_per_agent=envs_per_agent, update_freq=50, grad_steps=50, random_frames=100, act_noise=0.1, steps_per_epoch=1000,
              acm_update_freq=200, acm_update_batches=10, update_stats=True, seed=rank)
    st = pop.train_synthetic(warm_frames, **kw)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    st = pop.train_synthetic(frames, state=st, **kw)
    pop.sync(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    upd = sum(e[1] for e in st["launch_log"] if e[0] == "update") * agents_per_gpu * world
    out = {"workload": name, "scaling": "weak", "n_gpus": world, "agents_per_gpu": agents_per_gpu, "frames_per_agent": frames,
           "frames_per_s": agents_per_gpu * world * frames / dt, "updates_per_s": upd / dt, "seconds": dt,
           "launches": {k: sum(1 for e in st["launch_log"] if e[0] == k) for k in ("rollout", "update", "acm", "stats")},
           "ring_capacity": ring, "envs_per_agent": envs_per_agent}
    pop.close()
    return out


def main():
    global G
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents-per-gpu", type=int, default=0, help="population per GPU (default: one agent per SM)")
    ap.add_argument("--ring-capacity", type=int, default=RING, help="replay ring capacity per agent (profiling runs shrink it)")
    ap.add_argument("--grad-steps", type=int, default=G)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-variant", action="store_true", help="skip the reduced-precision (single tf32 pass) extra measurement")
    ap.add_argument("--no-rollout", action="store_true", help="skip the rollout (transitions/s) half of the metric")
    ap.add_argument("--no-ppo", action="store_true", help="skip the SPP-PPO config-4 whole-iteration block")
    ap.add_argument("--no-train-loop", action="store_true", help="skip the population train-loop blocks (configs 3 and 5)")
    ap.add_argument("--rollout-envs", type=int, default=256, help="vectorised synthetic environments per agent")
    ap.add_argument("--rollout-steps", type=int, default=32, help="environment steps per rollout launch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = args.gpus
    K, W = args.steps, max(args.warmup, 0)
    G = args.grad_steps
    RINGC = args.ring_capacity

    # ------------------------------------------------------------------ reference arm: the reference's CPU implementation of the path
    if args.impl == "reference":
        if rank != 0:
            return
        cores = host_cores()
        kind = cpu_kind()
        n = 20                      # updates per core and step: a bounded sample of the 7 400-update step of our arm
        arm = CpuArm(cores, 1, kind)
        try:
            for _ in range(W):
                arm.step(n)
            t0 = time.perf_counter()
            for _ in range(K):
                arm.step(n)
            total = time.perf_counter() - t0
        finally:
            arm.close()
        ups = cores * n * K / total
        sample = "%d single-thread processes x %d updates per step (%s), %d warm-up + %d timed steps" % (
            cores, n, "rltoolkit SAC_AcM.update from oracle/_ref, unmodified" if kind == "reference" else "oracle port of SAC_AcM.update", W, K)
        line = {
            "impl": "reference", "metric": METRIC, "value": ups, "unit": "updates/s", "n_gpus": N, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SPP-SAC Hopper shapes (ob 11, ac 3, hidden 256, B 256): SAC_AcM.update on the host cores, one "
                                   "single-thread run per core (the reference's mp.Pool model); one step = %d updates" % (cores * n),
                       "updates_per_step": cores * n, "batch": B},
            "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ CPU baseline first (fork before CUDA init)
    cpu_base = None
    ref_ppo = ref_roll = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_block(host_cores())
        if not args.no_ppo:
            ref_ppo = reference_ppo_block()
        if not args.no_rollout:
            ref_roll = reference_rollout_block()

    import numpy as np
    import torch

    import __graft_entry__
    if not os.path.exists(__graft_entry__.LIB):
        __graft_entry__.build()
    from spp_rl_b200 import kernel_launches, load_library

    lib = load_library()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    P = args.agents_per_gpu or sm_count
    pop = build_population(local_rank, P, args.ring_capacity)
    stream = torch.cuda.Stream(device=local_rank)
    sptr = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident value
    for w in range(W):
        pop.update_ring_device(G, seed=100 + w, stream=sptr)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = kernel_launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        t_all0.record(stream)
        for k in range(K):
            evs[k][0].record(stream)
            pop.update_ring_device(G, seed=1000 + k, stream=sptr)
            evs[k][1].record(stream)
        t_all1.record(stream)
    barrier()
    launches = kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    total_ms = t_all0.elapsed_time(t_all1)
    kern_ms = [a.elapsed_time(b) for a, b in evs]
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    updates_per_step = P * G * world
    value = updates_per_step * K / (total_ms_max * 1e-3)

    # ---- reduced-precision variant (north_star: "bf16 variants within a stated 1e-2"): the same bursts with ONE tf32 pass per product
    #      (spp_set_gemm_path(2)).  An extra, labelled number: `value`, `e2e` and `roofline` above / below are the fp32-accurate path.
    variant = None
    if not args.no_variant:
        from spp_rl_b200 import _lib as _splib
        lib = _splib.load_library()
        lib.spp_set_gemm_path(2)
        try:
            pop.update_ring_device(G, seed=7000, stream=sptr)
            v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                v0.record(stream)
                for k in range(K):
                    pop.update_ring_device(G, seed=7001 + k, stream=sptr)
                v1.record(stream)
            barrier()
            vt = torch.tensor([v0.elapsed_time(v1)], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(vt, op=dist.ReduceOp.MAX)
            variant = {"name": "single tf32 pass, whole K accumulated in TMEM (spp_set_gemm_path(2))", "dtype": "tf32",
                       "value": updates_per_step * K / (float(vt.item()) * 1e-3), "unit": "updates/s",
                       "stated_tolerance": "1e-2 relative on losses / post-step weights, 5e-2 on Adam moments (tests/test_gpu_update_parity.py)"}
        finally:
            lib.spp_set_gemm_path(1)

    # ---- e2e through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        rng = np.random.RandomState(5 + rank)
        mn, mx = np.array(MIN_OBS, np.float32), np.array(MAX_OBS, np.float32)

        def pinned(a):
            return torch.from_numpy(a).pin_memory()
        h_obs = pinned((rng.rand(P, G, B, OB) * (mx - mn) + mn).astype(np.float32))
        h_nobs = pinned((h_obs.numpy() + 0.02 * rng.randn(P, G, B, OB)).astype(np.float32))
        h_rew = pinned(rng.randn(P, G, B).astype(np.float32))
        h_done = pinned((rng.rand(P, G, B) < 1e-3).astype(np.int8))
        h_aacm = pinned(np.tanh(rng.randn(P, G, B, AC)).astype(np.float32))
        h_loss = torch.empty((P, G, 8), dtype=torch.float32).pin_memory()
        h2d = sum(x.numel() * x.element_size() for x in (h_obs, h_nobs, h_rew, h_done, h_aacm))
        d2h = h_loss.numel() * 4
        for w in range(max(W, 1)):
            pop.update_host(G, h_obs, h_nobs, None, h_rew, h_done, h_aacm, eps=None, seed=w, losses=h_loss)
        barrier()
        t0 = time.perf_counter()
        for k in range(K):
            pop.update_host(G, h_obs, h_nobs, None, h_rew, h_done, h_aacm, eps=None, seed=50 + k, losses=h_loss)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": updates_per_step * K / float(t.item()), "unit": "updates/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "loss_check": float(h_loss[0, -1, 2])}

    # ---- second half of the metric: ACM rollout transitions/s (A11: normalise -> actor -> noise/clip -> denormalise -> ACM ->
    #      ring write), E vectorised synthetic environments per agent, everything device-resident; and the same through the
    #      host-facing noise_action/process_action call with pinned host observations.
    rollout = None
    if not args.no_rollout:
        E, RS = args.rollout_envs, args.rollout_steps
        for w in range(max(W, 1)):
            pop.rollout_synthetic(E, RS, seed=300 + w, act_noise=0.1, stream=sptr)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lr0 = kernel_launches()
        with torch.cuda.stream(stream):
            r0.record(stream)
            for k in range(K):
                pop.rollout_synthetic(E, RS, seed=400 + k, act_noise=0.1, stream=sptr)
            r1.record(stream)
        barrier()
        r_launches = kernel_launches() - lr0
        t = torch.tensor([r0.elapsed_time(r1)], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r_ms = float(t.item())
        trans = P * E * RS * K * world
        f_tr = 2 * (OB * 256 + 256 * 256 + 256 * 2 * OB) + 2 * (2 * OB * 64 + 64 * 32 + 32 * AC)      # F_a + F_m = 155 072
        bytes_tr = (OB + OB + AC + 1) * 4 + 2 + OB * 4                                                # 106 B written + 44 B read
        rollout = {"value": trans / (r_ms * 1e-3), "unit": "transitions/s", "envs_per_agent": E, "steps_per_launch": RS,
                   "launches": int(r_launches), "ms_per_launch": r_ms / K,
                   "roofline": {"bound": "tensor", "flops_per_transition": f_tr, "achieved": f_tr * trans / (r_ms * 1e-3) / 1e12,
                                "unit": "TFLOP/s", "hbm_bytes_per_transition": bytes_tr,
                                "hbm_gbs": bytes_tr * trans / (r_ms * 1e-3) / 1e9}}
        if not args.no_e2e:
            rng = np.random.RandomState(9 + rank)
            mn, mx = np.array(MIN_OBS, np.float32), np.array(MAX_OBS, np.float32)
            pin = lambda a: torch.from_numpy(a).pin_memory().numpy()      # noqa: E731  (views of pinned host buffers)
            h_o = pin((rng.rand(P, E, OB) * (mx - mn) + mn).astype(np.float32))
            h_n = pin(rng.randn(P, E, OB).astype(np.float32))
            h_e = pin(rng.randn(P, E, OB).astype(np.float32))
            h_out = (pin(np.zeros((P, E, OB), np.float32)), pin(np.zeros((P, E, AC), np.float32)))
            pop.rollout_step(h_o, h_n, h_e, out=h_out)
            barrier()
            t0 = time.perf_counter()
            for k in range(K * 4):
                tgt, act = pop.rollout_step(h_o, h_n, h_e, out=h_out)
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rollout["e2e"] = {"value": P * E * K * 4 * world / float(t.item()), "unit": "transitions/s",
                              "h2d_bytes_per_step": int(3 * h_o.nbytes), "d2h_bytes_per_step": int(tgt.nbytes + act.nbytes),
                              "note": "spp_rollout_step_host: one vectorised noise_action + process_action call per step"}

    # ---- the minibatch gather alone (A1), HBM side of sample_batch: population-wide, device-drawn indices, dense output
    gather = None
    if not args.no_rollout:
        NB = 20
        for _ in range(2):
            gbytes = pop.ring_gather_bench(NB, seed=1, stream=sptr)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            g0.record(stream)
            for k in range(K):
                gbytes = pop.ring_gather_bench(NB, seed=2 + k, stream=sptr)
            g1.record(stream)
        barrier()
        g_ms = g0.elapsed_time(g1) / K
        try:
            _hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6500.0)
        except Exception:
            _hbm_peak = 6500.0
        _g_traffic = 675.0e6 * (P * NB * B) / 757760.0      # PROFILE CONSTANT: ncu dram__bytes_read + write of this kernel per 757 760 rows
        gather = {"rows_per_launch": P * NB * B, "ms_per_launch": g_ms, "algorithmic_bytes_per_launch": gbytes,
                  "algorithmic_gbs": gbytes / (g_ms * 1e-3) / 1e9, "rows_per_s": P * NB * B / (g_ms * 1e-3),
                  "roofline": {"bound": "hbm", "achieved": gbytes / (g_ms * 1e-3) / 1e9, "peak": _hbm_peak, "unit": "GB/s",
                               "frac": gbytes / (g_ms * 1e-3) / 1e9 / _hbm_peak, "traffic": _g_traffic,
                               "traffic_gbs": _g_traffic / (g_ms * 1e-3) / 1e9,
                               "note": "achieved = ALGORITHMIC bytes / time; traffic is a profile constant (profiles/r02_gather_summary.md: "
                                       "4.1x the algorithmic bytes, a random transition touches seven arrays of the ring)"},
                  "note": "ring_gather_bench_kernel: B*((2 ob + ac + 1)*4 + 1) bytes read and written per row + two 4-byte index "
                          "words; random transitions of a %.1f GB ring" % (P * RINGC * 114 / 1e9)}

    # ---- BASELINE configs 3 / 5 through the population train loop, config 4 as whole SPP-PPO iterations (all ranks take part)
    pop.close()
    pop = None
    train_loop = None
    if not args.no_train_loop:
        train_loop = [population_train_block(local_rank, rank, world, dist, "config 3: SPP-DDPG HalfCheetah shapes (ob 17, ac 6), 256 independent agents per GPU",
                                             "ddpg", 17, 6, 256, 1000, 500),
                      population_train_block(local_rank, rank, world, dist, "config 5: SPP-SAC Ant shapes (ob 111, ac 8), 128 agents per GPU (1024 over 8 GPUs)",
                                             "sac", 111, 8, 128, 500, 500)]
    ppo_cfg4 = None
    if not args.no_ppo:
        ppo_cfg4 = ppo_cfg4_block(local_rank, rank, world, dist, max(1, min(K, 3)), 1)
        if ref_ppo is not None and isinstance(ppo_cfg4, dict):
            ppo_cfg4["cpu_reference"] = ref_ppo      # the reference's own PPO_AcM phases on one host core, per transition (SURVEY 8d iv)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16 sustained)" if peaks else "fallback (B200_PROFILING.md sustained)"
        f_upd = flops_per_update()
        kt = statistics.mean(kern_ms) * 1e-3
        achieved = f_upd * P * G / kt / 1e12
        traffic = None
        try:
            per_update = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("update_burst_kernel_bytes_per_update")
            traffic = int(per_update * P * G)      # ncu dram bytes per update (profiles/) x the updates of one launch here
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6500.0)
        hbm = None if traffic is None else {"measured_traffic_gbs": traffic / kt / 1e9, "peak_gbs": hbm_peak, "frac": traffic / kt / 1e9 / hbm_peak,
                                            "note": "traffic is a PROFILE CONSTANT, not measured in this run: ncu dram__bytes_read + write per update "
                                                    "(profiles/traffic.json, one --set full capture of this kernel) x the updates of one launch here"}
        line = {
            "metric": METRIC, "value": value, "unit": "updates/s", "n_gpus": world if world > 1 else N, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SPP-SAC Hopper shapes (ob 11, ac 3, hidden 256, B 256) fused update bursts, population of "
                                   "%d independent agents per GPU (one per SM), each with its own 1M-transition device replay "
                                   "ring; one step = %d updates per agent" % (P, G),
                       "agents_per_gpu": P, "grad_steps": G, "batch": B, "ring_capacity": RINGC, "ring_fill": RINGC * 999 // 1000,
                       "updates_per_step": updates_per_step, "l2": "working set (rings %.1f GB + agent state) >> 126 MB L2, no flush needed"
                       % (P * RINGC * 114 / 1e9), "parallelism": "independent agents sharded over GPUs, no collective"},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "hbm": hbm, "peak_source": peak_src, "flops_per_update": f_upd,
                         "kernel": "update_burst_kernel<SAC>", "kernel_ms": statistics.mean(kern_ms),
                         "note": "256-wide GEMMs on tcgen05 kind::tf32, 3-pass hi/lo split with per-chunk fp32 drain (1e-5 parity); "
                                 "achieved counts algorithmic fp32 FLOPs (each is 3 tensor-core passes); peak is the dense bf16 figure"},
            "rollout": rollout, "gather": gather, "train_loop": train_loop, "ppo_cfg4": ppo_cfg4, "reduced_precision_variant": variant, "cpu_baseline": cpu_base, "clocks": clocks,
        }
        if rollout is not None and ref_roll is not None:
            rollout["cpu_reference"] = ref_roll      # the reference's frame-loop body on one host core (SURVEY 8d iii)
        if rollout is not None:
            rollout["roofline"]["peak"] = peak
            rollout["roofline"]["frac"] = rollout["roofline"]["achieved"] / peak
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if pop is not None:
        pop.close()


if __name__ == "__main__":
    main()
