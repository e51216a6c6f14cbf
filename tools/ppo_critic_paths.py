"""SPP-PPO critic step: the tcgen05 kernel (ppo_critic_tc.cu) against the FFMA tile kernel -- gradient agreement and time per step."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spp_rl_b200.ppo import PpoPolicy

E, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 1024)
ob, ac = 17, 6
N = E * T
rng = np.random.RandomState(0)
obs = rng.randn(N, ob).astype(np.float32); nobs = (obs + 0.05 * rng.randn(N, ob)).astype(np.float32)
act = rng.randn(N, ob).astype(np.float32); logp = rng.randn(N).astype(np.float32); rew = rng.randn(N).astype(np.float32)
done = (rng.rand(N) < 0.02).astype(np.float32); end = done.copy(); end[(T - 1) * E:] = 1
res = {}
for tc in (False, True):
    pol = PpoPolicy(ob, ac, max_rows=N, max_batch_rows=1024, min_max_denormalize=True)
    pol.set_critic_path(tc)
    pol.set_norm_stats(-3 * np.ones(ob, np.float32), 3 * np.ones(ob, np.float32))
    torch.manual_seed(1)
    sd = {}
    for name, (o, i) in (("fc1", (64, ob)), ("fc2", (64, 64)), ("fc3", (1, 64))):
        b = 1 / np.sqrt(i)
        sd[name + ".weight"] = torch.empty(o, i).uniform_(-b, b).numpy(); sd[name + ".bias"] = torch.empty(o).uniform_(-b, b).numpy()
    pol.load_state_dict("critic", sd)
    pol.load_rollout(obs, nobs, act, logp, rew, done, end, np.arange(E), np.full(E, T), traj_stride=E)
    pol.update_critic(1, 2)                     # warm-up
    pol.sync()
    t0 = time.perf_counter()
    loss = pol.update_critic(2, 10)
    pol.sync()
    dt = (time.perf_counter() - t0) / 20
    res[tc] = (loss, pol.state_dict("critic"))
    print("critic path %-8s %.3f ms per step (%d rows, incl. target / reduce / Adam launches)  loss %.8g" % ("tcgen05" if tc else "FFMA", dt * 1e3, N, loss))
    pol.close()
for k in res[True][1]:
    a, b = res[True][1][k], res[False][1][k]
    print("  %-12s tcgen05 vs FFMA after 22 steps: %.2e" % (k, np.linalg.norm(a - b) / np.linalg.norm(b)))
