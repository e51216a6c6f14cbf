"""-m gpu, needs 2 GPUs: SPP-PPO data-parallel iteration (collectives inside the library: the gradient all-reduce as the fused kernel over
NVLink peer memory, and as NCCL; tools/ppo_dp.py) against the same iteration on
one GPU AND against the CPU oracle on the same data: all ranks end bit-identical, and the post-iteration weights of both the
data-parallel and the single-GPU run lie within 1e-5 (norm-relative, stated) of oracle/ppo.py's."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_iteration(E, T, batch, epochs, critic_targets, critic_steps):
    sys.path.insert(0, ROOT)
    from oracle import ppo as P
    from oracle.norm import NormStats, denormalize, normalize
    from tools.ppo_dp import OB, synth

    d, w = synth(E, T)
    N = E * T
    s = {k: torch.from_numpy(v.copy()) for k, v in w.items()}
    st = NormStats(True, torch.from_numpy(d["mn"]), torch.from_numpy(d["mx"]))
    x, xn = normalize(st, torch.from_numpy(d["obs"]), True), normalize(st, torch.from_numpy(d["nobs"]), True)
    tr, td, te = torch.from_numpy(d["rew"]), torch.from_numpy(d["done"]), torch.from_numpy(d["end"])
    closs = P.update_critic(s, x, xn, tr, td, 0.99, 3e-4, critic_targets, critic_steps)
    from oracle import nets
    from oracle.offpolicy import sub
    v = nets.ppo_critic_fwd(sub(s, "critic"), x)[0].squeeze(-1)
    nv = nets.ppo_critic_fwd(sub(s, "critic"), xn)[0].squeeze(-1)
    q = P.q_values(tr, td, nv, 0.99)
    adv = torch.empty(N)
    for e in range(E):      # step-major rows: environment e owns rows e, e + E, ...
        rows = torch.arange(e, N, E)
        adv[rows] = P.gae(q[rows], v[rows], nv[rows], td[rows], te[rows], 0.99, 0.95)
    advn = P.normalize_adv(adv)
    rng = np.random.RandomState(11)
    perms = [torch.from_numpy(rng.permutation(N)) for _ in range(epochs)]
    P.update_actor_acm(s, x, denormalize(st, torch.from_numpy(d["act"])), denormalize(st, xn), torch.from_numpy(d["logp"]), advn,
                       perms, 1.0, 3e-4, 0.2, 1e9, epochs, batch, 0.0, 0.1)
    return closs, {k: v.numpy() for k, v in s.items() if k in w}


def _relnorm(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("p2p", [True, False], ids=["allreduce-nvlink-peer-memory", "allreduce-nccl"])
def test_ppo_data_parallel_matches_single_gpu_and_oracle(tmp_path, p2p):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dump = str(tmp_path / "dp.npz")
    E, T, batch = 64, 128, 2048
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "ppo_dp.py"), "--envs", str(E), "--steps", str(T), "--batch", str(batch),
           "--dump", dump]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, SPP_PPO_P2P="1" if p2p else "0"))
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    # the gradient all-reduces (critic steps + actor minibatches) took the path under test; the fp64 statistics always go through NCCL
    assert out["p2p"] == p2p and (out["p2p_steps"] >= out["allreduces"] - 1 if p2p else out["p2p_steps"] == 0)
    assert out["ranks_bit_identical"] and out["dp_vs_single_worst_relnorm"] < 1e-5
    assert out["allreduces"] == 2 * 5 + 1 + 2 * ((E * T + batch - 1) // batch)      # critic steps + advantage statistics + actor minibatches
    closs, ref = _oracle_iteration(E, T, batch, 2, 2, 5)
    got = np.load(dump)
    assert float(got["critic_loss_dp"]) == pytest.approx(closs, rel=1e-5)
    assert float(got["critic_loss_one"]) == pytest.approx(closs, rel=1e-5)
    for k, v in ref.items():
        for arm in ("dp", "one"):
            e = _relnorm(got["%s:%s" % (arm, k)], v)
            assert e < 1e-5, (arm, k, e)


def test_ppo_single_gpu_iteration_path_matches_oracle():
    """The same iteration code path (iteration_dp with world = 1: epoch-at-once actor loop, device-recorded scalars) on ONE GPU vs
    the oracle; runs on the driver's 1-GPU box where the two-rank test above is skipped."""
    sys.path.insert(0, ROOT)
    from tools.ppo_dp import make_policy, synth

    E, T, batch = 48, 96, 1000
    d, w = synth(E, T)
    N = E * T
    pol = make_policy(0, N, batch, d, w)
    pol.load_rollout(d["obs"], d["nobs"], d["act"], d["logp"], d["rew"], d["done"], d["end"], np.arange(E), np.full(E, T), traj_stride=E)
    rng = np.random.RandomState(11)
    perms = np.stack([rng.permutation(N) for _ in range(2)]).astype(np.int64)
    res = pol.iteration_dp(perms, batch, E, 1e9, 2, 5, 0, 1)
    closs, ref = _oracle_iteration(E, T, batch, 2, 2, 5)
    assert res["critic_loss"] == pytest.approx(closs, rel=1e-5)
    assert res["epochs"] == 2
    for net in ("actor", "critic"):
        for k, v in pol.state_dict(net).items():
            e = _relnorm(v, ref["%s.%s" % (net, k)])
            assert e < 1e-5, (net, k, e)
    pol.close()
