"""-m gpu: the CUDA path (through the C ABI) against the golden fixtures produced by the unmodified reference,
the replay ring against the reference's index stream bit for bit, and size-independent properties at full size."""
import os

import numpy as np
import pytest
import torch

from spp_rl_b200 import Population, init_state
from tests.parity_util import make_batches, make_stats, relnorm, upload_state

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cmp_gold(pop, gold, nets, tol=1e-5):
    for net in nets:
        sd = pop.state_dict(net, agent=0)
        for k, v in sd.items():
            e = relnorm(v, gold[net + "." + k])
            assert e < tol, (net, k, e)
        if not net.endswith("_targ"):
            ad, step = pop.adam_state(net, agent=0)
            assert step == int(gold[net + "#step"])
            for k, (m, v) in ad.items():
                lim = tol
                assert relnorm(m, gold[net + "." + k + "#m"]) < lim, (net, k)
                assert relnorm(v, gold[net + "." + k + "#v"]) < lim, (net, k)


def test_sac_update_matches_reference_fixture():
    gold = np.load(os.path.join(G, "sac_hopper_g3.npz"))
    ob, ac, B, Gs, seed = [int(x) for x in gold["meta"]]
    mn, mx, mean, std = make_stats(ob, seed, True)
    obs, nobs, act, rew, done, aacm, eps = make_batches(ob, ac, 1, Gs, B, seed, mn, mx)
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=1, update_batch_size=B, gamma=0.99, custom_loss=0.2,
                     acm_critic=True, norm_closs=False, min_max_denormalize=True, alpha=0.2, target_entropy=-float(ac))
    pop.set_norm_stats(mn, mx, mean, std)
    upload_state(pop, init_state("sac", ob, ac, seed * 100), 0, "sac")
    losses = pop.update_host(Gs, obs, nobs, act, rew, done, aacm, eps=eps)
    for g in range(Gs):
        ref = gold["losses"][g]       # critic_1, critic_2, actor, sac, dist, alpha
        mine = [losses[0, g, 0], losses[0, g, 1], losses[0, g, 2], losses[0, g, 3], losses[0, g, 4], losses[0, g, 6]]
        for v, r in zip(mine, ref):
            assert v == pytest.approx(r, rel=1e-5, abs=1e-6)
    assert pop.alpha(0)[0] == pytest.approx(float(gold["log_alpha"]), rel=1e-6)
    _cmp_gold(pop, gold, ["actor", "critic_1", "critic_2", "critic_1_targ", "critic_2_targ"])
    pop.close()


def test_ddpg_update_matches_reference_fixture():
    gold = np.load(os.path.join(G, "ddpg_hcheetah_g2.npz"))
    ob, ac, B, Gs, seed = [int(x) for x in gold["meta"]]
    mn, mx, mean, std = make_stats(ob, seed, True)
    obs, nobs, act, rew, done, aacm, _ = make_batches(ob, ac, 1, Gs, B, seed, mn, mx)
    pop = Population(algo="ddpg", ob_dim=ob, ac_dim=ac, population=1, update_batch_size=B, gamma=0.95, custom_loss=1.0,
                     acm_kind="basic", acm_critic=True, norm_closs=False, min_max_denormalize=True, actor_lr=5e-4, critic_lr=5e-4)
    pop.set_norm_stats(mn, mx, mean, std)
    s0 = init_state("ddpg", ob, ac, seed * 100, "basic", True)
    s0["acm.t"][:] = 0.7
    s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
    upload_state(pop, s0, 0, "ddpg")
    losses = pop.update_host(Gs, obs, nobs, act, rew, done, aacm)
    for g in range(Gs):
        ref = gold["losses"][g]       # critic, actor, ddpg, dist
        for v, r in zip([losses[0, g, 0], losses[0, g, 2], losses[0, g, 3], losses[0, g, 4]], ref):
            assert v == pytest.approx(r, rel=1e-5, abs=1e-6)
    _cmp_gold(pop, gold, ["actor", "critic", "actor_targ", "critic_targ"])
    pop.close()


def test_ring_matches_reference_index_stream_bit_exact():
    g = np.load(os.path.join(G, "ring_ops.npz"))
    size, ob, ac = int(g["size"]), int(g["ob"]), int(g["ac"])
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=2, update_batch_size=16, buffer_size=size, acm_critic=False)
    for agent in (0, 1):          # two agents: rings are independent
        for kind, f, iv, stt in zip(g["kinds"], g["fvals"], g["ivals"], g["states"]):
            if kind == 0:
                assert pop.ring_add_obs(agent, f[:ob]) == iv[0]
            else:
                pop.ring_add_acm_action(agent, f[2 * ob:2 * ob + ac])
                assert pop.ring_add_obs(agent, f[:ob]) == iv[1]
                pop.ring_add_timestep(agent, iv[0], iv[1], f[ob:2 * ob], f[-1], bool(iv[2]), bool(iv[3]))
            assert pop.ring_state(agent) == tuple(int(x) for x in stt)      # incl. the shrinking current_len quirk
        out = pop.ring_sample_batch(agent, g["idx"])
        for mine, ref in zip(out, (g["s_obs"], g["s_next"], g["s_act"], g["s_rew"], g["s_done"], g["s_aacm"])):
            assert np.array_equal(mine, ref)
        assert out[4].dtype == np.int8
    with pytest.raises(Exception):
        pop.ring_sample_batch(0, np.array([pop.ring_state(0)[2]]))          # index == len(buffer) is out of range
    pop.close()


def test_ring_update_equals_host_batch_update():
    """sample_batch + update through the device ring == update(...) on the same rows handed over explicitly."""
    ob, ac, B, Gs, P = 11, 3, 64, 2, 2
    mn, mx, mean, std = make_stats(ob, 3, True)
    rng = np.random.RandomState(9)

    def fresh():
        pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, update_batch_size=B, buffer_size=400, gamma=0.99,
                         custom_loss=0.2, acm_critic=True, alpha=0.2, target_entropy=-float(ac))
        pop.set_norm_stats(mn, mx, mean, std)
        for a in range(P):
            upload_state(pop, init_state("sac", ob, ac, 40 + a), a, "sac")
        return pop
    pa, pb = fresh(), fresh()
    rows = {}
    for a in range(P):
        prev = None
        for t in range(300):
            o = rng.randn(ob).astype(np.float32)
            if prev is None:
                prev = pa.ring_add_obs(a, o); rows[(a, "first")] = o
                continue
            aa = np.tanh(rng.randn(ac)).astype(np.float32)
            pa.ring_add_acm_action(a, aa)
            nxt = pa.ring_add_obs(a, o)
            pa.ring_add_timestep(a, prev, nxt, rng.randn(ob).astype(np.float32), float(rng.randn()), rng.rand() < 0.05, False)
            prev = nxt
    idx = np.stack([rng.randint(0, pa.ring_state(a)[2], size=(Gs, B)) for a in range(P)]).astype(np.int64)
    eps = rng.randn(P, Gs, 2, B, ob).astype(np.float32)
    l_ring = pa.update_ring(Gs, idx=idx, eps=eps)
    batches = [pa.ring_sample_batch(a, idx[a].reshape(-1)) for a in range(P)]
    obs = np.stack([b[0].reshape(Gs, B, ob) for b in batches]); nobs = np.stack([b[1].reshape(Gs, B, ob) for b in batches])
    act = np.stack([b[2].reshape(Gs, B, ob) for b in batches]); rew = np.stack([b[3].reshape(Gs, B) for b in batches])
    done = np.stack([b[4].reshape(Gs, B) for b in batches]); aacm = np.stack([b[5].reshape(Gs, B, ac) for b in batches])
    l_host = pb.update_host(Gs, obs, nobs, act, rew, done, aacm, eps=eps)
    assert np.array_equal(l_ring, l_host)                                    # bitwise: same kernel, same rows
    for a in range(P):
        for net in ("actor", "critic_1", "critic_2_targ"):
            sa, sb = pa.state_dict(net, a), pb.state_dict(net, a)
            for k in sa:
                assert np.array_equal(sa[k], sb[k]), (a, net, k)
    pa.close(); pb.close()


def test_full_size_properties():
    """Hopper shapes at BASELINE size (B=256, bursts of 50, 1M-capacity rings, device sampler and noise):
    determinism, independence from the population size, finite losses, untouched frozen ACM, pads stay zero."""
    ob, ac, B, Gs = 11, 3, 256, 50

    def run(P):
        pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, update_batch_size=B, buffer_size=1_000_000,
                         store_actions=False, gamma=0.99, custom_loss=0.2, acm_critic=True, alpha=0.2, target_entropy=-3.0)
        pop.set_norm_stats(-np.ones(ob, np.float32), np.ones(ob, np.float32))
        for a in range(P):
            upload_state(pop, init_state("sac", ob, ac, 7 + a), a, "sac")
        pop.ring_fill_synthetic(seed=1, n=999_000, episode_len=1000)
        losses = torch.zeros((P, Gs, 8), dtype=torch.float32, device="cuda")
        pop.update_ring_device(Gs, seed=5, losses_dev_ptr=losses.data_ptr())
        pop.sync()
        sd = {net: pop.state_dict(net, 0) for net in ("actor", "critic_1", "critic_2", "critic_1_targ", "acm")}
        out = losses.cpu().numpy(), sd, pop.alpha(0)
        pop.close()
        return out
    l4, sd4, al4 = run(4)
    l4b, sd4b, al4b = run(4)
    l2, sd2, al2 = run(2)
    assert np.isfinite(l4).all()
    assert np.array_equal(l4, l4b) and al4 == al4b                           # run-to-run determinism (no atomics races)
    assert np.array_equal(l4[:2], l2) and al4 == al2                         # agent 0/1 do not depend on the population size
    for net in sd4:
        for k in sd4[net]:
            assert np.array_equal(sd4[net][k], sd4b[net][k]) and np.array_equal(sd4[net][k], sd2[net][k]), (net, k)
    ref_acm = {k[4:]: v for k, v in init_state("sac", ob, ac, 7).items() if k.startswith("acm.")}
    for k, v in ref_acm.items():
        assert np.array_equal(sd4["acm"][k], v)                               # the ACM is frozen during update()
    assert l4[0, -1, 0] != l4[0, 0, 0]                                        # and the critics did move


@pytest.mark.parametrize("kind", ["acm", "basic"])
def test_acm_regression_matches_reference_fixture(kind):
    """AcMTrainer.batch_update x3 (rltoolkit/acm/acm.py:246-258) incl. BasicAcM's learnable gains t, t1."""
    g = np.load(os.path.join(G, "acm_regress.npz"))
    ob, ac, P, n, Bm = 17, 6, 2, 3, 100
    pop = Population(algo="ddpg", ob_dim=ob, ac_dim=ac, population=P, acm_kind=kind, acm_batch_size=Bm, acm_lr=1e-3,
                     update_batch_size=64)
    s0 = init_state("ddpg", ob, ac, 7, kind, True)
    if kind == "basic":
        s0["acm.t"][:] = 0.7
        s0["acm.t1"][:] = np.linspace(0.5, 1.5, ac)
    pop.load_state_dict("acm", {k[4:]: v for k, v in s0.items() if k.startswith("acm.")}, agent=-1)
    rng = np.random.RandomState(11)
    xs, ys = [], []
    for _ in range(n):
        xs.append(rng.randn(Bm, 2 * ob).astype(np.float32)); ys.append(np.tanh(rng.randn(Bm, ac)).astype(np.float32))
    x = np.ascontiguousarray(np.broadcast_to(np.stack(xs), (P, n, Bm, 2 * ob)))
    y = np.ascontiguousarray(np.broadcast_to(np.stack(ys), (P, n, Bm, ac)))
    losses = pop.acm_update_host(n, x, y)
    for a in range(P):
        for i in range(n):
            assert losses[a, i] == pytest.approx(float(g[kind + ":losses"][i]), rel=1e-5)
        sd = pop.state_dict("acm", agent=a)
        ad, step = pop.adam_state("acm", agent=a)
        assert step == n
        for k, v in sd.items():
            lim = 1e-5
            assert relnorm(v, g[kind + ":acm." + k]) < lim, (k, relnorm(v, g[kind + ":acm." + k]))
            assert relnorm(ad[k][0], g[kind + ":acm." + k + "#m"]) < lim, k
            assert relnorm(ad[k][1], g[kind + ":acm." + k + "#v"]) < lim, k
    pop.close()


def test_acm_ring_regression_equals_host_form():
    """update_acm_batches from the device ring (rbuffer_sample_acm + acm_cat) == batch_update on the gathered rows."""
    ob, ac, P, n, Bm = 11, 3, 2, 2, 64
    rng = np.random.RandomState(21)

    def fresh():
        pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, acm_batch_size=Bm, acm_lr=3e-3, buffer_size=300,
                         update_batch_size=64)
        for a in range(P):
            s0 = init_state("sac", ob, ac, 60 + a)
            pop.load_state_dict("acm", {k[4:]: v for k, v in s0.items() if k.startswith("acm.")}, agent=a)
        return pop
    pa, pb = fresh(), fresh()
    for a in range(P):
        prev = pa.ring_add_obs(a, rng.randn(ob).astype(np.float32))
        for t in range(200):
            pa.ring_add_acm_action(a, np.tanh(rng.randn(ac)).astype(np.float32))
            nxt = pa.ring_add_obs(a, rng.randn(ob).astype(np.float32))
            pa.ring_add_timestep(a, prev, nxt, rng.randn(ob).astype(np.float32), 0.0, False, False)
            prev = nxt
    idx = rng.randint(0, 200, size=(P, n, Bm)).astype(np.int64)
    l_ring = pa.acm_update_ring(n, idx=idx)
    rows = [pa.ring_sample_batch(a, idx[a].reshape(-1)) for a in range(P)]
    x = np.stack([np.concatenate([r[0], r[1]], axis=1).reshape(n, Bm, 2 * ob) for r in rows])
    y = np.stack([r[5].reshape(n, Bm, ac) for r in rows])
    l_host = pb.acm_update_host(n, np.ascontiguousarray(x), np.ascontiguousarray(y))
    assert np.array_equal(l_ring, l_host)
    for a in range(P):
        sa, sb = pa.state_dict("acm", a), pb.state_dict("acm", a)
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), (a, k)
    pa.close(); pb.close()


def _philox4x32_10(seed, stream, ctr):
    """Philox4x32-10 as csrc/common.cuh: key = seed, counter = (ctr lo, ctr hi, stream lo, stream hi) -> four 32-bit words."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    k = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
    c = [ctr & 0xFFFFFFFF, (ctr >> 32) & 0xFFFFFFFF, stream & 0xFFFFFFFF, (stream >> 32) & 0xFFFFFFFF]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_gather_bench_kernel_gathers_what_sample_batch_gathers():
    """A1: the population-wide gather kernel bench.py times (ring_gather_bench_kernel, device-drawn indices, two-phase / four lanes
    per row) against spp_ring_sample_batch on the recomputed indices: every column bit-exact, incl. a ragged last warp."""
    from spp_rl_b200 import Population

    P, B, NB, n = 3, 50, 3, 4321               # 450 rows: 14 full warps of 32 rows + a ragged one
    for ob, ac in ((11, 3), (17, 6)):          # rows of 12 floats (one float4 per lane) and of 20 (the strided form)
        pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, update_batch_size=B, buffer_size=5000, store_actions=True)
        pop.ring_fill_synthetic(seed=3, n=n, episode_len=100)
        seed = 77
        pop.ring_gather_bench(NB, seed=seed)
        obs, nobs, aacm, rew, done = pop.ring_gather_bench_rows(NB, 0, P * NB * B)
        for a in range(P):
            idx = np.array([(((x[0] << 32) | x[1]) * n) >> 64 for x in (_philox4x32_10(seed, a, t) for t in range(NB * B))], np.int64)
            assert idx.min() >= 0 and idx.max() < n and len(set(idx.tolist())) > NB * B // 2
            ro, rn, _, rr, rd, ra = pop.ring_sample_batch(a, idx)
            sl = slice(a * NB * B, (a + 1) * NB * B)
            assert np.array_equal(obs[sl], ro) and np.array_equal(nobs[sl], rn) and np.array_equal(aacm[sl], ra)
            assert np.array_equal(rew[sl], rr) and np.array_equal(done[sl], rd)
        pop.close()
