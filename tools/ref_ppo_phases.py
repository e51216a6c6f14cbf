"""the REFERENCE's own PPO_AcM, phase by phase, on the host -- the CPU number that sits beside bench.py's config-4
block.  The reference's Python-list Memory cannot hold 8.4 M steps, so one iteration is timed at a reduced batch (default 16 384
transitions, Walker2d shapes through the gym shim's shape-only environment) and reported per transition; the phases are the calls of
AcMOnPolicyTrainer.perform_iteration (rltoolkit/acm/on_policy.py:55-86) in its order, each the reference's own method:
    collect_batch -> update_critic -> update_actor (update_actor_acm) -> replay_buffer.add_buffer -> update_acm_batches.
Single thread (the reference's evals.py:26 model).  ORACLE-SIDE infrastructure: imports oracle/_ref (or /root/reference).
    python tools/ref_ppo_phases.py [transitions] -> one JSON line"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n=16384, acm_batches=100, seed=0):
    import torch

    torch.set_num_threads(1)
    from oracle.ref_import import import_reference
    rl = import_reference(scratch_dir=tempfile.mkdtemp(prefix="spp_ref_ppo_"))
    from rltoolkit.buffer import MemoryAcM
    torch.manual_seed(seed)
    # kwargs of train/spp_ppo_hopper.py:72-103 (gamma 0.99, lr 3e-4, KL 0.1, <= 10 epochs, min-max denormalisation), config 4's custom loss
    # and ACM batches; a short ACM pre-training (it is not part of the iteration that is timed)
    m = rl.PPO_AcM(env_name="Walker2d-v2", iterations=1, gamma=0.99, actor_lr=3e-4, critic_lr=3e-4, batch_size=n, ppo_batch_size=min(n, 65536),
                   kl_div_threshold=0.1, max_ppo_epochs=2, entropy_coef=0.0, obs_norm_alpha=None, denormalize_actor_out=True,
                   min_max_denormalize=True, custom_loss=0.5, norm_closs=True, acm_update_freq=1, acm_update_batches=acm_batches, acm_batch_size=256,
                   acm_lr=1e-4, acm_pre_train_samples=2000, acm_pre_train_epochs=1, tensorboard_dir=None, log_dir=None, verbose=0)
    m.pre_train()
    t = [time.perf_counter()]
    m.buffer = MemoryAcM(obs_mean=m.obs_mean, obs_std=m.obs_std, device=m.device, alpha=m.obs_norm_alpha, max_obs=m.max_obs,
                         min_obs=m.min_obs, min_max_denormalize=m.min_max_denormalize)
    m.collect_batch(m.buffer); t.append(time.perf_counter())
    adv = m.update_critic(m.buffer); t.append(time.perf_counter())
    m.update_actor(adv, m.buffer); t.append(time.perf_counter())
    m.replay_buffer.add_buffer(m.buffer); t.append(time.perf_counter())
    m.update_acm_batches(m.acm_update_batches); t.append(time.perf_counter())
    rows = len(m.buffer)
    names = ["rollout", "critic_fit", "actor_epochs", "add_buffer", "acm_update"]
    ph = {k: (t[i + 1] - t[i]) * 1e3 for i, k in enumerate(names)}
    tot = t[-1] - t[0]
    return {"impl": "reference PPO_AcM (rltoolkit, unmodified), 1 thread", "transitions": rows, "ms_per_iteration": tot * 1e3, "phases_ms": ph,
            "us_per_transition": tot * 1e6 / rows, "transitions_per_s": rows / tot,
            "note": "one perform_iteration at a reduced batch (the Python-list Memory cannot hold 8.4 M steps); critic fit 10 x 10 full-batch steps, "
                    "actor epochs until KL 0.1 or 2, %d ACM batches of 256" % acm_batches}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 16384)))
