"""the REFERENCE's rollout inner step on the host, environment excluded -- the CPU number beside the rollout
(transitions/s) half of the bench metric.  The calls of the frame loop of DDPG.collect_batch_and_train
(rltoolkit/algorithms/ddpg/ddpg.py:205-220) for a SAC_AcM at Hopper shapes, in its order, each the reference's own method, with a
no-op environment (a fresh random observation per step):
    replay_buffer.normalize -> noise_action -> process_action (ACM + add_acm_action) -> process_obs -> add_obs -> add_timestep.
Single thread, one environment (the reference steps one).  ORACLE-SIDE infrastructure: imports oracle/_ref (or /root/reference).
    python tools/ref_rollout_step.py [transitions] -> one JSON line"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n=20000, seed=0):
    import numpy as np
    import torch

    torch.set_num_threads(1)
    from oracle.ref_import import import_reference
    rl = import_reference(scratch_dir=tempfile.mkdtemp(prefix="spp_ref_roll_"))
    torch.manual_seed(seed)
    m = rl.SAC_AcM(env_name="Hopper-v2", update_batch_size=256, custom_loss=0.2, acm_critic=True, norm_closs=False, denormalize_actor_out=True,
                   min_max_denormalize=True, acm_pre_train_samples=100, acm_val_buffer_size=None, buffer_size=max(2 * n, 1000), tensorboard_dir=None,
                   log_dir=None, verbose=0)
    ob = m.ob_dim
    m.replay_buffer.min_obs, m.replay_buffer.max_obs = -torch.ones(ob), torch.ones(ob)
    rng = np.random.RandomState(seed)
    raw = rng.uniform(-1, 1, (n + 1, ob)).astype(np.float32)
    obs = m.process_obs(raw[0])
    prev = m.replay_buffer.add_obs(obs)

    def steps(k0, k1, prev, obs):
        for k in range(k0, k1):
            o = m.replay_buffer.normalize(obs)
            action = m.noise_action(o, m.act_noise)
            m.process_action(action, o)
            obs = m.process_obs(raw[k + 1])
            nxt = m.replay_buffer.add_obs(obs)
            m.replay_buffer.add_timestep(prev, nxt, action, 0.0, False, (k + 1) % 1000 == 0)
            prev = nxt
        return prev, obs

    warm = min(200, n // 10)
    prev, obs = steps(0, warm, prev, obs)
    t0 = time.perf_counter()
    prev, obs = steps(warm, n, prev, obs)
    dt = time.perf_counter() - t0
    k = n - warm
    return {"impl": "reference SAC_AcM frame-loop body (rltoolkit, unmodified), 1 thread, 1 environment, no-op env", "transitions": k,
            "us_per_transition": dt * 1e6 / k, "transitions_per_s": k / dt}


if __name__ == "__main__":
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 20000)))
