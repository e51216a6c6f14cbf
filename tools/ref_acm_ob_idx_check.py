"""Does the REFERENCE itself run with an `acm_ob_idx` subset (rltoolkit/acm/acm.py:94-96,109,148)?  Build container only (imports the
unmodified reference through oracle/ref_import.py).  Result here (torch 2.11 CPU): no class gets through pre_train() --
    SAC_AcM / DDPG_AcM: ValueError in collect_samples -> replay_buffer.add_timestep (the ring's action rows are [len(idx)] wide, the
                        stored action is the [ob] observation-sized target);
    PPO_AcM:            RuntimeError in the ACM regression: features cat[obs[:, idx], next_obs[:, idx]] are 2 len(idx) wide, the ACM's
                        first layer expects ob + len(idx) (acm.py:148 vs :262).
So the subset branch is dead code in the reference; spp_rl_b200 refuses it in the constructors instead of guessing a meaning."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_import import import_reference

import_reference()
from rltoolkit import DDPG_AcM, PPO_AcM, SAC_AcM  # noqa: E402

for cls, kw in ((SAC_AcM, dict(random_frames=10, buffer_size=1000, update_batch_size=16)),
                (DDPG_AcM, dict(random_frames=10, buffer_size=1000, update_batch_size=16)),
                (PPO_AcM, dict(custom_loss=0.1, denormalize_actor_out=True, min_max_denormalize=True))):
    try:
        m = cls(env_name="Pendulum-v0", acm_ob_idx=[0, 1], acm_pre_train_samples=200, acm_pre_train_epochs=1, iterations=1, batch_size=50,
                tensorboard_dir=None, **kw)
        m.pre_train()
        m.train()
        print(cls.__name__, "runs")
    except Exception as e:      # noqa: BLE001
        print(cls.__name__, "fails:", type(e).__name__, str(e)[:120])
