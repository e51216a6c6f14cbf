"""Where one actor epoch's time goes on ONE GPU at config-4 shapes: the permutation's H2D copy, the epoch-at-once device loop
(spp_ppo_actor_epoch_device), and the same epochs through spp_ppo_update_actor.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from tools.ppo_dp import make_policy, synth


def main():
    E, T, batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 2048, 65536
    N = E * T
    d, w = synth(E, T)
    pol = make_policy(0, N, batch, d, w)
    pol.load_rollout(d["obs"], d["nobs"], d["act"], d["logp"], d["rew"], d["done"], d["end"], np.arange(E), np.full(E, T), traj_stride=E)
    pol.update_critic(1, 1); pol.advantages(want_host=False); pol.normalize_adv(); pol.sync()
    rng = np.random.RandomState(11)
    perms = np.stack([rng.permutation(N) for _ in range(2)]).astype(np.int64)
    nb = (N + batch - 1) // batch
    off = [min(k * batch, N) for k in range(nb + 1)]
    ng = np.minimum(batch, N - batch * np.arange(nb)).astype(np.int64)
    out = {}
    st = pol._ext_stream()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.cuda.stream(st):
            pd = torch.from_numpy(perms[0]).to("cuda")
        st.synchronize(); t1 = time.perf_counter()
        with torch.cuda.stream(st):
            pol.actor_epoch_device(pd, off, ng)
        t2 = time.perf_counter()
        pol.update_actor(perms, batch, 1e9, 2); pol.sync()
        t3 = time.perf_counter()
        out["rep%d" % rep] = {"perm_h2d_ms": (t1 - t0) * 1e3, "epoch_device_ms": (t2 - t1) * 1e3, "update_actor_2_epochs_ms": (t3 - t2) * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
