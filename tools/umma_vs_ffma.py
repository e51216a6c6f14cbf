"""Kernel bisection: run one update of one agent on the FFMA tiles and on the tcgen05 tiles with identical inputs and print
the norm-relative difference of every scratch buffer (GPU only; debugging aid, not part of the product path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spp_rl_b200 import Population, init_state, _lib
from tests.parity_util import make_stats, make_batches, upload_state, relnorm

NAMES = ["xo", "xn", "xc", "ha1", "ha2", "ml", "xcp", "xm", "hm1", "hm2", "tm3", "hc1_0", "hc1_1", "hc2_0", "hc2_1", "vec", "dz2_0", "dz2_1",
         "dz1_0", "dz1_1", "dxc", "dm3", "dm2", "dm1", "dxm", "dml", "dza2", "dza1", "gvec"]


def run(path, algo="sac", ob=11, ac=3, B=256, seed=0, G=1):
    _lib.load_library().spp_set_gemm_path(path)
    mn, mx, mean, std = make_stats(ob, seed, True)
    obs, nobs, act, rew, done, aacm, eps = make_batches(ob, ac, 1, G, B, seed, mn, mx)
    pop = Population(algo=algo, ob_dim=ob, ac_dim=ac, population=1, update_batch_size=B, gamma=0.99, custom_loss=0.2,
                     acm_critic=True, norm_closs=False, min_max_denormalize=True, alpha=0.2, target_entropy=-float(ac))
    pop.set_norm_stats(mn, mx, mean, std)
    upload_state(pop, init_state(algo, ob, ac, seed * 100), 0, algo)
    losses = pop.update_host(G, obs, nobs, act, rew, done, aacm, eps=eps if algo == "sac" else None)
    out = {n: pop.debug_scratch(0, n).copy() for n in NAMES}
    nets = ["actor", "critic_1", "critic_2", "critic_1_targ"] if algo == "sac" else ["actor", "critic"]
    for net in nets:
        for k, v in pop.state_dict(net, agent=0).items():
            out[net + "." + k] = v.copy()
    pop.close()
    return losses, out


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    l0, o0 = run(0, B=B, G=G)
    l1, o1 = run(1, B=B, G=G)
    for g in range(G):
        print("losses ffma", g, l0[0, g]); print("losses umma", g, l1[0, g])
    for n in o0:
        d = relnorm(o1[n], o0[n])
        print("%-28s relnorm %.3e  max|d| %.3e%s" % (n, d, np.abs(o1[n] - o0[n]).max(), "   <<<" if d > 1e-5 else ""))
