"""Deterministic parameter initialisation in the reference's state_dict layout.

Mirrors torch.nn.Linear's default init (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias) with a
numpy RandomState so the same numbers are produced on any box; shapes follow
rltoolkit/algorithms/sac/models.py:8-22,72-79, rltoolkit/algorithms/ddpg/models.py:5-15,31-37,
rltoolkit/basic_model.py:108-116 and rltoolkit/acm/models/basic_acm.py:11-22.
"""
from collections import OrderedDict

import numpy as np

HIDDEN = 256


def net_shapes(algo, ob, ac, acm_kind="acm", acm_critic=True):
    """-> OrderedDict net -> OrderedDict tensor name -> shape (reference state_dict order)."""
    def lin(d, name, out_f, in_f):
        d[name + ".weight"] = (out_f, in_f)
        d[name + ".bias"] = (out_f,)

    actor = OrderedDict()
    lin(actor, "fc1", HIDDEN, ob); lin(actor, "fc2", HIDDEN, HIDDEN)
    if algo == "sac":
        lin(actor, "fc_prob", ob, HIDDEN); lin(actor, "fc_scale", ob, HIDDEN)
    else:
        lin(actor, "fc3", ob, HIDDEN)
    critic = OrderedDict()
    lin(critic, "fc1", HIDDEN, ob + (ac if acm_critic else ob)); lin(critic, "fc2", HIDDEN, HIDDEN); lin(critic, "fc3", 1, HIDDEN)
    acm = OrderedDict()
    if acm_kind in ("acm", "mlp"):
        lin(acm, "fc1", 64, 2 * ob); lin(acm, "fc2", 32, 64); lin(acm, "fc3", ac, 32)
    else:
        acm["t"] = (1,); acm["t1"] = (ac,)
        lin(acm, "fc1", 100, 2 * ob); lin(acm, "fc2", 50, 100); lin(acm, "fc21", 50, 2 * ob); lin(acm, "fc3", ac, 50)
    nets = OrderedDict()
    nets["actor"] = actor
    if algo == "sac":
        nets["critic_1"] = critic; nets["critic_2"] = critic
    else:
        nets["critic"] = critic
    nets["acm"] = acm
    return nets


def init_state(algo, ob, ac, seed, acm_kind="acm", acm_critic=True):
    """-> dict '<net>.<tensor>' -> float32 ndarray, including target copies (deep copies at creation)."""
    rng = np.random.RandomState(seed)
    out = OrderedDict()
    for net, tensors in net_shapes(algo, ob, ac, acm_kind, acm_critic).items():
        for name, shape in tensors.items():
            if name in ("t", "t1"):
                out[net + "." + name] = np.ones(shape, np.float32)
                continue
            layer = name.rsplit(".", 1)[0]
            fan_in = tensors[layer + ".weight"][1]
            bound = 1.0 / np.sqrt(fan_in)
            out[net + "." + name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    targets = {"sac": [("critic_1", "critic_1_targ"), ("critic_2", "critic_2_targ")],
               "ddpg": [("critic", "critic_targ"), ("actor", "actor_targ")]}[algo]
    for src, dst in targets:
        for k in [k for k in out if k.startswith(src + ".")]:
            out[dst + k[len(src):]] = out[k].copy()
    return out
