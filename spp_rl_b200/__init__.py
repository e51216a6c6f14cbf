"""spp_rl_b200: B200-native (sm_100a) implementation of SPP-RL's update-and-rollout hot path.

The arithmetic lives in hand-written CUDA behind the C ABI of include/spp_rl_b200.h
(libspp_rl_b200.so, built in-tree by __graft_entry__.build()).  `Population` is the host handle;
`spp_rl_b200.rltoolkit_api` mirrors the reference's algorithm classes on top of it.
"""
from ._lib import SppError, load_library  # noqa: F401
from .init import init_state, net_shapes  # noqa: F401
from .population import Population, kernel_launches  # noqa: F401

__all__ = ["Population", "SppError", "load_library", "init_state", "net_shapes", "kernel_launches"]
