"""CPU-only: the C-ABI library loads and exports every symbol include/spp_rl_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "spp_rl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import ctypes

    from spp_rl_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "library does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes binding and header disagree"
    assert lib.spp_abi_version() == _lib.ABI_VERSION


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product path must raise, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from spp_rl_b200 import Population, SppError

    with pytest.raises(SppError):
        Population(algo="sac", ob_dim=11, ac_dim=3, population=1)
