"""Import the UNMODIFIED reference (rltoolkit) -- ORACLE INFRASTRUCTURE.

From /root/reference where that exists (the build container: tests/golden/make_golden.py generates the fixtures from it), else
from oracle/_ref (the byte-for-byte copy made by oracle/build_ref.py, which travels to the GPU box so that bench.py's CPU
baseline can time the reference's own code there).  Installs
the three shims SURVEY.md section 8c lists: a gym stub, a pyvirtualdisplay stub, and the removed
`numpy.int` alias (rltoolkit/buffer/replay_buffer.py:29-30,106).
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference/rltoolkit"
if not os.path.isdir(os.path.join(REF_ROOT, "rltoolkit")):
    REF_ROOT = os.path.join(_HERE, "_ref")
_SHIMS = os.path.join(_HERE, "shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "rltoolkit"))


def import_reference(scratch_dir: str = "/tmp/spp_ref_scratch"):
    """Return the imported `rltoolkit` module (reference), with shims installed."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import numpy as np

    if not hasattr(np, "int"):
        np.int = int  # alias removed in numpy 1.24
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if REF_ROOT not in sys.path:
        sys.path.insert(1, REF_ROOT)
    # rltoolkit/logger.py:11-18 creates <dir of sys.argv[0]>/logs at import time.
    os.makedirs(scratch_dir, exist_ok=True)
    old_argv0 = sys.argv[0]
    sys.argv[0] = os.path.join(scratch_dir, "run.py")
    try:
        import rltoolkit  # noqa: F401
    finally:
        sys.argv[0] = old_argv0
    import logging

    logging.getLogger().setLevel(logging.WARNING)
    for h in logging.getLogger().handlers:
        h.setLevel(logging.WARNING)
    return rltoolkit
