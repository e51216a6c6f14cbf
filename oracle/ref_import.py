"""Import the UNMODIFIED reference (rltoolkit) in the build container -- ORACLE INFRASTRUCTURE.

Only usable where /root/reference exists (this container, never the GPU box).  Used by
tests/golden/make_golden.py to generate fixtures and by oracle validation scripts.  Installs
the three shims SURVEY.md section 8c lists: a gym stub, a pyvirtualdisplay stub, and the removed
`numpy.int` alias (rltoolkit/buffer/replay_buffer.py:29-30,106).
"""
import os
import sys

REF_ROOT = "/root/reference/rltoolkit"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


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
