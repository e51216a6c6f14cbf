#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference package, copied where it lies so that it travels to the GPU box.

The reference (raznem/spp-rl, `rltoolkit`) is pure Python -- there is nothing to compile.  This script copies the package
directory /root/reference/rltoolkit/rltoolkit to oracle/_ref/rltoolkit byte for byte (tests and __pycache__ excluded) and writes a
manifest with the sha256 of every file.  oracle/_ref/ is git-ignored (never part of the history, never edited) but not
gpurun-ignored, so `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's OWN SAC_AcM.update on the GPU
box's host cores (cpu_baseline.kind = "reference").  It is test / measurement infrastructure: nothing under spp_rl_b200/ imports it.

    python oracle/build_ref.py          (also run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/rltoolkit/rltoolkit"
DST = os.path.join(HERE, "_ref", "rltoolkit")


def build(verbose=False):
    if not os.path.isdir(SRC):
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "test", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    # every copied file is identical to its source
    for rel, h in manifest.items():
        assert hashlib.sha256(open(os.path.join(SRC, rel), "rb").read()).hexdigest() == h, rel
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d files copied from %s" % (len(manifest), SRC))
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    sys.exit(0 if ok else 1)
