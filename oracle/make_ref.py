"""Stage the UNMODIFIED reference next to the oracle so that it can travel to the GPU box.

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_ref      (also run by __graft_entry__.build())

/root/reference (CoderWZW/ARLib) exists in the builder container only; a ``gpurun`` box gets a snapshot of this
repository.  This recipe copies the reference's Python files and its complete ml-100k split into ``oracle/_ref/``,
byte for byte.  ``oracle/_ref/`` is listed in .gitignore (reference sources never enter this repository's history)
but not in .gpurunignore, so the copy ships with the snapshot exactly like the built ``libagcf.so``.

What uses it (and nothing else may): ``oracle/ref_loader.py`` -- the drop-in tests that run the reference's own
``ARLib`` driver and attack loops against ``arlib_b200.recommender.*`` (tests/test_gpu_dropin_ref.py), and
``bench.py``'s baselines (``--impl reference`` on the host cores, ``gpu_eager_baseline`` = the reference's own
eager-PyTorch CUDA path on the same B200).  Nothing under ``arlib_b200/`` reads it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

SRC = os.environ.get("ARLIB_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# python sources of the whole harness (callers included: they are what the drop-in tests run) + the one complete dataset
_DIRS = ("attack", "conf", "recommender", "util")
_FILES = ("ARLib.py", "main.py", "README.md")
_DATA = ("data/clean/ml-100k/train.txt", "data/clean/ml-100k/val.txt", "data/clean/ml-100k/test.txt")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def make_ref(verbose=True):
    """Copy the reference into oracle/_ref/ (idempotent).  Returns the destination, or None when the reference is
    not mounted (GPU box: the shipped copy is used as is)."""
    if not os.path.isfile(os.path.join(SRC, "recommender", "LightGCN.py")):
        return None
    manifest = {}
    rels = list(_FILES) + list(_DATA)
    for d in _DIRS:
        for root, _, files in os.walk(os.path.join(SRC, d)):
            for f in files:
                if f.endswith(".py"):
                    rels.append(os.path.relpath(os.path.join(root, f), SRC))
    for rel in sorted(rels):
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        digest = _sha(src)
        if not (os.path.isfile(dst) and _sha(dst) == digest):
            shutil.copyfile(src, dst)
        manifest[rel] = digest
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref: %d reference files staged (unmodified, sha256 in MANIFEST.json)" % len(manifest))
    return DST


if __name__ == "__main__":
    if make_ref() is None:
        sys.exit("reference not mounted at %s" % SRC)
