"""Recipe for ``oracle/_ref`` (git-ignored, NOT gpurun-ignored: it travels to the GPU box like the built .so files).

The reference is pure Python (no native code to compile); its implementation of the hot path is the four files of
/root/reference/models (532 lines, MIT).  This recipe places an unmodified snapshot of that package under oracle/_ref/models so that
``bench.py --impl reference`` and ``cpu_baseline`` time the REFERENCE's own modules on the GPU box's host cores (kind "reference"), where
/root/reference does not exist.  Nothing from it enters the repository's history, and no product module imports it.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/models"
DST = os.path.join(HERE, "_ref", "models")


def build_ref(verbose: bool = False) -> bool:
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for name in sorted(os.listdir(SRC)):
        if not name.endswith(".py"):
            continue
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
        manifest[name] = hashlib.sha256(open(os.path.join(DST, name), "rb").read()).hexdigest()
    for extra in ("LICENSE", "LICENSE.txt", "LICENSE.md"):
        p = os.path.join(os.path.dirname(SRC), extra)
        if os.path.exists(p):
            shutil.copyfile(p, os.path.join(os.path.dirname(DST), extra))
    with open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} files from {SRC}")
    return True


if __name__ == "__main__":
    ok = build_ref(verbose=True)
    sys.exit(0 if ok else 1)
