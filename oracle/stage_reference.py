"""ORACLE / test infrastructure — stage the reference's model file for the GPU box.

`bench.py --impl reference` must time the UNMODIFIED reference forward on the box's host cores, but /root/reference exists in
the build container only.  The reference's hot path is one pure-Python file (MIT licence, `MPL/lib/models/multiview_mpl.py`);
this recipe copies it, byte for byte, to `oracle/_ref/MPL/lib/models/multiview_mpl.py`.  `oracle/_ref/` is git-ignored (no
reference source enters the history) but not gpurun-ignored, so the copy travels to the GPU box with the built library.
Run by `__graft_entry__.build()` whenever /root/reference is present.

    python -m oracle.stage_reference
"""
from __future__ import annotations

import filecmp
import os
import shutil

from . import ref_loader

REL = "MPL/lib/models/multiview_mpl.py"


def stage() -> str | None:
    """Copy the model file if the reference is mounted; returns the staged path (or None when nothing is staged)."""
    src = os.path.join(ref_loader.REF_ROOT, REL)
    dst = os.path.join(ref_loader.STAGE_ROOT, REL)
    if os.path.isfile(src):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    return dst if os.path.isfile(dst) else None


if __name__ == "__main__":
    print(stage())
