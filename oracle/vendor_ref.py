"""Vendor the UNMODIFIED reference model files into the git-ignored ``baseline/_ref/``.

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` exists only in the build container; the GPU boxes receive a snapshot
of ``/root/repo``.  ``baseline/_ref/`` is git-ignored (no reference source enters the history) but travels with the
snapshot, so the ``-m gpu`` drop-in tests can build the reference's own networks (``RR``, ``Model.forward_chop``) and
load its shipped checkpoints on the box.  Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present:

    python oracle/vendor_ref.py

Copied verbatim, per task directory (DN_Gray, CAR, Demosaic): ``model/{__init__,dagl,common}.py`` and, where the
reference ships one, ``exp/model/model_best.pt``.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
TASKS = ("DN_Gray", "CAR", "Demosaic")
FILES = ("model/__init__.py", "model/dagl.py", "model/common.py", "exp/model/model_best.pt")


def vendor(verbose: bool = False) -> bool:
    if not os.path.isdir(SRC):
        return False
    for task in TASKS:
        for rel in FILES:
            s = os.path.join(SRC, task, rel)
            d = os.path.join(DST, task, rel)
            if not os.path.exists(s):
                continue
            if os.path.exists(d) and os.path.getsize(d) == os.path.getsize(s):
                continue
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
            if verbose:
                print("vendored", os.path.relpath(d, ROOT))
    return True


if __name__ == "__main__":
    ok = vendor(verbose=True)
    print("baseline/_ref ready" if ok else "no /root/reference here", file=sys.stderr)
