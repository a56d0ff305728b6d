"""Import the UNMODIFIED reference ``model`` package of one task directory (DN_Gray / CAR / Demosaic).

TEST INFRASTRUCTURE ONLY (tests/, golden generators).  The reference's task directories each hold a top-level package
called ``model`` (``import model.common as common`` inside ``dagl.py``), so two of them cannot be imported side by side
under that name; ``load_task`` imports one, then re-registers its modules under ``ref_<task>.*`` and clears the
``model*`` entries so the next task can be loaded.

Search order: ``baseline/_ref/<task>`` (vendored by oracle/vendor_ref.py; what a GPU box has), then
``/root/reference/<task>`` (the build container).
"""
import contextlib
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = {}


def task_root(task: str):
    for base in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        p = os.path.join(base, task)
        if os.path.isfile(os.path.join(p, "model", "dagl.py")):
            return p
    return None


def available(task: str = "DN_Gray") -> bool:
    return task_root(task) is not None


def checkpoint(task: str):
    r = task_root(task)
    p = os.path.join(r, "exp", "model", "model_best.pt") if r else None
    return p if p and os.path.isfile(p) else None


def load_task(task: str) -> types.SimpleNamespace:
    """-> namespace(pkg=<model/__init__>, dagl=<model.dagl>, common=<model.common>, root=<dir>)"""
    if task in _CACHE:
        return _CACHE[task]
    root = task_root(task)
    if root is None:
        raise FileNotFoundError(f"reference task {task} not found under baseline/_ref or /root/reference")
    stale = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "model" or k.startswith("model.")}
    sys.path.insert(0, root)
    try:
        pkg = importlib.import_module("model")
        dagl = importlib.import_module("model.dagl")
        common = importlib.import_module("model.common")
    finally:
        sys.path.remove(root)
        for k in list(sys.modules):
            if k == "model" or k.startswith("model."):
                sys.modules[f"ref_{task}." + k] = sys.modules.pop(k)
        sys.modules.update(stale)
    ns = types.SimpleNamespace(pkg=pkg, dagl=dagl, common=common, root=root)
    _CACHE[task] = ns
    return ns


def rr_args(task: str, n_colors: int = None):
    """The five options that reach ``RR`` (dagl.py:15-16,24,29,38,41) with each task's defaults."""
    nc = n_colors if n_colors is not None else (3 if task == "Demosaic" else 1)
    return types.SimpleNamespace(n_resblocks=16, n_feats=64, n_colors=nc, res_scale=1, rgb_range=1.0)


def wrapper_args(task: str, cpu: bool = True, chop: bool = True, n_colors: int = None):
    """Options ``Model.__init__`` reads (model/__init__.py:78-112)."""
    a = rr_args(task, n_colors)
    a.__dict__.update(scale=[1], self_ensemble=False, chop=chop, precision="single", cpu=cpu, n_GPUs=1,
                      save_models=False, model="dagl", pre_train=".", resume=0, print_model=False, seed=1)
    return a


@contextlib.contextmanager
def as_model_package(ns):
    """Temporarily expose a loaded task as the top-level package ``model`` again: the reference's ``Model.__init__``
    resolves its plugin by name (``import_module('model.' + args.model.lower())``, model/__init__.py:92-93)."""
    names = {"model": ns.pkg, "model.dagl": ns.dagl, "model.common": ns.common}
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.update(names)
    try:
        yield ns
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
