import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_ROOT = "/root/reference/DN_Gray"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_npz(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def rand_weights():
    return load_npz("ce_rand_w.npz")


def have_reference() -> bool:
    return os.path.isdir(REF_ROOT)


def import_reference():
    """The unmodified reference model package (only in the build container)."""
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import model.dagl as ref  # noqa
    return ref


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure the in-tree library exists (nvcc cross-compiles without a GPU)."""
    from dagl_b200 import build
    if build.needs_build():
        try:
            build.build()
        except Exception as e:  # pragma: no cover
            pytest.exit(f"could not build libdagl_b200.so: {e}")
    yield
