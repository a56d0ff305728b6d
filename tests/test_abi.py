"""CPU tests of the boundary: the in-tree C-ABI library loads and exports every
symbol include/dagl_b200.h declares; host-side argument checking fails loudly;
no compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "dagl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dagl_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from dagl_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dagl_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_abi_version_and_sizes():
    from dagl_b200 import _lib
    L = _lib.lib()
    assert L.dagl_abi_version() == 5
    small = L.dagl_ce_workspace_bytes(1, 64, 64, 64)
    big = L.dagl_ce_workspace_bytes(1, 64, 256, 256)
    assert 0 < small < big < 2 ** 33
    assert L.dagl_ce_workspace_bytes(0, 64, 64, 64) == 0
    assert L.dagl_ce_host_staging_bytes(1, 64, 64, 64) >= 4 * (64 + 16) * 64 * 64
    # K embeddings [Nk,196] fp32 must fit
    assert big >= 4 * 65536 * 196


def test_argument_validation_without_gpu():
    from dagl_b200 import _lib
    L = _lib.lib()
    w = _lib.DaglCEWeights()
    rc = L.dagl_ce_forward_f32(ctypes.byref(w), None, None, 1, 8, 8, None, 0, 0, None)
    assert rc == -1 and b"null" in L.dagl_last_error()
    # unsupported configuration is refused, not emulated
    fake = ctypes.c_void_p(16)
    w2 = _lib.DaglCEWeights(*([16] * 12), 64, 16, 5, 4, 1, 10.0)
    rc = L.dagl_ce_forward_f32(ctypes.byref(w2), fake, fake, 1, 8, 8, fake, 1 << 30, 0, None)
    assert rc == -2 and b"unsupported" in L.dagl_last_error()
    w3 = _lib.DaglCEWeights(*([16] * 12), 64, 16, 7, 4, 1, 10.0)
    rc = L.dagl_ce_forward_f32(ctypes.byref(w3), fake, fake, 1, 8, 8, fake, 16, 0, None)
    assert rc == -3 and b"workspace" in L.dagl_last_error()


def test_module_has_no_cpu_path():
    import dagl_b200
    ce = dagl_b200.CE(in_channels=64)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        ce(torch.zeros(1, 64, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):          # also under autograd: the forward is always the CUDA path
        ce(torch.zeros(1, 64, 8, 8))
    with torch.no_grad(), pytest.raises(RuntimeError, match="fp32"):
        ce(torch.zeros(1, 64, 8, 8, dtype=torch.float64))


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (task rule ③)."""
    pkg = os.path.join(ROOT, "dagl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_constructor_signature_matches_reference():
    import inspect
    import dagl_b200
    sig = inspect.signature(dagl_b200.CE.__init__)
    names = list(sig.parameters)[1:13]
    assert names == ["ksize", "stride_1", "stride_2", "softmax_scale", "shape", "p_len", "in_channels",
                     "inter_channels", "use_multiple_size", "use_topk", "add_SE", "num_edge"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["ksize"], d["stride_1"], d["stride_2"], d["softmax_scale"], d["in_channels"],
            d["inter_channels"], d["num_edge"]) == (7, 4, 1, 10, 64, 16, 50)
