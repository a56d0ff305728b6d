"""Golden fixture for the Demosaic / DN_Real wrapper variants of the tile scheduler (shave_size_max = 12 and the
``self_ensemble`` branch), made by running the UNMODIFIED ``Demosaic/model/__init__.py`` on the CPU.

TEST INFRASTRUCTURE ONLY -- run once in the build container:   python oracle/make_golden_chop_demosaic.py

``Model.forward_chop`` (Demosaic/model/__init__.py:179-235) hard-codes shave_size_max = 12; ``Model.forward_x8``
(:265-299) is the wrapper's own 8-fold ensemble (flip / flip / transpose lists, mean over the batch axis).  Its
``_transform`` only assigns its result on the ``not self.cpu`` path (``ret = torch.Tensor(tfnp).cuda()``), i.e. the
reference code needs a GPU there; the generator runs it with ``cpu=False`` and ``torch.Tensor.cuda`` temporarily bound to
the identity, so the reference's lines execute unmodified on CPU tensors.  The network inside the wrapper is the
position-dependent stand-in of make_golden_chop.py.   -> tests/golden/chop_probe_demosaic.npz
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader as R  # noqa: E402


def probe_net(t: torch.Tensor) -> torch.Tensor:
    n, c, h, w = t.shape
    yy = torch.arange(h, dtype=t.dtype).view(1, 1, h, 1)
    xx = torch.arange(w, dtype=t.dtype).view(1, 1, 1, w)
    return t * (1.0 + 0.01 * yy + 0.0003 * xx) + 0.125 * yy - 0.0625 * xx


def main():
    ref = R.load_task("Demosaic")
    args = R.wrapper_args("Demosaic", cpu=True, chop=True)
    ckp = types.SimpleNamespace(dir="/tmp", log_file=open(os.devnull, "w"))
    with R.as_model_package(ref):
        m = ref.pkg.Model(args, ckp)
    m.eval()

    class Probe(torch.nn.Module):
        def forward(self, t):
            return probe_net(t)
    m.model = Probe()
    out = {}
    torch.manual_seed(11)
    for i, (b, h, w) in enumerate([(2, 100, 104), (1, 130, 97)]):
        x = torch.rand(b, 3, h, w)
        m.ensemble, m.idx_scale = False, 0
        with torch.no_grad():
            y = m.forward_chop(x)
        out[f"x{i}"] = x.numpy(); out[f"y{i}"] = y.numpy()
    # self_ensemble branch: forward_x8 over forward_chop, one image (the mean runs over the batch axis)
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        m.cpu = False
        x = torch.rand(1, 3, 100, 104)
        with torch.no_grad():
            y = m.forward_x8(x, m.forward_chop)
            y2 = m.forward_x8(x[:, :, :40, :52], m.model.forward)
    finally:
        torch.Tensor.cuda = saved
    out["xe"] = x.numpy(); out["ye"] = y.numpy(); out["ye_direct"] = y2.numpy()
    path = os.path.join(ROOT, "tests", "golden", "chop_probe_demosaic.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
