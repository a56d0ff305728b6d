"""Golden fixtures for the tile scheduler (dagl_b200/chop.py), made by running the UNMODIFIED reference
wrapper ``Model.forward_chop`` / ``test_x8`` (DN_Gray/model/__init__.py:179-231, 53-62) on CPU.

TEST INFRASTRUCTURE ONLY — run once in the build container:   python oracle/make_golden_chop.py

The network inside the wrapper is replaced by a cheap, *position-dependent* stand-in
(``probe_net``): its output at a pixel depends on the pixel's coordinates inside the tile, so the
stitched result records which pixel of which tile the reference copied where, and an asymmetric
ramp makes every flip/rotation of the x8 ensemble distinguishable.

  chop_probe.npz   for several (h, w, shave_size_max): input, reference forward_chop output,
                   and the reference x8 output on one tile
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/DN_Gray"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)
import model as ref_model  # noqa: E402   (the reference package: model/__init__.py)


def probe_net(t: torch.Tensor) -> torch.Tensor:
    """Stand-in network: [n, C, h, w] -> [n, C, h, w]; depends on tile-local coordinates (same function as in
    tests/test_chop.py)."""
    n, c, h, w = t.shape
    yy = torch.arange(h, dtype=t.dtype).view(1, 1, h, 1)
    xx = torch.arange(w, dtype=t.dtype).view(1, 1, 1, w)
    return t * (1.0 + 0.01 * yy + 0.0003 * xx) + 0.125 * yy - 0.0625 * xx


def make_wrapper(shave_size_max):
    args = types.SimpleNamespace(
        model="dagl", scale=[1], self_ensemble=False, chop=True, precision="single", cpu=True, n_GPUs=1,
        save_models=False, print_model=False, pre_train=".", resume=0, n_resblocks=2, n_feats=64, n_colors=1,
        res_scale=1, rgb_range=1, test_only=True)
    ckp = types.SimpleNamespace(dir="/tmp", log_file=open(os.devnull, "w"))
    m = ref_model.Model(args, ckp)
    m.eval()
    class Probe(torch.nn.Module):            # the wrapper only ever calls self.model(batch)
        def forward(self, t):
            return probe_net(t)
    m.model = Probe()
    assert shave_size_max == 24, "DN_Gray wrapper hard-codes 24 (model/__init__.py:187)"
    return m


def main():
    torch.manual_seed(7)
    m = make_wrapper(24)
    out = {}
    cases = [(100, 120), (256, 256), (321, 481), (130, 97)]
    for i, (h, w) in enumerate(cases):
        x = torch.rand(2 if i == 0 else 1, 1, h, w)
        m.ensemble = False
        m.idx_scale = 0
        with torch.no_grad():
            y = m.forward_chop(x)
        out[f"x{i}"] = x.numpy()
        out[f"y{i}"] = y.numpy()
    # x8 ensemble on a non-square tile batch
    xt = torch.rand(2, 1, 20, 28)
    with torch.no_grad():
        yt = ref_model.test_x8(lambda t: t * torch.linspace(0.5, 1.5, t.shape[-1]).view(1, 1, 1, -1)
                               + torch.linspace(-1, 1, t.shape[-2]).view(1, 1, -1, 1), xt)
    out["x8_in"] = xt.numpy()
    out["x8_out"] = yt.numpy()
    # chop + ensemble together (what `test.py --ensemble` runs)
    x = torch.rand(1, 1, 100, 120)
    m.ensemble = True
    with torch.no_grad():
        y = m.forward_chop(x)
    out["xe"] = x.numpy()
    out["ye"] = y.numpy()
    path = os.path.join(OUT, "chop_probe.npz")
    np.savez_compressed(path, **out)
    print("chop_probe.npz", os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
