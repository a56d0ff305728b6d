"""Golden fixture for the LEGACY fixed-top-k variant, made by running the reference's own leftover class
(DN_Gray/model/.ipynb_checkpoints/GReccR2b_3mh_1-checkpoint.py, class CE) on the CPU.

TEST INFRASTRUCTURE ONLY — run once in the build container:   python oracle/make_golden_topk.py

That class ends with ``y = self.W(y); y = b + y`` (:262-263).  To expose the graph stage alone, W is set to a channel
selector (W.weight[c, c] = 1 for c < 16, else 0; W.bias = 0), so that ``out[:, :16] - b[:, :16]`` is exactly the folded
aggregation the shipping CE would return.  Weights g / theta / fc1 / fc2 are the committed random-init head
(tests/golden/ce_rand_w.npz); inputs are seeded.  -> tests/golden/ce_topk_legacy.npz
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader as R  # noqa: E402

LEGACY = "/root/reference/DN_Gray/model/.ipynb_checkpoints/GReccR2b_3mh_1-checkpoint.py"


def main():
    ref = R.load_task("DN_Gray")
    with R.as_model_package(ref):                      # the leftover file does `import model.common as common`
        spec = importlib.util.spec_from_file_location("legacy_topk_model", LEGACY)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    with np.load(os.path.join(ROOT, "tests", "golden", "ce_rand_w.npz")) as z:
        w = {k: torch.from_numpy(z[k]) for k in z.files}
    out = {}
    gen = torch.Generator().manual_seed(77)
    for tag, (shape, k) in {"a": ((1, 64, 32, 36), 8), "b": ((2, 64, 30, 41), 50), "c": ((1, 64, 12, 10), 50)}.items():
        ce = mod.CE(in_channels=64, num_edge=k).eval()
        sd = ce.state_dict()
        for name in ("g.weight", "g.bias", "theta.weight", "theta.bias", "fc1.0.weight", "fc1.0.bias", "fc2.0.weight", "fc2.0.bias"):
            sd[name].copy_(w[name])
        sd["W.weight"].zero_(); sd["W.bias"].zero_()
        for c in range(16):
            sd["W.weight"][c, c, 0, 0] = 1.0
        x = torch.randn(*shape, generator=gen)
        with torch.no_grad():
            y = ce(x)[:, :16] - x[:, :16]
        out[f"x_{tag}"] = x.numpy(); out[f"y_{tag}"] = y.numpy(); out[f"k_{tag}"] = np.array(k)
        print(tag, shape, k, float(y.abs().max()))
    path = os.path.join(ROOT, "tests", "golden", "ce_topk_legacy.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
