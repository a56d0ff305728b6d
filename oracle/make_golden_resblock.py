"""Golden fixture for the ResBlock chain, made by running the UNMODIFIED reference class on the CPU.

TEST INFRASTRUCTURE ONLY -- run once in the build container:   python oracle/make_golden_resblock.py
Builds ``nn.Sequential`` of the reference's ``common.ResBlock`` (DN_Gray/model/common.py:59-79) exactly as ``CES`` does
(dagl.py:86-101: default_conv, n_feats 64, kernel 3, act PReLU, res_scale 1), loads seeded weights and records
input / output (and a checksum of the seeded weights) for two small chains.  -> tests/golden/resblock_chain.npz
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader as R  # noqa: E402
from oracle import resblock_oracle as RB  # noqa: E402


def main():
    common = R.load_task("DN_Gray").common
    out = {}
    for tag, (shape, nb, res_scale) in {"a": ((1, 64, 12, 10), 2, 1), "b": ((2, 64, 9, 17), 3, 0.5)}.items():
        seq = nn.Sequential(*[common.ResBlock(common.default_conv, n_feats=64, kernel_size=3, act=nn.PReLU(), res_scale=res_scale)
                              for _ in range(nb)]).eval()
        for i, blk in enumerate(seq):
            blk.load_state_dict(RB.init_resblock_params(100 * ord(tag) + i))
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(ord(tag)))
        with torch.no_grad():
            y = seq(x)
        out[f"{tag}_x"] = x.numpy(); out[f"{tag}_y"] = y.numpy()
        out[f"{tag}_res_scale"] = np.float32(res_scale); out[f"{tag}_nb"] = np.int32(nb)
        # the weights are re-made from their seeds by the tests (RB.init_resblock_params); a checksum detects RNG drift
        out[f"{tag}_wsum"] = np.float64(sum(float(v.double().abs().sum()) for blk in seq for v in blk.state_dict().values()))
    path = os.path.join(ROOT, "tests", "golden", "resblock_chain.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
