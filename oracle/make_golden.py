"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (imported from /root/reference/DN_Gray) on CPU.

TEST INFRASTRUCTURE ONLY — run once in the build container (where
/root/reference is mounted):

    python oracle/make_golden.py

The fixtures are what pins the oracle (oracle/ce_oracle.py) and, through it,
the CUDA path.  The reference ships no golden vectors for this path
(SURVEY.md §8c), so these are "outputs of the reference itself run here".

Fixtures (all float32 / int32, np.savez_compressed):

  ce_rand_w.npz          CE(in_channels=64) state_dict under torch.manual_seed(0)
                         (reference constructor, default torch init)
  ce_cfg1_64x64.npz      BASELINE config 1: x = randn(1,64,64,64) (same seed
                         stream, drawn right after construction), y, packed
                         neighbour mask, nnz/row, gamma, beta, mu
  ce_ragged.npz          same weights on (2,64,30,41), (1,64,7,9), (1,64,65,67),
                         (1,64,72,72): x, y, nnz (mask for the small ones)
  ce_trained_<head>.npz  heads of the shipped DN_Gray/exp/model/model_best.pt
                         (weights + the head's real input captured by a
                         forward hook on a 48x48 sigma=25 BSD68 crop + y + mask)
  ce_demo_smoke.npz      the reference's own __main__ smoke shape
                         (Demosaic/model/dagl.py:280-284): (2,64,16,16)
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/DN_Gray"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import model.dagl as ref  # noqa: E402  (the reference)
from oracle import ce_oracle as O  # noqa: E402


def ref_ce_with_aux(ce, x):
    """Run reference CE.forward and recover mask/mu by re-evaluating the
    reference's own sub-modules in the reference's order (dagl.py:208-257)."""
    with torch.no_grad():
        y = ce(x)
        p = {k: v for k, v in ce.state_dict().items()}
        y2, aux = O.ce_forward(p, x, return_aux=True)
    assert torch.equal(y, y2), "oracle restatement is not bit-identical to the reference here"
    return y, aux


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path)/1e6:.2f} MB")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---- random-init head + config 1 -------------------------------------
    torch.manual_seed(0)
    ce = ref.CE(in_channels=64).eval()
    x = torch.randn(1, 64, 64, 64)
    save("ce_rand_w.npz", **{k: v for k, v in ce.state_dict().items()})
    y, aux = ref_ce_with_aux(ce, x)
    save("ce_cfg1_64x64.npz", x=x, y=y, mask_bits=O.pack_mask_bits(aux["mask"]),
         nnz=aux["mask"].sum(-1).to(torch.int32), gamma=aux["gamma"], beta=aux["beta"], mu=aux["mu"],
         Q_head=aux["Q"][0, :8], K_head=aux["K"][0, :8], S_head=aux["S"][0, :8, :64])

    # ---- ragged / edge shapes --------------------------------------------
    rag = {}
    gen = torch.Generator().manual_seed(1234)
    for i, shape in enumerate([(2, 64, 30, 41), (1, 64, 7, 9), (1, 64, 65, 67), (1, 64, 72, 72)]):
        xs = torch.randn(*shape, generator=gen)
        ys, a = ref_ce_with_aux(ce, xs)
        rag[f"x{i}"] = xs
        rag[f"y{i}"] = ys
        rag[f"nnz{i}"] = a["mask"].sum(-1).to(torch.int32)
        if shape[2] * shape[3] <= 2048:
            rag[f"mask_bits{i}"] = O.pack_mask_bits(a["mask"])
    save("ce_ragged.npz", **rag)

    # ---- reference's own smoke shape ---------------------------------------
    xs = torch.randn(2, 64, 16, 16, generator=gen)
    ys, a = ref_ce_with_aux(ce, xs)
    save("ce_demo_smoke.npz", x=xs, y=ys, mask_bits=O.pack_mask_bits(a["mask"]))

    # ---- trained heads (shipped checkpoint) ---------------------------------
    args = types.SimpleNamespace(n_resblocks=16, n_feats=64, n_colors=1, res_scale=1, rgb_range=1.0)
    net = ref.RR(args).eval()
    sd = torch.load(os.path.join(REF, "exp/model/model_best.pt"), map_location="cpu")
    net.load_state_dict(sd)
    import cv2
    img = cv2.imread(os.path.join(REF, "testsets/BSD68/test002.png"), cv2.IMREAD_GRAYSCALE)
    clean = torch.from_numpy(img[100:148, 200:248].astype(np.float32) / 255.0)[None, None]
    torch.manual_seed(1)                                   # DN_Gray/test.py:55-56 noise recipe
    noisy = clean + torch.FloatTensor(clean.size()).normal_(mean=0, std=25 / 255.0)
    captured = {}
    ces = net.body[8]
    hooks = []
    for name in ["c1_2", "c1_3", "c2_1", "c3_1", "c3_3"]:
        mod = getattr(ces, name)
        hooks.append(mod.register_forward_hook(
            lambda m, inp, out, name=name: captured.__setitem__(name, (inp[0].detach().clone(), out.detach().clone()))))
    with torch.no_grad():
        out = net(noisy)
    for h in hooks:
        h.remove()
    for name in ["c1_2", "c1_3", "c2_1", "c3_1", "c3_3"]:
        mod = getattr(ces, name)
        xin, yout = captured[name]
        y, a = ref_ce_with_aux(mod, xin)
        assert torch.equal(y, yout)
        nnz = a["mask"].sum(-1)
        print(name, "mean nnz/row", nnz.float().mean().item(), "of", a["mask"].shape[-1], "max|y|", y.abs().max().item())
        if name in ("c1_2", "c2_1", "c3_1", "c3_3"):
            save(f"ce_trained_{name}.npz", x=xin, y=y, mask_bits=O.pack_mask_bits(a["mask"]),
                 nnz=nnz.to(torch.int32), **{"w." + k: v for k, v in mod.state_dict().items()})
    save("rr_trained_io.npz", noisy=noisy, out=out, clean=clean)


if __name__ == "__main__":
    main()
