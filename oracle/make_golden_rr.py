"""Golden fixtures for the FULL-NETWORK configs of BASELINE.json (cfg2, cfg3, cfg4), made by running the UNMODIFIED
reference networks (``RR`` of DN_Gray / CAR / Demosaic, imported from /root/reference) on the CPU.

TEST INFRASTRUCTURE ONLY — run once in the build container:

    python oracle/make_golden_rr.py [cfg2] [cfg4] [cfg3]

  rr_cfg2_dn256.npz     DN_Gray RR, shipped checkpoint, 1x1x256x256: 256^2 crop of BSD68/test002 + sigma=25 noise
                        (DN_Gray/test.py:55-56 recipe, seed 1) -> reference output (direct, no chop)
  rr_cfg2_dn256_chop.npz  the same input through the reference's REAL inference path, Model.forward -> forward_chop
                        (model/__init__.py:114-125,179-231: 64 overlapping 72x72 leaf tiles, 4 per call), on the CPU
  rr_cfg4_dm256.npz     Demosaic RR (32 ResBlocks, 3 colours, torch.manual_seed(0) default init: the reference ships no
                        Demosaic checkpoint), 4x3x256x256 synthetic GRBG mosaics -> rows [96,160) of the reference output
  rr_cfg3_car512.npz    CAR RR, shipped checkpoint (1 colour, CAR/option.py:45-46), 1x1x512x512 = Classic5 lena at JPEG q10
                        -> rows [192,320) of the output.  The reference's own CE.forward needs ~130 GB at 512^2
                        (SURVEY §6), so every CE head of the unmodified RR is evaluated by the query-chunked oracle
                        (oracle/ce_oracle.py:ce_forward_chunked, pinned to the reference by tests/test_oracle.py);
                        everything else (ResBlocks, 1x1 merges, head/tail convs) is the reference's own code.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import ce_oracle as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402


def save(name, **arrs):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path)/1e6:.2f} MB", flush=True)


def weight_checksum(net) -> float:
    return float(sum(v.double().abs().sum() for v in net.state_dict().values()))


def synthetic_mosaic(batch: int, size: int, seed: int) -> torch.Tensor:
    """Seeded smooth-ish RGB images quantised to 8 bits and masked to the GRBG pattern of the bundled Demosaic inputs
    ((0,0)=G, (0,1)=R, (1,0)=B, (1,1)=G).  Returns float32 [batch,3,size,size] in [0,1]."""
    gen = torch.Generator().manual_seed(seed)
    low = torch.rand(batch, 3, size // 8, size // 8, generator=gen)
    img = torch.nn.functional.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)
    img = (img + 0.05 * torch.randn(batch, 3, size, size, generator=gen)).clamp(0, 1)
    img = torch.round(img * 255.0) / 255.0
    m = torch.zeros(3, size, size)
    m[1, 0::2, 0::2] = 1; m[0, 0::2, 1::2] = 1; m[2, 1::2, 0::2] = 1; m[1, 1::2, 1::2] = 1
    return (img * m).contiguous()


def cfg2():
    import cv2
    ref = R.load_task("DN_Gray")
    net = ref.dagl.RR(R.rr_args("DN_Gray")).eval()
    net.load_state_dict(torch.load(R.checkpoint("DN_Gray"), map_location="cpu"))
    img = cv2.imread("/root/reference/DN_Gray/testsets/BSD68/test002.png", cv2.IMREAD_GRAYSCALE)
    crop = np.ascontiguousarray(img[100:356, 32:288])
    clean = torch.from_numpy(crop.astype(np.float32) / 255.0)[None, None]
    torch.manual_seed(1)
    noisy = clean + torch.FloatTensor(clean.size()).normal_(mean=0, std=25 / 255.0)
    t0 = time.time()
    with torch.no_grad():
        out = net(noisy)
    print(f"cfg2 reference RR 256^2 direct: {time.time()-t0:.1f} s", flush=True)
    save("rr_cfg2_dn256.npz", clean_u8=crop, noisy=noisy, out=out, wsum=weight_checksum(net))


def cfg2chop():
    import types
    ref = R.load_task("DN_Gray")
    g = np.load(os.path.join(OUT, "rr_cfg2_dn256.npz"))
    noisy = torch.from_numpy(g["noisy"])
    with R.as_model_package(ref):
        model = ref.pkg.Model(R.wrapper_args("DN_Gray", cpu=True, chop=True), types.SimpleNamespace(dir="."))
    model.model.load_state_dict(torch.load(R.checkpoint("DN_Gray"), map_location="cpu"))
    model.eval()
    t0 = time.time()
    with torch.no_grad():
        out = model(noisy, 0)
    print(f"cfg2 reference Model.forward_chop 256^2: {time.time()-t0:.1f} s", flush=True)
    save("rr_cfg2_dn256_chop.npz", out_chop=out)


def cfg4():
    ref = R.load_task("Demosaic")
    torch.manual_seed(0)
    net = ref.dagl.RR(R.rr_args("Demosaic")).eval()
    x = synthetic_mosaic(4, 256, seed=4)
    t0 = time.time()
    with torch.no_grad():
        out = net(x)
    print(f"cfg4 reference Demosaic RR 4x3x256^2: {time.time()-t0:.1f} s", flush=True)
    save("rr_cfg4_dm256.npz", x_u8=torch.round(x * 255).to(torch.uint8), out_rows=out[:, :, 96:160].contiguous(),
         rows=np.array([96, 160]), wsum=weight_checksum(net), xsum=float(x.double().sum()),
         out_absmax=float(out.abs().max()))


def cfg3():
    import cv2
    ref = R.load_task("CAR")
    net = ref.dagl.RR(R.rr_args("CAR")).eval()
    net.load_state_dict(torch.load(R.checkpoint("CAR"), map_location="cpu"))
    img = cv2.imread("/root/reference/CAR/testsets/Classic5/lr_10/lena.jpeg")[:, :, 0]         # CAR/test.py:60,66
    x = torch.from_numpy(img.astype(np.float32) / 255.0)[None, None]
    # the reference CE.forward cannot run at 512^2 on this host: evaluate the heads with the chunked oracle
    ref.dagl.CE.forward = lambda self, b: O.ce_forward_chunked({k: v for k, v in self.state_dict().items()}, b, chunk=256)
    t0 = time.time()
    with torch.no_grad():
        out = net(x)
    print(f"cfg3 CAR RR 512^2 (reference network, chunked-oracle heads): {time.time()-t0:.1f} s", flush=True)
    save("rr_cfg3_car512.npz", x_u8=np.ascontiguousarray(img), out_rows=out[:, :, 192:320].contiguous(),
         rows=np.array([192, 320]), wsum=weight_checksum(net), out_absmax=float(out.abs().max()))


if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    which = sys.argv[1:] or ["cfg2", "cfg4", "cfg3"]
    for w in which:
        {"cfg2": cfg2, "cfg2chop": cfg2chop, "cfg4": cfg4, "cfg3": cfg3}[w]()
