"""ResBlock chain kernel vs torch (development check + timing).  usage: resblock_check.py [single|pair] [time]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import dagl_b200
from dagl_b200.resblock import ResBlock, resblocks_forward

mode = sys.argv[1] if len(sys.argv) > 1 else "single"
dev = torch.device("cuda")
torch.manual_seed(0)


def ref64(blocks, x):            # fp64 on the CPU: ground truth
    x = x.double().cpu()
    for b in blocks:
        c1, a, c2 = b.body
        h = torch.nn.functional.conv2d(x, c1.weight.double().cpu(), c1.bias.double().cpu(), padding=1)
        h = torch.nn.functional.prelu(h, a.weight.double().cpu())
        h = torch.nn.functional.conv2d(h, c2.weight.double().cpu(), c2.bias.double().cpu(), padding=1)
        x = h * b.res_scale + x
    return x


ok = True
CASES = [((1, 64, 64, 64), 1), ((2, 64, 37, 41), 2), ((1, 64, 5, 3), 1), ((3, 64, 72, 72), 4), ((2, 64, 40, 250), 2), ((1, 64, 33, 128), 1), ((1, 64, 256, 256), 4)]
if len(sys.argv) > 2 and sys.argv[2] == "timeline":
    CASES = []
for (shape, nb) in CASES:
    blocks = [ResBlock(64).to(dev).eval() for _ in range(nb)]
    for b in blocks:
        b.body[1].weight.data.fill_(0.1 + 0.3 * torch.rand(1).item())
    x = torch.randn(*shape, device=dev)
    with torch.no_grad():
        y = resblocks_forward(blocks, x, mode)
        torch.cuda.synchronize()
        torch.backends.cudnn.allow_tf32 = False
        yt = x
        for b in blocks:
            yt = b.body(yt).mul(b.res_scale) + yt
        torch.backends.cudnn.allow_tf32 = True
        y32 = x
        for b in blocks:
            y32 = b.body(y32).mul(b.res_scale) + y32
    r = ref64(blocks, x)
    den = r.abs().max().item()
    e = (y.double().cpu() - r).abs().max().item() / den
    et = (yt.double().cpu() - r).abs().max().item() / den
    e32 = (y32.double().cpu() - r).abs().max().item() / den
    good = e <= 5e-6 and bool(torch.isfinite(y).all())
    ok &= good
    print(f"{shape} x{nb} mode {mode}: rel err chain {e:.2e}  (cuDNN fp32 {et:.2e}, cuDNN tf32 {e32:.2e})  {'ok' if good else 'FAIL'}", flush=True)

if len(sys.argv) > 2 and sys.argv[2] == "timeline":
    import ctypes
    from dagl_b200 import _lib
    L = _lib.lib()
    for shape in [(1, 64, 256, 256), (64, 64, 72, 72)]:
        blocks = [ResBlock(64).to(dev).eval() for _ in range(2)]
        x = torch.randn(*shape, device=dev)
        with torch.no_grad():
            for _ in range(3): resblocks_forward(blocks, x, mode)
            acc = None
            for _ in range(10):
                L.dagl_profile_enable(2)
                resblocks_forward(blocks, x, mode)
                buf = (ctypes.c_float * 256)()
                n = L.dagl_profile_read(buf, 256)
                L.dagl_profile_enable(0)
                v = [buf[i] * 1e3 for i in range(n)]
                acc = v if acc is None else [a + b for a, b in zip(acc, v)]
        print(os.path.basename(_lib.LIB_PATH), shape, mode, "per-launch us (pack, conv1, conv2, conv1, conv2):", " ".join(f"{t / 10:.1f}" for t in acc), flush=True)
    sys.exit(0)
if len(sys.argv) > 2:
    for shape in [(1, 64, 256, 256), (64, 64, 72, 72), (1, 64, 64, 64), (1, 64, 512, 512)]:
        blocks = [ResBlock(64).to(dev).eval() for _ in range(4)]
        def seq(t):                                      # torch modules (cuDNN), not the fused ResBlock.forward
            for b in blocks:
                t = b.body(t).mul(b.res_scale) + t
            return t
        x = torch.randn(*shape, device=dev)
        with torch.no_grad():
            for name, fn in (("chain " + mode, lambda: resblocks_forward(blocks, x, mode)), ("torch/cuDNN tf32", lambda: seq(x))):
                for _ in range(3): fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20): fn()
                e1.record(); torch.cuda.synchronize()
                print(f"{shape} 4 ResBlocks {name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
print("RESBLOCK_CHECK", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
