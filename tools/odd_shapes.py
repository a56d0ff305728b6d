"""Cross-implementation agreement (simt / tc / tc4) on odd shapes: narrow, tiny, ragged, batched (development aid; the\nparity tests proper are in tests/test_ce_gpu.py)."""
import sys, torch
sys.path.insert(0, ".")
import dagl_b200
from oracle import ce_oracle as O
dev = torch.device("cuda:0")
params = O.init_ce_params(31)
ok = True
for shape in [(3, 64, 9, 200), (1, 64, 200, 9), (5, 64, 33, 31), (1, 64, 128, 130), (2, 64, 8, 8), (1, 64, 7, 7), (1, 64, 97, 101), (16, 64, 24, 24)]:
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape))).to(dev)
    ys = {}
    for impl in ("simt", "tc", "tc4"):
        ce = dagl_b200.CE(in_channels=64, impl=impl); ce.load_state_dict(params); ce = ce.to(dev).eval()
        with torch.no_grad():
            ys[impl] = ce(x)
    torch.cuda.synchronize()
    d = ys["simt"].abs().max().item()
    e2 = (ys["tc"] - ys["simt"]).abs().max().item() / d
    e4 = (ys["tc4"] - ys["simt"]).abs().max().item() / d
    good = e2 <= 1e-3 and e4 <= 1e-3 and bool(torch.isfinite(ys["tc4"]).all())
    ok &= good
    print(shape, f"tc vs simt {e2:.2e}  tc4 vs simt {e4:.2e}", "OK" if good else "FAIL", flush=True)
print("ODD_SHAPES", "OK" if ok else "FAIL")
