"""tc vs tc4 at 320^2 .. 512^2 (development aid: where the auto dispatch switches kernels)."""
import sys, torch
sys.path.insert(0, ".")
import dagl_b200
from oracle import ce_oracle as O
dev = torch.device("cuda:0")
params = O.init_ce_params(1000)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for hw in (320, 384, 448, 512):
    x = torch.randn(1, 64, hw, hw, generator=torch.Generator().manual_seed(1)).to(dev)
    res = {}
    for impl in ("tc", "tc4"):
        ce = dagl_b200.CE(in_channels=64, impl=impl); ce.load_state_dict(params); ce = ce.to(dev).eval()
        with torch.no_grad():
            for _ in range(2): y = ce(x)
            ts = []
            for _ in range(3):
                flush.zero_(); torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); y = ce(x); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        res[impl] = (sorted(ts)[1], y)
    err = (res["tc"][1] - res["tc4"][1]).abs().max().item() / res["tc"][1].abs().max().item()
    print(hw, f"tc {res['tc'][0]:.3f} ms  tc4 {res['tc4'][0]:.3f} ms  rel diff {err:.1e}", flush=True)
