"""CE.forward timing over the shapes of BASELINE.json's configs (development aid): direct 64^2 / 256^2 / 512^2 and the
chop-leaf batches (64 x 72^2, 256 x 76^2)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from oracle import ce_oracle as O
dev = torch.device("cuda:0")
params = O.init_ce_params(1000)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for shape in [(1, 64, 64, 64), (32, 64, 64, 64), (1, 64, 256, 256), (4, 64, 256, 256), (64, 64, 72, 72), (256, 64, 76, 76), (1, 64, 512, 512)]:
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).to(dev)
    ce = dagl_b200.CE(in_channels=64); ce.load_state_dict(params); ce = ce.to(dev).eval()
    with torch.no_grad():
        for _ in range(2): y = ce(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); y = ce(x); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
    B, _, H, W = shape
    nq, nk = ((H + 3) // 4) * ((W + 3) // 4), H * W
    ms = sorted(ts)[len(ts) // 2]
    print(f"{str(shape):22s} impl {ce.last_impl:4s} {ms:8.3f} ms  {B * nq / ms * 1e3 / 1e6:7.2f} M patches/s  "
          f"{B * 1960.0 * nq * nk / ms / 1e9:7.1f} TFLOP/s alg.  finite={bool(torch.isfinite(y).all())}", flush=True)
