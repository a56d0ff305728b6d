"""One small graph-block forward per implementation, for compute-sanitizer (development aid):
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from oracle import ce_oracle as O

dev = torch.device("cuda:0")
params = O.init_ce_params(123)
x = torch.randn(2, 64, 40, 36, generator=torch.Generator().manual_seed(321))
yref = O.ce_forward(params, x)
for impl in sys.argv[1:] or ["simt", "tc", "tc4"]:
    ce = dagl_b200.CE(in_channels=64, impl=impl); ce.load_state_dict(params); ce = ce.to(dev).eval()
    with torch.no_grad():
        y = ce(x.to(dev)); yd, bits, nnz = ce.forward_debug(x.to(dev))
    torch.cuda.synchronize()
    err = (y.cpu() - yref).abs().max().item() / yref.abs().max().item()
    print(f"sanitize[{impl}]: rel_err={err:.3e} debug_equal={bool(torch.equal(y, yd))}", flush=True)
