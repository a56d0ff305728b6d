import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from oracle import ce_oracle as O
dev = torch.device("cuda:0")
params = O.init_ce_params(123)
for shape in [(1, 64, 7, 9), (1, 64, 40, 36)]:
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape)))
    for scale in (0.0, 1.0, 10.0):
        O.SOFTMAX_SCALE = scale
        yref = O.ce_forward(params, x)
        for impl in ("tc1", "tc"):
            ce = dagl_b200.CE(in_channels=64, impl=impl, softmax_scale=scale); ce.load_state_dict(params); ce = ce.to(dev).eval()
            with torch.no_grad(): y = ce(x.to(dev))
            print(shape, "scale", scale, impl, "rel_err %.3e" % ((y.cpu() - yref).abs().max().item() / yref.abs().max().item()))
