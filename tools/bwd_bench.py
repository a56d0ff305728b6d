"""Training-step timing of one CE head at the reference's training shape (32 x 64 x 64 x 64, trainer.py:51-57):
forward + backward through the CUDA path vs autograd through the all-torch recompute (development aid)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from dagl_b200.autograd import ce_recompute
from oracle import ce_oracle as O
dev = torch.device("cuda:0")
B = int(os.environ.get("B", "32")); HW = int(os.environ.get("HW", "64"))
params = O.init_ce_params(3)
ce = dagl_b200.CE(in_channels=64); ce.load_state_dict(params); ce = ce.to(dev).train()
x = torch.randn(B, 64, HW, HW, device=dev, requires_grad=True)
w = torch.randn(B, 16, HW, HW, device=dev)

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def step_cuda():
    x.grad = None; ce.zero_grad(set_to_none=True)
    (ce(x) * w).sum().backward()

ORDER = ["g.weight", "g.bias", "theta.weight", "theta.bias", "fc1.0.weight", "fc1.0.bias", "fc2.0.weight", "fc2.0.bias",
         "thr_conv.weight", "thr_conv.bias", "bias_conv.weight", "bias_conv.bias"]
leaves = [dict(ce.named_parameters())[k] for k in ORDER]
def step_torch():
    x.grad = None; ce.zero_grad(set_to_none=True)
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    (ce_recompute(x, leaves) * w).sum().backward()
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev

with torch.no_grad():
    fwd = t(lambda: ce(x))
print(f"{B}x64x{HW}x{HW}: forward (CUDA path) {fwd:.3f} ms")
print(f"  forward + backward, CUDA graph-stage backward kernels + torch prologue autograd: {t(step_cuda):.3f} ms")
print(f"  forward + backward, all-torch recompute (fp32 matmuls, the round-1 backward):   {t(step_torch):.3f} ms")

# breakdown of the CUDA-path backward
from dagl_b200.autograd import ce_prologue, graph_stage_backward
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
xs = x.detach().requires_grad_(True)
def pro():
    return ce_prologue(xs, leaves)
mids = pro()
dmids = graph_stage_backward(*[m.detach() for m in mids], w, 10.0)
print(f"  breakdown: torch prologue forward {t(pro):.3f} ms; CUDA graph-stage backward {t(lambda: graph_stage_backward(*[m.detach() for m in mids], w, 10.0)):.3f} ms; "
      f"torch autograd through the prologue {t(lambda: torch.autograd.grad(list(pro()), [xs] + leaves, dmids, allow_unused=True)):.3f} ms (incl. its forward)")
