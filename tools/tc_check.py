"""Quick on-GPU comparison of the tc / simt kernels against the CPU oracle (development aid)."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from oracle import ce_oracle as O

dev = torch.device("cuda:0")
params = O.init_ce_params(123)
shapes = [(1, 64, 7, 9), (1, 64, 16, 16), (2, 64, 30, 41), (1, 64, 64, 64), (1, 64, 72, 72), (1, 64, 128, 128)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for shape in shapes:
    gen = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=gen)
    if shape[2] * shape[3] <= 72 * 72:
        yref, aux = O.ce_forward(params, x, return_aux=True)
        mask_ref = aux["mask"]
    else:
        yref = O.ce_forward_chunked(params, x, chunk=256); mask_ref = None
    for impl in ("tc", "tc4"):
        ce = dagl_b200.CE(in_channels=64, impl=impl); ce.load_state_dict(params); ce = ce.to(dev).eval()
        print(f"{shape} {impl}: launching", flush=True)
        with torch.no_grad():
            y, bits, nnz = ce.forward_debug(x.to(dev))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(3): ce(x.to(dev))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        err = (y.cpu() - yref).abs().max().item() / yref.abs().max().item()
        flips = -1
        if mask_ref is not None:
            flips = int((O.unpack_mask_bits(bits.cpu(), shape[2] * shape[3]) != mask_ref).sum())
        print(f"{shape} {impl}: rel_err={err:.3e} flips={flips} nan={int(torch.isnan(y).sum())} {dt*1e3:.2f} ms/fwd", flush=True)
