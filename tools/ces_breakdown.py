"""Where the time of a CES forward goes (development aid): stage calls vs cuDNN ResBlocks vs 1x1 merges."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from dagl_b200.ce import stage_heads_forward
dev = torch.device("cuda:0")
B, HW = int(os.environ.get("B", "1")), int(os.environ.get("HW", "256"))
torch.manual_seed(5)
ces = dagl_b200.CES(in_channels=64).to(dev).eval()
x = torch.randn(B, 64, HW, HW, device=dev)

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - w0) / n * 1e3

with torch.no_grad():
    heads = [getattr(ces, f"c1_{h}") for h in (1, 2, 3, 4)]
    print("stage call (4 heads)      gpu %.3f ms  wall %.3f ms" % t(lambda: stage_heads_forward(heads, x)))
    print("one head                  gpu %.3f ms  wall %.3f ms" % t(lambda: heads[0](x)))
    print("RBS1 (4 ResBlocks)        gpu %.3f ms  wall %.3f ms" % t(lambda: ces.RBS1(x)))
    print("one 3x3 conv 64->64       gpu %.3f ms  wall %.3f ms" % t(lambda: ces.RBS1[0].body[0](x)))
    print("1x1 merge conv            gpu %.3f ms  wall %.3f ms" % t(lambda: ces.c1_c(x)))
    print("whole CES                 gpu %.3f ms  wall %.3f ms" % t(lambda: ces(x)))
    print("cudnn.allow_tf32", torch.backends.cudnn.allow_tf32, "benchmark", torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    print("RBS1, cudnn.benchmark     gpu %.3f ms  wall %.3f ms" % t(lambda: ces.RBS1(x)))
    print("whole CES, benchmark      gpu %.3f ms  wall %.3f ms" % t(lambda: ces(x)))
