import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import dagl_b200
dev = torch.device("cuda:0")
torch.manual_seed(5)
ces = dagl_b200.CES(in_channels=64).to(dev).eval()
for B, HW in ((1, 64), (4, 72)):
    x = torch.randn(B, 64, HW, HW, device=dev)
    with torch.no_grad():
        for _ in range(3): y0 = ces(x)
        torch.cuda.synchronize()
        def t(fn, n=20):
            for _ in range(3): fn()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(n): fn()
            torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
        eager = t(lambda: ces(x))
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): ces(x)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            yg = ces(x)
        graphed = t(lambda: g.replay())
        g.replay(); torch.cuda.synchronize()
        print(f"CES {B}x64x{HW}x{HW}: eager {eager:.3f} ms, CUDA graph {graphed:.3f} ms, max|diff| {float((yg - y0).abs().max()):.2e}")
