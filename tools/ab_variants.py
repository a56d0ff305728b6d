"""A/B timing of compile-time variants of the library (development aid).
  build (here, no GPU):  python tools/ab_variants.py build name1="-DV4_PDEPTH=3 ..." name2="..."
  run (on the GPU box):  python tools/ab_variants.py run [HW] [B]
Variants are built into dagl_b200/variants/lib<name>.so (git-ignored, travels with the gpurun snapshot)."""
import ctypes, glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, "dagl_b200", "variants")

def build(specs):
    os.makedirs(VDIR, exist_ok=True)
    for spec in specs:
        name, flags = spec.split("=", 1)
        env = dict(os.environ, DAGL_B200_LIB=os.path.join(VDIR, f"lib{name}.so"), DAGL_NVCC_EXTRA=flags)
        r = subprocess.run([sys.executable, "-m", "dagl_b200.build", "--force"], cwd=ROOT, env=env, capture_output=True, text=True)
        print(name, flags, "->", "ok" if r.returncode == 0 else r.stderr[-2000:])

def run_one():
    import torch
    sys.path.insert(0, ROOT)
    import dagl_b200
    from dagl_b200 import _lib
    from oracle import ce_oracle as O
    HW, B = int(os.environ.get("HW", "256")), int(os.environ.get("B", "1"))
    dev = torch.device("cuda:0")
    params = O.init_ce_params(1000)
    x = torch.randn(B, 64, HW, HW, generator=torch.Generator().manual_seed(2000)).to(dev)
    ce = dagl_b200.CE(in_channels=64); ce.load_state_dict(params); ce = ce.to(dev).eval()
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3): y = ce(x)
        torch.cuda.synchronize()
        L.dagl_profile_enable(1)
        tot = 0.0
        n = 20
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); y = ce(x); b.record(); torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        buf = (ctypes.c_float * 256)()
        k = L.dagl_profile_read(buf, 256)
        L.dagl_profile_enable(0)
    kms = sum(buf[i] for i in range(k)) / max(k, 1)
    print(f"{os.path.basename(_lib.LIB_PATH):28s} {B}x64x{HW}x{HW}: forward {tot / n * 1e3:8.1f} us, graph kernel {kms * 1e3:8.1f} us, "
          f"checksum {float(y.double().abs().sum()):.6f} finite={bool(torch.isfinite(y).all())}", flush=True)

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "run":
        libs = [os.path.join(ROOT, "dagl_b200", "libdagl_b200.so")] + sorted(glob.glob(os.path.join(VDIR, "lib*.so")))
        for lib in libs:
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, DAGL_B200_LIB=lib), timeout=90)
    else:
        run_one()
