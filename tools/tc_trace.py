"""Build a -DDAGL_TC_TRACE variant of the library and print per-role wait-cycle counters of the
tensor-core graph kernel for the bench shape (development aid; not part of the product)."""
import ctypes, os, subprocess, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dagl_b200 import build as B
lib_trace = os.path.join(ROOT, "dagl_b200", "libdagl_b200_trace.so")
if "--build" in sys.argv or not os.path.exists(lib_trace):
    cmd = ["nvcc"] + B.NVCC_FLAGS + ["-DDAGL_TC_TRACE", "-o", lib_trace] + B.sources()
    subprocess.check_call(cmd)
    if "--build" in sys.argv:
        sys.exit(0)
from dagl_b200 import _lib
_lib.LIB_PATH = lib_trace
import dagl_b200
from oracle import ce_oracle as O
dev = torch.device("cuda:0")
H = W = int(os.environ.get("HW", "256"))
params = O.init_ce_params(1000)
x = torch.randn(1, 64, H, W, generator=torch.Generator().manual_seed(2000)).to(dev)
ce = dagl_b200.CE(in_channels=64, impl=os.environ.get("IMPL", "tc")); ce.load_state_dict(params); ce = ce.to(dev).eval()
L = _lib.lib()
for mode in [int(m) for m in os.environ.get('MODES', '0,1,2,3').split(',')]:
  L.dagl_debug_set_tc_mode(mode)
  print(f"=== dbg mode {mode} (bit0: no S MMAs, bit1: no P.V MMAs, 4: tiny P forwards, 8: tiny theta loads, 16: tiny K loads) ===")
  with torch.no_grad():
    for _ in range(3): ce(x)
  torch.cuda.synchronize()
  buf = np.zeros((1024, 24), dtype=np.int64)
  rc = L.dagl_debug_read_tc_trace(buf.ctypes.data_as(ctypes.c_void_p))
  assert rc == 0
  n = int((buf[:, 12] > 0).sum())
  b = buf[:n]
  nt = b[:, 12].astype(float)
  print(f"{n} CTAs, tiles/CTA mean {nt.mean():.1f}")
  def per_tile(col): return (b[:, col] / nt)
  print("per tile (cycles), mean over CTAs:")
  print(f"  producer : wait k_empty {per_tile(0).mean():8.0f}  wait t_empty {per_tile(1).mean():8.0f}  total {per_tile(3).mean():8.0f}")
  print(f"  mma      : wait k_full(+s_free) {per_tile(4).mean():8.0f}  wait p_full  {per_tile(5).mean():8.0f}  wait t_full {per_tile(6).mean():8.0f}  total {per_tile(7).mean():8.0f}")
  print(f"  softmax  : wait s_full  {per_tile(8).mean():8.0f}  named bar / p_free {per_tile(9).mean():8.0f}  total {per_tile(11).mean():8.0f}")
  print(f"  score iss: wait k_full {per_tile(16).mean():8.0f}  wait s_free {per_tile(17).mean():8.0f}  total {per_tile(19).mean():8.0f}")
  t0 = b[:, 14].min(); 
  print("kernel span cycles:", (b[:, 14] + b[:, 3]).max() - t0, " CTA total mean", b[:, 3].mean())
  if os.environ.get("TIMELINE"):
    tl = np.zeros((4, 32, 24), dtype=np.int64)
    assert L.dagl_debug_read_tc_timeline(tl.ctypes.data_as(ctypes.c_void_p)) == 0
    names = {0: "sm:s_full", 1: "sm:ld_done", 2: "sm:computed", 3: "sm:p_free", 4: "sm:stored", 5: "sc:k_full", 6: "sc:s_free",
             7: "sc:issued", 8: "pv0:p_full", 9: "pv0:issued", 10: "pv1:p_full", 11: "pv1:issued", 12: "pv2:p_full",
             13: "pv2:issued", 14: "pv3:p_full", 15: "pv3:issued", 16: "pv:t_full", 17: "fw:p_full", 18: "fw:issued",
             19: "kl:issued", 20: "tl:issued"}
    sel = [int(e) for e in os.environ.get("EVENTS", "3,4,18,8,9,10,11,12,13,14,15").split(",")]
    print("--- cluster 0, all four ranks on a common clock (cycles since the cluster barrier); rounds 100..103; event = R<rank>:r<round>:<name>")
    ev = []
    for rank in range(4):
      for r in range(0, 4):
        for e in sel:
          if tl[rank, r, e] > 0: ev.append((int(tl[rank, r, e]), f"R{rank}:r{r}:{names[e]}"))
    ev.sort()
    t0 = ev[0][0]
    line = [f"{t - t0:6d} {nm}" for t, nm in ev]
    for k in range(0, len(line), 4): print("   ".join(f"{x:28s}" for x in line[k:k + 4]))
