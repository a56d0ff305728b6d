"""Warm, in-pipeline per-launch timeline of one CE.forward (dagl_profile_enable(2): an event after every launch).
ncu's launch list is cold-cache and serialised; this one keeps L2 state and back-to-back launches (development aid)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from dagl_b200 import _lib
from oracle import ce_oracle as O

NAMES_TC = ["absmax_img", "pack_b+gamma_beta", "featmap_tc(+G/theta img)", "gather_qpatch", "embed_tc<K>(+tiles)", "kbar",
            "embed_tc<Q>(+tiles,thr)", "rowmax_tc", "rowmax_exact", "attend_tc4", "fold_partials"]
NAMES_TC_HYBRID = NAMES_TC[:10] + ["attend_tc2 tail (concurrent: its mark = end of both)", "fold_partials"]
NAMES_STAGE = ["absmax_img", "gamma_beta_heads", "pack_b", "featmap_tc(+G/theta img)", "gather_qpatch", "embed_tc<K>(+tiles)", "kbar",
               "embed_tc<Q>(+tiles,thr)", "rowmax_tc", "rowmax_exact", "attend_tc4", "fold_partials"]
HEADS = int(os.environ.get("HEADS", "1"))        # > 1: one CES stage call (heads as a grid dimension)
dev = torch.device("cuda:0")
H = W = int(os.environ.get("HW", "256"))
B = int(os.environ.get("B", "1"))
params = O.init_ce_params(1000)
x = torch.randn(B, 64, H, W, generator=torch.Generator().manual_seed(2000)).to(dev)
ce = dagl_b200.CE(in_channels=64, impl=os.environ.get("IMPL", "auto")); ce.load_state_dict(params); ce = ce.to(dev).eval()
if HEADS > 1:
    from dagl_b200.ce import stage_heads_forward
    heads = [ce]
    for h in range(1, HEADS):
        c = dagl_b200.CE(in_channels=64, impl=os.environ.get("IMPL", "auto")); c.load_state_dict(O.init_ce_params(1000 + h)); heads.append(c.to(dev).eval())
    run = lambda t: stage_heads_forward(heads, t)
else:
    run = ce
L = _lib.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(3): run(x)
    acc = None
    reps = 10
    for _ in range(reps):
        flush.zero_()
        L.dagl_profile_enable(2)
        run(x)
        buf = (ctypes.c_float * 256)()
        n = L.dagl_profile_read(buf, 256)
        L.dagl_profile_enable(0)
        v = [buf[i] * 1e3 for i in range(n)]
        acc = v if acc is None else [a + b for a, b in zip(acc, v)]
names = (NAMES_TC if len(acc) == len(NAMES_TC) and HEADS == 1 else NAMES_TC_HYBRID if len(acc) == len(NAMES_TC_HYBRID) and HEADS == 1 else NAMES_STAGE if len(acc) == len(NAMES_STAGE) and HEADS > 1 else
         [f"launch {i}" for i in range(len(acc))])
tot = 0.0
for nm, t in zip(names, acc):
    print(f"{nm:26s} {t / reps:8.1f} us")
    tot += t / reps
print(f"{'TOTAL':16s} {tot:8.1f} us   ({B}x64x{H}x{W}, heads {HEADS}, impl {ce.last_impl}, L2 flushed before each forward, mean of {reps})")
