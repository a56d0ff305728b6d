"""2-GPU check of single-image query sharding (run under torchrun with 2+ ranks):
the sharded forward must reproduce the single-GPU forward (same kernels, only the key-split factor differs)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dagl_b200
from oracle import ce_oracle as O

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
params = O.init_ce_params(5)
ok = True
for shape in [(1, 64, 64, 64), (2, 64, 72, 60), (1, 64, 256, 256)]:
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(9)).to(dev)
    ce = dagl_b200.CE(in_channels=64, impl="tc"); ce.load_state_dict(params); ce = ce.to(dev).eval()
    with torch.no_grad():
        y1 = ce(x)
        ys = ce.forward_query_sharded(x)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        for _ in range(5): ce.forward_query_sharded(x)
        torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(5): ce(x)
        torch.cuda.synchronize(); dt1 = (time.perf_counter() - t0) / 5
    err = (ys - y1).abs().max().item() / y1.abs().max().item()
    ok &= err <= 1e-3
    if rank == 0:
        print(f"{shape}: sharded vs single rel_err={err:.2e}  sharded {dt*1e3:.2f} ms  single {dt1*1e3:.2f} ms  (world {world})", flush=True)
if rank == 0:
    print("SHARDED_CHECK", "OK" if ok else "FAIL", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
