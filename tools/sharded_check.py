"""Multi-GPU check of single-image query sharding (run under torchrun with 2+ ranks): the sharded forward
(CE.forward_query_sharded: dagl_ce_forward_rows_f32 on this rank's query tiles + ONE NCCL all-gather of the aggregation
rows + dagl_ce_fold_rows_f32) must reproduce the CPU ORACLE (reference math), and the single-GPU forward."""
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
# 64^2: 2 query tiles; 72x60 batch 2: 3 tiles (uneven over 2 ranks, batch > 1 staging path); 100x80: 4 tiles with a ragged
# last one; 256^2: the north-star shape (chunked oracle)
for shape in [(1, 64, 64, 64), (2, 64, 72, 60), (1, 64, 100, 80), (1, 64, 256, 256)]:
    x_cpu = torch.randn(*shape, generator=torch.Generator().manual_seed(9))
    x = x_cpu.to(dev)
    ce = dagl_b200.CE(in_channels=64); ce.load_state_dict(params); ce = ce.to(dev).eval()
    with torch.no_grad():
        y1 = ce(x)
        ys = ce.forward_query_sharded(x)
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        for _ in range(5): ce.forward_query_sharded(x)
        torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for _ in range(5): ce(x)
        torch.cuda.synchronize(); dt1 = (time.perf_counter() - t0) / 5
        yref = O.ce_forward_chunked(params, x_cpu, chunk=256) if rank == 0 else None
    err = (ys - y1).abs().max().item() / y1.abs().max().item()
    ok &= err <= 1e-5                      # same kernels and operands; only the key-split factor (fp32 summation order) differs
    if rank == 0:
        err_o = (ys.cpu() - yref).abs().max().item() / yref.abs().max().item()
        ok &= err_o <= 1e-3
        print(f"{shape}: sharded vs ORACLE rel_err={err_o:.2e}, vs single-GPU {err:.2e}; sharded {dt*1e3:.2f} ms  single {dt1*1e3:.2f} ms  "
              f"(world {world}, impl {ce.last_impl})", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_CHECK", "OK" if int(flag.item()) else "FAIL", flush=True)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
