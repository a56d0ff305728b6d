#!/usr/bin/env python
"""bench.py — headline benchmark of the dynamic attentive graph block on B200.

Contract (see task brief):  python bench.py --gpus N --steps K --warmup W
prints ONE JSON line on rank 0.  A "step" is one graph-block forward
(``CE.forward``, DN_Gray/model/dagl.py:207-275) on one 64-channel 256x256
feature map per GPU — the shape BASELINE.json's metric is quoted on
("patches/sec + ms/graph-block fwd, 64ch 256x256").

  value        query patches/s (B*Nq / t) with the input resident in HBM, all GPUs
  e2e          the same through the host-buffer C-ABI entry (pinned host input,
               H2D + forward + D2H of the result inside the timed region)
  roofline     dominant fused graph kernel: algorithmic FLOPs / measured kernel
               time against the measured bf16 tensor peak (the path is a dense
               contraction pair) + algorithmic HBM GB/s for context
  cpu_baseline the CPU oracle port of the reference timed on this box's cores
  configs      (N = 1) the other shapes of the path: cfg1 64^2, the reference's chop batches, 512^2, the CES caller
  strong       (N > 1) ONE image query-sharded over the N GPUs with an NCCL all-gather: ms, speed-up, in-run error

``--impl reference`` times the oracle port (reference op order, torch CPU,
all host threads) on the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 256
C_IN = 64
B_PER_GPU = 1
NQ = ((H + 3) // 4) * ((W + 3) // 4)
NK = H * W
FLOPS_ALG = 2.0 * NQ * NK * (196 + 784)                       # 526.1 GFLOP / image
BYTES_ALG = 4.0 * (NQ * 196 + NK * 196 + 196 + 2 * NQ + 32 * NK)  # 63.0 MB / image


def workload_config():
    """The `config` object: identical in the product arm and in the reference arm (same workload, same per-GPU unit)."""
    return {"workload": f"CE.forward {B_PER_GPU}x{C_IN}x{H}x{W} per GPU (one graph block, direct/no-chop), random-init head",
            "Nq": NQ, "Nk": NK,
            "l2": "GPU arm: L2 flushed between timed iterations (256 MiB memset); CPU arm: n/a",
            "timing": "GPU arm: CUDA events per step, summed, max over ranks; CPU arm: perf_counter around the timed steps"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained"), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


CE_SHAPES = {"g.weight": (16, C_IN, 3, 3), "g.bias": (16,), "W.weight": (C_IN, 16, 1, 1), "W.bias": (C_IN,),
             "theta.weight": (16, C_IN, 1, 1), "theta.bias": (16,), "fc1.0.weight": (196, 784), "fc1.0.bias": (196,),
             "fc2.0.weight": (196, 784), "fc2.0.bias": (196,), "thr_conv.weight": (1, C_IN, 7, 7), "thr_conv.bias": (1,),
             "bias_conv.weight": (1, C_IN, 7, 7), "bias_conv.bias": (1,)}


def workload_tensors(seed: int):
    """Synthetic, seeded: random-init head (torch's default Conv2d / Linear init statistics: weight and bias
    U(-1/sqrt(fan_in), 1/sqrt(fan_in))) and a randn feature map.  Self-contained: the product arm does not touch oracle/."""
    import math
    gen = torch.Generator().manual_seed(1000 + seed)
    params = {}
    for name, shape in CE_SHAPES.items():
        wshape = CE_SHAPES[name.rsplit(".", 1)[0] + ".weight"]
        bound = 1.0 / math.sqrt(math.prod(wshape[1:]))
        params[name] = ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * bound).float()
    gen = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(B_PER_GPU, C_IN, H, W, generator=gen)
    return params, x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = None
        self.gpu = gpu_index

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    @staticmethod
    def _epoch(ts: str):
        import datetime
        try:
            return datetime.datetime.strptime(ts.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0: float = None, t1: float = None):
        """Summarise the samples taken between wall-clock t0 and t1 (time.time())."""
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) < 9:
                    continue
                ts = self._epoch(f[0])
                if t0 is not None and ts is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def _oracle_pass(O, params, x, nq_sample):
    """One pass of the reference path on the CPU: the whole graph block when ``nq_sample`` is None, else the
    prologue + both embeddings + the score / neighbour-selection / aggregation rows of the first ``nq_sample`` queries
    (rows are independent, dagl.py:250-264), in the reference's op order."""
    if nq_sample is None:
        return O.ce_forward(params, x)
    import torch.nn.functional as F
    G, Th, gamma, beta, qp, kp, vp, fold_pad = O._prologue(params, x)
    Q = F.relu(F.linear(qp[0].t()[:nq_sample], params["fc1.0.weight"], params["fc1.0.bias"]))
    K = F.relu(F.linear(kp[0].t(), params["fc2.0.weight"], params["fc2.0.bias"]))
    S = torch.matmul(Q, K.t())
    mu = S.mean(dim=1)
    P, _ = O._edge_weights(S, mu, gamma[0, :nq_sample], beta[0, :nq_sample])
    return torch.mm(P, vp[0].t())


def run_reference_arm(args):
    """The reference's own CPU implementation of the path (oracle port, reference op order), all host threads.
    A step is the full workload when K steps of it fit in ~3 minutes, else a bounded sample of its query rows."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ce_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params, x = workload_tensors(0)
    with torch.no_grad():
        t0 = time.perf_counter()
        O.ce_forward(params, x)                                   # probe: one full pass (also warms the allocator)
        probe = time.perf_counter() - t0
        budget = 180.0 / max(1, args.steps + args.warmup)
        nq_s = None if probe <= budget else max(128, int(NQ * budget / probe) // 64 * 64)
        for _ in range(args.warmup):
            _oracle_pass(O, params, x, nq_s)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _oracle_pass(O, params, x, nq_s)
        dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    units = B_PER_GPU * (NQ if nq_s is None else nq_s)
    value = units / (ms * 1e-3)
    sample = ("full workload per step" if nq_s is None else
              f"bounded sample per step: prologue + embeddings of the whole image and the graph rows of the first {nq_s} of {NQ} "
              f"query patches (full pass measured at {probe*1e3:.0f} ms)")
    out = {
        "impl": "reference", "metric": "graph-block query patches/s (64ch 256x256)", "value": value,
        "unit": "patches/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample}; {args.steps} steps after {args.warmup} warm-up, oracle port (reference op order, torch CPU)"},
        "e2e": {"value": value, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def cpu_baseline_sample():
    from oracle import ce_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params, x = workload_tensors(0)
    with torch.no_grad():
        O.ce_forward(params, x)                       # warm-up
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            O.ce_forward(params, x)
            ts.append(time.perf_counter() - t0)
    best = min(ts)
    return {"value": B_PER_GPU * NQ / best, "unit": "patches/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"full workload (1x{C_IN}x{H}x{W}), 1 warm-up + best of 3, oracle.ce_forward (reference op order, torch CPU); {best*1e3:.0f} ms",
            "ms_per_step": best * 1e3}


def time_call(fn, iters, flush, warm=2):
    """Mean ms of fn() over `iters` calls, CUDA events per call, L2 flushed (untimed) before each."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def other_configs(ce, dev, flush, peaks):
    """The other shapes the path really runs at (BASELINE cfg1; the reference's chop leaves, model/__init__.py:179-214;
    512^2 direct; the CES caller row), timed the same way: reported next to the headline, not as bench lines."""
    import dagl_b200
    out = []
    gen = torch.Generator().manual_seed(77)

    def shape_line(name, B, Hh, Ww, iters, fn=None, heads=1):
        x = torch.randn(B, C_IN, Hh, Ww, generator=gen).to(dev)
        f = (lambda: fn(x)) if fn else (lambda: ce(x))
        with torch.no_grad():
            ms = time_call(f, iters, flush)
        nq, nk = ((Hh + 3) // 4) * ((Ww + 3) // 4), Hh * Ww
        flops = heads * B * 2.0 * nq * nk * 980
        out.append({"workload": name, "ms": ms, "patches_per_s": heads * B * nq / (ms * 1e-3),
                    "tflops_alg": flops / (ms * 1e-3) / 1e12, "frac_of_bf16_peak": flops / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"]})

    shape_line("cfg1: CE.forward 1x64x64x64", 1, 64, 64, 50)
    # the same forward replayed from a caller-side CUDA graph (11 dependent launches: the eager number is launch-latency-bound)
    try:
        xg1 = torch.randn(1, C_IN, 64, 64, generator=gen).to(dev)
        with torch.no_grad():
            side1 = torch.cuda.Stream()
            side1.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side1):
                for _ in range(2):
                    ce(xg1)
            torch.cuda.current_stream().wait_stream(side1)
            graph1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph1, stream=side1):
                ce(xg1)
        shape_line("cfg1: CE.forward 1x64x64x64, replay of a caller-side CUDA graph", 1, 64, 64, 50, fn=lambda _x: graph1.replay())
    except Exception as e:                                            # a bench extra, never fatal
        out.append({"workload": "cfg1: CE.forward 1x64x64x64, CUDA graph", "error": str(e)[:200]})
    shape_line("chop leaves of 256^2: CE.forward 64x64x72x72", 64, 72, 72, 10)
    shape_line("chop leaves of 512^2: CE.forward 256x64x76x76", 256, 76, 76, 5)
    shape_line("cfg3/cfg5 head: CE.forward 1x64x512x512", 1, 512, 512, 5)
    torch.manual_seed(5)
    ces = dagl_b200.CES(in_channels=C_IN).to(dev).eval()
    # CES = twelve heads (3 stage calls) + RBS1 / RBS2 (two chains of four ResBlocks on the tcgen05 convolution kernel) + three
    # 1x1 merge convolutions (cuDNN)
    shape_line("CES.forward 1x64x64x64 (12 heads as 3 stage calls + 2 ResBlock chain calls)", 1, 64, 64, 20, fn=ces, heads=12)
    # the same forward captured in a CUDA graph by the caller (the library allocates nothing and never synchronises)
    try:
        xg = torch.randn(1, C_IN, 64, 64, generator=gen).to(dev)
        with torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    ces(xg)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                ces(xg)
        shape_line("CES.forward 1x64x64x64, replay of a caller-side CUDA graph", 1, 64, 64, 20, fn=lambda _x: graph.replay(), heads=12)
    except Exception as e:                                            # a bench extra, never fatal
        out.append({"workload": "CES.forward 1x64x64x64, CUDA graph", "error": str(e)[:200]})
    shape_line("CES.forward 64x64x72x72 (chop batch)", 64, 72, 72, 3, fn=ces, heads=12)
    shape_line("CES.forward 1x64x256x256 (direct)", 1, 256, 256, 5, fn=ces, heads=12)

    # the callers either side of the graph blocks: a chain of four ResBlocks (CES.RBS1; common.py:59-79), this repo's tcgen05
    # convolution (fp32-accurate) next to torch / cuDNN under its TF32 default and with TF32 off (what "fp32" means to cuDNN)
    def chain_line(B, Hh, Ww, iters):
        x = torch.randn(B, C_IN, Hh, Ww, generator=gen).to(dev)
        seq = ces.RBS1
        def plain(t):                                                             # the reference's ResBlock.forward (common.py:75-79) on cuDNN
            for blk in seq:
                t = blk.body(t).mul(blk.res_scale) + t
            return t
        rec = {"workload": f"4-ResBlock chain {B}x64x{Hh}x{Ww} (8 convolutions 64->64 3x3 + PReLU + residual)"}
        with torch.no_grad():
            rec["ms"] = time_call(lambda: seq(x), iters, flush)
            rec["ms_cudnn_tf32"] = time_call(lambda: plain(x), iters, flush)
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            try:
                rec["ms_cudnn_fp32"] = time_call(lambda: plain(x), iters, flush)
                ref = plain(x)
            finally:
                torch.backends.cudnn.allow_tf32 = prev
            if not os.environ.get("DAGL_BENCH_NOPOISON"):
                for _ in range(4):                                                # poison the allocator's free blocks of this size:
                    torch.empty_like(x).fill_(float("nan"))                      # an output that is not written would show
            y = seq(x)
            d_async = (y - ref).abs().max() / ref.abs().max()                    # consumed by torch at once, no synchronisation
            rec["impl"] = dagl_b200._lib.lib().dagl_last_impl().decode()
            torch.cuda.synchronize()
            rec["rel_err_vs_cudnn_fp32"] = float((y - ref).abs().max() / ref.abs().max())
            rec["rel_err_unsynchronised"] = float(d_async)
        flops = 8 * 2.0 * B * Hh * Ww * 64 * 64 * 9
        rec["tflops_alg"] = flops / (rec["ms"] * 1e-3) / 1e12
        out.append(rec)

    chain_line(1, 256, 256, 10)
    chain_line(64, 72, 72, 5)
    # the whole network (dagl.py:10-54; random init): head conv, 8 ResBlocks, CES, 8 ResBlocks, conv, tail conv, global skip
    try:
        torch.manual_seed(6)
        rr = dagl_b200.RR().to(dev).eval()
        xi = torch.rand(1, 1, 256, 256, generator=gen).to(dev)
        with torch.no_grad():
            ms = time_call(lambda: rr(xi), 5, flush)
        out.append({"workload": "RR.forward 1x1x256x256 (BASELINE cfg2 shape, direct; 12 heads + 24 ResBlocks + 3 plain convolutions)", "ms": ms,
                    "images_per_s": 1.0 / (ms * 1e-3)})
    except Exception as e:                                            # a bench extra, never fatal
        out.append({"workload": "RR.forward 1x1x256x256", "error": str(e)[:200]})
    return out


def strong_scaling(ce, dev, flush, world, rank, params):
    """ONE image sharded over all ranks by query tiles (CE.forward_query_sharded: redundant prologue, fused graph stage
    on this rank's tiles, one NCCL all-gather of the aggregation rows, fold): ms (max over ranks), speed-up against the
    same image on one GPU, and the in-run difference between the sharded and the single-GPU result."""
    import torch.distributed as dist
    import dagl_b200
    from dagl_b200 import parallel
    rec = []
    # every rank must hold the SAME head here (the weak-scaling arm above gives each rank its own image and head)
    params0, _ = workload_tensors(0)
    ce = dagl_b200.CE(in_channels=C_IN, impl=ce.impl)
    ce.load_state_dict(params0)
    ce = ce.to(dev).eval()
    for size, iters in ((256, 20), (512, 5)):
        gen = torch.Generator().manual_seed(4242 + size)                 # the same image on every rank
        x = torch.randn(1, C_IN, size, size, generator=gen).to(dev)
        with torch.no_grad():
            y1 = ce(x)
            ys = ce.forward_query_sharded(x)
            torch.cuda.synchronize()
            err = float((ys - y1).abs().max() / y1.abs().max())
            dist.barrier()
            ms_single = time_call(lambda: ce(x), iters, flush)
            dist.barrier()
            ms_shard = time_call(lambda: ce.forward_query_sharded(x), iters, flush)
        ms_shard = parallel.max_over_ranks(ms_shard, dev)
        ms_single = parallel.max_over_ranks(ms_single, dev)
        err = parallel.max_over_ranks(err, dev)
        nq = (size // 4) ** 2
        rec.append({"workload": f"CE.forward 1x{C_IN}x{size}x{size}, ONE image over {world} GPUs (query-tile sharding + all-gather of rows)",
                    "ms_sharded": ms_shard, "ms_single_gpu": ms_single, "speedup": ms_single / ms_shard,
                    "patches_per_s": nq / (ms_shard * 1e-3), "rel_err_vs_single_gpu": err,
                    "allgather_bytes": nq * 784 * 4})
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="auto", choices=["auto", "simt", "tc", "tc4", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other-shapes table (N = 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist
    import dagl_b200
    from dagl_b200 import _lib, parallel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    L = _lib.lib()
    params, x_host = workload_tensors(rank)
    ce = dagl_b200.CE(in_channels=C_IN, impl=args.impl)
    ce.load_state_dict(params)
    ce = ce.to(dev).eval()
    x_dev = x_host.to(dev)
    x_pin = x_host.pin_memory()
    y_pin = torch.empty(B_PER_GPU, 16, H, W).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value) --------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        for _ in range(args.warmup):
            ce(x_dev)
        barrier()
        L.dagl_profile_enable(1)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        launches = 0
        barrier()
        t_epoch0 = time.time()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()                              # L2 flush between timed iterations (untimed)
            starts[i].record()
            ce(x_dev)
            stops[i].record()
            launches += ce.last_launches
            launches_per_step = ce.last_launches
        barrier()
        t_wall = time.perf_counter() - t_wall0
        t_epoch1 = time.time()
        kbuf = (ctypes.c_float * 256)()
        nk = L.dagl_profile_read(kbuf, 256)
        L.dagl_profile_enable(0)
        clocks = sampler.stop(t_epoch0, t_epoch1) if rank == 0 else None
        step_ms = [s.elapsed_time(e) for s, e in zip(starts, stops)]
        total_ms = sum(step_ms)
        impl_used = ce.last_impl

        # ---- end-to-end through the host-buffer entry (e2e) -----------------------
        for _ in range(2):
            ce.forward_host(x_pin, y_pin, device=dev)
        barrier()
        e_starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        e_stops = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.zero_()
            e_starts[i].record()
            ce.forward_host(x_pin, y_pin, device=dev, sync=False)
            e_stops[i].record()
        barrier()
        e2e_total_ms = sum(s.elapsed_time(e) for s, e in zip(e_starts, e_stops))
        checksum = float(y_pin.double().abs().sum())

        # ---- the same requests through the streaming host entry (CE.host_pipeline: the H2D copy of request i+1 and the D2H
        # copy of result i-1 overlap the kernels of request i).  Reported NEXT to e2e, not as e2e: one region around all K
        # requests (they overlap by design), every request with its own host->device and device->host copy
        pipe_rec = None
        try:
            pipe = ce.host_pipeline(B_PER_GPU, H, W, depth=2, device=dev)
            x_pins = [x_pin, x_pin.clone().pin_memory()]
            y_pins = [torch.empty_like(y_pin).pin_memory() for _ in range(2)]
            for i in range(4):
                pipe.submit(x_pins[i & 1], y_pins[i & 1])
            pipe.drain()
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                pipe.submit(x_pins[i & 1], y_pins[i & 1])
            pipe.drain()
            t_pipe = time.perf_counter() - t0
            same = bool(torch.equal(y_pins[0], y_pin)) and bool(torch.equal(y_pins[1], y_pin))
            pipe_rec = {"ms_per_step": t_pipe / args.steps * 1e3, "depth": 2, "results_equal_serial_entry": same,
                        "timing": "host wall clock around all K requests (submit ... drain), L2 not flushed"}
        except Exception as e:                                            # an extra, never fatal
            pipe_rec = {"error": str(e)[:200]}

    # ---- beside the headline: other shapes (N = 1), single-image strong scaling (N > 1) ----
    peaks0 = measured_peaks()
    extra = other_configs(ce, dev, flush, peaks0) if (world == 1 and not args.no_extra) else None
    strong = strong_scaling(ce, dev, flush, world, rank, params) if world > 1 else None

    total_ms = parallel.max_over_ranks(total_ms, dev)
    e2e_total_ms = parallel.max_over_ranks(e2e_total_ms, dev)
    launches_all = int(parallel.sum_over_ranks(launches, dev))
    ms_per_step = total_ms / args.steps
    patches_per_step = world * B_PER_GPU * NQ
    value = patches_per_step / (ms_per_step * 1e-3)
    e2e_value = patches_per_step / (e2e_total_ms / args.steps * 1e-3)

    if rank == 0:
        peaks = measured_peaks()
        kern_ms = [kbuf[i] for i in range(nk)]
        k_avg = sum(kern_ms) / len(kern_ms) if kern_ms else float("nan")
        achieved_tflops = B_PER_GPU * FLOPS_ALG / (k_avg * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                with open(tpath) as f:
                    traffic = json.load(f).get(impl_used, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "bound": "tensor", "achieved": achieved_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved_tflops / peaks["bf16_tflops"], "traffic": traffic,
            "peak_source": f"{peaks['source']} bf16 dense burst (MEASURED_PEAKS.json)",
            "kernel": {"tc": "attend_tc2_kernel", "tc4": "attend_tc4_kernel", "simt": "attend_simt_kernel"}.get(impl_used, impl_used), "kernel_ms": k_avg, "kernel_share_of_step": k_avg / ms_per_step,
            "kernel_note": ("the graph stage = attend_tc4_kernel on the first ~94 % of the keys + attend_tc2_kernel on the rest, launched as its "
                            "programmatic dependent and running concurrently on the SMs the 4-CTA clusters cannot use; kernel_ms spans both, "
                            "flops_per_launch is their sum (DAGL_HYBRID=0: the 4-CTA kernel alone)") if (impl_used == "tc4" and launches_per_step == 12) else None,
            "flops_per_launch": B_PER_GPU * FLOPS_ALG, "bytes_per_launch": B_PER_GPU * BYTES_ALG,
            "hbm_gbs_algorithmic": B_PER_GPU * BYTES_ALG / (k_avg * 1e-3) / 1e9,
            "hbm_frac_of_measured": B_PER_GPU * BYTES_ALG / (k_avg * 1e-3) / 1e9 / peaks["hbm_gbs"],
        }
        out = {
            "metric": "graph-block query patches/s (64ch 256x256)", "value": value, "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 io; split-fp16 x3 MMAs (fp32-accurate) for feature maps / embeddings / scores, fp16 P.V, fp32 accumulate"
                     if impl_used != "simt" else "f32",
            "data": "synthetic",
            "config": workload_config(), "impl_used": impl_used,
            "ms_per_graph_block": ms_per_step, "pairs_per_s": world * B_PER_GPU * NQ * NK / (ms_per_step * 1e-3),
            "wall_s_timed_region": t_wall,
            "e2e": {"value": e2e_value, "unit": "patches/s", "h2d_bytes_per_step": B_PER_GPU * C_IN * H * W * 4,
                    "d2h_bytes_per_step": B_PER_GPU * 16 * H * W * 4, "ms_per_step": e2e_total_ms / args.steps,
                    "api": "CE.forward_host -> dagl_ce_forward_host_f32 (pinned host buffers)", "checksum": checksum,
                    "pipelined": pipe_rec},
            "gpu_launches": launches_all, "roofline": roofline, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_sample()
        if world == 1 and not args.no_extra:
            out["configs"] = extra
        if strong is not None:
            out["strong"] = strong
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
