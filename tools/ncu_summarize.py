"""Summaries of the ncu captures for profiles/ (run here, after gpurun brought the files back).
  python tools/ncu_summarize.py launches gpurun_out/<tag>_launches.csv  <forwards in the capture>
  python tools/ncu_summarize.py full     gpurun_out/<tag>_prof_attend.ncu-rep"""
import csv, subprocess, sys, collections, io

def launches(path, nfwd):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
    tot = collections.OrderedDict(); cnt = collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv: continue
        name = r[ik].split("(")[0]
        v = float(r[iv].replace(",", "")); u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        tot[name] = tot.get(name, 0.0) + us; cnt[name] += 1
    total = 0.0
    for name, us in tot.items():
        if not name.startswith("dagl::") and "dagl::" not in name: continue
        per = us / cnt[name]
        print(f"{name:58s} launches/forward {cnt[name] / nfwd:4.1f} {per:9.1f} us")
        if cnt[name] >= nfwd: total += per * (cnt[name] / nfwd)
    print(f"{'TOTAL (kernels launched once per forward or more)':76s} {total:9.1f} us")

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out))); hdr, units = rows[0], rows[1]
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]
    for n, r in enumerate(rows[2:]):
        print(f"---- launch {n + 1}: {r[hdr.index('Kernel Name')]}")
        for k in keys:
            if k in hdr:
                i = hdr.index(k); print(f"  {k:72s} {r[i]:>18s} {units[i]}")

if __name__ == "__main__":
    if sys.argv[1] == "launches": launches(sys.argv[2], int(sys.argv[3]))
    else: full(sys.argv[2])
