"""Text summary of an .ncu-rep (run here, no GPU needed): per captured kernel the duration, DRAM traffic, L2 hit rate,
issue / tensor / XU pipe utilisation, registers, and the top stall reasons from the source page.
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/r1_x_ncu.txt"""
import csv, subprocess, sys, io
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, data = rows[0], rows[1], rows[2:]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
print(f"# ncu summary of {rep}  (ncu --set full --clock-control none; caches flushed per replay by ncu)")
for r in data:
    print("\nkernel:", r[h.index("Kernel Name")])
    for k in keys:
        if k in h:
            i = h.index(k)
            print(f"  {k:72s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kernel = None
hdr = None
agg = Counter()
def flush():
    if kernel and agg:
        tot = sum(agg.values())
        print(f"\nstall reasons (warp samples) for {kernel[:90]}:")
        for k, v in agg.most_common(8):
            print(f"  {k:24s} {v:8d}  {100.0 * v / tot:5.1f}%")
for r in rows:
    if r and r[0] == "Kernel Name":
        flush()
        kernel = r[1]; agg = Counter(); hdr = None
        continue
    if r and r[0] == "Address":
        hdr = r
        cols = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr and len(r) == len(hdr):
        for i, name in cols:
            if r[i] not in ("", "0"):
                agg[name] += int(r[i])
flush()
