"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + warp-stall breakdown + hottest SASS lines.
usage: python profiles/ncu_summary.py gpurun_out/<name>.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__shared_mem_per_block_dynamic",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("== kernel:", d.get("Kernel Name", "?")[:90])
    for w in WANT:
        if w in d:
            print("  %-72s %s %s" % (w, d[w], units[hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0
    tot = sum(f(r, "# Samples") for r in data) or 1.0
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((sum(f(r, k) for r in data), k) for k in reasons), reverse=True)
    print("== warp stall sampling (all samples = %d)" % tot)
    for v, k in agg[:8]:
        print("  %-28s %5.1f%%" % (k, 100 * v / tot))
    print("== hottest instructions")
    for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top_n]:
        rs = sorted(((f(r, k), k) for k in reasons), reverse=True)[:2]
        print("  %5.1f%% inst=%10.0f thr=%4.1f | %-58s | %s" % (100 * f(r, "# Samples") / tot, f(r, "Instructions Executed"),
              f(r, "Avg. Threads Executed"), r[ix["Source"]][:58], ", ".join("%s %.0f" % (k[6:], v) for v, k in rs)))
