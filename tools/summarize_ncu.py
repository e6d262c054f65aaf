"""Turn ncu artefacts (brought back in gpurun_out/) into the small text summaries kept in profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches_X.csv  > profiles/X_launches.txt
    python tools/summarize_ncu.py full gpurun_out/prof_X.ncu-rep      > profiles/X_full.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[1:]:
        if len(r) > vi:
            try:
                agg[r[ki]].append(float(r[vi].replace(",", "")))
            except ValueError:
                pass
    tot = sum(sum(v) for v in agg.values())
    print("# per-launch gpu__time_duration (ncu, cold cache, serialised): compare SHARES, not absolutes")
    print("%-100s %5s %12s %8s" % ("kernel", "n", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-100s %5d %12.1f %7.1f%%" % (k[:100], len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==== %s  (id %s)" % (r[idx["Kernel Name"]], r[idx["ID"]]))
        for k in KEYS:
            if k in idx:
                print("  %-80s %s %s" % (k, r[idx[k]], units[idx[k]]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
