"""Turns the ncu artefacts brought back in gpurun_out/ into the small text summaries kept under
profiles/.  Usage: python profiles/summarize.py <launches.csv> <prof.ncu-rep> <out-prefix>"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, out):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, data = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("unnamed>::", "")
        v = float(d["Metric Value"].replace(",", ""))
        v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : %d launches, %.1f us total\n" % (len(data), tot))
        f.write("# (cold-cache, serialised: compare SHARES, not absolutes)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-48s launches=%5d  time_us=%12.1f  share=%5.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))


WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    h, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on  (%s)\n" % path)
        for r in rows[2:]:
            for w in WANT:
                if w in h:
                    i = h.index(w)
                    f.write("%-80s %s %s\n" % (w, r[i], units[i]))
            f.write("\n")


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[3] + "_launches.txt")
    full(sys.argv[2], sys.argv[3] + "_top_kernel.txt")
