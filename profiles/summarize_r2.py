"""profiles/r2_*: summaries of the round-2 ncu captures brought back by profiles/collect.sh.
    python profiles/summarize_r2.py gpurun_out r2
Reads <dir>/<tag>_conv_raw.csv (ncu --set full of 60 consecutive conv_tc_kernel launches = one whole eps evaluation of the
bench's 10,000-row pass), <tag>_aux_raw.csv (step / swept / kuka kernels), <tag>_launches.csv (launch list) and a per-shape
launch sequence (cls,R,Hout,N,C1,C2,taps,flags rows of one evaluation, from a `make diag` PMP_PROFILE_DUMP run) to put the
algorithmic bytes beside the measured DRAM bytes."""
import collections
import csv
import json
import re
import sys

import summarize

import os
HERE = os.path.dirname(os.path.abspath(__file__))
d, tag = sys.argv[1], sys.argv[2]
ROWS = 10000      # rows per launch in the bench (5,000 plans x CFG pair per family)


def load(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    h, u = rows[0], rows[1]
    return h, u, rows[2:]


def load_sections(path):
    """(header, units, row) of every kernel of a file that is several `ncu --page raw --csv` reports back to back"""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    out, h, u = [], None, None
    i = 0
    while i < len(rows):
        if "Kernel Name" in rows[i] and "ID" in rows[i]:
            h, u = rows[i], rows[i + 1]
            i += 2
            continue
        out.append((h, u, rows[i]))
        i += 1
    return out


def val(h, u, r, name):
    i = h.index(name)
    v = float(r[i].replace(",", "") or 0)
    unit = u[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "%": 1.0}.get(unit, 1.0)
    return v * scale


# launch sequence of one evaluation (shapes), scaled to the bench's rows
seq = []
try:
    rows = list(csv.DictReader(open("%s/r2j_dump_ng1.csv" % d)))[:60]
    for r in rows:
        seq.append({k: int(r[k]) for k in ("R", "Hout", "N", "C1", "C2", "taps", "flags")})
except Exception:
    seq = []


def algo_bytes(s):
    """bytes one launch has to move: BF16 hi+lo planes (4 B/element) of its inputs and of its result, the fp32 yhat it
    saves (forward GroupNorm) or re-reads (backward), the identity planes it adds; weights are negligible"""
    f = s["flags"]
    gnf, gnb, ident, s_in, s_out = f & 1, (f >> 1) & 1, (f >> 2) & 1, (f >> 3) & 1, (f >> 4) & 1
    hin = s["Hout"] * (2 if s_in else 1) // (2 if s_out else 1)
    n = ROWS * s["Hout"] * s["N"]
    b = 4.0 * ROWS * hin * s["C1"] + 4.0 * ROWS * s["Hout"] * s["C2"]
    b += 4.0 * n * (1 + gnf + gnb + ident)
    return b


h, u, data = load("%s/%s_conv_raw.csv" % (d, tag))
out = open("%s/%s_conv_kernels.txt" % (HERE, tag), "w")
out.write("# ncu --set full --clock-control none, 60 consecutive conv_tc_kernel launches = ONE eps evaluation (U-Net forward +\n"
          "# analytic backward) of the bench step (`python bench.py`: 10,000 rows per launch, Maze2D Static2 model), launch order.\n"
          "# algo_MB: BF16 hi+lo planes of the inputs and of the result, fp32 yhat saved / re-read, identity planes (see\n"
          "# profiles/summarize_r2.py); backward launches that also keep the raw gradient planes write one more plane pair.\n")
out.write("%-3s %-44s %9s %9s %9s %8s %7s %8s %8s %5s  %s\n" % ("#", "kernel", "time_us", "dram_rd", "dram_wr", "algo_MB", "ratio",
                                                             "dram_%", "tensor_%", "regs", "shape R,H,N,C1,C2,taps,flags"))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
per = []
for i, r in enumerate(data):
    nm = re.search(r"conv_tc_kernel<[^>]*>", r[h.index("Kernel Name")]).group(0).replace("(int)", "").replace("(bool)", "")
    t = val(h, u, r, "gpu__time_duration.sum")
    rd, wr = val(h, u, r, "dram__bytes_read.sum"), val(h, u, r, "dram__bytes_write.sum")
    dp = val(h, u, r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    tp = val(h, u, r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    regs = r[h.index("launch__registers_per_thread")]
    ab = algo_bytes(seq[i]) if i < len(seq) else 0.0
    shape = ",".join(str(seq[i][k]) for k in ("R", "Hout", "N", "C1", "C2", "taps", "flags")).replace("7400", str(ROWS)) if i < len(seq) else ""
    out.write("%-3d %-44s %9.1f %9.1f %9.1f %8.1f %7.2f %8.1f %8.1f %5s  %s\n" % (i, nm, t, rd / 1e6, wr / 1e6, ab / 1e6,
                                                                              (rd + wr) / ab if ab else 0, dp, tp, regs, shape))
    a = agg[nm]
    a[0] += 1; a[1] += t; a[2] += rd + wr; a[3] += ab; a[4] += dp * t; a[5] += tp * t
    per.append(dict(kernel=nm, us=t, dram=rd + wr, algo=ab, dram_pct=dp, tensor_pct=tp))
tot = sum(a[1] for a in agg.values())
out.write("\n# per template instance (time-weighted means)\n")
for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.write("%-44s launches=%2d time_us=%9.1f share=%5.1f%% dram_MB/launch=%8.1f algo_MB/launch=%8.1f ratio=%5.2f dram_%%=%5.1f tensor_%%=%5.1f\n" % (
        nm, a[0], a[1], 100 * a[1] / tot, a[2] / a[0] / 1e6, a[3] / a[0] / 1e6, a[2] / a[3] if a[3] else 0, a[4] / a[1], a[5] / a[1]))
out.write("# whole evaluation: %.1f us, DRAM %.1f GB, algorithmic %.1f GB\n" % (tot, sum(a[2] for a in agg.values()) / 1e9,
                                                                                 sum(a[3] for a in agg.values()) / 1e9))
out.close()

# traffic of the dominant kernel class (what bench.py's roofline.traffic quotes) and of the 64-channel forward kernel
dom = max(agg.items(), key=lambda kv: kv[1][1])[0]
c64 = [p for p in per if p["kernel"].startswith("conv_tc_kernel<64, 8, 0")]
json.dump({
    "source": "ncu --set full --clock-control none of `python bench.py --steps 1 --warmup 1 --no-graph` (10,000 rows per launch, "
              "the bench's own pass size), launches 120-179 of conv_tc_kernel = one eps evaluation; profiles/%s_conv_kernels.txt" % tag,
    "kernel": dom,
    "dram_bytes_per_launch": agg[dom][2] / agg[dom][0],
    "algorithmic_bytes_per_launch": agg[dom][3] / agg[dom][0],
    "duration_us": [p["us"] for p in per if p["kernel"] == dom],
    "tensor_pipe_active_pct": [p["tensor_pct"] for p in per if p["kernel"] == dom],
    "kernel_64ch": {"launches": len(c64), "dram_bytes_per_launch": sum(p["dram"] for p in c64) / max(len(c64), 1),
                    "algorithmic_bytes_per_launch": sum(p["algo"] for p in c64) / max(len(c64), 1),
                    "dram_throughput_pct": [p["dram_pct"] for p in c64], "tensor_pipe_active_pct": [p["tensor_pct"] for p in c64],
                    "duration_us": [p["us"] for p in c64]},
    "whole_evaluation": {"dram_bytes": sum(a[2] for a in agg.values()), "algorithmic_bytes": sum(a[3] for a in agg.values()),
                         "time_us": tot},
}, open("%s/%s_traffic.json" % (HERE, tag), "w"), indent=1)

# HBM-bound kernels of the path
data = load_sections("%s/%s_aux_raw.csv" % (d, tag))
with open("%s/%s_aux_kernels.txt" % (HERE, tag), "w") as f:
    f.write("# ncu --set full --clock-control none of tests/gpu_profile_aux.py: fused sampler update (400,000 plans), Maze2D swept\n"
            "# check (1,000,000 trajectories, 10 candidates per environment), Kuka sphere-link check (200,000 trajectories)\n")
    names = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
             "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
    for h, u, r in data:
        nm = re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("void ", "").replace("pmp::<unnamed>::", "")
        f.write("%s\n" % nm)
        for n in names:
            if n in h:
                f.write("    %-72s %s %s\n" % (n, r[h.index(n)], u[h.index(n)]))
        t = val(h, u, r, "gpu__time_duration.sum")
        by = val(h, u, r, "dram__bytes_read.sum") + val(h, u, r, "dram__bytes_write.sum")
        f.write("    %-72s %.1f GB/s\n\n" % ("achieved DRAM bandwidth (bytes / duration)", by / (t * 1e-6) / 1e9))

summarize.launches("%s/%s_launches.csv" % (d, tag), "%s/%s_launches.txt" % (HERE, tag))


def train_kernels(path, out):
    """table of an `ncu --set full` capture of training-step kernels (raw csv)"""
    h, u, data = load(path)
    names = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none of `python bench.py --workload dualkuka_train` (batch 256, H = 56), one window of\n"
                "# the reverse pass: time_us, grid, registers, DRAM %, SM %, tensor-pipe %, DRAM MB read / written\n")
        for r in data:
            nm = re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("void ", "").replace("pmp::<unnamed>::", "")
            f.write("%-36s %9.1f %6s %4s %6.1f %6.1f %6.1f %9.1f %9.1f\n" % (
                nm[:36], val(h, u, r, names[0]), r[h.index(names[1])], r[h.index(names[2])], val(h, u, r, names[3]),
                val(h, u, r, names[4]), val(h, u, r, names[5]), val(h, u, r, names[6]) / 1e6, val(h, u, r, names[7]) / 1e6))


if os.path.exists("%s/%s_train_raw.csv" % (d, tag)):
    train_kernels("%s/%s_train_raw.csv" % (d, tag), "%s/%s_train_kernels.txt" % (HERE, tag))
if os.path.exists("%s/%s_train_launches.csv" % (d, tag)):
    summarize.launches("%s/%s_train_launches.csv" % (d, tag), "%s/%s_train_launches.txt" % (HERE, tag))
