"""Summaries of ncu captures for profiles/ (run here, no GPU needed):

    python tools/ncu_summary.py launches gpurun_out/r2_y_launches.csv profiles/r2_c3_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/r2_y_null2_c3.ncu-rep profiles/r2_c3_null2_kernel.txt c3 <launches per step>

`kernel` also updates profiles/k2_traffic.json (DRAM bytes per step of the K2 launches), which bench.py reads for
roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        if any(t in r[ik] for t in ("native::", "at_cuda", "<unnamed>", "at::")):
            continue  # the synthetic-data generator's torch kernels are not part of the step
        name = r[ik].split("(")[0].replace("void ", "").replace("dfb::", "")
        t = float(r[iv].replace(",", "")) / 1e6  # ns -> ms
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        total += t
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)\n")
        f.write("# source: %s; %d launches, %.3f ms in total\n" % (os.path.basename(src), sum(a[0] for a in agg.values()), total))
        f.write("%-72s %8s %12s %8s\n" % ("kernel", "launches", "ms", "share"))
        for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %8d %12.3f %7.1f%%\n" % (name[:72], c, t, 100 * t / total))
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def kernel(src, dst, workload=None, launches_per_step=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, source: %s\n" % os.path.basename(src))
        traffic = []
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write("\nkernel: %s\n" % d.get("Kernel Name", "?")[:160])
            for h, u in zip(hdr, units):
                if h in WANT:
                    f.write("  %-82s %-12s %s\n" % (h, u, d[h]))
            def val(k):
                i = hdr.index(k)
                x = float(d[k].replace(",", ""))
                return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
            traffic.append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
    print(open(dst).read())
    if workload:
        p = os.path.join(ROOT, "profiles", "k2_traffic.json")
        tj = json.load(open(p)) if os.path.exists(p) else {}
        n = float(launches_per_step or 1)
        tj[workload] = {"dram_bytes_per_launch": sum(traffic) / len(traffic), "dram_bytes_per_step": sum(traffic) / len(traffic) * n,
                        "k2_launches_per_step": n, "source": os.path.relpath(dst, ROOT)}
        json.dump(tj, open(p, "w"), indent=1)
        print("updated", p, tj[workload])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(*sys.argv[2:6])
