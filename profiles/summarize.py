"""Digest ncu outputs brought back in gpurun_out/ into the small text summaries kept in profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv  > profiles/rN_X_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_X.ncu-rep  > profiles/rN_X_kernel.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ci = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0][:70]
        v = float(r[ci["Metric Value"]]) * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1}[r[ci["Metric Unit"]]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms total (ncu: cold-cache, serialised; compare shares)")
    print(f"{'kernel':72s} {'count':>6s} {'ms':>10s} {'share':>7s}")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {c:6d} {v:10.3f} {100 * v / tot:6.1f}%")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:85s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        print()


def table(path):
    """One line per kernel (mean over the captured launches): duration, DRAM traffic and achieved GB/s,
    pipe utilisation -- the figures DESIGN.md and bench.py's roofline.traffic quote."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        v = float(r[ci[k]].replace(",", "") or 0)
        u = units[ci[k]]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1,
                 "ns": 1e-6, "nsecond": 1e-6, "second": 1e3}.get(u, 1)
        return v * scale

    agg = collections.OrderedDict()
    cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size"]
    for r in rows[2:]:
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")[:58]
        a = agg.setdefault(name, [0] + [0.0] * len(cols))
        a[0] += 1
        for j, k in enumerate(cols):
            a[j + 1] += val(r, k)
    print(f"# {path}: ncu --set full --clock-control none, mean over the captured launches of each kernel")
    print(f"{'kernel':58s} {'n':>2s} {'ms':>7s} {'rd MB':>7s} {'wr MB':>7s} {'GB/s':>6s} {'dram%':>6s} {'sm%':>5s} {'issue%':>6s} "
          f"{'tensor%':>7s} {'fp64%':>6s} {'warps%':>6s} {'regs':>4s} {'grid':>6s} {'blk':>4s}")
    for name, a in agg.items():
        n = a[0]
        m = [x / n for x in a[1:]]
        gbs = (m[1] + m[2]) / (m[0] * 1e-3) / 1e9 if m[0] > 0 else 0
        print(f"{name:58s} {n:2d} {m[0]:7.3f} {m[1] / 1e6:7.1f} {m[2] / 1e6:7.1f} {gbs:6.0f} {m[3]:6.1f} {m[4]:5.1f} {m[5]:6.1f} "
              f"{m[6]:7.1f} {m[7]:6.1f} {m[8]:6.1f} {m[9]:4.0f} {m[10]:6.0f} {m[11]:4.0f}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "table": table}[sys.argv[1]](sys.argv[2])
