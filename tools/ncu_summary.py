"""Condense ncu outputs brought back from the GPU box into small text summaries for profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches_<case>.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_full_<case>.txt
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(r[ki].split("(")[0][:64], float(r[vi].replace(",", "")), r[ui]) for r in rows[1:]]
    print("# every launch, in order (ncu --metrics gpu__time_duration.sum --clock-control none):")
    print("# cold-cache and serialised - compare SHARES, not absolutes")
    for name, v, u in seq:
        print("%-66s %12.1f %s" % (name, v, u))
    tot = OrderedDict()
    for name, v, _ in seq:
        c, t = tot.get(name, (0, 0.0))
        tot[name] = (c + 1, t + v)
    total = sum(t for _, t in tot.values())
    print("\n# totals by kernel")
    for name, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1]):
        print("%-66s launches=%4d total=%12.1f ns share=%5.1f%%" % (name, c, t, 100 * t / total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("# ncu --set full --clock-control none (one replayed launch per block below)")
    for r in rows[2:]:
        print("\n== %s" % r[ki])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                print("  %-66s %s %s" % (k, r[i], units[i]))


def dispatch(path):
    """CSV of `tools/dispatch_probe.py` under ncu: per (variant, batch) the dominant kernel's time, DRAM bytes,
    achieved GB/s and tensor-pipe utilisation.  first_match_kernel launches separate the cases."""
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from dispatch_probe import CASES
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ii, ki, mi, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    scale = {"nsecond": 1, "ns": 1, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6, "second": 1e9, "s": 1e9,
             "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}             # -> ns and bytes
    launches = OrderedDict()
    for r in rows[1:]:
        launches.setdefault(int(r[ii]), {"name": r[ki].split("(")[0]})[r[mi]] = (
            float(r[vi].replace(",", "")) * scale.get(r[ui], 1))
    cases, cur = [], None
    for _, L in sorted(launches.items()):
        if "first_match" in L["name"]:
            cur = []
            cases.append(cur)
        elif cur is not None:
            cur.append(L)
    print("# 1 M x 512 gallery, top-5, one match per case under ncu --clock-control none (cold cache, serialised:")
    print("# compare variants at the same batch, not absolutes).  HBM-bound cases: algorithmic bytes = 2.048 GB (fp32")
    print("# master, scan_f32: once per 4 queries) / 1.024 GB (bf16 plane, tc_exact: once per 128 queries)")
    print("%-9s %5s | %-26s %4s %10s %9s %8s %8s %8s" % ("variant", "F", "dominant kernel", "n", "total us", "GB read",
                                                        "GB/s", "dram %", "tensor %"))
    for (variant, F), ks in zip(CASES, cases):
        if not ks:
            continue
        by = OrderedDict()
        for L in ks:
            by.setdefault(L["name"], []).append(L)
        name, group = max(by.items(), key=lambda kv: sum(x["gpu__time_duration.sum"] for x in kv[1]))
        t = sum(x["gpu__time_duration.sum"] for x in group)                    # ns
        b = sum(x["dram__bytes_read.sum"] for x in group)
        dr = sum(x["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"] * x["gpu__time_duration.sum"] for x in group) / t
        tp = sum(x["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] * x["gpu__time_duration.sum"] for x in group) / t
        other = sum(x["gpu__time_duration.sum"] for L in by.values() for x in L) - t
        print("%-9s %5d | %-26s %4d %10.1f %9.3f %8.0f %8.1f %8.1f   (+%.1f us in the other scan launches)" % (
            variant, F, name[-26:], len(group), t / 1e3, b / 1e9, b / t, dr, tp, other / 1e3))


if __name__ == "__main__":
    {"launches": launches, "full": full, "dispatch": dispatch}[sys.argv[1]](sys.argv[2])
