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


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
