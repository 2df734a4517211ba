"""ncu evidence for the dispatch table (north_star: "ncu counters - tensor-pipe utilisation, achieved HBM GB/s -
justify each variant at each batch size").  Run on the GPU box under

    ncu --clock-control none --print-units base --csv --log-file gpurun_out/dispatch.csv \
        --kernel-name regex:'scan_f32_kernel|tc_scan_kernel|first_match_kernel' \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed \
        python tools/dispatch_probe.py

One match per (variant, batch) against the 1 M x 512 gallery; every case is preceded by one first_match launch,
which serves as the separator `tools/ncu_summary.py dispatch` uses to attribute kernels to cases."""
import sys

import numpy as np

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from oracle import synth

CASES = [(v, F) for F in (1, 8, 64, 128, 256, 1024) for v in ("scan_f32", "tc_exact") if not (v == "scan_f32" and F > 64)]

if __name__ == "__main__":
    n, d, k = 1_000_000, 512, 5
    store = frg.GalleryStore(dim=d, capacity=n)
    store.fill_synthetic(n, 0, synth.GALLERY_SEED)
    m = frg.Matcher(store)
    for variant, F in CASES:
        Q, _ = synth.queries(F, n, d)
        m.first_above(Q[:1], 2.0)                      # separator launch (first_match_kernel)
        r = m.match(Q, k, 0.45, variant=variant, with_ids=False)
        print("case %-9s F=%-5d launches=%d" % (variant, F, r.launches), flush=True)
    store.close()
