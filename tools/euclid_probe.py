"""BASELINE config 3 on the GPU box: 10 M x 128 raw rows, Euclidean, batch 256, top-1.
Times the tensor-core path (filter over the augmented bf16 plane + exact rescoring) with its per-stage
split, checks it bit for bit against the exact fp32 scan on the same store, and reports both against
the HBM roofline of SURVEY.md section 8d (fp32 gallery bytes: 5.12 GB -> 0.782 ms at 6546.6 GB/s).
    python tools/euclid_probe.py [rows] [batch] [k]
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from facerecognition_infrenceengine_b200 import _native as N
from oracle import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
d = 128
peaks = json.load(open("MEASURED_PEAKS.json")) if len(sys.argv) < 5 else {}
hbm = 6546.6

store = frg.GalleryStore(dim=d, capacity=n, raw=True)          # raw + scan plane = Euclidean plane
store.fill_synthetic(n, 0, 99)
m = frg.Matcher(store, metric="euclidean")
Qh, tgt = synth.queries(F, n, d, seed=5, gallery_seed=99)
Q = torch.from_numpy(Qh).cuda()
out = {"rows": n, "dim": d, "batch": F, "k": k}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(variant, steps, warm):
    res = None
    for _ in range(warm):
        res = m.match_device(Q, k, 0.6, variant=variant, out=res)
    torch.cuda.synchronize()
    N.profile_enable(True)
    N.profile_collect()
    e0.record()
    for _ in range(steps):
        res = m.match_device(Q, k, 0.6, variant=variant, out=res)
    e1.record()
    torch.cuda.synchronize()
    N.profile_collect()
    st = {a: round(b / steps, 4) for a, b in N.profile_stages().items()}
    N.profile_enable(False)
    return e0.elapsed_time(e1) / steps, st, [t.cpu().numpy() for t in res]


ms, stages, tc = timed("tc_exact", 20, 5)
roof_f32 = n * d * 4 / (hbm * 1e9) * 1e3
roof_plane = n * (d + 64) * 2 / (hbm * 1e9) * 1e3
out["tc_exact"] = {"ms_per_batch": ms, "queries_per_s": F / ms * 1e3, "stage_ms": stages,
                   "roofline_ms_fp32_gallery_bytes": roof_f32, "frac_of_fp32_roofline": roof_f32 / ms,
                   "plane_bytes_roofline_ms": roof_plane, "filter_frac_of_plane_roofline": roof_plane / max(stages.get("dominant", ms), 1e-9)}
print(json.dumps(out["tc_exact"]), flush=True)
ms2, _, sc = timed("scan_f32", 2, 1)
out["scan_f32"] = {"ms_per_batch": ms2, "queries_per_s": F / ms2 * 1e3}
out["tc_equals_scan_bit_for_bit"] = bool(np.array_equal(tc[0], sc[0]) and
                                         np.array_equal(tc[1].view(np.uint32), sc[1].view(np.uint32)) and
                                         np.array_equal(tc[2], sc[2]))
hit = tgt >= 0
out["genuine_found"] = bool((tc[0][hit, 0] == tgt[hit]).all())
out["accepts"] = int(tc[2].sum())
print(json.dumps(out), flush=True)
json.dump(out, open("gpurun_out/euclid_probe.json", "w"), indent=1)
