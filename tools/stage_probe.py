"""Warm per-stage device timings of the TC pipeline for a few batch sizes (run on the GPU box)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from facerecognition_infrenceengine_b200 import _native as N
from oracle import synth

n, d, k = 1_000_000, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 5
store = frg.GalleryStore(dim=d, capacity=n)
store.fill_synthetic(n, 0, 1234)
m = frg.Matcher(store)
for F in (1, 64, 128, 1024):
    Q = torch.from_numpy(synth.queries(F, n, d)[0]).cuda()
    out = None
    for _ in range(5):
        out = m.match_device(Q, k, 0.45, out=out)
    torch.cuda.synchronize()
    N.profile_enable(True)
    N.profile_collect()
    steps = 30
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.match_device(Q, k, 0.45, out=out)
    e1.record()
    torch.cuda.synchronize()
    N.profile_collect()
    st = {a: round(b / steps * 1e3, 1) for a, b in N.profile_stages().items()}
    N.profile_enable(False)
    print("F=%4d k=%d step=%.1f us (with events)  stages(us)=%s  sum=%.1f" % (
        F, k, e0.elapsed_time(e1) / steps * 1e3, st, sum(st.values())), flush=True)
