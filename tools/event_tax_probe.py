"""How much do the per-stage profiling events cost inside a timed region?  (run on the GPU box)"""
import sys

import torch

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from facerecognition_infrenceengine_b200 import _native as N
from oracle import synth

n, d, k = 1_000_000, 512, 5
store = frg.GalleryStore(dim=d, capacity=n)
store.fill_synthetic(n, 0, 1234)
m = frg.Matcher(store)
for F in (1, 64, 128, 256, 1024):
    Q = torch.from_numpy(synth.queries(F, n, d)[0]).cuda()
    out = None
    for _ in range(5):
        out = m.match_device(Q, k, 0.45, out=out)
    torch.cuda.synchronize()
    res = {}
    for mode in (0, 1, 2, 0, 1, 2):
        N.check(N.lib.frg_profile_enable(mode))
        N.profile_collect()
        steps = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            m.match_device(Q, k, 0.45, out=out)
        e1.record()
        torch.cuda.synchronize()
        N.profile_collect()
        res.setdefault(mode, []).append(round(e0.elapsed_time(e1) / steps * 1e3, 1))
    N.check(N.lib.frg_profile_enable(0))
    print("F=%4d  step us: no events %s | all stages %s | dominant only %s" % (F, res[0], res[1], res.get(2)), flush=True)
