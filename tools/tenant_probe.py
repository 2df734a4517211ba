"""Tenant-filtered matching (infrenceServer.py:343-380 filters EVERY frame by company) when
a company's rows sit in one contiguous block of the gallery (run on the GPU box):
    python tools/tenant_probe.py [rows_per_tenant] [far]
"far": one more row of the probed tenant sits at the very end of the gallery, so its extent spans almost everything
while its rows stay concentrated in one CTA's chunk (the private candidate segments spill into the dense list)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from oracle import matcher_oracle as mo
from oracle import synth

n, d = 1_000_000, 512
T = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
store = frg.GalleryStore(dim=d, capacity=n)
for t0 in range(0, n, T):
    store.fill_synthetic(min(T, n - t0), t0, synth.GALLERY_SEED, tag=1 + t0 // T)
store._tenants = {"c%d" % i: i for i in range(1, n // T + 2)}
m = frg.Matcher(store)
tenant = 37 if n // T > 37 else 1
lo = (tenant - 1) * T
far = len(sys.argv) > 2 and sys.argv[2] == "far"
if far:
    store.overwrite_rows([n - 1], synth.unit_rows([n - 1], d), np.array([tenant], np.int32), prenormalised=True)
G, tags = store.read_rows(lo, T)
for F in (8, 64, 256):
    rng = np.random.default_rng(F)
    rows = rng.integers(lo, lo + T, size=F)
    Q = synth.unit_rows(rows, d) + np.float32(0.03) * rng.standard_normal((F, d)).astype(np.float32)
    Q[F // 2:] = rng.standard_normal((F - F // 2, d)).astype(np.float32)
    ref = mo.match_topk(Q, G, 6, 0.4)
    Qd = torch.from_numpy(Q).cuda()
    for company in ("c%d" % tenant, None):
        r = m.match(Q, 5, 0.4, company_id=company)
        if company is not None and not far:
            want = np.where(ref[0] >= 0, ref[0] + lo, -1)
            assert mo.ids_match_with_gap(want, ref[1], r.rows, 1e-4).all(), "tenant-filtered result differs from the oracle"
        out = None
        for _ in range(5):
            out = m.match_device(Qd, 5, 0.4, company_id=company, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            m.match_device(Qd, 5, 0.4, company_id=company, out=out)
        e1.record()
        torch.cuda.synchronize()
        print("rows/tenant=%d%s F=%3d company=%-5s  %.1f us/step" % (T, " +far" if far else "", F, company, e0.elapsed_time(e1) / 20 * 1e3), flush=True)
store.close()
