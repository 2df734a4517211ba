"""A small tour of every kernel for compute-sanitizer (run on the GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_probe.py
(compute-sanitizer is closed on this GPU pool - the call is refused there; the tour still runs as a plain
parity pass over every kernel: python tools/sanitize_probe.py)
cosine and Euclidean TC pipelines (single-CTA and CTA-pair forms, fused probe form), tenant masks, tombstones,
candidate overflow -> exact fallback, exact scan, first_match, ingest / overwrite / compaction."""
import sys

import numpy as np

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from oracle import matcher_oracle as mo
from oracle import synth

n, d = 6000, 512
G = synth.gallery(n, d, 31)
G[np.arange(100, 3100, 2)] = G[7]                       # 1500 duplicates: overflow -> fallback
store = frg.GalleryStore(dim=d, capacity=1024)          # grows
tags = ["T%d" % (i % 3) for i in range(n)]
store.upsert(["p%d" % i for i in range(n)], G, tags, prenormalised=True)
store.remove(["p%d" % i for i in range(0, 50)])
Gd, tg = store.read_rows()
m = frg.Matcher(store)
for F in (3, 40, 200):                                   # fused probe, single CTA, CTA pair
    Q, _ = synth.queries(F, n, d, seed=8, gallery_seed=31)
    Q[1] = G[7]
    for company in (None, "T1"):
        tenant = None if company is None else store.tenant_code(company, create=False)
        ref = mo.match_topk(Q, Gd, 6, 0.4, tg, tenant)
        for variant in ("tc_exact", "scan_f32", "tc_bf16"):
            r = m.match(Q, 5, 0.4, company_id=company, variant=variant)
            if variant != "tc_bf16":
                assert mo.ids_match_with_gap(ref[0], ref[1], r.rows, 1e-4).all(), (F, company, variant)
rows, sc = m.first_above(Q[:4], 0.4)
store.compact()
m.match(Q, 5, 0.4)
store.close()

n, d = 5000, 128
rng = np.random.default_rng(1)
G = (rng.standard_normal((n, d)) * 0.1).astype(np.float32)
G[1000:2500] = G[3]
store = frg.GalleryStore(dim=d, capacity=n, raw=True)
store.upsert(["e%d" % i for i in range(n)], G)
store.upsert(["e9"], (rng.standard_normal((1, d)) * 0.3).astype(np.float32))      # in place, new norm
store.remove(["e10", "e11"])
me = frg.Matcher(store, metric="euclidean")
Gd, tg = store.read_rows()
for F in (2, 50, 150):
    Q = np.concatenate([G[[3, 77]] + np.float32(0.01), (rng.standard_normal((F - 2, d)) * 0.1).astype(np.float32)])
    ref = mo.euclidean_topk(Q, Gd, 4, 0.6, tg, None)
    for variant in ("tc_exact", "scan_f32"):
        r = me.match(Q, 3, 0.6, variant=variant)
        assert mo.ids_match_with_gap(ref[0], -ref[1], r.rows, 1e-4).all(), (F, variant)
store.compact()
me.match(Q, 3, 0.6)
store.close()
print("sanitize probe ok")
