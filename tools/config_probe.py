"""Timings of the BASELINE.json configs that are parity-test cases rather than the bench line
(run on the GPU box): config 3 (128-d Euclidean, 10 M rows, batch 256, top-1) and config 5
(video stream: <= 1600 faces per batch against 1 M rows with online enrol/remove between batches)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from oracle import matcher_oracle as mo
from oracle import synth

out = {}

# ---- config 5: 32 frames x 50 faces, 1 M gallery, upserts / removes between batches
n, d, F = 1_000_000, 512, 1600
store = frg.GalleryStore(dim=d, capacity=n + 16384)
store.fill_synthetic(n, 0, 1234)
m = frg.Matcher(store)
Q = torch.from_numpy(synth.queries(F, n, d)[0]).cuda()
steps, warm = 100, 10
new = synth.unit_rows(np.arange(64 * (steps + warm)), d, 777, synth.STREAM_IMPOSTOR).reshape(steps + warm, 64, d)
res = None
for i in range(5):
    res = m.match_device(Q, 1, 0.45, out=res)


def step(i):
    global res
    res = m.match_device(Q, 1, 0.45, out=res)
    store.upsert(["new%d_%d" % (i, j) for j in range(64)], new[i])          # 64 enrolments per batch
    if i:
        store.remove(["new%d_%d" % (i - 1, j) for j in range(0, 64, 4)])     # 16 evictions per batch


for i in range(warm):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(warm, warm + steps):
    step(i)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
steps_total = warm + steps
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    res = m.match_device(Q, 1, 0.45, out=res)
e1.record()
torch.cuda.synchronize()
match_ms = e0.elapsed_time(e1) / steps
# the freshly enrolled rows are found, the evicted ones are not
probe = np.stack([new[steps_total - 1][1], new[steps_total - 2][0], new[steps_total - 2][1]])
r = m.match(probe, 1, 0.45)
out["config5"] = {"faces_per_batch": F, "gallery_rows": n, "match_ms": match_ms, "queries_per_s_match_only": F / match_ms * 1e3,
                  "step_ms_with_64_upserts_16_removes": dt * 1e3, "queries_per_s_with_updates": F / dt,
                  "fresh_row_found": bool(r.accept[0] and r.accept[2]), "evicted_row_rejected": bool(not r.accept[1]),
                  "masked_kernel": True}
print(json.dumps(out["config5"]), flush=True)
store.close()

# ---- config 3: 10 M x 128 Euclidean, batch 256, top-1 (ours; parity unpinned by the reference)
n, d, F = 10_000_000, 128, 256
store = frg.GalleryStore(dim=d, capacity=n, raw=True)          # raw + scan plane: the Euclidean TC filter
store.fill_synthetic(n, 0, 99)           # unit rows; the Euclidean store keeps them as given
m = frg.Matcher(store, metric="euclidean")
Qh = synth.queries(F, n, d, seed=5, gallery_seed=99)[0]
Q = torch.from_numpy(Qh).cuda()
res = None
for i in range(2):
    res = m.match_device(Q, 1, 0.6, out=res)
torch.cuda.synchronize()
e0.record()
for i in range(5):
    res = m.match_device(Q, 1, 0.6, out=res)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
rows = res[0].cpu().numpy()
tgt = synth.queries(F, n, d, seed=5, gallery_seed=99)[1]
hit = tgt >= 0
out["config3"] = {"rows": n, "dim": d, "batch": F, "ms_per_batch": ms, "queries_per_s": F / ms * 1e3,
                  "variant": frg._native.last_variant(),
                  "genuine_found": bool((rows[hit, 0] == tgt[hit]).all())}
print(json.dumps(out["config3"]), flush=True)
json.dump(out, open("gpurun_out/config_probe.json", "w"), indent=1)
