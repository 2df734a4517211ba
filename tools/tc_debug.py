"""Small diagnostic for the tensor-core variants (run on the GPU box under `timeout`)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import facerecognition_infrenceengine_b200 as frg
from oracle import matcher_oracle as mo
from oracle import synth


def run(n, f, k, d=512):
    store = frg.GalleryStore(dim=d, capacity=n)
    store.fill_synthetic(n, 0, 1234)
    G, _ = store.read_rows()
    Q, target = synth.queries(f, n, d)
    ref_rows, ref_scores, ref_acc = mo.match_topk(Q, G, k + 1, 0.45)
    m = frg.Matcher(store)
    for variant in ("tc_bf16", "tc_exact"):
        t0 = time.time()
        try:
            r = m.match(Q, k, 0.45, variant=variant)
        except Exception as e:
            print("n=%d f=%d k=%d %s FAILED: %s" % (n, f, k, variant, e), flush=True)
            return False
        tol = 4e-3 if variant == "tc_bf16" else 1e-4
        ids_ok = mo.ids_match_with_gap(ref_rows, ref_scores, r.rows, 2 * tol if variant == "tc_bf16" else tol)
        ds = np.abs(r.scores - ref_scores[:, :k])
        print("n=%d f=%d k=%d %-8s launches=%d ids_ok=%d/%d max|ds|=%.3e accept_ok=%s t=%.3fs" % (
            n, f, k, variant, r.launches, ids_ok.sum(), f, ds.max(), (r.accept == ref_acc).all(), time.time() - t0),
            flush=True)
        if not ids_ok.all():
            bad = np.nonzero(~ids_ok)[0][:3]
            for b in bad:
                print("  q=%d got rows %s scores %s" % (b, r.rows[b], r.scores[b]))
                print("       ref rows %s scores %s" % (ref_rows[b], ref_scores[b]))
    store.close()
    return True


if __name__ == "__main__":
    for n, f, k in [(4096, 8, 1), (20000, 40, 5), (20000, 130, 5), (200001, 64, 10), (1000000, 256, 5)]:
        if not run(n, f, k):
            break
