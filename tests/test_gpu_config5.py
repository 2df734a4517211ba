"""BASELINE config 5 at full size: the peopleCount video stream - batches of 32 frames x up to 50 faces
matched against a 1 M x 512 gallery through CameraProcessor (three-way decision of peopleCount.py:876-887),
with online enrolment / update / eviction between the batches (infrenceServer.py:185-258 semantics).

The gallery is too large for the per-face oracle loop, so the checks are the size-independent properties the
domain offers: a face of an enrolled person is recognised as that person from the very next batch on, an
evicted person is never reported again, an updated template replaces the old one in place, ragged frames
(0 faces included) split back exactly, and every reported score equals the exact fp32 dot product of the
query with the reported row (numpy, |delta| <= 1e-4)."""
import numpy as np
import pytest

from oracle import matcher_oracle as mo
from oracle import synth

pytestmark = pytest.mark.gpu

TOL = 1e-4       # north_star: |delta score| <= 1e-4 for fp32 storage


@pytest.fixture(scope="module")
def frg():
    import __graft_entry__ as g
    g.build()
    import facerecognition_infrenceengine_b200 as frg
    return frg


def _noisy(v, rng, sigma=0.03):
    return (v + np.float32(sigma) * rng.standard_normal(v.shape).astype(np.float32)).astype(np.float32)


def test_config5_stream_with_online_enrolment(frg):
    n, d = 1_000_000, 512
    rng = np.random.default_rng(55)
    store = frg.GalleryStore(dim=d, capacity=n + 4096)
    store.fill_synthetic(n, 0, synth.GALLERY_SEED)
    cam = frg.CameraProcessor(store)                     # 0.45 / 0.35, all tenants (peopleCount.py:829-830, 848)
    enrolled = {}                                        # id -> unit template (the reference's dict values)
    evicted_ids, evicted_rows = set(), set()
    next_id = 0

    for batch in range(5):
        counts = rng.integers(0, 51, size=32)            # faces per frame, ragged, zeros allowed
        counts[rng.integers(0, 32)] = 0
        F = int(counts.sum())
        kind = rng.integers(0, 4, size=F)                # 0 synthetic person, 1 enrolled person, 2 impostor, 3 evicted
        want, Q = [], np.empty((F, d), np.float32)
        live_ids = [p for p in enrolled if p not in evicted_ids]
        for f in range(F):
            if kind[f] == 1 and live_ids:
                p = live_ids[rng.integers(0, len(live_ids))]
                Q[f] = _noisy(enrolled[p], rng)
                want.append(("recognized", p))
            elif kind[f] == 3 and (evicted_ids or evicted_rows):
                if evicted_ids and (not evicted_rows or rng.integers(0, 2)):
                    p = sorted(evicted_ids)[rng.integers(0, len(evicted_ids))]
                    Q[f] = _noisy(enrolled[p], rng)
                else:
                    r = sorted(evicted_rows)[rng.integers(0, len(evicted_rows))]
                    Q[f] = _noisy(synth.unit_rows([r], d)[0], rng)
                want.append(("unknown", None))           # nobody else is within 0.35 of an evicted person
            elif kind[f] == 0:
                r = int(rng.integers(0, n))
                while r in evicted_rows:
                    r = int(rng.integers(0, n))
                Q[f] = _noisy(synth.unit_rows([r], d)[0], rng)
                want.append(("recognized", "%024x" % r))
            else:
                Q[f] = rng.standard_normal(d).astype(np.float32)
                want.append(("unknown", None))           # max cosine of a random vector over 1 M rows ~ 0.22
        events, stats = cam.process(Q)
        assert stats["faces"] == F and len(events) == F
        for f, (ev, w) in enumerate(zip(events, want)):
            assert ev[0] == w[0] and ev[1] == w[1], (batch, f, kind[f], ev, w)
        assert stats["recognized"] == sum(w[0] == "recognized" for w in want)
        # frames split back exactly (ragged, empty frames included)
        cuts = np.cumsum(counts)[:-1]
        frames = np.split(np.arange(F), cuts)
        assert [len(x) for x in frames] == list(counts)
        # exact scores: reported score == fp32 dot of the normalised query with the reported row
        r = frg.Matcher(store).match(Q, 5, 0.45)
        assert (np.diff(r.scores, axis=1) <= 0).all() and all(len(set(x)) == 5 for x in r.rows)
        for f in rng.choice(F, size=min(F, 40), replace=False):
            for j in (0, 4):
                row = int(r.rows[f, j])
                g, tag = store.read_rows(row, 1)
                assert tag[0] >= 0                       # a tombstone never comes back
                assert abs(float(np.dot(mo.normalise(Q[f]), g[0])) - float(r.scores[f, j])) <= TOL
            assert events[f][0] != "recognized" or abs(events[f][2] - float(r.scores[f, 0])) <= 1e-6

        # ---- online enrolment between the batches: 64 new people, 8 updated templates, 16 evictions
        new_ids = ["person-%06d" % (next_id + i) for i in range(64)]
        next_id += 64
        V = rng.standard_normal((64, d)).astype(np.float32) * np.float32(0.7)      # raw means, norm != 1
        store.upsert(new_ids, V, ["campus-a"] * 64)
        for p, v in zip(new_ids, V):
            enrolled[p] = mo.normalise(v)
        live_ids = [p for p in enrolled if p not in evicted_ids]
        upd = [live_ids[i] for i in rng.choice(len(live_ids), size=8, replace=False)]
        U = rng.standard_normal((8, d)).astype(np.float32)
        rows_before = [store.row_of(p) for p in upd]
        store.upsert(upd, U, ["campus-a"] * 8)
        assert [store.row_of(p) for p in upd] == rows_before         # dict assignment keeps the position
        for p, v in zip(upd, U):
            enrolled[p] = mo.normalise(v)
        live_ids = [p for p in enrolled if p not in evicted_ids]
        gone = [live_ids[i] for i in rng.choice(len(live_ids), size=8, replace=False)]
        assert store.remove(gone) == 8
        evicted_ids.update(gone)
        rows_gone = [int(x) for x in rng.integers(0, n, size=8)]
        store.remove_rows(rows_gone)
        evicted_rows.update(rows_gone)

    st = store.stats()
    assert st.rows == n + 5 * 64 and st.live == n + 5 * 64 - len(evicted_ids) - len(evicted_rows)
    store.close()
