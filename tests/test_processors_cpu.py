"""Host logic of the two drop-in processors on CPU: the decision code of FaceRecognitionProcessor
(infrenceServer.py:545-552) and CameraProcessor (peopleCount.py:876-887) over the reference's own outputs
(tests/golden/edge_cases.npz, cfg1_10k_x_64.npz: scores within 1 ulp of every threshold, ties, a NaN row, the
nothing-matches frame).  The matcher is injected and answers with the reference's best row / score from the golden
file (test infrastructure), so what is under test is only what the product does AFTER the kernels."""
import numpy as np

import facerecognition_infrenceengine_b200 as frg
from facerecognition_infrenceengine_b200.matcher import MatchResult

KIND = {"recognized": 0, "unknown": 1, "ignored": 2}


class _Store:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def id_of(self, row):
        return None if row < 0 else "%024x" % row

    def metadata(self, pid):
        return {"name": "n" + pid[-4:], "type": "employee"}


class _GoldenMatcher:
    """match_host() answers with the reference's own best row and score."""

    def __init__(self, rows, scores):
        self.rows, self.scores = np.asarray(rows, np.int64), np.asarray(scores, np.float32)
        self.calls = []

    def match_host(self, Q, k, threshold, company_id=None, variant="auto", with_ids=True):
        self.calls.append((len(Q), k, float(threshold), company_id, with_ids))
        rows = self.rows[:len(Q), None].copy()
        scores = np.where(rows[:, 0] >= 0, self.scores[:len(Q)], np.float32(-1.0)).astype(np.float32)[:, None]
        accept = (rows[:, 0] >= 0) & (scores[:, 0] >= np.float32(threshold))      # the kernels' fp32 compare
        return MatchResult(rows, scores, accept)


def _check(g, n):
    Q = np.zeros((len(g["ref_best_row"]), 8), np.float32)
    m = _GoldenMatcher(g["ref_best_row"], g["ref_best_score"])
    store = _Store(n)
    events, stats = frg.CameraProcessor(store, matcher=m).process(Q)
    assert [KIND[e[0]] for e in events] == list(g["ref_campus_kind"])
    rec = np.asarray(g["ref_campus_kind"]) == 0
    assert [e[1] for e, ok in zip(events, rec) if ok] == ["%024x" % r for r in np.asarray(g["ref_best_row"])[rec]]
    assert all(isinstance(e[2], float) for e, ok in zip(events, rec) if ok)       # float(best_score), :881
    assert all(e[1] is None and e[2] is None for e, ok in zip(events, rec) if not ok)
    assert stats["faces"] == len(Q) and stats["recognized"] == int(rec.sum())
    assert stats["unknown"] == int((np.asarray(g["ref_campus_kind"]) == 1).sum())
    live = frg.FaceRecognitionProcessor(store, matcher=m).recognize(Q, "acme")
    got_row = np.array([int(x["person_id"], 16) if x["person_id"] else -1 for x in live])
    assert (got_row == g["ref_live_row"]).all()
    for x, row in zip(live, g["ref_live_row"]):
        if row < 0:
            assert x["person_info"] == {"name": "Unknown", "type": "unknown"} and x["recognition_score"] == 0
        else:
            assert x["person_info"]["type"] == "employee"
    assert m.calls[0][1:] == (1, 0.45, None, False) and m.calls[1][3] == "acme"
    assert abs(m.calls[1][2] - 0.4) < 1e-12


def test_decisions_at_the_threshold_edges(golden):
    g = golden("edge_cases.npz")
    _check(g, len(g["gallery"]))


def test_decisions_config1(golden):
    g = golden("cfg1_10k_x_64.npz")
    _check(g, int(g["n"]))
    assert np.abs(np.array(g["ref_live_score"], np.float32)[g["ref_live_row"] >= 0]
                  - np.array(g["ref_best_score"], np.float32)[g["ref_live_row"] >= 0]).max() == 0


def test_empty_gallery_and_empty_frame():
    m = _GoldenMatcher([], [])
    assert frg.FaceRecognitionProcessor(_Store(0), matcher=m).recognize(np.zeros((3, 8), np.float32)) == []
    assert frg.CameraProcessor(_Store(0), matcher=m).process(np.zeros((3, 8), np.float32)) == (
        [], {"faces": 0, "recognized": 0, "unknown": 0})                         # peopleCount.py:850-851
    ev, st = frg.CameraProcessor(_Store(5), matcher=m).process(np.zeros((0, 8), np.float32))
    assert ev == [] and st == {"faces": 0, "recognized": 0, "unknown": 0} and m.calls == []


def test_manager_compacts_after_heavy_eviction_only():
    """EmbeddingManager squeezes tombstones out once they are many AND a sizeable share of the gallery (the
    reference's `del self.embeddings[id]`, infrenceServer.py:248-251, frees the entry at once); the odd eviction
    never triggers it.  Host logic only: the store is a stand-in that counts."""
    from types import SimpleNamespace
    from facerecognition_infrenceengine_b200.manager import EmbeddingManager, ListSource

    class Store:
        def __init__(self):
            self.rows, self.live, self.compactions, self.ids_ = 0, 0, 0, set()

        def upsert(self, ids, vecs, comps=None, meta=None):
            new = [i for i in ids if i not in self.ids_]
            self.ids_.update(new); self.rows += len(new); self.live += len(new)

        def remove(self, ids):
            gone = [i for i in ids if i in self.ids_]
            self.ids_.difference_update(gone); self.live -= len(gone)
            return len(gone)

        def stats(self):
            return SimpleNamespace(rows=self.rows, live=self.live, capacity=self.rows, bytes=0, version=0)

        def compact(self):
            self.compactions += 1; self.rows = self.live

    def emp(i, status="active"):
        return {"_id": "%024x" % i, "embedding": np.ones(8, np.float32), "companyId": "c", "status": status,
                "blacklisted": False, "lastUpdated": None}

    E = [emp(i) for i in range(6000)]
    st = Store()
    m = EmbeddingManager(ListSource(E, []), mode="live", store=st)
    assert st.rows == 6000 and st.compactions == 0
    for e in E[:10]:
        e["status"] = "inactive"                    # a few evictions: tombstones stay
    m._remove_inactive_employees()
    assert st.live == 5990 and st.rows == 6000 and st.compactions == 0
    for e in E[:1200]:
        e["status"] = "inactive"                    # many, but under a quarter of the gallery
    m._remove_inactive_employees()
    assert st.live == 4800 and st.compactions == 0
    for e in E[:2000]:
        e["status"] = "inactive"                    # a third of the gallery is dead now
    m._remove_inactive_employees()
    assert st.live == 4000 and st.compactions == 1 and st.rows == 4000
    m._remove_inactive_employees()                  # nothing new: nothing happens
    assert st.compactions == 1


def _counting_store(dim=8, fail_compact=False):
    from types import SimpleNamespace

    class Store:
        def __init__(self):
            self.dim, self.rows, self.live, self.compactions, self.ids_ = dim, 0, 0, 0, {}

        def upsert(self, ids, vecs, comps=None, meta=None):
            assert vecs.shape == (len(ids), dim)
            for i in ids:
                if i not in self.ids_:
                    self.rows += 1; self.live += 1
                self.ids_[i] = True

        def remove(self, ids):
            gone = [i for i in ids if i in self.ids_]
            for i in gone:
                del self.ids_[i]
            self.live -= len(gone)
            return len(gone)

        def stats(self):
            return SimpleNamespace(rows=self.rows, live=self.live, capacity=self.rows, bytes=0, version=0)

        def compact(self):
            self.compactions += 1
            if fail_compact:
                raise RuntimeError("libfrg error 3: out of device memory")
            self.rows = self.live

    return Store()


def test_one_malformed_document_does_not_stop_the_sync():
    """The reference wraps every person in its own try/except and skips the bad ones (infrenceServer.py:264-341).
    Here a wrong-sized, a None and a ragged embedding are skipped and logged; the good documents of the same batch
    are enrolled, the constructor does not raise, and last_sync_time advances so the bad document is not refetched
    for ever."""
    from datetime import datetime
    from facerecognition_infrenceengine_b200.manager import EmbeddingManager, ListSource

    def emp(i, emb):
        return {"_id": "%024x" % i, "embedding": emb, "companyId": "c", "status": "active", "blacklisted": False,
                "lastUpdated": datetime(2026, 1, 1)}

    good = np.ones(8, np.float32)
    E = [emp(0, good), emp(1, np.ones(7, np.float32)), emp(2, good), emp(3, [[1.0, 2.0], [3.0]]), emp(4, "not a vector")]
    V = [{"_id": "%024x" % 100, "embedding": np.ones((2, 4), np.float32), "companyId": "c", "lastUpdated": datetime(2026, 1, 1)},
         {"_id": "%024x" % 101, "embedding": good, "companyId": "c", "lastUpdated": datetime(2026, 1, 1)}]
    st = _counting_store()
    m = EmbeddingManager(ListSource(E, V), mode="live", store=st)
    assert sorted(st.ids_) == ["%024x" % i for i in (0, 2, 101)]
    assert m.rejected_records == 4 and not m.is_initial_load and m.last_sync_time is not None
    before = m.last_sync_time
    E.append(emp(5, good)); E[-1]["lastUpdated"] = datetime(2999, 1, 1)
    E.append(emp(6, np.ones(3, np.float32))); E[-1]["lastUpdated"] = datetime(2999, 1, 1)
    m.force_sync()
    assert "%024x" % 5 in st.ids_ and "%024x" % 6 not in st.ids_
    assert m.last_sync_time >= before and m.rejected_records == 5


def test_failed_compaction_is_logged_backed_off_and_the_sync_goes_on():
    from datetime import datetime
    from facerecognition_infrenceengine_b200.manager import EmbeddingManager, ListSource

    def emp(i, status="active"):
        return {"_id": "%024x" % i, "embedding": np.ones(8, np.float32), "companyId": "c", "status": status,
                "blacklisted": False, "lastUpdated": datetime(2026, 1, 1)}

    E = [emp(i) for i in range(4000)]
    st = _counting_store(fail_compact=True)
    m = EmbeddingManager(ListSource(E, []), mode="live", store=st)
    for e in E[:2000]:
        e["status"] = "inactive"
    E.append(dict(emp(9000), lastUpdated=datetime(2999, 1, 1)))
    m.force_sync()                                   # eviction -> compaction fails inside: must not abort the cycle
    assert st.compactions == 1 and st.live == 2001 and "%024x" % 9000 in st.ids_
    t = m.last_sync_time
    m.force_sync()                                   # not retried every cycle
    assert st.compactions == 1 and m.last_sync_time >= t


def test_layout_lock_orders_compaction_against_read_sections():
    """gallery._LayoutLock: read sections run together and nest; a writer waits for them and holds new ones off."""
    import threading
    import time
    from facerecognition_infrenceengine_b200.gallery import _LayoutLock
    lk = _LayoutLock()
    log, in_read, go = [], threading.Event(), threading.Event()

    def reader():
        with lk.read():
            with lk.read():                          # nested (Matcher.match inside a processor's section)
                in_read.set()
                go.wait(5)
                log.append("read-done")

    def writer():
        with lk.write():
            log.append("write")

    def late_reader():
        with lk.read():
            log.append("late-read")

    tr = threading.Thread(target=reader); tr.start()
    assert in_read.wait(5)
    tw = threading.Thread(target=writer); tw.start()
    time.sleep(0.1)
    tl = threading.Thread(target=late_reader); tl.start()      # arrives while the writer waits: queues behind it
    time.sleep(0.1)
    assert log == []
    go.set()
    for t in (tr, tw, tl):
        t.join(5)
    assert log == ["read-done", "write", "late-read"]
    with lk.read():
        try:
            with lk.write():
                raise AssertionError("write inside read must be refused")
        except RuntimeError:
            pass
