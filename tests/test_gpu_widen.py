"""The rows next to the matching path, each against the reference's own outputs
(tests/golden) or the oracle: enrol-time duplicate / same-person checks, on-device unknown-person
clustering, the bulk snapshot format, and the frame->batch aggregator over the real matcher."""
import numpy as np
import pytest

from oracle import matcher_oracle as mo
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frg():
    import __graft_entry__ as g
    g.build()
    import facerecognition_infrenceengine_b200 as frg
    return frg


def hex_ids(n, base=0):
    return ["%024x" % (base + i) for i in range(n)]


def test_enrol_checks_golden(frg, golden):
    """trainingServer._check_duplicate_face / _check_image_similarity, reference outputs."""
    g = golden("enrol_checks.npz")
    n, d = int(g["n"]), int(g["dim"])
    G = synth.gallery(n, d, seed=int(g["seed"]))
    stored = (G * g["scale"]).astype(np.float32)              # raw templates, norm < 1: normalised on ingest
    store = frg.GalleryStore(dim=d, capacity=n)
    store.upsert(hex_ids(n), stored, ["acme"] * n)
    chk = frg.EnrolmentChecker(store)
    got = []
    for p in g["probes"]:
        dup, pid = chk.check_duplicate_face(p, "acme")
        got.append(int(pid, 16) if dup else -1)
    assert got == list(g["ref_duplicate"])
    assert chk.check_duplicate_face(g["probes"][0], "other-company") == (False, None)
    for tr, ok, pair in zip(g["triples"], g["ref_same_ok"], g["ref_same_pair"]):
        o, p = chk.check_image_similarity(list(tr))
        assert o == bool(ok) and (p if p else (-1, -1)) == tuple(pair)
    assert chk.check_image_similarity([g["triples"][0][0]]) == (True, None)
    store.close()


def test_enrolment_gallery_scans_what_the_reference_scans(frg):
    """trainingServer.py:170-200 scans ONE collection of the company (employees or visitors), every document with an
    embeddingId whatever its status, and returns doc[id_field].  The live matching store has evicted inactive /
    blacklisted employees (infrenceServer.py:234-258) and mixes both collections - an EnrolmentGallery does not."""
    rng = np.random.default_rng(8)
    d = 512
    V = mo.normalise_rows(rng.standard_normal((6, d)).astype(np.float32))
    gal = frg.EnrolmentGallery(dim=d, capacity=16)
    #        doc id   company  kind        ref_id (doc[id_field])
    docs = [("e1", "acme", "employee", "EMP-001"), ("e2", "acme", "employee", "EMP-002"),
            ("v1", "acme", "visitor", "VIS-001"), ("e3", "globex", "employee", "EMP-900"),
            ("e4", "acme", "employee", None)]
    gal.add([x[0] for x in docs], V[:5], [x[1] for x in docs], [x[2] for x in docs], [x[3] for x in docs])
    chk = frg.EnrolmentChecker(gal)
    near = lambda i: V[i] + np.float32(0.02) * rng.standard_normal(d).astype(np.float32)      # noqa: E731
    assert chk.check_duplicate_face(near(0), "acme", kind="employee") == (True, "EMP-001")
    # an inactive / blacklisted employee is still in the enrolment gallery: the re-enrolment is caught
    assert chk.check_duplicate_face(near(1), "acme", kind="employee") == (True, "EMP-002")
    # a visitor's template does not make an EMPLOYEE enrolment a duplicate, and the other way round
    assert chk.check_duplicate_face(near(2), "acme", kind="employee") == (False, None)
    assert chk.check_duplicate_face(near(2), "acme", kind="visitor") == (True, "VIS-001")
    assert chk.check_duplicate_face(near(0), "acme", kind="visitor") == (False, None)
    # another company's people never match; a document without the id field falls back to its _id
    assert chk.check_duplicate_face(near(3), "acme", kind="employee") == (False, None)
    assert chk.check_duplicate_face(near(3), "globex", kind="employee") == (True, "EMP-900")
    assert chk.check_duplicate_face(near(4), "acme", kind="employee") == (True, "e4")
    assert chk.check_duplicate_face(V[5], "acme", kind="employee") == (False, None)
    ok, pair = chk.check_image_similarity([V[0], near(0), V[1]])
    assert not ok and pair == (0, 2)
    gal.close()


def test_first_above_is_first_not_best(frg):
    d = 512
    G = synth.gallery(5000, d, 3)
    q = G[4000] + 0.0
    G[100] = mo.normalise(q + 0.9 * G[100])                  # cos ~ 0.74 at a LOWER row than the exact hit
    store = frg.GalleryStore(dim=d, capacity=5000)
    store.append_rows(G, prenormalised=True)
    m = frg.Matcher(store)
    rows, scores = m.first_above(np.stack([q, q, q]), 0.5)
    assert list(rows) == [100, 100, 100] and abs(scores[0] - float(np.dot(mo.normalise(q), G[100]))) < 1e-6
    assert m.first_above(q[None], 0.9)[0][0] == 4000            # only the exact row clears 0.9
    assert m.first_above(q[None], 1.5)[0][0] == -1
    # strict vs non-strict exactly at the score
    s100 = m.first_above(q[None], 0.5)[1][0]
    assert m.first_above(q[None], float(s100))[0][0] == 100
    assert m.first_above(q[None], float(s100), strict=True)[0][0] == 4000
    store.remove_rows([100])
    assert m.first_above(q[None], 0.5)[0][0] == 4000            # tombstones are skipped
    store.close()


def test_unknown_clustering_golden(frg, golden):
    """peopleCount.process_unknown_detection, reference outputs: same cluster sequence, same means."""
    g = golden("unknown_clusters.npz")
    Q, _ = synth.queries(int(g["f"]), int(g["people"]), int(g["dim"]), seed=int(g["seed"]),
                         genuine_every=1, noise=float(g["noise"]))
    uc = frg.UnknownClusterer(dim=int(g["dim"]), campus_id="campus0")
    res = [uc.observe(q) for q in mo.normalise_rows(Q)]
    assert [r[0] for r in res] == list(g["ref_cluster"])
    assert [r[1] for r in res] == list(g["ref_created"])
    assert np.array_equal(uc.representatives(), g["ref_avgs"])   # host mean == np.mean of the reference
    assert uc.unknown_id(0) == "unknown_campus0_1"
    uc.close()


def test_snapshot_roundtrip(frg, tmp_path):
    d, n = 512, 3000
    rng = np.random.default_rng(2)
    raw = rng.standard_normal((n, d)).astype(np.float32)
    store = frg.GalleryStore(dim=d, capacity=n)
    store.upsert(hex_ids(n), raw, ["A" if i % 4 else "B" for i in range(n)],
                 [{"name": "p%d" % i, "type": "employee"} for i in range(n)])
    store.remove(hex_ids(n)[10:20])
    path = str(tmp_path / "gallery.frgsnap")
    assert store.save(path, chunk_rows=700) == n - 10
    again = frg.GalleryStore.load(path, chunk_rows=512)
    ids0, G0, t0 = store.snapshot_arrays()
    ids1, G1, t1 = again.snapshot_arrays()
    assert ids0 == ids1 and np.array_equal(G0, G1) and (t0 == t1).all()          # bit-exact reload
    assert again.metadata(ids1[5]) == {"name": "p5", "type": "employee"}
    Q = raw[[3, 15, 777, 2999]] + 0.01
    for company in (None, "A", "B"):
        a = frg.Matcher(store).match(Q, 5, 0.4, company_id=company)
        b = frg.Matcher(again).match(Q, 5, 0.4, company_id=company)
        assert a.ids == b.ids and np.array_equal(a.scores, b.scores) and (a.accept == b.accept).all()
    store.close(); again.close()


def test_aggregator_over_the_real_matcher(frg):
    n, d = 50000, 512
    store = frg.GalleryStore(dim=d, capacity=n)
    store.fill_synthetic(n, 0, 1234)
    m = frg.Matcher(store)

    def match_fn(Q, tenant):
        r = m.match(Q, 1, 0.45, company_id=tenant, with_ids=False)
        return r.rows, r.scores, r.accept

    agg = frg.BatchAggregator(match_fn, max_batch=256, max_delay_ms=20)
    frames = [synth.queries(7, n, d, q0=100 * i) for i in range(12)]
    futs = [agg.submit(q) for q, _ in frames]
    for (q, target), f in zip(frames, futs):
        rows, scores, acc = f.result(timeout=30)
        hit = target >= 0
        assert (rows[hit, 0] == target[hit]).all() and acc[hit].all() and not acc[~hit].any()
    assert agg.batches < len(frames)
    agg.close(); store.close()
