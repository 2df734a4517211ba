"""Round-2 parity cases (VERDICT r01 "parity gaps at the headline shape"): the CTA-pair kernels that produce the
benchmark numbers, checked against the oracle at the benchmarked shapes; zero / NaN queries; tombstones + tenant
filter at F > 128; the adversarial rounding case of tests/rounding_case.py; chunked in-place compaction; matching
while the gallery is being compacted.  Tolerances: ids exact where adjacent oracle scores differ by more than
1e-4, |dscore| <= 1e-4, decisions identical outside 1e-4 of the threshold (BASELINE.json)."""
import threading

import numpy as np
import pytest

from oracle import matcher_oracle as mo
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def frg():
    import __graft_entry__ as g
    g.build()
    import facerecognition_infrenceengine_b200 as frg
    return frg


def compare(r_rows, r_scores, r_accept, ref, k, threshold):
    ref_rows, ref_scores, ref_acc = ref
    assert mo.ids_match_with_gap(ref_rows, ref_scores, r_rows, TOL).all()
    filled = ref_rows[:, :k] >= 0
    assert np.abs(r_scores[filled] - ref_scores[:, :k][filled]).max(initial=0) <= TOL
    assert (r_scores[~filled] == -1).all() and (r_rows[~filled] == -1).all()
    near = np.abs(ref_scores[:, 0].astype(np.float64) - threshold) <= TOL
    assert (np.asarray(r_accept, bool)[~near] == ref_acc[~near]).all()


@pytest.fixture(scope="module")
def big(frg):
    """1 M x 512 synthetic gallery on the device + its host copy (bit-identical to the CPU generator)."""
    n, d = 1_000_000, 512
    store = frg.GalleryStore(dim=d, capacity=2 * n)
    store.fill_synthetic(n, 0, synth.GALLERY_SEED)
    G, _ = store.read_rows()
    yield store, G
    store.close()


@pytest.mark.parametrize("k", [5, 10])
def test_headline_shape_every_query_against_the_oracle(frg, big, k):
    """configs[1] exactly as benchmarked: ALL 1024 queries x 1 M rows through the CTA-pair kernels, top-5 and
    top-10, against the sgemm oracle.  Three queries are degenerate: a zero vector and a vector with a NaN
    (normalise -> NaN scores: `s > best` is never true, peopleCount.py:871 - row -1, score -1, reject) and a
    gallery row itself (score 1)."""
    store, G = big
    n = len(G)
    Q, target = synth.queries(1024, n, 512)
    Q[7] = 0.0
    Q[9, 100] = np.nan
    Q[11] = G[123_456] * np.float32(3.0)
    ref = mo.match_topk_fast(Q, G, k + 1, 0.45)
    r = frg.Matcher(store).match(Q, k, 0.45, with_ids=False)
    assert r.variant == "tc_exact" and r.launches >= 4
    compare(r.rows, r.scores, r.accept, ref, k, 0.45)
    assert (r.rows[[7, 9]] == -1).all() and (r.scores[[7, 9]] == -1).all() and not r.accept[[7, 9]].any()
    assert r.rows[11, 0] == 123_456 and abs(r.scores[11, 0] - 1) < 2e-6
    ok = target >= 0
    ok[[7, 9, 11]] = False
    assert (r.rows[ok, 0] == target[ok]).all() and r.accept[ok].all()


def test_config4_shape_f4096_k10_on_2m_rows(frg, big):
    """The per-GPU shape of BASELINE configs[3] (batch 4096, top-10) on 2 M rows: every genuine query finds its
    target first, every 8th query is compared with the oracle in full."""
    store, G1 = big
    n = 2_000_000
    store.fill_synthetic(n - len(G1), len(G1), synth.GALLERY_SEED)           # rows 1 M .. 2 M of the same gallery
    try:
        Q, target = synth.queries(4096, n, 512, q0=50_000)
        r = frg.Matcher(store).match(Q, 10, 0.45, with_ids=False)
        hit = target >= 0
        assert (r.rows[hit, 0] == target[hit]).all() and r.accept[hit].all() and not r.accept[~hit].any()
        assert (np.diff(r.scores, axis=1) <= 0).all()
        G, _ = store.read_rows()
        sel = np.arange(0, 4096, 8)
        ref = mo.match_topk_fast(Q[sel], G, 11, 0.45)
        compare(r.rows[sel], r.scores[sel], r.accept[sel], ref, 10, 0.45)
    finally:
        # back to 1 M rows for the other tests of this module
        store.remove_rows(np.arange(len(G1), n))
        store.compact()


def test_tombstones_and_tenant_filter_with_pair_kernels(frg):
    """F > 128 (CTA-pair form) with removed rows and interleaved companies: masked epilogue, every variant."""
    rng = np.random.default_rng(17)
    n, d, f, k = 300_000, 512, 300, 5
    G = synth.gallery(n, d, 31)
    tags = (np.arange(n) % 3 + 1).astype(np.int32)          # three interleaved companies: no row window applies
    store = frg.GalleryStore(dim=d, capacity=n)
    store.append_rows(G, tags, prenormalised=True)
    for c in ("c1", "c2", "c3"):
        store.tenant_code(c)
    dead = rng.choice(n, 20_000, replace=False)
    store.remove_rows(dead)
    tags_now = tags.copy()
    tags_now[dead] = -1
    pick = rng.integers(0, n, f)
    Q = G[pick] + np.float32(0.03) * rng.standard_normal((f, d)).astype(np.float32)
    m = frg.Matcher(store)
    for company, code in ((None, None), ("c2", 2), ("c3", 3)):
        ref = mo.match_topk_fast(Q, G, k + 1, 0.4, tags_now, code)
        for variant in ("tc_exact", "scan_f32"):
            r = m.match(Q, k, 0.4, company_id=company, variant=variant, with_ids=False)
            compare(r.rows, r.scores, r.accept, ref, k, 0.4)
            assert not np.isin(r.rows[r.rows >= 0], dead).any()
    store.close()


@pytest.mark.parametrize("copies", [1, 200])
def test_adversarial_rounding_does_not_drop_the_true_best_row(frg, copies):
    """tests/rounding_case.py: exact A > B by 5e-4, bf16 B > A by 8.6e-3.  The filter's bound is measured from the
    rounding residuals of the stored rows and of the query, so A stays a candidate and wins the exact rescoring
    (a fixed 2 x 4e-3 margin returned B).  copies = 200: the same through the CTA-pair kernels."""
    from rounding_case import adversarial_pair
    rng = np.random.default_rng(2)
    n, d = 4094, 512
    G = mo.normalise_rows(rng.standard_normal((n, d)).astype(np.float32))
    A, B, q = adversarial_pair(0)
    G = np.concatenate([G[:1000], B[None], G[1000:3000], A[None], G[3000:]])       # B sits BEFORE A in gallery order
    rowA, rowB = 3001, 1000
    store = frg.GalleryStore(dim=d, capacity=len(G))
    store.append_rows(G, prenormalised=True)
    Q = np.concatenate([np.repeat(q[None], copies, axis=0), rng.standard_normal((5, d)).astype(np.float32)])
    for k in (1, 2, 5):
        ref = mo.match_topk(Q, G, k + 1, 0.45)
        assert ref[0][0, 0] == rowA and (k == 1 or ref[0][0, 1] == rowB)
        for variant in ("tc_exact", "auto", "scan_f32"):
            r = frg.Matcher(store).match(Q, k, 0.45, variant=variant, with_ids=False)
            compare(r.rows, r.scores, r.accept, ref, k, 0.45)
            assert (r.rows[:copies, 0] == rowA).all()
    store.close()


def test_compaction_in_place_over_several_chunks(frg):
    """frg_store_compact moves rows chunk by chunk (65 536 rows) through a bounce buffer, in place: 300 k rows with
    random tombstones, a long surviving prefix that must not move, master / plane / tags / tenant windows intact."""
    rng = np.random.default_rng(23)
    n, d = 300_000, 256
    G = synth.gallery(n, d, 5)
    tags = np.where(np.arange(n) < 150_000, 1, 2).astype(np.int32)
    store = frg.GalleryStore(dim=d, capacity=n + 10)
    store.append_rows(G, tags, prenormalised=True)
    store.tenant_code("first"); store.tenant_code("second")
    dead = np.unique(np.concatenate([rng.choice(np.arange(70_000, n), 90_000, replace=False), [n - 1]]))
    store.remove_rows(dead)
    keep = np.setdiff1d(np.arange(n), dead)
    store.compact()
    assert store.rows == len(keep) and store.layout_version == 1
    Gc, tc = store.read_rows()
    assert np.array_equal(Gc, G[keep]) and np.array_equal(tc, tags[keep])
    Q = G[keep[[5, 69_999, 70_000, 150_000, len(keep) - 1]]]
    for company, code in ((None, None), ("first", 1), ("second", 2)):
        ref = mo.match_topk(Q, G[keep], 3, 0.45, tags[keep], code)
        for variant in ("tc_exact", "scan_f32"):          # the plane moved with the master
            r = frg.Matcher(store).match(Q, 2, 0.45, company_id=company, variant=variant, with_ids=False)
            compare(r.rows, r.scores, r.accept, ref, 2, 0.45)
    store.compact()                                        # nothing to do: no renumbering, version unchanged
    assert store.stats().live == len(keep)
    store.close()


def test_threads_match_while_the_gallery_is_compacted(frg):
    """compaction renumbers rows while other threads match and translate rows to ids.  Every
    accepted face must be attributed to the person it was enrolled as - never to whoever now sits at the stale row."""
    rng = np.random.default_rng(29)
    n, d = 40_000, 512
    G = synth.gallery(n, d, 77)
    ids = ["p%06d" % i for i in range(n)]
    store = frg.GalleryStore(dim=d, capacity=n)
    store.upsert(ids, G, prenormalised=True)
    victims = rng.permutation(np.arange(0, n // 2))           # removed over time, from the front half
    probe = np.arange(n // 2, n // 2 + 96)                    # never removed: their rows shift down at every compaction
    Q = G[probe]
    want = [ids[i] for i in probe]
    stop, errors, rounds = threading.Event(), [], [0, 0]

    def matcher_loop(which):
        import torch
        torch.cuda.set_device(0)
        m = frg.Matcher(store)
        proc = frg.FaceRecognitionProcessor(store, recognition_threshold=0.45)
        try:
            while not stop.is_set():
                if which == 0:
                    r = m.match(Q, 1, 0.45)                  # with ids: one read section inside
                    got = [x[0] for x in r.ids]
                else:
                    got = [x["person_id"] for x in proc.recognize(Q)]
                if got != want:
                    errors.append((which, sum(a != b for a, b in zip(got, want))))
                    return
                rounds[which] += 1
        except Exception as ex:            # noqa: BLE001
            errors.append(repr(ex))

    th = [threading.Thread(target=matcher_loop, args=(w,)) for w in (0, 1)]
    [t.start() for t in th]
    for step in range(12):
        store.remove([ids[i] for i in victims[step * 1500:(step + 1) * 1500]])
        store.compact()
    stop.set()
    [t.join(60) for t in th]
    assert not errors, errors[:3]
    assert min(rounds) > 0 and store.layout_version == 12
    # a result kept across a compaction is refused, not mistranslated
    r = frg.Matcher(store).match(Q[:4], 1, 0.45, with_ids=False)
    store.remove([ids[int(victims[-1])]]); store.compact()
    with pytest.raises(frg.StaleRows):
        store.ids_of(r.rows, layout_version=r.layout_version)
    store.close()


def _stretched(frg, n, T, tenant, alternate):
    """Companies enrolled as blocks of T rows; ONE person of `tenant` re-enrolled at the very end of the gallery, so
    that company's row window spans almost everything while its rows stay concentrated in one CTA's chunk.
    alternate: inside its block the company's rows alternate with another company's (thousands of 1-row intervals:
    the store gives up on an interval list for it - "scattered")."""
    d = 512
    G = synth.gallery(n, d, 41)
    tags = (1 + np.arange(n) // T).astype(np.int32)
    lo = (tenant - 1) * T
    if alternate:
        tags[lo + 1:lo + T:2] = 9999
    tags[n - 1] = tenant                                 # the late re-enrolment
    store = frg.GalleryStore(dim=d, capacity=n)
    store.append_rows(G, tags, prenormalised=True)
    store._tenants = {"c%d" % i: i for i in range(1, n // T + 2)}
    return store, G, tags, lo


@pytest.mark.parametrize("alternate", [False, True])
def test_stretched_tenant_window(frg, alternate):
    """alternate=False: the kernels walk only the TILES the company's row intervals touch (tile list, DESIGN.md
    section 4.4) - no overflow, no fallback.  alternate=True: no interval list; the masked scan of the wide window
    overflows the private candidate segments of the one chunk that holds the company and those queries are redone
    exactly by the tag-first ("sparse") fallback scan.  Both against the oracle's masked scan; F = 64 (single-CTA
    kernels) and F = 300 (CTA pairs)."""
    n, T, tenant = 400_000, 12_000, 7
    store, G, tags, lo = _stretched(frg, n, T, tenant, alternate)
    rng = np.random.default_rng(3)
    m = frg.Matcher(store)
    mine = np.nonzero(tags == tenant)[0]
    for F in (64, 300):
        pick = rng.choice(mine, size=F)
        pick[0] = n - 1
        Q = G[pick] + np.float32(0.03) * rng.standard_normal((F, 512)).astype(np.float32)
        Q[F // 2:] = rng.standard_normal((F - F // 2, 512)).astype(np.float32)
        ref = mo.match_topk_fast(Q, G, 6, 0.4, tags, tenant)
        for variant in ("auto", "tc_bf16"):
            r = m.match(Q, 5, 0.4, company_id="c%d" % tenant, variant=variant, with_ids=False)
            if variant == "auto":
                compare(r.rows, r.scores, r.accept, ref, 5, 0.4)
            else:                                        # bf16 scores: own tolerance
                assert np.abs(r.scores - ref[1][:, :5]).max() <= 8e-3
            assert r.rows[0, 0] == n - 1
    # after more enrolment and a compaction the lists follow (store version)
    store.remove_rows(np.arange(0, 50_000))
    extra = synth.gallery(100, 512, 43)
    store.append_rows(extra, np.full(100, tenant, np.int32), prenormalised=True)
    G2 = np.concatenate([G, extra]); tags2 = np.concatenate([tags, np.full(100, tenant, np.int32)])
    tags2[:50_000] = -1
    Q = np.concatenate([extra[:3], G[mine[:5]]])
    ref = mo.match_topk_fast(Q, G2, 3, 0.4, tags2, tenant)
    r = m.match(Q, 2, 0.4, company_id="c%d" % tenant, with_ids=False)
    compare(r.rows, r.scores, r.accept, ref, 2, 0.4)
    store.compact()
    keep = np.nonzero(tags2 >= 0)[0]
    ref = mo.match_topk_fast(Q, G2[keep], 3, 0.4, tags2[keep], tenant)
    r = m.match(Q, 2, 0.4, company_id="c%d" % tenant, with_ids=False)
    compare(r.rows, r.scores, r.accept, ref, 2, 0.4)
    store.close()


def test_small_frames_across_mutations(frg):
    """The reference's own operating point (one frame of a few faces per call against some ten thousand templates,
    infrenceServer.py:603-622), call after call with the same shapes while the gallery changes underneath: enrol,
    evict, overwrite in place, growth of the arrays, compaction - a match never sees a stale gallery (every answer
    against the reference's per-face loop over the equivalent dict)."""
    rng = np.random.default_rng(31)
    n, d = 10_000, 512
    G = synth.gallery(n, d, 51)
    ids = ["p%05d" % i for i in range(n)]
    store = frg.GalleryStore(dim=d, capacity=n)              # capacity n: the appends below also grow the arrays
    store.upsert(ids, G, ["acme" if i % 2 else "globex" for i in range(n)], prenormalised=True)
    m = frg.Matcher(store)
    live = dict(zip(ids, G))
    comp = {p: ("acme" if i % 2 else "globex") for i, p in enumerate(ids)}
    comp["fresh"] = "acme"

    def check(Q, k, thr, company=None):
        r = m.match(Q, k, thr, company_id=company)
        sub = [p for p in live if company is None or comp[p] == company]
        for f in range(len(Q)):
            want = mo.scan_best(mo.normalise(Q[f]), {p: live[p] for p in sub})
            assert r.ids[f][0] == want[0] or (want[0] is None and r.ids[f][0] is None), (f, r.ids[f], want)
            if want[0] is not None:
                assert abs(r.scores[f, 0] - want[1]) <= 1e-4 and bool(r.accept[f]) == bool(np.float32(want[1]) >= np.float32(thr))
        return r

    Q8 = G[rng.integers(0, n, 8)] + np.float32(0.03) * rng.standard_normal((8, d)).astype(np.float32)
    Q8[5:] = rng.standard_normal((3, d)).astype(np.float32)
    check(Q8, 1, 0.4)
    r2 = check(Q8, 1, 0.4)
    r3 = check(Q8[::-1].copy(), 1, 0.4)                      # other queries, same shape
    assert r3.rows[0, 0] == r2.rows[7, 0] and r2.launches >= 4
    check(Q8, 5, 0.4)
    check(Q8, 1, 0.4, "acme")
    check(Q8, 1, 0.4, "globex")
    # mutations: the gallery of the old version must never answer
    victim = ids[int(r2.rows[0, 0])]
    store.remove([victim]); del live[victim]
    r = check(Q8, 1, 0.4)
    assert r.ids[0][0] != victim
    check(Q8, 1, 0.4)
    newv = mo.normalise(Q8[6])
    store.upsert(["fresh"], newv[None], ["acme"]); live["fresh"] = newv          # append: also grows the arrays
    r = check(Q8, 1, 0.4)
    assert r.ids[6][0] == "fresh"
    r = check(Q8, 1, 0.4)
    assert r.ids[6][0] == "fresh"
    store.upsert([ids[3]], mo.normalise(Q8[7])[None], ["globex"]); live[ids[3]] = mo.normalise(Q8[7])   # overwrite in place
    comp[ids[3]] = "globex"
    assert check(Q8, 1, 0.4).ids[7][0] == ids[3]
    assert check(Q8, 1, 0.4).ids[7][0] == ids[3]
    store.compact()
    check(Q8, 1, 0.4)
    check(Q8, 1, 0.4)
    for f in range(1, 9):                                    # ragged frames
        check(Q8[:f], 1 + f % 3, 0.45)
    # a second thread has its own stream
    import threading
    err = []

    def other():
        try:
            import torch
            torch.cuda.set_device(0)
            for _ in range(4):
                check(Q8, 1, 0.4)
        except Exception as ex:            # noqa: BLE001
            err.append(repr(ex))

    t = threading.Thread(target=other); t.start(); t.join(60)
    assert not err, err
    store.close()


def test_stretched_tenant_window_euclidean(frg):
    """Tile lists serve the Euclidean scan plane too (raw store, 128-d, pad block): a company's block plus one far
    row, against the fp64 direct-difference oracle; bit-identical to the exact scan."""
    rng = np.random.default_rng(37)
    n, d, T, tenant = 100_000, 128, 9_000, 4
    G = (synth.gallery(n, d, 61) * np.float32(1.3)).astype(np.float32)
    tags = (1 + np.arange(n) // T).astype(np.int32)
    tags[n - 1] = tenant
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.append_rows(G, tags)
    store._tenants = {"c%d" % i: i for i in range(1, n // T + 2)}
    mine = np.nonzero(tags == tenant)[0]
    m = frg.Matcher(store, metric="euclidean")
    for F in (40, 130):
        pick = rng.choice(mine, size=F)
        pick[0] = n - 1
        Q = G[pick] + np.float32(0.02) * rng.standard_normal((F, d)).astype(np.float32)
        ref_r, ref_d, ref_a = mo.euclidean_topk(Q, G, 3, 0.6, tags, tenant)
        a = m.match(Q, 2, 0.6, company_id="c%d" % tenant, variant="tc_exact", with_ids=False)
        b = m.match(Q, 2, 0.6, company_id="c%d" % tenant, variant="scan_f32", with_ids=False)
        assert np.array_equal(a.rows, b.rows) and np.array_equal(a.scores.view(np.uint32), b.scores.view(np.uint32))
        assert np.array_equal(a.accept, b.accept)
        assert mo.ids_match_with_gap(ref_r, -ref_d, a.rows, TOL).all()
        assert np.abs(a.scores - ref_d[:, :2]).max() <= TOL and a.rows[0, 0] == n - 1
    store.close()


@pytest.mark.parametrize("seed", range(4))
def test_random_life_of_a_multi_tenant_gallery(frg, seed):
    """Companies enrolled as blocks, then a random sequence of late enrolments (stretch the row windows), in-place
    moves between companies, evictions and compactions, with company-filtered matches in between (single-CTA and pair
    kernels, both exact variants) against the oracle's masked scan over a host mirror: row windows, span intervals,
    the tile-list cache and its versioning can never change an answer, only the time it takes."""
    rng = np.random.default_rng(500 + seed)
    d, T, C = 512, 9_000, 8
    n0 = T * C
    G = synth.gallery(n0 + 4_000, d, 70 + seed)
    vec = {"r%d" % i: G[i] for i in range(n0)}
    comp = {"r%d" % i: "c%d" % (i // T) for i in range(n0)}
    order = list(vec)
    store = frg.GalleryStore(dim=d, capacity=n0 + 8_000)
    store.upsert(order, G[:n0], [comp[p] for p in order], prenormalised=True)
    m = frg.Matcher(store)
    fresh = n0

    def check():
        company = "c%d" % rng.integers(0, C)
        F = int(rng.choice([3, 40, 150]))
        mine = [p for p in order if comp[p] == company]
        pick = rng.choice(len(mine), size=F)
        Q = np.stack([vec[mine[i]] for i in pick]) + np.float32(0.03) * rng.standard_normal((F, d)).astype(np.float32)
        Q[F // 2:] = rng.standard_normal((F - F // 2, d)).astype(np.float32)
        Gm = np.stack([vec[p] for p in order])
        tg = np.array([int(comp[p][1:]) for p in order], np.int32)
        ref = mo.match_topk_fast(Q, Gm, 4, 0.4, tg, int(company[1:]))
        for variant in ("auto", "scan_f32"):
            r = m.match(Q, 3, 0.4, company_id=company, variant=variant)
            got = np.array([[order.index(x) if x is not None else -1 for x in row] for row in r.ids]) if F <= 40 else None
            # rows are positions in the store (tombstones included): compare through ids for small F, scores always
            filled = ref[0][:, :3] >= 0
            assert np.abs(r.scores[filled] - ref[1][:, :3][filled]).max(initial=0) <= TOL, (variant, company, F)
            assert (r.scores[~filled] == -1).all()
            if got is not None:
                assert mo.ids_match_with_gap(ref[0], ref[1], got, TOL).all(), (variant, company, F)
            near = np.abs(ref[1][:, 0].astype(np.float64) - 0.4) <= TOL
            assert (r.accept[~near] == ref[2][~near]).all()

    check()
    for step in range(14):
        op = rng.choice(["late", "move", "evict", "compact", "late"])
        if op == "late":                                   # a few people of one company enrolled at the end
            c = "c%d" % rng.integers(0, C)
            cnt = int(rng.integers(1, 40))
            ids = ["n%d" % (fresh + i) for i in range(cnt)]
            store.upsert(ids, G[fresh:fresh + cnt], [c] * cnt, prenormalised=True)
            for i, p in enumerate(ids):
                vec[p] = G[fresh + i]; comp[p] = c; order.append(p)
            fresh += cnt
        elif op == "move":                                 # existing ids change company, in place
            ids = [order[i] for i in rng.choice(len(order), size=5, replace=False)]
            c = "c%d" % rng.integers(0, C)
            store.upsert(ids, np.stack([vec[p] for p in ids]), [c] * len(ids), prenormalised=True)
            for p in ids:
                comp[p] = c
        elif op == "evict":
            ids = [order[i] for i in rng.choice(len(order), size=int(rng.integers(1, 300)), replace=False)]
            store.remove(ids)
            for p in ids:
                order.remove(p); del vec[p], comp[p]
        else:
            store.compact()
        check()
    assert len(store) == len(order) and store.ids() == order
    store.close()
