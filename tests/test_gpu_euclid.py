"""BASELINE config 3 - 128-d Euclidean (dlib / face_recognition style) - on the tensor-core path.

NOT in the reference: the metric, its oracle (oracle.matcher_oracle.euclidean_topk) and these
expectations are OUR definition ("parity unpinned").  What is checked:
  * FRG_VARIANT_TC_EXACT with the Euclidean metric (tcgen05 filter on S = q.g - 0.5*||g||^2 over the raw
    store's augmented scan plane + exact direct-difference rescoring) against the fp64 oracle:
    rows exact wherever adjacent oracle distances differ by more than TOL = 1e-4, |delta d| <= TOL;
  * BIT-identical rows, distances and decisions between the tensor-core path and the exact fp32 scan
    (same rescoring arithmetic and order), on every shape the tile code covers;
  * the awkward cases: exact duplicates (ties -> lowest rows), d = 0, rows of very different norms,
    NaN rows, removed rows, tenants, in-place overwrite (norm terms follow), compaction,
    candidate overflow -> exact fallback on the device.
"""
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


def check(frg, store, Q, G, k, tol_d, tags=None, tenant=None, company=None, variant="tc_exact"):
    m = frg.Matcher(store, metric="euclidean")
    r = m.match(Q, k, tol_d, company_id=company, variant=variant)
    assert r.variant == ("scan_f32" if variant == "scan_f32" else "tc_exact")
    kk = min(k + 1, max(len(G), 1))
    ref_rows, ref_d, ref_acc = mo.euclidean_topk(Q, G, max(kk, k), tol_d, tags, tenant)
    assert mo.ids_match_with_gap(ref_rows, -ref_d, r.rows, TOL).all()
    filled = ref_rows[:, :k] >= 0
    assert np.abs(r.scores[filled] - ref_d[:, :k][filled]).max(initial=0) <= TOL
    assert (r.rows[~filled] == -1).all() and np.isinf(r.scores[~filled]).all()
    near = np.abs(ref_d[:, 0].astype(np.float64) - tol_d) <= TOL
    assert (r.accept[~near] == ref_acc[~near]).all()
    return r


def same_as_scan(frg, store, Q, k, tol_d, company=None):
    m = frg.Matcher(store, metric="euclidean")
    a = m.match(Q, k, tol_d, company_id=company, variant="tc_exact")
    b = m.match(Q, k, tol_d, company_id=company, variant="scan_f32")
    assert a.variant == "tc_exact" and b.variant == "scan_f32"
    assert np.array_equal(a.rows, b.rows)
    assert np.array_equal(a.scores.view(np.uint32), b.scores.view(np.uint32))      # bit for bit
    assert np.array_equal(a.accept, b.accept)
    return a


@pytest.mark.parametrize("variant", ["scan_f32", "tc_exact", "auto"])
@pytest.mark.parametrize("k", [1, 3])
def test_config3_small(frg, k, variant):
    rng = np.random.default_rng(3)
    n, d = 30000, 128
    G = (rng.standard_normal((n, d)) * 0.1).astype(np.float32)
    Q = (G[[5, 999, 12345]] + 0.02 * rng.standard_normal((3, d))).astype(np.float32)
    Q = np.concatenate([Q, (rng.standard_normal((2, d)) * 0.1).astype(np.float32)])
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.append_rows(G)
    r = check(frg, store, Q, G, k, 0.6, variant=variant)
    assert list(r.rows[:3, 0]) == [5, 999, 12345]
    assert r.accept[:3].all() and not r.accept[3:].any()
    store.close()


@pytest.mark.parametrize("d,n,f,k", [(128, 200_003, 300, 5), (128, 50_001, 7, 1), (256, 70_000, 129, 16),
                                     (256, 40_000, 64, 2), (128, 1, 3, 2), (128, 129, 1, 16)])
def test_tc_equals_exact_scan_bit_for_bit(frg, d, n, f, k):
    """Synthetic unit rows kept raw (what config 3's probe uses); queries half genuine, half impostor."""
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.fill_synthetic(n, 0, 99)
    Q, target = synth.queries(f, n, d, seed=5, gallery_seed=99)
    r = same_as_scan(frg, store, Q, k, 0.6)
    hit = target >= 0
    if n > 1000:
        assert (r.rows[hit, 0] == target[hit]).all()
    G, _ = store.read_rows()
    Q, _ = synth.queries(min(f, 24), n, d, seed=6, gallery_seed=99)
    check(frg, store, Q, G, k, 0.6)
    store.close()


def test_rows_of_very_different_norms(frg):
    """The filter's bound scales with ||q|| * max||g||: rows with norms 0.05 .. 8 in one gallery,
    queries of norms 0.1 .. 6, near and far."""
    rng = np.random.default_rng(11)
    n, d = 60_000, 128
    G = rng.standard_normal((n, d)).astype(np.float32)
    G *= (rng.uniform(0.05, 8.0, (n, 1)) / np.linalg.norm(G, axis=1, keepdims=True)).astype(np.float32)
    near = rng.integers(0, n, 40)
    Q = G[near] + (0.01 * rng.standard_normal((40, d))).astype(np.float32)
    far = rng.standard_normal((24, d)).astype(np.float32) * rng.uniform(0.01, 0.6, (24, 1)).astype(np.float32)
    Q = np.concatenate([Q, far]).astype(np.float32)
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.append_rows(G)
    for k in (1, 5):
        r = check(frg, store, Q, G, k, 0.6)
        same_as_scan(frg, store, Q, k, 0.6)
    assert (r.rows[:40, 0] == near).all()
    store.close()


def test_ties_zero_distance_nan_rows_removed_rows_tenants(frg):
    rng = np.random.default_rng(12)
    n, d = 20_011, 128
    G = (rng.standard_normal((n, d)) * 0.1).astype(np.float32)
    dup = np.array([17, 300, 301, 4000, 19_999])
    G[dup[1:]] = G[dup[0]]                                  # exact ties: lowest rows first
    G[123] = np.nan                                         # never matches (exact scan: NaN never inserted)
    tags = rng.integers(0, 3, n).astype(np.int32)
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.upsert(["id%d" % i for i in range(n)], G, ["T%d" % t for t in tags])
    removed = [300, 5, 6]
    store.remove(["id%d" % i for i in removed])
    Gd, tg = store.read_rows()
    Q = np.stack([G[17], G[17] + 1e-3, G[5], G[9000], rng.standard_normal(d) * 0.1]).astype(np.float32)
    for company in (None, "T1"):
        tenant = None if company is None else store.tenant_code(company, create=False)
        for k in (1, 4, 16):
            r = check(frg, store, Q, Gd, k, 0.6, tg, tenant, company)
            same_as_scan(frg, store, Q, k, 0.6, company)
            assert not np.isin(r.rows, removed + [123]).any()
    r = frg.Matcher(store, metric="euclidean").match(Q, 4, 0.6, variant="tc_exact")
    assert list(r.rows[0]) == [17, 301, 4000, 19_999] and (r.scores[0] == 0).all()
    assert list(r.rows[1]) == [17, 301, 4000, 19_999]
    store.close()


def test_overwrite_in_place_and_compaction_keep_the_norm_terms(frg):
    rng = np.random.default_rng(13)
    n, d = 10_000, 128
    G = (rng.standard_normal((n, d)) * 0.1).astype(np.float32)
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.upsert(["p%d" % i for i in range(n)], G)
    m = frg.Matcher(store, metric="euclidean")
    # same id, new template with a 3x larger norm: the row keeps its position, its bias columns follow
    big = (rng.standard_normal(d) * 0.3).astype(np.float32)
    store.upsert(["p77"], big[None])
    G[77] = big
    Q = np.stack([big + 0.01, G[78], G[76]]).astype(np.float32)
    r = check(frg, store, Q, G, 3, 0.6)
    assert list(r.rows[:, 0]) == [77, 78, 76]
    same_as_scan(frg, store, Q, 3, 0.6)
    # evict a third of the gallery, squeeze it: rows move, the augmented plane moves with them
    gone = list(range(0, n, 3))
    store.remove(["p%d" % i for i in gone])
    store.compact()
    Gc, tags = store.read_rows()
    keep = np.setdiff1d(np.arange(n), gone)
    assert np.array_equal(Gc, G[keep])
    r = check(frg, store, Q, Gc, 3, 0.6, tags)
    assert r.ids[0][0] == "p77"
    same_as_scan(frg, store, Q, 3, 0.6)
    store.close()


def test_one_outlier_row_of_huge_norm_does_not_loosen_the_other_rows_bound(frg):
    """The filter's bound is per ROW (r2): every row carries its own rounding residual, norm and squared norm in the
    plane's pad columns, the query image the matching coefficients, so the tensor core itself returns upper (filter)
    and lower (pre-pass) bounds of the exact score.  One row 30x longer than the rest used to loosen a store-wide
    bound until every row was a candidate of every query and all of them fell back to the exact scan; now it
    only widens its own interval: results right, bit-identical to the exact scan, and the fallback stays idle."""
    from facerecognition_infrenceengine_b200 import _native as N
    rng = np.random.default_rng(14)
    n, d = 40_000, 128
    G = (rng.standard_normal((n, d)) * 0.1).astype(np.float32)
    G[31_000] = (rng.standard_normal(d) * 3.0).astype(np.float32)
    G[100] = (rng.standard_normal(d) * 20.0).astype(np.float32)
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.append_rows(G)
    Q = np.concatenate([G[[31_000, 4, 39_999, 100]] + np.float32(0.01),
                        (rng.standard_normal((150, d)) * 0.1).astype(np.float32)])
    for k in (1, 5):
        r = check(frg, store, Q, G, k, 0.6)
        same_as_scan(frg, store, Q, k, 0.6)
    assert list(r.rows[:4, 0]) == [31_000, 4, 39_999, 100]
    # the exact fallback had nothing to redo: its launch stays at the idle few microseconds
    m = frg.Matcher(store, metric="euclidean")
    m.match(Q, 5, 0.6, variant="tc_exact")
    N.profile_enable(True)
    N.profile_collect()
    for _ in range(5):
        m.match(Q, 5, 0.6, variant="tc_exact")
    N.profile_collect()
    st = N.profile_stages()
    N.profile_enable(False)
    assert st["fallback"] / 5 < 0.05, st            # ms per match (a redo of 154 queries takes ~0.5 ms here)
    store.close()


def test_candidate_overflow_falls_back_to_exact_scan(frg):
    d, n = 128, 60_000
    G = synth.gallery(n, d, 31)
    dup = np.arange(1000, 9000, 2)                          # 4000 copies of row 7
    G[dup] = G[7]
    store = frg.GalleryStore(dim=d, capacity=n, raw=True)
    store.append_rows(G)
    Q, _ = synth.queries(6, n, d, seed=8, gallery_seed=31)
    Q[2] = G[7] + 0.001
    Q[4] = G[7]
    for k in (1, 5, 16):
        r = check(frg, store, Q, G, k, 0.6)
        want = sorted([7] + list(dup))[:k]
        assert list(r.rows[4]) == want and list(r.rows[2]) == want
    store.close()


def test_refusals(frg):
    d = 128
    raw = frg.GalleryStore(dim=d, capacity=64, raw=True)
    raw.append_rows(np.eye(8, d, dtype=np.float32))
    unit = frg.GalleryStore(dim=d, capacity=64)
    unit.append_rows(np.eye(8, d, dtype=np.float32))
    wide = frg.GalleryStore(dim=512, capacity=64, raw=True)       # 512 + 64 columns: no tile shape
    wide.append_rows(np.eye(8, 512, dtype=np.float32))
    q = np.eye(2, d, dtype=np.float32)
    with pytest.raises(frg.NativeError):                          # distances from bf16 scores would cancel
        frg.Matcher(raw, metric="euclidean").match(q, 1, 0.6, variant="tc_bf16")
    with pytest.raises(frg.NativeError):                          # unit store: no norm terms in its plane
        frg.Matcher(unit, metric="euclidean").match(q, 1, 0.6, variant="tc_exact")
    with pytest.raises(frg.NativeError):                          # raw store: the cosine bound needs unit rows
        frg.Matcher(raw, metric="cosine").match(q, 1, 0.4, variant="tc_exact")
    with pytest.raises(frg.NativeError):
        frg.Matcher(wide, metric="euclidean").match(np.eye(2, 512, dtype=np.float32), 1, 0.6, variant="tc_exact")
    # AUTO still answers all of them, by the exact scan
    for st, mt, qq in ((unit, "euclidean", q), (raw, "cosine", q), (wide, "euclidean", np.eye(2, 512, dtype=np.float32))):
        r = frg.Matcher(st, metric=mt).match(qq, 1, 0.6)
        assert r.variant == "scan_f32" and list(r.rows[:, 0]) == [0, 1]
    r = frg.Matcher(raw, metric="euclidean").match(q, 1, 0.6)
    assert r.variant == "tc_exact" and list(r.rows[:, 0]) == [0, 1] and (r.scores[:, 0] == 0).all()
    for st in (raw, unit, wide):
        st.close()
