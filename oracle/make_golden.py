"""Regenerates tests/golden/*.npz from the reference's own code (via oracle/ref_harness.py).
TEST INFRASTRUCTURE - run in the build container only:  ``python -m oracle.make_golden``

The reference has no tests or fixtures of its own, so these files ARE the
pin: every array named ``ref_*`` was produced by unmodified reference code under stubs.
Inputs are either stored in the file (small, adversarial cases) or re-derivable from the
stored seeds through ``oracle.synth`` (larger cases), so the fixtures stay small.
"""
from __future__ import annotations

import os
from datetime import datetime, timedelta

import numpy as np

from . import matcher_oracle as mo
from . import ref_harness as rh
from . import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KIND = {"recognized": 0, "unknown": 1, "ignored": 2}


def hex_ids(n, base=0):
    return ["%024x" % (base + i) for i in range(n)]


def _campus(ids, G, Q):
    """Decisions at the shipped thresholds + the raw best (id, score) of every face, obtained by
    re-running the same verbatim code with ``recognition_threshold`` lowered to -2 (an instance
    attribute, peopleCount.py:829 - configuration, not a code change)."""
    pc = rh.load("peopleCount")
    res, stats = rh.run_campus_frame(ids, G, Q)
    kind = np.array([KIND[r["kind"]] for r in res], np.int8)
    pos = {i: r for r, i in enumerate(ids)}
    rec_row = np.array([pos[r["id"]] if r["id"] else -1 for r in res], np.int64)
    rec_score = np.array([r["score"] if r["score"] is not None else np.nan for r in res], np.float64)

    best_row = np.full(len(Q), -1, np.int64)
    best_score = np.full(len(Q), -1.0, np.float32)
    em = rh._DictManager({i: g for i, g in zip(ids, G)}, {i: {"name": i, "type": "employee"} for i in ids})
    for fi, e in enumerate(Q):
        mgr = rh._CapturingManager()
        proc = pc.CameraProcessor(em, mgr)
        proc.recognition_threshold = -2.0
        proc.face_detector = rh.FakeDetector()
        proc.face_detector.faces = [rh.FakeFace(e, fi)]
        proc.process_frame(np.zeros((4, 4, 3), np.uint8), "cam0")
        ev = [x for x in mgr.events if x["kind"] == "recognized"]
        if ev:
            best_row[fi] = pos[ev[0]["id"]]
            best_score[fi] = np.float32(ev[0]["score"])
    return dict(ref_campus_kind=kind, ref_campus_row=rec_row, ref_campus_score=rec_score,
                ref_best_row=best_row, ref_best_score=best_score,
                ref_campus_stats=np.array([stats["faces"], stats["recognized"], stats["unknown"]]))


def _live(ids, G, Q):
    pos = {i: r for r, i in enumerate(ids)}
    cap = rh.run_live_frame(ids, G, Q)
    row = np.array([pos.get(c["name"], -1) for c in cap], np.int64)
    score = np.array([np.float32(c["score"]) for c in cap], np.float32)
    return dict(ref_live_row=row, ref_live_score=score)


def golden_cfg1():
    """BASELINE config 1: 10 000 x 512 gallery, 64 faces, top-1 + threshold (0.45 / 0.4)."""
    n, f, d = 10000, 64, 512
    G = synth.gallery(n, d)
    Q, target = synth.queries(f, n, d)
    ids = hex_ids(n)
    out = dict(n=n, f=f, dim=d, gallery_seed=synth.GALLERY_SEED, query_seed=synth.QUERY_SEED,
               target=target, gallery_checksum=np.float64(G.astype(np.float64).sum()),
               query_checksum=np.float64(Q.astype(np.float64).sum()))
    out.update(_campus(ids, G, Q))
    out.update(_live(ids, G, Q))
    np.savez_compressed(os.path.join(OUT, "cfg1_10k_x_64.npz"), **out)
    print("cfg1: recognized/unknown/ignored =", np.bincount(out["ref_campus_kind"], minlength=3))


def _ulp_neighbours(x: float, span: int = 2):
    c = np.float32(x)
    lo = c
    vals = [c]
    for _ in range(span):
        lo = np.nextafter(lo, np.float32(-2))
        vals.insert(0, lo)
    hi = c
    for _ in range(span):
        hi = np.nextafter(hi, np.float32(2))
        vals.append(hi)
    return vals


def golden_edges():
    """Adversarial, self-contained: exact ties, NaN row, score == -1, scores within 2 ulp of every
    threshold the reference uses.  Row j of the 'edge' block is (c_j, sqrt(1-c_j^2)) on its own
    coordinate pair and query j is the basis vector of the first coordinate, so the score is
    exactly c_j in ANY fp32 arithmetic (one non-zero product)."""
    d = 512
    rows, queries = [], []
    # block A: threshold edges
    cs = []
    for thr in (mo.LIVE_THRESHOLD, mo.CAMPUS_THRESHOLD, mo.CAMPUS_UNKNOWN, mo.UNKNOWN_CLUSTER):
        cs += _ulp_neighbours(thr, 2)
    for j, c in enumerate(cs):
        g = np.zeros(d, np.float32)
        g[2 * j] = c
        g[2 * j + 1] = np.float32(np.sqrt(1.0 - float(c) ** 2))
        rows.append(g)
        q = np.zeros(d, np.float32)
        q[2 * j] = 1.0
        queries.append(q)
    n_edge = len(rows)
    base = 2 * n_edge
    # block B: three identical rows (exact tie) + a slightly worse one; query hits them at 0.9
    t = np.zeros(d, np.float32)
    t[base] = 0.9
    t[base + 1] = np.float32(np.sqrt(1 - 0.81))
    worse = np.zeros(d, np.float32)
    worse[base] = 0.8
    worse[base + 2] = 0.6
    rows += [worse.copy(), t.copy(), t.copy(), worse.copy(), t.copy()]
    q = np.zeros(d, np.float32)
    q[base] = 3.0                      # un-normalised on purpose: the matcher normalises (a1)
    queries.append(q)
    # block C: a NaN row (what a zero template becomes at load) and the antipode of a query
    nan_row = np.full(d, np.nan, np.float32)
    anti = np.zeros(d, np.float32)
    anti[base + 8] = -1.0
    rows += [nan_row, anti]
    q = np.zeros(d, np.float32)
    q[base + 8] = 1.0                  # best score over the whole gallery is... see below
    queries.append(q)
    G = np.stack(rows)
    Q = np.stack(queries)
    ids = hex_ids(len(G))
    out = dict(gallery=G, queries=Q, n_edge=n_edge, edge_scores=np.array(cs, np.float32))
    out.update(_campus(ids, G, Q))
    out.update(_live(ids, G, Q))
    # a gallery holding ONLY the antipode and the NaN row: nothing can ever match (score -1 / NaN)
    G2 = np.stack([anti, nan_row])
    out2 = {("nomatch_" + k): v for k, v in {**_campus(hex_ids(2), G2, Q[-1:]), **_live(hex_ids(2), G2, Q[-1:])}.items()}
    out.update(out2)
    np.savez_compressed(os.path.join(OUT, "edge_cases.npz"), **out)
    print("edges: kinds =", out["ref_campus_kind"], "nomatch kind =", out["nomatch_ref_campus_kind"])


def golden_managers():
    """Gallery residency / enrol / update / evict / tenant subset, through both EmbeddingManagers."""
    d = 512
    raw = synth.raw_rows(np.arange(40), d, 777, synth.STREAM_GALLERY)
    # what the enrol worker stores: fp32 mean of 3 unit pose vectors -> norm < 1 (trainingServer.py:355)
    unit = raw / np.linalg.norm(raw, axis=1, keepdims=True)
    pose_noise = synth.raw_rows(np.arange(120), d, 778, synth.STREAM_NOISE).reshape(40, 3, d) / np.float32(synth.SIGMA)
    poses = unit[:, None, :] + np.float32(0.02) * pose_noise
    poses = poses / np.linalg.norm(poses, axis=2, keepdims=True)
    stored = np.mean(poses.astype(np.float32), axis=1).astype(np.float32)
    stored[2] = 0.0                                     # a zero template -> NaN row at load
    A, B = "a" * 24, "b" * 24
    t0 = datetime(2026, 1, 1)

    def emp(i, **kw):
        return dict(id="%024x" % i, vec=stored[i], company=A if i < 8 or i >= 12 else B, lastUpdated=t0, **kw)

    def vis(i, **kw):
        return dict(id="%024x" % (100 + i), vec=stored[20 + i], company=A if i < 3 else B, lastUpdated=t0, **kw)

    employees = [emp(i) for i in range(12)]
    employees[3]["status"] = "inactive"
    employees[5]["blacklisted"] = True
    employees[6]["emb_status"] = "pending"
    visitors = [vis(i) for i in range(6)]
    visitors[4]["emb_status"] = "pending"

    out = dict(stored=stored)

    def snap(prefix, m):
        ids = list(m.embeddings.keys())
        out[prefix + "_ids"] = np.array(ids)
        out[prefix + "_G"] = np.stack([m.embeddings[i] for i in ids]).astype(np.float32)

    # ---- manager A (infrenceServer.EmbeddingManager)
    rh.reset_database()
    db = rh.seed_people(employees, visitors)
    m = rh.live_manager()
    snap("ref_live_load", m)
    st = m.get_stats()
    out["ref_live_load_stats"] = np.array([st["total_embeddings"], st["employees"], st["visitors"]])

    later = datetime.utcnow() + timedelta(seconds=5)
    docs = {str(x["_id"]): x for x in db["employeeInfo"].docs}
    vdocs = {str(x["_id"]): x for x in db["visitors"].docs}
    efs = rh.FakeGridFS(db, collection="employee_embeddings")
    vfs = rh.FakeGridFS(db, collection="visitor_embeddings")
    import pickle

    def set_vec(doc, field, fs, v):
        fid = fs.put(pickle.dumps(v))
        doc[field]["buffalo_l"]["embeddingId"] = fid
        doc["lastUpdated"] = later

    # step 1: e1 re-enrolled; e0 deactivated; e12 new; v1 re-enrolled; e3 re-activated
    set_vec(docs["%024x" % 1], "employeeEmbeddings", efs, stored[30])
    docs["%024x" % 0]["status"] = "inactive"
    rh.seed_people([dict(id="%024x" % 12, vec=stored[12], company=A, lastUpdated=later)], [])
    set_vec(vdocs["%024x" % 101], "visitorEmbeddings", vfs, stored[31])
    docs["%024x" % 3]["status"] = "active"
    docs["%024x" % 3]["lastUpdated"] = later
    m.force_sync()
    snap("ref_live_sync1", m)
    # step 2: e0 comes back -> re-appended at the END of the order
    later2 = datetime.utcnow() + timedelta(seconds=10)
    docs["%024x" % 0]["status"] = "active"
    docs["%024x" % 0]["lastUpdated"] = later2
    m.force_sync()
    snap("ref_live_sync2", m)
    ea, _ = m.get_embeddings_for_company(A)
    eb, _ = m.get_embeddings_for_company(B)
    out["ref_live_tenant_a"] = np.array(sorted(ea.keys()))
    out["ref_live_tenant_b"] = np.array(sorted(eb.keys()))
    # recognise through the real manager with the tenant filter (infrenceServer.py:515-563)
    srv = rh.load("infrenceServer")
    Qm = (unit[[1, 30, 9, 21, 12, 0, 33]] + np.float32(0.01) * pose_noise[[1, 30, 9, 21, 12, 0, 33], 0]).astype(np.float32)
    out["mgr_queries"] = Qm
    for tag, comp in (("a", A), ("b", B)):
        proc = srv.FaceRecognitionProcessor(m)
        proc.face_detector = rh.FakeDetector()
        proc.face_detector.faces = [rh.FakeFace(e, i) for i, e in enumerate(Qm)]
        cap = []
        proc.draw_enhanced_bounding_box = lambda fr, bb, col, info, ds, rs, cap=cap: (cap.append((info["name"], rs)), fr)[1]
        proc.recognize_faces(np.zeros((4, 4, 3), np.uint8), comp)
        names = {str(x["_id"]): x.get("employeeName", x.get("visitorName")) for x in db["employeeInfo"].docs + db["visitors"].docs}
        # metadata 'name' is the display name, not the id; map back through the manager's table
        ids_now = list(m.embeddings.keys())
        best = []
        for (nm, rs) in cap:
            best.append(np.float32(rs))
        out["ref_live_match_%s_score" % tag] = np.array(best, np.float32)
        out["ref_live_match_%s_known" % tag] = np.array([nm != "Unknown" for nm, _ in cap])

    # ---- manager B (peopleCount.EmbeddingManager): full reload, never evicts
    rh.reset_database()
    db = rh.seed_people(employees, visitors)
    mb = rh.campus_manager()
    snap("ref_campus_load", mb)
    docs = {str(x["_id"]): x for x in db["employeeInfo"].docs}
    efs = rh.FakeGridFS(db, collection="employee_embeddings")
    set_vec(docs["%024x" % 1], "employeeEmbeddings", efs, stored[30])
    docs["%024x" % 0]["status"] = "inactive"
    rh.seed_people([dict(id="%024x" % 12, vec=stored[12], company=A)], [])
    # body of the sync loop, peopleCount.py:771-773
    mb._load_embeddings(mb._get_all_active_employees(), mb._get_all_visitors())
    snap("ref_campus_sync1", mb)
    e_all, _ = mb.get_all()
    out["ref_campus_get_all_ids"] = np.array(list(e_all.keys()))
    np.savez_compressed(os.path.join(OUT, "managers.npz"), **out)
    print("managers: live load ids", len(out["ref_live_load_ids"]), "-> sync1", len(out["ref_live_sync1_ids"]),
          "-> sync2", len(out["ref_live_sync2_ids"]), "| campus", len(out["ref_campus_load_ids"]),
          "->", len(out["ref_campus_sync1_ids"]))


def golden_unknown_clusters():
    f, people, d = 60, 7, 512
    Q, target = synth.queries(f, people, d, seed=99, genuine_every=1, noise=0.02)
    Qn = mo.normalise_rows(Q)
    res, avgs = rh.run_unknown_clustering(Qn)
    np.savez_compressed(os.path.join(OUT, "unknown_clusters.npz"), f=f, people=people, dim=d, seed=99,
                        noise=0.02, target=target, ref_cluster=np.array([r[0] for r in res]),
                        ref_created=np.array([r[1] for r in res]), ref_avgs=np.stack(avgs).astype(np.float32))
    print("unknown clusters:", len(avgs), "clusters from", f, "observations")


def golden_enrol_checks():
    d = 512
    G = synth.gallery(300, d, seed=55)
    stored = (G * np.float32(0.93)).astype(np.float32)      # raw templates have norm < 1
    probes = []
    dup = []
    for j, (row, noise) in enumerate([(17, 0.02), (250, 0.05), (-1, 0.0), (5, 0.08), (299, 0.0)]):
        if row >= 0:
            nz = synth.raw_rows([1000 + j], d, 56, synth.STREAM_NOISE)[0] / np.float32(synth.SIGMA)
            p = (G[row] + np.float32(noise) * nz).astype(np.float32) * np.float32(0.9)
        else:
            p = synth.unit_rows([777], d, 57, synth.STREAM_IMPOSTOR)[0]
        probes.append(p)
        r, _, _ = rh.run_enrol_checks(p, list(stored), [p])
        dup.append(r)
    # same-person check on pose triples
    triples, ok, pair = [], [], []
    for j, (a, b, c) in enumerate([(1, 1, 1), (1, 1, 2), (3, 4, 3)]):
        nz = synth.raw_rows(np.arange(3) + 10 * j, d, 58, synth.STREAM_NOISE) / np.float32(synth.SIGMA)
        tr = (G[[a, b, c]] + np.float32(0.03) * nz).astype(np.float32)
        _, o, p = rh.run_enrol_checks(tr[0], [], list(tr))
        triples.append(tr)
        ok.append(o)
        pair.append(p if p else (-1, -1))
    np.savez_compressed(os.path.join(OUT, "enrol_checks.npz"), n=300, dim=d, seed=55, scale=np.float32(0.93),
                        probes=np.stack(probes), ref_duplicate=np.array(dup), triples=np.stack(triples),
                        ref_same_ok=np.array(ok), ref_same_pair=np.array(pair))
    print("enrol checks: duplicates", dup, "same-person", ok, pair)


def main():
    if not rh.available():
        raise SystemExit("reference tree not found at %s" % rh.REFERENCE_ROOT)
    os.makedirs(OUT, exist_ok=True)
    golden_cfg1()
    golden_edges()
    golden_managers()
    golden_unknown_clusters()
    golden_enrol_checks()


if __name__ == "__main__":
    main()
