"""Row-sharded match on real GPUs.  world=1 runs everywhere; the 2-rank NCCL case needs >= 2 devices
(it is launched as its own torchrun job) and is skipped otherwise."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import matcher_oracle as mo
from oracle import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_world1_equals_plain_matcher():
    import torch
    import facerecognition_infrenceengine_b200 as frg
    from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
    n, d, f, k = 50000, 512, 24, 5
    store = frg.GalleryStore(dim=d, capacity=n)
    g = ShardedGallery(dim=d, device=0, store=store, rank=0, world=1)
    g.fill_synthetic(n, 1234)
    Q, _ = synth.queries(f, n, d)
    rows, scores, acc = ShardedMatcher(g).match(torch.from_numpy(Q).cuda(), k, 0.45)
    torch.cuda.synchronize()
    ref = frg.Matcher(store).match(Q, k, 0.45)
    assert (rows.cpu().numpy() == ref.rows).all() and (scores.cpu().numpy() == ref.scores).all()
    assert (acc.cpu().numpy().astype(bool) == ref.accept).all()
    store.close()


def test_sharded_world1_enrolment_and_tenants():
    """ShardedGallery's id-level API on one rank equals GalleryStore's (same dict semantics), odd F*k."""
    import torch
    import facerecognition_infrenceengine_b200 as frg
    from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
    rng = np.random.default_rng(3)
    n, d = 3001, 512
    ids = ["%024x" % (i + 1) for i in range(n)]
    V = rng.standard_normal((n, d)).astype(np.float32)
    comp = ["acme" if i % 2 else "globex" for i in range(n)]
    plain = frg.GalleryStore(dim=d, capacity=n + 8)
    plain.upsert(ids, V, comp)
    store = frg.GalleryStore(dim=d, capacity=n + 8)
    g = ShardedGallery(dim=d, device=0, store=store, rank=0, world=1)
    g.load(ids, V, comp)
    for tgt in (plain, g):
        tgt.upsert([ids[4], "x", "y"], V[:3], ["globex", "acme", "acme"])
        tgt.remove([ids[9], "ghost"])
        tgt.upsert([ids[9]], V[5:6], ["acme"])
    Q = (V[[4, 0, 5]] + np.float32(0.02) * rng.standard_normal((3, d)).astype(np.float32))
    m, pm = ShardedMatcher(g), frg.Matcher(plain)
    for company in (None, "acme", "globex", "nobody"):
        rows, scores, acc = m.match(torch.from_numpy(Q).cuda(), 1, 0.4, company_id=company)
        torch.cuda.synchronize()
        ref = pm.match(Q, 1, 0.4, company_id=company)
        assert (rows.cpu().numpy() == ref.rows).all() and (scores.cpu().numpy() == ref.scores).all()
        assert m.ids_of(rows) == ref.ids
    store.close(); plain.close()


def _manager_scenario(frg, gal, matcher, src_factory, g, rank=0):
    """The reference's own EmbeddingManager scenario (tests/golden/managers.npz, produced by the unmodified
    infrenceServer.py under stubs) over a sharded gallery: ids order, company subsets and the match results of
    FaceRecognitionProcessor.recognize_faces.  Only `rank` 0 edits the documents (BroadcastSource ships them)."""
    from datetime import datetime, timedelta, timezone
    st = g["stored"]
    A, B = "a" * 24, "b" * 24
    t0 = datetime(2026, 1, 1)

    def emp(i):
        return {"_id": "%024x" % i, "embedding": st[i], "companyId": A if i < 8 or i >= 12 else B, "status": "active",
                "blacklisted": False, "lastUpdated": t0, "employeeName": "e%d" % i}

    def vis(i):
        return {"_id": "%024x" % (100 + i), "embedding": st[20 + i], "companyId": A if i < 3 else B,
                "lastUpdated": t0, "visitorName": "v%d" % i}

    E = [emp(i) for i in range(12)]
    E[3]["status"] = "inactive"; E[5]["blacklisted"] = True; E[6]["embedding_status"] = "pending"
    V = [vis(i) for i in range(6)]
    V[4]["embedding_status"] = "pending"
    m = frg.EmbeddingManager(src_factory(E, V), mode="live", store=gal)

    def same(prefix):
        ref_ids, ref_G = list(g[prefix + "_ids"]), g[prefix + "_G"]
        assert gal.ids() == ref_ids
        lo, hi = gal.bounds[gal.rank]
        vecs, tags = gal.store.read_rows(0, hi - lo)
        for i, pid in enumerate(ref_ids):
            r = gal.row_of(pid)
            if lo <= r < hi:
                a, b = vecs[r - lo], ref_G[i]
                assert tags[r - lo] >= 0 and (np.isnan(a) == np.isnan(b)).all()
                assert np.isnan(b).all() or np.nanmax(np.abs(a - b)) <= 2e-7

    same("ref_live_load")
    s = m.get_stats()
    assert [s["total_embeddings"], s["employees"], s["visitors"]] == list(g["ref_live_load_stats"])
    later = datetime.now(timezone.utc).replace(tzinfo=None) + timedelta(seconds=5)
    if rank == 0:
        E[1]["embedding"] = st[30]; E[1]["lastUpdated"] = later
        E[0]["status"] = "inactive"
        E.append(dict(emp(12), lastUpdated=later))
        V[1]["embedding"] = st[31]; V[1]["lastUpdated"] = later
        E[3]["status"] = "active"; E[3]["lastUpdated"] = later
    m.force_sync()
    same("ref_live_sync1")
    if rank == 0:
        E[0]["status"] = "active"
        E[0]["lastUpdated"] = datetime.now(timezone.utc).replace(tzinfo=None) + timedelta(seconds=10)
    m.force_sync()
    same("ref_live_sync2")
    proc = frg.FaceRecognitionProcessor(gal, matcher=matcher)
    for comp, tag in ((A, "a"), (B, "b")):
        view = m.get_embeddings_for_company(comp)
        assert len(view) == len(g["ref_live_tenant_%s" % tag])
        out = proc.recognize(g["mgr_queries"], view.company_id)
        assert [x["person_id"] is not None for x in out] == list(g["ref_live_match_%s_known" % tag])
        assert np.abs(np.array([x["recognition_score"] for x in out], np.float32)
                      - g["ref_live_match_%s_score" % tag]).max() <= 1e-4


def test_sharded_world1_embedding_manager_golden(golden):
    import facerecognition_infrenceengine_b200 as frg
    from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
    g = golden("managers.npz")
    store = frg.GalleryStore(dim=g["stored"].shape[1], capacity=64)
    gal = ShardedGallery(dim=store.dim, device=0, store=store, rank=0, world=1)
    _manager_scenario(frg, gal, ShardedMatcher(gal), lambda E, V: frg.ListSource(E, V), g)
    store.close()


WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import facerecognition_infrenceengine_b200 as frg
from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
from oracle import matcher_oracle as mo, synth
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d, f, k = 300001, 512, 64, 10
store = frg.GalleryStore(dim=d, capacity=n, device=local)
g = ShardedGallery(dim=d, device=local, store=store)
g.fill_synthetic(n, 1234)
Q, target = synth.queries(f, n, d)
Qd = torch.from_numpy(Q).cuda()
m_nccl = ShardedMatcher(g, exchange="nccl")
rows, scores, acc = m_nccl.match(Qd, k, 0.45)
torch.cuda.synchronize()
assert m_nccl.exchange == "nccl"
if dist.get_rank() == 0:
    G = synth.gallery(n, d)
    rr, rs, ra = mo.match_topk(Q, G, k + 1, 0.45)
    assert mo.ids_match_with_gap(rr, rs, rows.cpu().numpy(), 1e-4).all()
    assert np.abs(scores.cpu().numpy() - rs[:, :k]).max() <= 1e-4
    assert (acc.cpu().numpy().astype(bool) == ra).all()
    print("SHARDED_OK world=%%d" %% dist.get_world_size())
# the peer-memory exchange (pushes from the match kernels, a poll-only merge kernel, no collective call) gives the same
# answer bit for bit, call after call (epoch parity reuse), for changing batch sizes (buffer regrowth), and
# when one rank runs late
m_p2p = ShardedMatcher(g, exchange="p2p")
for it, (ff, kk) in enumerate([(64, 10), (64, 10), (64, 10), (8, 1), (200, 5), (64, 10), (2, 16)] * 3):
    Qi = Qd[:ff] if ff <= f else torch.from_numpy(synth.queries(ff, n, d, seed=77 + it)[0]).cuda()
    if it %% 4 == dist.get_rank():
        torch.cuda._sleep(200_000_000)              # ~0.1 s of GPU time: this rank arrives late
    a = m_p2p.match(Qi, kk, 0.45)
    b = m_nccl.match(Qi, kk, 0.45)
    torch.cuda.synchronize()
    assert m_p2p.exchange == "p2p", m_p2p.p2p_error
    for x, y in zip(a, b):
        assert torch.equal(x, y), (it, ff, kk)
# variants without a select stage: the exchange kernel pushes every query itself
a = m_p2p.match(Qd, 5, 0.45, variant="scan_f32"); b = m_nccl.match(Qd, 5, 0.45, variant="scan_f32")
torch.cuda.synchronize()
for x, y in zip(a, b):
    assert torch.equal(x, y)
# adversarial shard: 3000 copies of one template land on the LAST rank; the queries near it overflow that
# rank's candidate lists, are redone by its exact fallback, and reach the other ranks through the late push
dupe = synth.unit_rows(np.arange(1), d, 4242, synth.STREAM_IMPOSTOR)
g.append_local(np.repeat(dupe, 3000, axis=0), prenormalised=True)
Qx = torch.from_numpy(np.concatenate([dupe, dupe + np.float32(1e-3), Q[:14]])).cuda()
for kk in (1, 16):
    a = m_p2p.match(Qx, kk, 0.45); b = m_nccl.match(Qx, kk, 0.45)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y), kk
    assert a[0][0].tolist() == list(range(n, n + kk)), a[0][0].tolist()      # ties -> earliest rows, global numbering
# id-level enrolment over the shards (SURVEY 8e): load in balanced blocks, overwrite on the owner rank, append
# on the last rank, tombstone, tenant filter - against the reference's dict semantics, with odd F*k
rng = np.random.default_rng(9)
n2 = 5001
ids2 = ["%%024x" %% (i + 7) for i in range(n2)]
V2 = rng.standard_normal((n2, d)).astype(np.float32)
comp2 = ["acme" if i %% 3 else "globex" for i in range(n2)]
store2 = frg.GalleryStore(dim=d, capacity=n2 + 64, device=local)
g2 = ShardedGallery(dim=d, device=local, store=store2)
g2.load(ids2, V2, comp2)
refv = {p: mo.normalise(v) for p, v in zip(ids2, V2)}; refc = dict(zip(ids2, comp2)); order = list(ids2)
up = [ids2[3], ids2[n2 - 3], "new-a", "new-b", "new-a"]; upv = rng.standard_normal((5, d)).astype(np.float32)
upc = ["acme", "globex", "acme", "initech", "acme"]
g2.upsert(up, upv, upc)
for p_, v_, c_ in zip(up, upv, upc):
    if p_ not in refv: order.append(p_)
    refv[p_] = mo.normalise(v_); refc[p_] = c_
assert g2.remove([ids2[10], ids2[n2 - 10], "nobody"]) == 2
for p_ in (ids2[10], ids2[n2 - 10]):
    order.remove(p_); del refv[p_], refc[p_]
g2.upsert([ids2[10]], upv[:1], ["acme"]); order.append(ids2[10]); refv[ids2[10]] = mo.normalise(upv[0]); refc[ids2[10]] = "acme"
assert g2.total_rows == n2 + 3 and len(g2) == len(order)
probe = [ids2[3], "new-a", ids2[10]]                       # F = 3, k = 1: odd F*k
Qp = torch.from_numpy(np.stack([refv[p_] + np.float32(0.01) * rng.standard_normal(d).astype(np.float32) for p_ in probe])).cuda()
for mm in (ShardedMatcher(g2, exchange="p2p"), ShardedMatcher(g2, exchange="nccl")):
    for company in (None, "acme", "globex", "initech", "no-such-company"):
        rows, scores, acc = mm.match(Qp, 1, 0.4, company_id=company)
        torch.cuda.synchronize()
        got = mm.ids_of(rows)
        sub = {p_: refv[p_] for p_ in order if company is None or refc[p_] == company}
        for fi in range(3):
            want = mo.scan_best(mo.normalise(Qp[fi].cpu().numpy()), sub)
            assert got[fi][0] == want[0], (company, fi, got[fi], want)
            if want[0] is not None:
                assert abs(float(scores[fi, 0]) - float(want[1])) <= 1e-4
                assert bool(acc[fi]) == bool(np.float32(want[1]) >= np.float32(0.4))
# enrol-time duplicate check over the shards (trainingServer.py:170-200): first hit in GLOBAL order
chk = frg.EnrolmentChecker(g2, matcher=ShardedMatcher(g2, exchange="nccl"))
two_hits = np.float32(0.6) * refv[ids2[n2 - 3]] + np.float32(0.8) * refv[ids2[20]]     # rows on different ranks
for company in (None, "acme", "globex", "nobody"):
    sub = [p_ for p_ in order if company is None or refc[p_] == company]
    for probe_v in (two_hits, refv["new-b"], V2[77], rng.standard_normal(d).astype(np.float32)):
        where = mo.duplicate_check(probe_v, [refv[p_] for p_ in sub], 0.4)
        dup, pid = chk.check_duplicate_face(probe_v, company)
        assert dup == (where >= 0) and (pid == sub[where] if dup else pid is None), (company, dup, pid, where)
if dist.get_rank() == 0:
    print("ENROL_OK world=%%d" %% dist.get_world_size())
# EmbeddingManager + FaceRecognitionProcessor over the sharded gallery, documents visible to rank 0 only
sys.path.insert(0, os.path.join(%r, "tests"))
from test_gpu_sharded import _manager_scenario
gm = np.load(os.path.join(%r, "tests", "golden", "managers.npz"))
ctl = dist.new_group(backend="gloo")
store3 = frg.GalleryStore(dim=gm["stored"].shape[1], capacity=64, device=local)
g3 = ShardedGallery(dim=store3.dim, device=local, store=store3)
_manager_scenario(frg, g3, ShardedMatcher(g3),
                  lambda E, V: frg.BroadcastSource(frg.ListSource(E, V) if dist.get_rank() == 0 else frg.ListSource([], []), ctl),
                  gm, rank=dist.get_rank())
if dist.get_rank() == 0:
    print("MANAGER_OK world=%%d" %% dist.get_world_size())
# soak: 200 collective calls with random per-rank HOST jitter (the ranks enqueue at different times, sometimes
# 50 ms apart) - p2p against NCCL every time.  Then the round-1 scaling crash on purpose: rank 1 skips a call.
import random, time
rnd = random.Random(1000 + dist.get_rank())
shapes = random.Random(7)
Qs = torch.from_numpy(synth.queries(320, n, d, seed=99)[0]).cuda()
for it in range(200):
    ff, kk = shapes.randint(1, 320), shapes.choice([1, 5, 10, 16])
    time.sleep(rnd.random() * (0.05 if rnd.random() < 0.05 else 0.002))
    a = m_p2p.match(Qs[:ff], kk, 0.45)
    if it %% 10 == 0:
        b = m_nccl.match(Qs[:ff], kk, 0.45)
        torch.cuda.synchronize()
        for x, y in zip(a, b):
            assert torch.equal(x, y), (it, ff, kk)
torch.cuda.synchronize()
m_p2p.check_exchange()
dist.barrier()
if dist.get_rank() == 0:
    a = m_p2p.match(Qs[:32], 5, 0.45)              # rank 1 does not take part: bounded wait, status, no trap
    t0 = time.perf_counter()
    try:
        m_p2p.check_exchange()
        raise SystemExit("a missing rank went unnoticed")
    except frg.NativeError as ex:
        assert ex.code == frg._native.ERR_STATE and "rank 1 never announced" in str(ex), str(ex)
    assert time.perf_counter() - t0 < 10 and (a[0] == -1).all()
dist.barrier()
m_p2p.reset_exchange()                             # collective recovery: fresh buffer, epochs restart
a = m_p2p.match(Qd, 5, 0.45); b = m_nccl.match(Qd, 5, 0.45)
torch.cuda.synchronize()
m_p2p.check_exchange()
for x, y in zip(a, b):
    assert torch.equal(x, y)
if dist.get_rank() == 0:
    print("SOAK_OK world=%%d" %% dist.get_world_size())
# every rank holds the same merged result
chk = a[0].clone()
dist.broadcast(chk, src=0)
assert torch.equal(chk, a[0])
if dist.get_rank() == 0:
    print("P2P_OK world=%%d epochs=%%d" %% (dist.get_world_size(), m_p2p._epoch))
dist.destroy_process_group()
'''


def test_sharded_two_ranks_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (ROOT, ROOT, ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, FRG_EXCHANGE_TIMEOUT_MS="500"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_OK world=2" in r.stdout
    assert "P2P_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ENROL_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MANAGER_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SOAK_OK world=2" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
