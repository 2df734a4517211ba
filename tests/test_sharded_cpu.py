"""Host-side logic of the row-sharded match on CPU: world_size-2 `gloo` processes.  The per-shard
match and the merge are injected from the oracle (test infrastructure); what is under test is the
product's shard planning, global-row offsets, packed all-gather layout and its strides - the parts
of sharded.py that do not need a GPU.  The CUDA path of the same class runs in tests/test_gpu_*.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import matcher_oracle as mo
from oracle import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _merge_numpy(gathered, parts, F, k, threshold, out):
    """Reads the packed [rows | scores] blocks exactly as frg_merge_topk_strided does."""
    from facerecognition_infrenceengine_b200.sharded import block_layout
    rb, block = block_layout(F, k)
    assert block % 8 == 0 and block >= F * k * 12
    buf = gathered.numpy()
    rows = np.stack([buf[p * block:p * block + rb].view(np.int64).reshape(F, k) for p in range(parts)])
    scores = np.stack([buf[p * block + rb:p * block + rb + F * k * 4].view(np.float32).reshape(F, k)
                       for p in range(parts)])
    o_r, o_s, o_a = (t.numpy() for t in out)      # views onto the caller's tensors
    for f in range(F):
        cand = [(-(scores[p, f, j]), rows[p, f, j]) for p in range(parts) for j in range(k) if rows[p, f, j] >= 0]
        cand.sort()
        for j in range(k):
            if j < len(cand):
                o_s[f, j] = -cand[j][0]
                o_r[f, j] = cand[j][1]
            else:
                o_s[f, j] = -1.0
                o_r[f, j] = -1
        o_a[f] = 1 if (o_r[f, 0] >= 0 and np.float32(o_s[f, 0]) >= np.float32(threshold)) else 0


def _worker(rank, world, port, n, F, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher, shard_bounds
        d = 512
        G = synth.gallery(n, d)
        Q, _ = synth.queries(F, n, d)
        tie = F > 1        # (a 1-row batch takes another BLAS path per shard shape: scores differ by an ulp)
        if tie:
            G[n - 3] = G[5]                   # an exact tie that straddles the two shards
            Q[0] = G[5]
        g = ShardedGallery(dim=d, store=None)
        lo, hi = g.plan(n)
        assert g.bounds == shard_bounds(n, world) and g.offset == lo and g.total_rows == n

        def local_match(Qt, k_, thr, variant, rows_out, scores_out):
            r, s, _ = mo.match_topk(Qt.numpy(), G[lo:hi], k_, thr)
            rows_out.copy_(torch.from_numpy(np.where(r >= 0, r + lo, -1)))
            scores_out.copy_(torch.from_numpy(s))

        m = ShardedMatcher(g, local_match=local_match, merge=_merge_numpy)
        Qt = torch.from_numpy(Q.copy())
        if rank != 0:
            Qt.zero_()                        # only rank 0 holds the batch; broadcast ships it
        rows, scores, acc = m.match(Qt, k, 0.45, broadcast=True)
        ref_r, ref_s, ref_a = mo.match_topk(Q, G, k, 0.45)
        assert (rows.numpy() == ref_r).all()
        assert np.abs(scores.numpy() - ref_s).max() < 1e-6
        assert (acc.numpy().astype(bool) == ref_a).all()
        if tie:
            assert list(rows.numpy()[0, :2]) == [5, n - 3][:k]    # the tie resolves to the lower GLOBAL row
        # appends go to the last rank and extend the global order
        g.append_local(G[:4])
        assert g.total_rows == n + 4 and g.bounds[0] == (0, n // 2)
        _enrolment_scenario(rank, world, F, k)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


class _HostShard:
    """Stands in for the device store of one rank (append / overwrite / tombstone of LOCAL rows), so that the
    routing of ShardedGallery.upsert / remove can be checked without a GPU.  Normalises on ingest like the
    real store (oracle arithmetic: test infrastructure)."""

    def __init__(self, dim):
        self.vecs = np.zeros((0, dim), np.float32)
        self.tags = np.zeros(0, np.int32)

    def append_rows(self, vecs, tags=None, prenormalised=False):
        first = len(self.vecs)
        v = np.asarray(vecs, np.float32) if prenormalised else mo.normalise_rows(np.asarray(vecs, np.float32))
        self.vecs = np.concatenate([self.vecs, v])
        self.tags = np.concatenate([self.tags, np.zeros(len(v), np.int32) if tags is None else np.asarray(tags, np.int32)])
        return first

    def overwrite_rows(self, rows, vecs, tags=None, prenormalised=False):
        v = np.asarray(vecs, np.float32) if prenormalised else mo.normalise_rows(np.asarray(vecs, np.float32))
        for j, r in enumerate(rows):
            assert 0 <= r < len(self.vecs), "overwrite routed to a rank that does not own the row"
            self.vecs[r] = v[j]
            self.tags[r] = 0 if tags is None else tags[j]

    def remove_rows(self, rows):
        for r in rows:
            assert 0 <= r < len(self.vecs), "removal routed to a rank that does not own the row"
            self.tags[r] = -1

    def stats(self):
        from types import SimpleNamespace
        return SimpleNamespace(rows=len(self.vecs), live=int((self.tags >= 0).sum()), capacity=len(self.vecs),
                               bytes=self.vecs.nbytes, version=0)


def _enrolment_scenario(rank, world, F, k):
    """Sharded upsert / remove / tenant filter against the reference's dict semantics (GalleryOracle follows
    infrenceServer.py:260-341): every rank makes the same calls, only owners touch their rows."""
    from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
    d, n0 = 64, 41
    rng = np.random.default_rng(5)
    ids0 = ["%024x" % (i + 1000) for i in range(n0)]
    V0 = rng.standard_normal((n0, d)).astype(np.float32)
    comp = ["acme" if i % 3 else "globex" for i in range(n0)]
    g = ShardedGallery(dim=d, store=_HostShard(d))
    g.load(ids0, V0, comp)        # contiguous balanced blocks: each rank ingests only its own
    assert len(g.store.vecs) == g.local_rows and g.total_rows == n0 and g.row_of(ids0[-1]) == n0 - 1
    ref_ids, ref_vecs, ref_comp = list(ids0), {p: mo.normalise(v) for p, v in zip(ids0, V0)}, dict(zip(ids0, comp))

    def ref_apply_upsert(ids, vecs, comps):
        for p, v, c in zip(ids, vecs, comps):
            if p not in ref_vecs:
                ref_ids.append(p)
            ref_vecs[p] = mo.normalise(v)
            ref_comp[p] = c

    def ref_apply_remove(ids):
        for p in ids:
            if p in ref_vecs:
                ref_ids.remove(p)
                del ref_vecs[p], ref_comp[p]

    # 1. overwrite ids owned by different ranks + two new ids (one repeated inside the batch: position of
    #    the first occurrence, content of the last)
    up_ids = [ids0[2], ids0[n0 - 2], "new-a", "new-b", "new-a"]
    up_v = rng.standard_normal((5, d)).astype(np.float32)
    up_c = ["acme", "globex", "acme", "initech", "acme"]
    g.upsert(up_ids, up_v, up_c, meta=[{"name": p} for p in up_ids])
    ref_apply_upsert(up_ids, up_v, up_c)
    # 2. remove one id per rank and an unknown one; re-enrol one of them (goes to the END)
    assert g.remove([ids0[5], ids0[n0 - 5], "nobody"]) == 2
    ref_apply_remove([ids0[5], ids0[n0 - 5]])
    g.upsert([ids0[5]], up_v[:1], ["acme"])
    ref_apply_upsert([ids0[5]], up_v[:1], ["acme"])
    assert g.total_rows == n0 + 3 and len(g) == len(ref_ids) and ids0[5] in g and "nobody" not in g
    assert g.row_of(ids0[5]) == n0 + 2 and g.id_of(5) is None and g.metadata("new-b") == {"name": "new-b"}
    assert g.bounds[0] == (0, n0 // world)                       # only the last block grew

    def local_match(Qt, k_, thr, variant, rows_out, scores_out, tenant=-1):
        lo_ = g.offset
        r, s, _ = mo.match_topk(Qt.numpy(), g.store.vecs, k_, thr, tags=g.store.tags, tenant=tenant)
        rows_out.copy_(torch.from_numpy(np.where(r >= 0, r + lo_, -1)))
        scores_out.copy_(torch.from_numpy(s))

    m = ShardedMatcher(g, local_match=local_match, merge=_merge_numpy)

    # first row in GLOBAL order above a threshold (trainingServer.py:170-200 over the shards): each rank scans
    # its block with the oracle's duplicate_check, the lowest global row over the ranks wins
    def local_first(Qn, thr, strict, tenant):
        rows, scores = np.full(len(Qn), -1, np.int64), np.full(len(Qn), -1.0, np.float32)
        for f, q in enumerate(Qn):
            qn = mo.normalise(q)
            for r in range(len(g.store.vecs)):
                if g.store.tags[r] < 0 or (tenant >= 0 and g.store.tags[r] != tenant):
                    continue
                sc = np.dot(qn, g.store.vecs[r])
                if (sc > np.float32(thr)) if strict else (sc >= np.float32(thr)):
                    rows[f], scores[f] = r + g.offset, sc
                    break
        return rows, scores

    probe = np.stack([0.6 * ref_vecs[ids0[n0 - 2]] + 0.8 * ref_vecs[ids0[3]],       # two hits, on different ranks
                      ref_vecs["new-b"], rng.standard_normal(d).astype(np.float32)])
    for company in (None, "acme", "globex"):
        rows_f, sc_f = m.first_above(probe, 0.4, strict=True, company_id=company, local_first=local_first)
        sub = [p for p in ref_ids if company is None or ref_comp[p] == company]
        for f in range(len(probe)):
            where = mo.duplicate_check(probe[f], [ref_vecs[p] for p in sub], 0.4)       # the reference's scan
            dup = where >= 0
            assert (rows_f[f] >= 0) == dup, (company, f)
            if dup:
                assert m.g.id_of(rows_f[f]) == sub[where]
                assert abs(sc_f[f] - np.dot(mo.normalise(probe[f]), ref_vecs[sub[where]])) < 1e-6
            else:
                assert sc_f[f] == np.float32(-1.0)

    Q = np.stack([ref_vecs[p] for p in (ids0[2], "new-a", ids0[5], ids0[n0 - 2], ids0[7])][:max(1, F)])
    for company in (None, "acme", "globex", "initech", "no-such-company"):
        rows, scores, acc = m.match(torch.from_numpy(Q.copy()), k, 0.4, company_id=company)
        sub = [p for p in ref_ids if company is None or ref_comp[p] == company]
        for f in range(len(Q)):
            want = mo.scan_best(Q[f], {p: ref_vecs[p] for p in sub})     # the reference's own loop
            got = m.ids_of(rows)[f][0]
            assert got == want[0], (company, f, got, want)
            if want[0] is not None:
                assert abs(float(scores[f, 0]) - float(want[1])) < 1e-6


@pytest.mark.parametrize("n,F,k", [(4001, 6, 5), (2000, 4, 1), (2000, 3, 1), (1500, 1, 3)])
def test_sharded_match_world2_gloo(n, F, k):
    import facerecognition_infrenceengine_b200  # noqa: F401  (builds / loads the library before forking)
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, F, k, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: 1, 1: 1}


def test_shard_bounds_cover_everything():
    from facerecognition_infrenceengine_b200.sharded import owner_of, shard_bounds
    for n in (0, 1, 7, 100_000_000):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    b = shard_bounds(10, 4)
    assert [owner_of(r, b) for r in range(10)] == [0, 0, 1, 1, 1, 2, 2, 3, 3, 3]


def test_matcher_wiring_without_injected_pieces():
    """The CUDA pieces are bound methods of the matcher itself; world = 1 never sets up an exchange."""
    from facerecognition_infrenceengine_b200.sharded import ShardedGallery, ShardedMatcher
    g = ShardedGallery(dim=512, device=0, store=None, rank=0, world=1)
    m = ShardedMatcher(g)
    assert m.exchange == "nccl" and callable(m._local) and callable(m._merge)
    for name in ("_local_cuda", "_merge_cuda", "_accept_scratch", "_p2p_setup", "_p2p_ready", "_match_exchange_p2p"):
        assert callable(getattr(m, name))
    import pytest
    with pytest.raises(ValueError):
        ShardedMatcher(g, exchange="smoke-signals")


# ---- EmbeddingManager over a sharded gallery: the reference's own manager scenario (tests/golden/managers.npz,
#      produced by the unmodified infrenceServer.EmbeddingManager under stubs) replayed on two ranks
def _manager_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from datetime import datetime, timedelta, timezone
        import facerecognition_infrenceengine_b200 as frg
        from facerecognition_infrenceengine_b200.sharded import ShardedGallery
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "managers.npz"))
        st = g["stored"]
        A, B = "a" * 24, "b" * 24
        t0 = datetime(2026, 1, 1)

        def emp(i):
            return {"_id": "%024x" % i, "embedding": st[i], "companyId": A if i < 8 or i >= 12 else B,
                    "status": "active", "blacklisted": False, "lastUpdated": t0, "employeeName": "e%d" % i}

        def vis(i):
            return {"_id": "%024x" % (100 + i), "embedding": st[20 + i], "companyId": A if i < 3 else B,
                    "lastUpdated": t0, "visitorName": "v%d" % i}

        E = [emp(i) for i in range(12)]
        E[3]["status"] = "inactive"; E[5]["blacklisted"] = True; E[6]["embedding_status"] = "pending"
        V = [vis(i) for i in range(6)]
        V[4]["embedding_status"] = "pending"
        # only rank 0 can see the documents; the others receive them through the broadcast
        src = frg.BroadcastSource(frg.ListSource(E, V) if rank == 0 else frg.ListSource([], []))
        gal = ShardedGallery(dim=st.shape[1], store=_HostShard(st.shape[1]))
        m = frg.EmbeddingManager(src, mode="live", store=gal)

        def same(prefix):
            ref_ids, ref_G = list(g[prefix + "_ids"]), g[prefix + "_G"]
            assert gal.ids() == ref_ids, (prefix, gal.ids(), ref_ids)
            lo, hi = gal.bounds[rank]
            assert len(gal.store.vecs) == hi - lo
            mine = 0
            for i, pid in enumerate(ref_ids):                  # every live id sits on exactly its owner rank
                r = gal.row_of(pid)
                if lo <= r < hi:
                    mine += 1
                    a, b = gal.store.vecs[r - lo], ref_G[i]
                    assert gal.store.tags[r - lo] >= 0
                    assert (np.isnan(a) == np.isnan(b)).all()         # a zero template is a NaN row on both sides
                    assert np.isnan(b).all() or np.nanmax(np.abs(a - b)) <= 2e-7
            assert int((gal.store.tags >= 0).sum()) == mine     # and nothing else is live here
            tot = torch.tensor([mine]); dist.all_reduce(tot)
            assert int(tot) == len(ref_ids)

        same("ref_live_load")
        s = m.get_stats()
        assert [s["total_embeddings"], s["employees"], s["visitors"]] == list(g["ref_live_load_stats"])
        assert gal.bounds[0][1] - gal.bounds[0][0] in (len(gal) // 2, (len(gal) + 1) // 2)   # balanced initial load
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
        for comp, key in ((A, "ref_live_tenant_a"), (B, "ref_live_tenant_b")):
            assert len(m.get_embeddings_for_company(comp)) == len(g[key])
            code = gal.tenant_code(comp, create=False)
            assert sorted(p for p in gal.ids() if gal._tenant_of[p] == code) == list(g[key])
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharded_embedding_manager_replays_reference_scenario_gloo():
    import facerecognition_infrenceengine_b200  # noqa: F401
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_manager_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {0: 1, 1: 1}


# ---- property: any sequence of upserts / removals over W ranks leaves the shards holding exactly what the
#      reference's dict holds, in the dict's order (no communication is involved: the ranks are emulated in
#      one process, each with its own gallery object and host shard, all receiving the same calls)
from hypothesis import given, settings          # noqa: E402
from hypothesis import strategies as st         # noqa: E402


@st.composite
def _op_sequences(draw):
    world = draw(st.sampled_from([1, 2, 3, 8]))
    n_ops = draw(st.integers(1, 12))
    ops = []
    for _ in range(n_ops):
        if draw(st.integers(0, 3)) == 0:
            ops.append(("remove", draw(st.lists(st.integers(0, 15), min_size=1, max_size=4))))
        else:
            ids = draw(st.lists(st.integers(0, 15), min_size=1, max_size=6))
            ops.append(("upsert", ids, draw(st.integers(0, 2 ** 31 - 1))))
    return world, ops


@settings(max_examples=120, deadline=None)
@given(_op_sequences())
def test_sharded_upserts_follow_the_dict(data):
    from facerecognition_infrenceengine_b200.sharded import ShardedGallery
    world, ops = data
    d = 8
    gals = [ShardedGallery(dim=d, store=_HostShard(d), rank=r, world=world) for r in range(world)]
    ref = {}                                                   # the reference's Dict[str, np.ndarray]
    comp = {}
    for op in ops:
        if op[0] == "remove":
            ids = ["p%d" % i for i in op[1]]
            for g in gals:
                g.remove(ids)
            for p in ids:
                ref.pop(p, None)
                comp.pop(p, None)
        else:
            ids = ["p%d" % i for i in op[1]]
            rng = np.random.default_rng(op[2])
            V = rng.integers(1, 9, size=(len(ids), d)).astype(np.float32)
            C_ = ["c%d" % (i % 3) for i in op[1]]
            for g in gals:
                g.upsert(ids, V, C_)
            for p, v, c in zip(ids, V, C_):
                ref[p] = mo.normalise(v)                       # dict assignment: position of the first, value of the last
                comp[p] = c
    for g in gals:
        assert g.ids() == list(ref) and len(g) == len(ref)
        assert g.bounds == gals[0].bounds and g.bounds[-1][1] == g.total_rows
        assert all(b[1] == g.bounds[i + 1][0] for i, b in enumerate(g.bounds[:-1]))
    live = 0
    for r, g in enumerate(gals):
        lo, hi = g.bounds[r]
        assert len(g.store.vecs) == hi - lo
        live += int((g.store.tags >= 0).sum())
    assert live == len(ref)
    for p, v in ref.items():
        row = gals[0].row_of(p)
        owner = next(r for r, (lo, hi) in enumerate(gals[0].bounds) if lo <= row < hi)
        sh = gals[owner].store
        assert np.array_equal(sh.vecs[row - gals[owner].bounds[owner][0]], v)
        assert sh.tags[row - gals[owner].bounds[owner][0]] == gals[owner].tenant_code(comp[p], create=False)
