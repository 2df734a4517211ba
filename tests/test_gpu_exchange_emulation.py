"""frg_match_exchange on ONE device: the ranks of a row-sharded gallery are emulated one after the other on a
single stream (FRG_XCHG_PUSH_ONLY for every rank, then FRG_XCHG_MERGE_ONLY for every rank), so that the whole
peer-memory protocol - select-stage pushes, hello packets, the exact fallback's late pushes, the push kernel of
variants without a select stage, epoch parity reuse, the poll + merge kernel and its failure reports - is
covered by the driver's single-GPU `pytest -m gpu` run.  No kernel here ever waits for another kernel (the
profiling guide forbids that on one GPU): every packet a merge polls for was pushed earlier in stream order,
and the two failure cases poll for packets nobody sends with a bound of 100-200 ms."""
import ctypes as C

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


class Ranks:
    """`world` shards of one synthetic gallery on device 0 + their exchange buffers (plain device memory:
    on one device every buffer is "peer-mapped" already)."""

    def __init__(self, frg, n, world, dim=512, max_slots=4096 * 16, seed=1234):
        import torch
        from facerecognition_infrenceengine_b200 import _native as N
        from facerecognition_infrenceengine_b200.sharded import shard_bounds
        self.N, self.torch, self.frg = N, torch, frg
        self.n, self.world, self.dim = n, world, dim
        self.bounds = shard_bounds(n, world)
        self.stores = []
        for lo, hi in self.bounds:
            st = frg.GalleryStore(dim=dim, capacity=max(hi - lo, 1))
            if hi > lo:
                st.fill_synthetic(hi - lo, lo, seed)
            self.stores.append(st)
        self.full = frg.GalleryStore(dim=dim, capacity=n)
        self.full.fill_synthetic(n, 0, seed)
        cap, total = C.c_int64(), C.c_int64()
        N.check(N.lib.frg_exchange_bytes(world, max_slots, 1, C.byref(cap), C.byref(total)))
        self.cap = int(cap.value)
        self.bufs = [torch.zeros(int(total.value), dtype=torch.uint8, device="cuda") for _ in range(world)]
        self.ptrs = torch.tensor([b.data_ptr() for b in self.bufs], dtype=torch.int64, device="cuda")
        self.epoch = 0

    def close(self):
        for s in self.stores + [self.full]:
            s.close()

    def _x(self, rank, epoch, flags):
        return self.N.Exchange(rank=rank, world=self.world, peer_bufs=int(self.ptrs.data_ptr()), block_cap=self.cap,
                               epoch=epoch, flags=flags)

    def call(self, rank, Qd, k, flags, epoch, threshold=0.45, variant="auto", out=None, tenant=-1):
        torch, N = self.torch, self.N
        F = Qd.shape[0]
        loc = (torch.empty((F, k), dtype=torch.int64, device="cuda"), torch.empty((F, k), dtype=torch.float32, device="cuda"))
        if out is None:
            out = (torch.full((F, k), -7, dtype=torch.int64, device="cuda"),
                   torch.full((F, k), -7.0, dtype=torch.float32, device="cuda"),
                   torch.full((F,), 9, dtype=torch.uint8, device="cuda"))
        p = N.MatchParams(metric=0, variant=N.VARIANTS[variant], threshold=float(np.float32(threshold)), tenant=tenant,
                          row_offset=self.bounds[rank][0], flags=0, reserved=0)
        x = self._x(rank, epoch, flags)
        stream = torch.cuda.current_stream().cuda_stream
        N.check(N.lib.frg_match_exchange(self.stores[rank].handle, C.c_void_p(Qd.data_ptr()), F, k, C.byref(p), C.byref(x),
                                         C.c_void_p(loc[0].data_ptr()), C.c_void_p(loc[1].data_ptr()),
                                         C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
                                         C.c_void_p(out[2].data_ptr()), C.c_void_p(stream)))
        return out, loc

    def collective(self, Qd, k, order=None, variant="auto", threshold=0.45, timeout_ms=0):
        """One collective call: every rank pushes (in `order`), then every rank merges."""
        N = self.N
        self.epoch += 1
        keep = []
        for r in (order or range(self.world)):
            keep.append(self.call(r, Qd, k, N.XCHG_PUSH_ONLY, self.epoch, threshold, variant))
        outs = [self.call(r, Qd, k, N.XCHG_MERGE_ONLY | (timeout_ms << 8), self.epoch, threshold, variant)[0]
                for r in range(self.world)]
        self.torch.cuda.synchronize()
        return outs

    def status(self, rank, clear=False):
        N = self.N
        st = N.ExchangeStatus()
        rc = N.lib.frg_exchange_status(0, C.c_void_p(self.bufs[rank].data_ptr()), 1 if clear else 0, None, C.byref(st))
        return rc, st, (N.lib.frg_last_error() or b"").decode()

    def reference(self, Qd, k, variant="auto", threshold=0.45):
        r = self.frg.Matcher(self.full).match_device(Qd, k, threshold, variant=variant)
        self.torch.cuda.synchronize()
        return r


def same(torch, outs, ref):
    for o in outs:
        for a, b in zip(o, ref):
            if not torch.equal(a, b):
                return False
    return True


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_ranks_equal_the_single_store(frg, world):
    import torch
    R = Ranks(frg, 200_001, world)
    Q, target = synth.queries(300, R.n, R.dim)
    Qd = torch.from_numpy(Q).cuda()
    # changing shapes from call to call (epoch parity reuse, slot planes that move with nq*k), both push orders,
    # pair kernels (F > 128), the exact scan (push kernel instead of the select-stage push)
    plan = [(64, 10, "auto"), (64, 10, "auto"), (8, 1, "auto"), (300, 5, "auto"), (200, 5, "tc_exact"), (2, 16, "auto"),
            (64, 5, "scan_f32"), (1, 1, "auto"), (129, 10, "auto")]
    for i, (f, k, variant) in enumerate(plan):
        order = list(range(world)) if i % 2 == 0 else list(reversed(range(world)))
        outs = R.collective(Qd[:f], k, order, variant)
        ref = R.reference(Qd[:f], k, variant)
        assert same(torch, outs, ref), (i, f, k, variant)
        for r in range(world):
            rc, st, _ = R.status(r)
            assert rc == 0 and st.code == 0
    # ... and against the oracle once (the single store is itself checked against it elsewhere)
    G = synth.gallery(R.n, R.dim)
    rr, rs, ra = mo.match_topk(Q[:64], G, 6, 0.45)
    outs = R.collective(Qd[:64], 5)
    rows, scores, acc = (t.cpu().numpy() for t in outs[world - 1])
    assert mo.ids_match_with_gap(rr, rs, rows, 1e-4).all()
    assert np.abs(scores - rs[:, :5]).max() <= 1e-4 and (acc.astype(bool) == ra).all()
    assert (rows[target[:64] >= 0, 0] == target[:64][target[:64] >= 0]).all()
    R.close()


def test_emulated_late_pushes_from_the_exact_fallback(frg):
    """3000 copies of one template on the last rank: the queries near it overflow that rank's candidate lists,
    are redone by its exact fallback and reach every rank through the fallback kernel's own push."""
    import torch
    R = Ranks(frg, 100_000, 2)
    dupe = synth.unit_rows(np.arange(1), R.dim, 4242, synth.STREAM_IMPOSTOR)
    for st in (R.stores[1], R.full):
        st.append_rows(np.repeat(dupe, 3000, axis=0), prenormalised=True)
    Q, _ = synth.queries(14, R.n, R.dim)
    Qx = torch.from_numpy(np.concatenate([dupe, dupe + np.float32(1e-3), Q])).cuda()
    for k in (1, 16, 5):
        outs = R.collective(Qx, k)
        ref = R.reference(Qx, k)
        assert same(torch, outs, ref), k
        assert outs[0][0][0].tolist() == list(range(R.n, R.n + k))      # ties -> earliest rows, global numbering
    R.close()


def test_emulated_soak_200_calls(frg):
    """200 collective calls, random batch size / k / push order: bit-equal to the single store every time."""
    import torch
    rng = np.random.default_rng(5)
    R = Ranks(frg, 60_000, 2)
    Q, _ = synth.queries(320, R.n, R.dim)
    Qd = torch.from_numpy(Q).cuda()
    for i in range(200):
        f = int(rng.integers(1, 321))
        k = int(rng.choice([1, 3, 5, 10, 16]))
        a = int(rng.integers(0, 320 - f + 1))
        order = [0, 1] if rng.integers(2) else [1, 0]
        outs = R.collective(Qd[a:a + f], k, order)
        ref = R.reference(Qd[a:a + f], k)
        assert same(torch, outs, ref), (i, f, k)
    assert R.status(0)[0] == 0 and R.status(1)[0] == 0
    R.close()


def test_shape_mismatch_is_reported_not_waited_on(frg):
    import torch
    R = Ranks(frg, 50_000, 2)
    N = R.N
    Q, _ = synth.queries(32, R.n, R.dim)
    Qd = torch.from_numpy(Q).cuda()
    R.epoch += 1
    R.call(0, Qd, 5, N.XCHG_PUSH_ONLY, R.epoch)
    R.call(1, Qd, 3, N.XCHG_PUSH_ONLY, R.epoch)                   # rank 1 believes k = 3
    out, _ = R.call(0, Qd, 5, N.XCHG_MERGE_ONLY | (200 << 8), R.epoch)
    torch.cuda.synchronize()
    rc, st, msg = R.status(0)
    assert rc == N.ERR_STATE and st.code == 2 and st.peer == 1 and (st.nq, st.k, st.peer_nq, st.peer_k) == (32, 5, 32, 3)
    assert "rank 1 passed nq=32, k=3" in msg
    assert (out[0] == -1).all() and (out[1] == -1).all() and (out[2] == 0).all()      # a void call matches nothing
    # sticky until cleared
    assert R.status(0, clear=True)[0] == N.ERR_STATE
    assert R.status(0)[0] == 0
    R.close()


def test_missing_rank_times_out_with_a_status_not_a_trap(frg):
    """The round-1 scaling crash: one rank makes one more collective call than its peer.  The poll is bounded,
    the kernel ends normally, the context survives and the host gets a status that says what happened."""
    import time
    import torch
    R = Ranks(frg, 50_000, 2)
    N = R.N
    Q, _ = synth.queries(16, R.n, R.dim)
    Qd = torch.from_numpy(Q).cuda()
    outs = R.collective(Qd, 5)                                   # call 1: both ranks
    assert same(torch, outs, R.reference(Qd, 5))
    R.epoch += 1
    R.call(0, Qd, 5, N.XCHG_PUSH_ONLY, R.epoch)                   # call 2: rank 0 only
    t0 = time.perf_counter()
    out, _ = R.call(0, Qd, 5, N.XCHG_MERGE_ONLY | (100 << 8), R.epoch)
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 5.0
    rc, st, msg = R.status(0, clear=True)
    assert rc == N.ERR_STATE and st.code == 1 and st.peer == 1 and st.epoch == 2 and st.peer_epoch == 0
    assert "different numbers of collective calls" in msg
    assert (out[0] == -1).all() and (out[2] == 0).all()
    # the context is alive: the same stores keep matching, and once rank 1 catches up the exchange works again
    R.call(1, Qd, 5, N.XCHG_PUSH_ONLY, R.epoch)
    outs = [R.call(r, Qd, 5, N.XCHG_MERGE_ONLY, R.epoch)[0] for r in range(2)]
    torch.cuda.synchronize()
    assert same(torch, outs, R.reference(Qd, 5))
    R.close()


def test_emulated_ranks_with_a_stretched_tenant_window(frg):
    """Row-sharded AND company-filtered: on each shard the company is a block plus one late row at the shard's end,
    so every rank's local match walks its own tile list; the merged result equals the single store's (which walks
    its own) and the oracle's masked scan."""
    import torch
    from facerecognition_infrenceengine_b200 import _native as N
    n, d, T, tenant = 300_000, 512, 10_000, 3
    R = Ranks(frg, n, 2)
    for st in R.stores + [R.full]:
        st.close()
    G = synth.gallery(n, d)
    tags = (1 + np.arange(n) // T).astype(np.int32)
    tags[n // 2 - 1] = tenant                  # last row of shard 0
    tags[n - 1] = tenant                       # last row of shard 1 (the company has no block there: just this row)
    tags[200_000:205_000] = tenant             # ... and a second block on shard 1
    R.stores = []
    for lo, hi in R.bounds:
        st = frg.GalleryStore(dim=d, capacity=hi - lo)
        st.append_rows(G[lo:hi], tags[lo:hi], prenormalised=True)
        R.stores.append(st)
    R.full = frg.GalleryStore(dim=d, capacity=n)
    R.full.append_rows(G, tags, prenormalised=True)
    rng = np.random.default_rng(1)
    mine = np.nonzero(tags == tenant)[0]
    for F in (32, 200):
        pick = rng.choice(mine, size=F)
        pick[:2] = [n // 2 - 1, n - 1]
        Q = G[pick] + np.float32(0.03) * rng.standard_normal((F, d)).astype(np.float32)
        Qd = torch.from_numpy(Q).cuda()
        R.epoch += 1
        for r in (0, 1):
            R.call(r, Qd, 5, N.XCHG_PUSH_ONLY, R.epoch, 0.4, tenant=tenant)
        outs = [R.call(r, Qd, 5, N.XCHG_MERGE_ONLY, R.epoch, 0.4, tenant=tenant)[0] for r in (0, 1)]
        ref = frg.Matcher(R.full).match_device(Qd, 5, 0.4, tenant=tenant)
        torch.cuda.synchronize()
        assert same(torch, outs, ref), F
        rr, rs, ra = mo.match_topk_fast(Q, G, 6, 0.4, tags, tenant)
        rows, scores, acc = (t.cpu().numpy() for t in outs[0])
        assert mo.ids_match_with_gap(rr, rs, rows, 1e-4).all() and np.abs(scores - rs[:, :5]).max() <= 1e-4
        assert rows[0, 0] == n // 2 - 1 and rows[1, 0] == n - 1
    R.close()
