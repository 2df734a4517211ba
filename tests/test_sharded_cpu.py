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
    block = F * k * 12
    buf = gathered.numpy()
    rows = np.stack([buf[p * block:p * block + F * k * 8].view(np.int64).reshape(F, k) for p in range(parts)])
    scores = np.stack([buf[p * block + F * k * 8:(p + 1) * block].view(np.float32).reshape(F, k) for p in range(parts)])
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
        G[n - 3] = G[5]                       # an exact tie that straddles the two shards
        Q, _ = synth.queries(F, n, d)
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
        assert list(rows.numpy()[0, :2]) == [5, n - 3][:k]        # the tie resolves to the lower GLOBAL row
        # appends go to the last rank and extend the global order
        g.append_local(G[:4])
        assert g.total_rows == n + 4 and g.bounds[0] == (0, n // 2)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,F,k", [(4001, 6, 5), (2000, 4, 1)])
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
