"""Frame -> batch aggregator (host logic, no GPU): batching by size and by deadline, per-tenant
separation, result routing, failure isolation."""
import threading
import time

import numpy as np

from facerecognition_infrenceengine_b200.aggregator import BatchAggregator


def fake_match(calls):
    def fn(Q, tenant):
        calls.append((len(Q), tenant))
        rows = Q[:, :1].astype(np.int64)                 # echo the first component as the "row"
        return rows, Q[:, 1:2].copy(), np.ones(len(Q), bool)
    return fn


def frame(ident, faces, dim=8):
    e = np.zeros((faces, dim), np.float32)
    e[:, 0] = ident
    e[:, 1] = np.arange(faces)
    return e


def test_batches_fill_up_and_results_route_back():
    calls = []
    agg = BatchAggregator(fake_match(calls), max_batch=12, max_delay_ms=2000)
    futs = [agg.submit(frame(i, 4)) for i in range(3)]          # 12 faces -> one batch, no deadline wait
    res = [f.result(timeout=5) for f in futs]
    assert calls == [(12, None)]
    for i, (rows, scores, acc) in enumerate(res):
        assert (rows[:, 0] == i).all() and list(scores[:, 0]) == [0, 1, 2, 3] and acc.all()
    agg.close()


def test_deadline_flushes_a_partial_batch():
    calls = []
    agg = BatchAggregator(fake_match(calls), max_batch=1000, max_delay_ms=30)
    t0 = time.monotonic()
    rows, _, _ = agg.submit(frame(7, 2)).result(timeout=5)
    dt = time.monotonic() - t0
    assert calls == [(2, None)] and (rows == 7).all()
    assert 0.02 <= dt < 1.0
    agg.close()


def test_tenants_are_not_mixed_and_failures_are_isolated():
    calls = []

    def fn(Q, tenant):
        if tenant == "bad":
            raise RuntimeError("boom")
        return fake_match(calls)(Q, tenant)

    agg = BatchAggregator(fn, max_batch=64, max_delay_ms=20)
    fa = [agg.submit(frame(1, 3), "A"), agg.submit(frame(2, 3), "A")]
    fb = agg.submit(frame(3, 2), "B")
    fbad = agg.submit(frame(4, 1), "bad")
    assert (fa[0].result(5)[0] == 1).all() and (fa[1].result(5)[0] == 2).all() and (fb.result(5)[0] == 3).all()
    try:
        fbad.result(5)
        assert False
    except RuntimeError:
        pass
    assert sorted(calls) == [(2, "B"), (6, "A")]
    ok = agg.submit(frame(9, 1), "A").result(5)          # the loop survived the failed batch
    assert (ok[0] == 9).all()
    agg.close()


def test_many_producers():
    calls = []
    agg = BatchAggregator(fake_match(calls), max_batch=50, max_delay_ms=10)
    out = {}

    def cam(i):
        out[i] = [agg.submit(frame(i * 100 + j, 3)).result(10) for j in range(20)]

    th = [threading.Thread(target=cam, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(8):
        for j, (rows, _, _) in enumerate(out[i]):
            assert (rows == i * 100 + j).all()
    assert sum(c[0] for c in calls) == 8 * 20 * 3 and len(calls) < 8 * 20      # it did batch
    agg.close()


def test_two_workers_keep_two_batches_in_flight():
    """workers=2: while one batch is inside the matcher the next one is already running (its copies and
    launches overlap the first one's kernels on the GPU); results still route back to their requests."""
    lock = threading.Lock()
    state = {"now": 0, "peak": 0, "calls": 0}

    def slow(Q, tenant):
        with lock:
            state["now"] += 1
            state["calls"] += 1
            state["peak"] = max(state["peak"], state["now"])
        time.sleep(0.05)
        with lock:
            state["now"] -= 1
        return Q[:, :1].astype(np.int64), Q[:, 1:2].copy(), np.ones(len(Q), bool)

    agg = BatchAggregator(slow, max_batch=4, max_delay_ms=1, workers=2)
    futs = [agg.submit(frame(i, 4)) for i in range(8)]           # 8 full batches
    for i, f in enumerate(futs):
        rows, scores, acc = f.result(timeout=10)
        assert (rows[:, 0] == i).all() and list(scores[:, 0]) == [0, 1, 2, 3]
    agg.close()
    assert state["calls"] == 8 and state["peak"] == 2 and agg.batches == 8 and agg.faces == 32
