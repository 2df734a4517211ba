"""Property tests (hypothesis) of the oracle's own invariants - the tie / NaN / threshold edges that
SURVEY.md section 4 asks for - and of bench.py's reference arm (CPU only)."""
import json
import os
import subprocess
import sys

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import matcher_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@st.composite
def galleries(draw):
    n = draw(st.integers(1, 40))
    d = draw(st.sampled_from([4, 8, 16]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    G = rng.integers(-3, 4, size=(n, d)).astype(np.float32)          # small integers: exact ties are common
    if draw(st.booleans()) and n > 1:
        G[rng.integers(0, n)] = 0.0                                    # zero template -> NaN row
    Q = rng.integers(-3, 4, size=(draw(st.integers(1, 6)), d)).astype(np.float32)
    return G, Q


@settings(max_examples=150, deadline=None)
@given(galleries(), st.integers(1, 8))
def test_topk_extends_the_reference_scan(data, k):
    G, Q = data
    with np.errstate(all="ignore"):
        Gn = mo.normalise_rows(G)
        ids = ["%03d" % i for i in range(len(G))]
        rows1, scores1, acc1 = mo.match_frame(Q, ids, Gn, 0.4)
        S = np.stack([np.array([np.dot(mo.normalise(q), g) for g in Gn], np.float32) for q in Q])
        rk, sk = mo.topk_from_scores(S, k)
    # slot 0 of the top-k IS the reference's scan (first wins ties, NaN never wins, <= -1 never matches)
    assert (rk[:, 0] == rows1).all()
    filled = rows1 >= 0
    assert (sk[filled, 0] == scores1[filled]).all()
    for f in range(len(Q)):
        r, s = rk[f], sk[f]
        got = r[r >= 0]
        assert len(set(got)) == len(got)                               # distinct rows
        assert (np.diff(s[: len(got)]) <= 0).all()                     # non-increasing
        for a, b in zip(range(len(got) - 1), range(1, len(got))):      # ties -> lower row first
            if s[a] == s[b]:
                assert r[a] < r[b]
        assert not np.isnan(s).any() and (s[: len(got)] > -1).all()
        # nothing better was left out
        if len(got) < k:
            ok = (S[f] > -1)
            assert ok.sum() == len(got)
        else:
            rest = np.setdiff1d(np.nonzero(S[f] > -1)[0], got)
            assert (S[f][rest] <= s[len(got) - 1]).all()


@settings(max_examples=100, deadline=None)
@given(st.floats(-1.5, 1.5, allow_nan=False, width=32), st.sampled_from([0.35, 0.4, 0.45, 0.65]))
def test_threshold_is_compared_in_fp32(score, thr):
    s = np.float32(score)
    assert bool(mo.accept_fp32(np.array([s]), np.array([0]), thr)[0]) == bool(s >= np.float32(thr))
    assert mo.decide_live("id", s, thr)[0] == bool(s >= thr)           # numpy >= 2: same comparison
    assert not mo.decide_live(None, s, thr)[0]


def test_reference_arm_emits_the_contract_line():
    """bench.py --impl reference on a small gallery: one JSON line with the keys the driver reads."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000",
                        "--steps", "2", "--warmup", "1", "--cpu-procs", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["gallery_rows"] == 20000 and line["gpu_launches"] == 0


def test_adversarial_rounding_case_really_breaks_a_4e3_margin():
    """Premises of tests/rounding_case.py, checked in numpy: exact scores put A ahead of B by more than the 1e-4 id
    tolerance, the bf16 images put B ahead of A by more than 2 * 4e-3 - so a filter margin derived from a 2^-9
    unit roundoff would drop the true best row, while |error| stays within the rigorous 2^-7 + 2^-16."""
    from rounding_case import adversarial_pair, bf16_rn
    for seed in range(4):
        A, B, q = adversarial_pair(seed)
        n32 = np.sqrt(np.sum(q * q, dtype=np.float32))
        assert abs(float(n32) - 1.0) < 1e-6
        qn = (q / n32).astype(np.float32)
        sA, sB = float(qn.astype(np.float64) @ A), float(qn.astype(np.float64) @ B)
        SA = float(bf16_rn(qn).astype(np.float64) @ bf16_rn(A))
        SB = float(bf16_rn(qn).astype(np.float64) @ bf16_rn(B))
        assert sA - sB > 2e-4 and SB - SA > 8.2e-3
        assert abs(SA - sA) <= (2.0 ** -7 + 2.0 ** -16) * 1.005 and abs(SB - sB) <= (2.0 ** -7 + 2.0 ** -16) * 1.005


def test_fast_topk_oracle_equals_the_stable_sort_oracle():
    """match_topk_fast (partial selection, chunked) == match_topk (full stable sort) on data with exact ties, a NaN
    row, a zero query, removed rows and a tenant filter."""
    from oracle import matcher_oracle as mo
    rng = np.random.default_rng(11)
    n, d = 3000, 64
    G = mo.normalise_rows(rng.standard_normal((n, d)).astype(np.float32))
    G[100:140] = G[7]                       # exact ties
    G[500] = np.nan
    tags = rng.integers(0, 3, n).astype(np.int32)
    tags[rng.integers(0, n, 200)] = -1
    Q = rng.standard_normal((70, d)).astype(np.float32)
    Q[3] = G[7] * 2.5
    Q[5] = 0
    Q[6, 2] = np.nan
    for k in (1, 5, 16):
        for tg, tn in ((None, None), (tags, None), (tags, 1), (tags, 7)):
            a = mo.match_topk(Q, G, k, 0.45, tg, tn)
            b = mo.match_topk_fast(Q, G, k, 0.45, tg, tn, chunk=32)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
