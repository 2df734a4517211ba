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
