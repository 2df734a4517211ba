"""numpy restatement of the reference's matching path.  TEST INFRASTRUCTURE - never shipped.

Every function cites the reference lines it follows (paths relative to /root/reference).
The restatement is pinned by ``tests/golden/*.npz`` (outputs of the reference's own code,
executed verbatim by ``oracle/ref_harness.py``); see ``tests/test_oracle_golden.py``.

Conventions shared with the CUDA path
  * a *gallery* is the ordered sequence of (id, unit fp32 vector); order = dict insertion
    order of ``EmbeddingManager.embeddings`` (infrenceServer.py:49, peopleCount.py:707).
  * "row" = position in that order.  No match -> row -1, score -1.0 (the loop's initial
    ``best_score = -1``; infrenceServer.py:535-536, peopleCount.py:866-867).
  * comparisons are done in fp32 against ``np.float32(threshold)``: NumPy >= 2 (NEP 50)
    demotes the Python-float threshold to the fp32 score's dtype.
"""
from __future__ import annotations

from collections import OrderedDict, deque
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

LIVE_THRESHOLD = 0.4          # infrenceServer.py:407
CAMPUS_THRESHOLD = 0.45       # peopleCount.py:829
CAMPUS_UNKNOWN = 0.35         # peopleCount.py:830
UNKNOWN_CLUSTER = 0.65        # peopleCount.py:232
DUPLICATE_THRESHOLD = 0.4     # trainingServer.py:70-71 (duplicate / same-person)
SAME_PERSON_THRESHOLD = 0.4

NO_ROW = -1
NO_SCORE = np.float32(-1.0)


# ----------------------------------------------------------------------------- a1 / a2
def normalise(v: np.ndarray) -> np.ndarray:
    """``v / np.linalg.norm(v)`` exactly as the reference spells it.

    Query: infrenceServer.py:532, peopleCount.py:863.  Gallery row at load:
    infrenceServer.py:271,324; peopleCount.py:788,806.  A zero vector yields a NaN row,
    which can never win the strict ``>`` of the scan.
    """
    v = np.asarray(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        return v / np.linalg.norm(v)


def normalise_rows(m: np.ndarray) -> np.ndarray:
    """Row-wise :func:`normalise` (one np.linalg.norm call per row, like the per-row loaders)."""
    m = np.asarray(m, dtype=np.float32)
    out = np.empty_like(m)
    for i in range(m.shape[0]):
        out[i] = normalise(m[i])
    return out


# ----------------------------------------------------------------------------- a3 / a4
def scan_best(face_embedding: np.ndarray, embeddings: Dict[str, np.ndarray]):
    """The reference's per-face scan: infrenceServer.py:534-542 == peopleCount.py:865-873.

    Strict ``>`` from ``best_score = -1``: earliest entry wins exact ties, NaN never wins,
    scores <= -1 never match.
    """
    best_match_id = None
    best_score = -1
    for person_id, registered_embedding in embeddings.items():
        similarity = np.dot(face_embedding, registered_embedding)
        if similarity > best_score:
            best_score = similarity
            best_match_id = person_id
    return best_match_id, best_score


# ----------------------------------------------------------------------------- a5
def decide_live(best_match_id, best_score, threshold: float = LIVE_THRESHOLD):
    """infrenceServer.py:545-552: accept iff an id was found and score >= threshold;
    a rejected face is reported with score 0."""
    if best_match_id and best_score >= threshold:
        return True, best_score
    return False, 0


def decide_campus(best_match_id, best_score, threshold: float = CAMPUS_THRESHOLD,
                  unknown_threshold: float = CAMPUS_UNKNOWN) -> str:
    """peopleCount.py:876-887: 'recognized' | 'unknown' | 'ignored' (the grey band)."""
    if best_match_id and best_score >= threshold:
        return "recognized"
    elif best_score < unknown_threshold:
        return "unknown"
    return "ignored"


def accept_fp32(score: np.ndarray, row: np.ndarray, threshold: float) -> np.ndarray:
    """Vectorised a5 in the arithmetic the in-container oracle uses: fp32 score >= fp32(thr)."""
    score = np.asarray(score, dtype=np.float32)
    return (np.asarray(row) >= 0) & (score >= np.float32(threshold))


# ----------------------------------------------------------------------------- batch form
def match_frame(faces: np.ndarray, ids: Sequence[str], gallery: np.ndarray, threshold: float):
    """Run a1 + a3/a4 + a5 for a batch of raw face embeddings against an ordered gallery,
    through the very same per-face dict loop.  Returns (rows, scores fp32, accept)."""
    embeddings = OrderedDict((i, g) for i, g in zip(ids, gallery))
    pos = {i: r for r, i in enumerate(ids)}
    rows = np.full(len(faces), NO_ROW, dtype=np.int64)
    scores = np.full(len(faces), NO_SCORE, dtype=np.float32)
    accept = np.zeros(len(faces), dtype=bool)
    for f, e in enumerate(faces):
        q = normalise(e)
        bid, bs = scan_best(q, embeddings)
        if bid is not None:
            rows[f] = pos[bid]
            scores[f] = bs
        accept[f] = bool(bid and bs >= threshold)
    return rows, scores, accept


def cosine_scores(Q: np.ndarray, G: np.ndarray, renormalise: bool = True) -> np.ndarray:
    """fp32 score matrix ``normalise(Q) @ G.T``.  Vectorised stand-in for the F*N np.dot calls
    (sgemm instead of sdot: summation order differs, scores agree to ~1e-6, which is why the
    id comparison carries the gap exemption of BASELINE.json)."""
    Q = np.asarray(Q, dtype=np.float32)
    if renormalise:
        Q = normalise_rows(Q)
    return (Q @ np.asarray(G, dtype=np.float32).T).astype(np.float32)


def topk_from_scores(S: np.ndarray, k: int, mask: Optional[np.ndarray] = None):
    """a6 (OUR extension - the reference is top-1 only; parity unpinned for k > 1): the first
    k entries of a STABLE descending sort over gallery order, restricted to scores > -1
    (NaN excluded).  k = 1 reduces exactly to :func:`scan_best`.

    ``mask`` bool[N] selects participating rows (tenant filter a9 / live rows).
    Returns (rows int64[F,k], scores fp32[F,k]); unfilled slots are (-1, -1.0).
    """
    S = np.asarray(S, dtype=np.float32)
    F, N = S.shape
    rows = np.full((F, k), NO_ROW, dtype=np.int64)
    scores = np.full((F, k), NO_SCORE, dtype=np.float32)
    for f in range(F):
        s = S[f]
        ok = s > np.float32(-1)          # False for NaN
        if mask is not None:
            ok &= mask
        cand = np.nonzero(ok)[0]
        if cand.size == 0:
            continue
        order = cand[np.argsort(-s[cand], kind="stable")][:k]
        rows[f, :order.size] = order
        scores[f, :order.size] = s[order]
    return rows, scores


def match_topk(Q: np.ndarray, G: np.ndarray, k: int = 1, threshold: float = LIVE_THRESHOLD,
               tags: Optional[np.ndarray] = None, tenant: Optional[int] = None,
               renormalise: bool = True):
    """Whole-batch oracle: (rows[F,k], scores[F,k], accept[F]).  ``tags`` int32[N] holds the
    tenant of each row (-1 = removed row); ``tenant`` None/-1 = match across all tenants
    (peopleCount.py:848) else only rows of that tenant (infrenceServer.py:343-380)."""
    G = np.asarray(G, dtype=np.float32)
    F = len(Q)
    if G.shape[0] == 0:
        return (np.full((F, k), NO_ROW, np.int64), np.full((F, k), NO_SCORE, np.float32),
                np.zeros(F, bool))
    S = cosine_scores(Q, G, renormalise)
    mask = None
    if tags is not None:
        tags = np.asarray(tags)
        mask = tags >= 0
        if tenant is not None and tenant >= 0:
            mask &= tags == tenant
    rows, scores = topk_from_scores(S, k, mask)
    return rows, scores, accept_fp32(scores[:, 0], rows[:, 0], threshold)


def match_topk_fast(Q: np.ndarray, G: np.ndarray, k: int = 1, threshold: float = LIVE_THRESHOLD,
                    tags: Optional[np.ndarray] = None, tenant: Optional[int] = None, chunk: int = 128):
    """:func:`match_topk` for LARGE inputs (1024 x 1 M): same definition, same results (checked against it in
    tests/test_oracle_properties.py), but the score matrix is produced in query chunks and the k best of a row are
    found with a partial selection instead of a full stable sort: everything >= the k-th largest score is kept
    (ties included), then ordered by (score desc, row asc) - which IS the stable descending sort's prefix."""
    Q = np.asarray(Q, dtype=np.float32)
    G = np.asarray(G, dtype=np.float32)
    F, N = len(Q), len(G)
    rows = np.full((F, k), NO_ROW, dtype=np.int64)
    scores = np.full((F, k), NO_SCORE, dtype=np.float32)
    if N == 0:
        return rows, scores, np.zeros(F, bool)
    mask = None
    if tags is not None:
        tags = np.asarray(tags)
        mask = tags >= 0
        if tenant is not None and tenant >= 0:
            mask &= tags == tenant
    for a in range(0, F, chunk):
        S = cosine_scores(Q[a:a + chunk], G)
        S[~(S > np.float32(-1))] = -np.inf                 # scores <= -1 and NaN never match
        if mask is not None:
            S[:, ~mask] = -np.inf
        kk = min(k, N)
        part = np.partition(S, N - kk, axis=1)[:, N - kk]  # k-th largest per row
        for f in range(S.shape[0]):
            s = S[f]
            cand = np.nonzero((s >= part[f]) & (s > -np.inf))[0]
            order = cand[np.lexsort((cand, -s[cand]))][:k]
            rows[a + f, :order.size] = order
            scores[a + f, :order.size] = s[order]
    return rows, scores, accept_fp32(scores[:, 0], rows[:, 0], threshold)


def euclidean_topk(Q: np.ndarray, G: np.ndarray, k: int = 1, tolerance: float = 0.6,
                   tags: Optional[np.ndarray] = None, tenant: Optional[int] = None):
    """Config 3 (128-d Euclidean).  NOT in the reference - OUR definition, parity unpinned:
    d_j = ||g_j - q||_2 (direct difference form, dlib/face_recognition convention), the k
    smallest distances in stable gallery order, accept iff d <= tolerance.  Neither side is
    normalised.  Unfilled slots are (-1, +inf)."""
    Q = np.asarray(Q, dtype=np.float32)
    G = np.asarray(G, dtype=np.float32)
    F = len(Q)
    rows = np.full((F, k), NO_ROW, dtype=np.int64)
    dist = np.full((F, k), np.inf, dtype=np.float32)
    mask = np.ones(len(G), bool)
    if tags is not None:
        tags = np.asarray(tags)
        mask = tags >= 0
        if tenant is not None and tenant >= 0:
            mask &= tags == tenant
    for f in range(F):
        diff = (G.astype(np.float64) - Q[f].astype(np.float64))
        d = np.sqrt(np.sum(diff * diff, axis=1)).astype(np.float32)
        ok = mask & np.isfinite(d)
        cand = np.nonzero(ok)[0]
        order = cand[np.argsort(d[cand], kind="stable")][:k]
        rows[f, :order.size] = order
        dist[f, :order.size] = d[order]
    accept = (rows[:, 0] >= 0) & (dist[:, 0] <= np.float32(tolerance))
    return rows, dist, accept


# ----------------------------------------------------------------------------- a7 / a8 / a9
class GalleryOracle:
    """Restates the residency/update semantics of both ``EmbeddingManager`` variants as an
    ordered id -> (unit vector, tenant, kind) map.

    * upsert: ``self.embeddings[id] = v / ||v||`` - a new id appends, an existing id is
      overwritten IN PLACE (dict keeps its position): infrenceServer.py:273,326;
      peopleCount.py:790,808.
    * remove: ``del self.embeddings[id]`` (infrenceServer.py:248-251); a later re-enrol
      appends at the END.
    * subset(tenant): infrenceServer.py:343-380 (order inside the subset is set-iteration
      order in the reference, i.e. undefined; we define it as gallery order).
    * snapshot(): peopleCount.py:816-819 (dict copy under the lock).
    """

    def __init__(self, dim: int):
        self.dim = dim
        self.embeddings: "OrderedDict[str, np.ndarray]" = OrderedDict()
        self.meta: Dict[str, Dict] = {}

    def upsert(self, pid: str, vec: np.ndarray, tenant: int = 0, kind: str = "employee"):
        self.embeddings[pid] = normalise(np.asarray(vec, dtype=np.float32))
        self.meta[pid] = {"tenant": int(tenant), "type": kind}

    def remove(self, pid: str) -> bool:
        if pid in self.embeddings:
            del self.embeddings[pid]
            self.meta.pop(pid, None)
            return True
        return False

    def snapshot(self):
        ids = list(self.embeddings.keys())
        G = (np.stack([self.embeddings[i] for i in ids]).astype(np.float32)
             if ids else np.zeros((0, self.dim), np.float32))
        tags = np.array([self.meta[i]["tenant"] for i in ids], dtype=np.int32)
        return ids, G, tags

    def get_stats(self) -> Dict:
        """Counting part of infrenceServer.py:386-398."""
        return {
            "total_embeddings": len(self.embeddings),
            "employees": sum(1 for m in self.meta.values() if m["type"] == "employee"),
            "visitors": sum(1 for m in self.meta.values() if m["type"] == "visitor"),
        }

    def match(self, Q, k=1, threshold=LIVE_THRESHOLD, tenant: Optional[int] = None):
        ids, G, tags = self.snapshot()
        rows, scores, accept = match_topk(Q, G, k, threshold, tags, tenant)
        names = [[ids[r] if r >= 0 else None for r in rr] for rr in rows]
        return names, scores, accept


# ----------------------------------------------------------------------------- f1
class UnknownClusterOracle:
    """peopleCount.py:52-91 (UnknownPerson) + :432-500 (process_unknown_detection), the
    arithmetic only.  Each cluster keeps the last 10 embeddings and their plain mean (NOT
    re-normalised, :74); a new face joins the FIRST cluster, in creation order, whose raw dot
    with the mean is a running maximum AND >= 0.65 (:446-452), else founds a new cluster."""

    def __init__(self, threshold: float = UNKNOWN_CLUSTER, window: int = 10):
        self.threshold = threshold
        self.window = window
        self.members: List[deque] = []
        self.avg: List[np.ndarray] = []
        self.count: List[int] = []

    def observe(self, q: np.ndarray) -> Tuple[int, bool]:
        """Returns (cluster index, created)."""
        matched = None
        best_similarity = -1
        for ci, avg in enumerate(self.avg):
            similarity = np.dot(avg, q)
            if similarity > best_similarity:
                best_similarity = similarity
                if similarity >= self.threshold:
                    matched = ci
                    break
        if matched is not None:
            self.members[matched].append(q)
            self.avg[matched] = np.mean(list(self.members[matched]), axis=0)
            self.count[matched] += 1
            return matched, False
        d = deque(maxlen=self.window)
        d.append(q)
        self.members.append(d)
        self.avg.append(q)
        self.count.append(1)
        return len(self.avg) - 1, True


# ----------------------------------------------------------------------------- f2
def duplicate_check(new_embedding: np.ndarray, existing: Iterable[np.ndarray],
                    threshold: float = DUPLICATE_THRESHOLD) -> int:
    """trainingServer.py:170-200: first stored (RAW, un-normalised) embedding in cursor order
    with cos > threshold; returns its position or -1."""
    for j, existing_embedding in enumerate(existing):
        sim = np.dot(new_embedding, existing_embedding) / (
            np.linalg.norm(new_embedding) * np.linalg.norm(existing_embedding))
        if sim > threshold:
            return j
    return -1


def same_person_check(embeddings: Sequence[np.ndarray],
                      threshold: float = SAME_PERSON_THRESHOLD) -> Tuple[bool, Optional[Tuple[int, int]]]:
    """trainingServer.py:202-214: every pair of pose embeddings must have cos >= threshold."""
    if len(embeddings) < 2:
        return True, None
    for i in range(len(embeddings)):
        for j in range(i + 1, len(embeddings)):
            sim = np.dot(embeddings[i], embeddings[j]) / (
                np.linalg.norm(embeddings[i]) * np.linalg.norm(embeddings[j]))
            if sim < threshold:
                return False, (i, j)
    return True, None


def enrol_mean(pose_embeddings: Sequence[np.ndarray]) -> np.ndarray:
    """trainingServer.py:355: the stored template is the plain fp32 mean of <= 3 pose vectors."""
    return np.mean(pose_embeddings, axis=0)


# ----------------------------------------------------------------------------- parity helper
def ids_match_with_gap(ref_rows: np.ndarray, ref_scores: np.ndarray, got_rows: np.ndarray,
                       tol: float) -> np.ndarray:
    """BASELINE.json's id rule: slot j of a query must carry the same row wherever the oracle's
    score at j is separated from BOTH neighbours (j-1, j+1) by more than ``tol``; inside a
    near-tie cluster any permutation of the cluster is acceptable.  Returns bool[F] (per query).
    The oracle must be evaluated at k+1 so the last slot has a right-hand neighbour."""
    ref_rows = np.asarray(ref_rows)
    got_rows = np.asarray(got_rows)
    F, k = got_rows.shape
    ok = np.ones(F, bool)
    rs = np.asarray(ref_scores, dtype=np.float64)
    for f in range(F):
        for j in range(k):
            if ref_rows[f, j] < 0:
                if got_rows[f, j] >= 0:
                    ok[f] = False
                continue
            left = j == 0 or (rs[f, j - 1] - rs[f, j]) > tol
            right = (j + 1 >= rs.shape[1]) or ref_rows[f, j + 1] < 0 or (rs[f, j] - rs[f, j + 1]) > tol
            if left and right:
                if got_rows[f, j] != ref_rows[f, j]:
                    ok[f] = False
            else:
                # near-tie: the returned row must be one of the oracle rows within tol of slot j
                near = ref_rows[f][np.abs(rs[f] - rs[f, j]) <= tol]
                if got_rows[f, j] not in near:
                    ok[f] = False
    return ok
