"""The matching service: query embeddings in; identity ids, scores and accept/reject out.

Replaces the inline per-face loop of the reference -
``FaceRecognitionProcessor.recognize_faces`` (infrenceServer.py:530-552) and
``CameraProcessor.process_frame`` (peopleCount.py:860-887) - with one batched call into
libfrg.so.  The two thin classes at the bottom keep the reference's own names, thresholds and
result shapes so a call site can be switched over line for line.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _native as N
from .gallery import GalleryStore

LIVE_THRESHOLD = 0.4         # infrenceServer.py:407
CAMPUS_THRESHOLD = 0.45      # peopleCount.py:829
CAMPUS_UNKNOWN = 0.35        # peopleCount.py:830


@dataclass
class MatchResult:
    rows: np.ndarray      # int64 [F, k]   gallery rows, -1 = none
    scores: np.ndarray    # fp32  [F, k]   cosine score (or Euclidean distance), -1.0 / +inf = none
    accept: np.ndarray    # bool  [F]      decision on slot 0 at the call's threshold
    ids: Optional[List[List[Optional[str]]]] = None
    variant: str = ""
    launches: int = 0
    layout_version: int = 0      # GalleryStore.layout_version the rows belong to (compact() renumbers rows)


def _params(metric: str, variant: str, threshold: float, tenant: int, row_offset: int) -> N.MatchParams:
    return N.MatchParams(N.METRICS[metric], N.VARIANTS[variant], float(np.float32(threshold)),
                         int(tenant), int(row_offset), 0, 0)


def _reading(store):
    """The store's read section (GalleryStore.reading): match + row -> id translation are atomic with respect to
    compact().  Stores without one (sharded galleries never renumber rows; test doubles) need none."""
    fn = getattr(store, "reading", None)
    return fn() if fn is not None else contextlib.nullcontext()


class Matcher:
    """``match(Q)`` for host (numpy) batches, ``match_device(Q)`` for batches already on the GPU."""

    def __init__(self, store: GalleryStore, metric: str = "cosine"):
        self.store = store
        self.metric = metric

    # ---- host buffers in, host buffers out: H2D + kernels + D2H inside one C call
    def match(self, Q: np.ndarray, k: int = 1, threshold: float = LIVE_THRESHOLD,
              company_id: Optional[str] = None, variant: str = "auto", row_offset: int = 0,
              with_ids: bool = True, out: Optional[MatchResult] = None) -> MatchResult:
        Q = np.ascontiguousarray(Q, dtype=np.float32).reshape(-1, self.store.dim)
        F = len(Q)
        if out is None:
            out = MatchResult(np.empty((F, k), np.int64), np.empty((F, k), np.float32), np.zeros(F, np.uint8))
        tenant = -1 if company_id is None else self.store.tenant_code(company_id, create=False)
        p = _params(self.metric, variant, threshold, tenant, row_offset)
        # one read section: a compaction cannot renumber the rows between the match and their translation.
        # With with_ids=False the caller translates later - inside its own `with store.reading():` around this
        # call (the processors below), or checked through `ids_of(rows, layout_version=out.layout_version)`.
        with _reading(self.store):
            out.layout_version = getattr(self.store, "layout_version", 0)
            N.check(N.lib.frg_match_host(self.store.handle, Q.ctypes.data, F, int(k), C.byref(p),
                                         out.rows.ctypes.data, out.scores.ctypes.data,
                                         out.accept.ctypes.data))
            out.variant, out.launches = N.last_variant(), N.last_launch_count()
            if out.accept.dtype != np.bool_:
                out.accept = out.accept.view(np.bool_)
            if with_ids:
                out.ids = self.store.ids_of(out.rows)
        return out

    # ---- device tensors in/out, enqueued on the caller's stream (torch is only the allocator here)
    def match_device(self, Q, k: int = 1, threshold: float = LIVE_THRESHOLD, company_id: Optional[str] = None,
                     variant: str = "auto", row_offset: int = 0, out=None, stream: Optional[int] = None,
                     tenant: Optional[int] = None):
        """tenant: a tag code given directly (sharded galleries keep their own id / tenant tables)."""
        import torch
        assert Q.is_cuda and Q.dtype == torch.float32 and Q.is_contiguous()
        F = Q.shape[0]
        if out is None:
            out = (torch.empty((F, k), dtype=torch.int64, device=Q.device),
                   torch.empty((F, k), dtype=torch.float32, device=Q.device),
                   torch.empty((F,), dtype=torch.uint8, device=Q.device))
        rows, scores, accept = out
        if stream is None:
            stream = torch.cuda.current_stream(Q.device).cuda_stream
        if tenant is None:
            tenant = -1 if company_id is None else self.store.tenant_code(company_id, create=False)
        p = _params(self.metric, variant, threshold, tenant, row_offset)
        N.check(N.lib.frg_match(self.store.handle, C.c_void_p(Q.data_ptr()), F, int(k), C.byref(p),
                                C.c_void_p(rows.data_ptr()), C.c_void_p(scores.data_ptr()),
                                C.c_void_p(accept.data_ptr()), C.c_void_p(stream)))
        return rows, scores, accept

    # ---- first row, in gallery order, whose score reaches the threshold (exact fp32)
    def first_above(self, Q: np.ndarray, threshold: float, strict: bool = False, company_id: Optional[str] = None,
                    query_prenormalised: bool = False):
        """(rows int64 [F], scores fp32 [F]); row -1 / score -1.0 when no row qualifies.  The scan rule of
        the enrol-time duplicate check (trainingServer.py:170-200, strict) and of unknown-person
        clustering (peopleCount.py:446-452)."""
        Q = np.ascontiguousarray(Q, dtype=np.float32).reshape(-1, self.store.dim)
        F = len(Q)
        rows, scores = np.empty(F, np.int64), np.empty(F, np.float32)
        tenant = -1 if company_id is None else self.store.tenant_code(company_id, create=False)
        p = _params(self.metric, "scan_f32", threshold, tenant, 0)
        p.flags = (N.FIRST_STRICT if strict else 0) | (N.QUERY_PRENORMALISED if query_prenormalised else 0)
        N.check(N.lib.frg_first_match_host(self.store.handle, Q.ctypes.data, F, C.byref(p),
                                           rows.ctypes.data, scores.ctypes.data))
        return rows, scores


class FaceRecognitionProcessor:
    """Drop-in for the matching half of infrenceServer.FaceRecognitionProcessor (:400-563).

    ``recognize(embeddings, company_id)`` takes the ``face.normed_embedding`` of every detected
    face and returns, per face, what the reference passes to its draw call (:545-558):
    ``person_info`` (metadata, or ``{'name': 'Unknown', 'type': 'unknown'}``), the reported score
    (0 when rejected) and the matched id (None when rejected)."""

    def __init__(self, store, recognition_threshold: float = LIVE_THRESHOLD, matcher=None):
        """store: a ``GalleryStore`` - or a ``ShardedGallery`` together with ``matcher=ShardedMatcher(store)``
        (one process per GPU; every rank makes the same calls)."""
        self.store = store
        self.matcher = matcher if matcher is not None else Matcher(store)
        self._match = getattr(self.matcher, "match_host", None) or self.matcher.match
        self.recognition_threshold = recognition_threshold

    def recognize(self, embeddings: np.ndarray, company_id: Optional[str] = None) -> List[Dict]:
        store = self.store
        if len(embeddings) == 0 or len(store) == 0:      # `if not embeddings: return frame` (:523-525)
            return []
        out = []
        with _reading(store):        # match + id lookup are one read section (compact() renumbers rows)
            r = self._match(embeddings, 1, self.recognition_threshold, company_id, with_ids=False)
            for f in range(len(embeddings)):
                if r.accept[f]:
                    pid = store.id_of(r.rows[f, 0])            # ids only for the faces that matched
                    info = store.metadata(pid) or {"name": pid, "type": "employee"}
                    out.append({"person_id": pid, "person_info": info, "recognition_score": r.scores[f, 0]})
                else:
                    out.append({"person_id": None, "person_info": {"name": "Unknown", "type": "unknown"},
                                "recognition_score": 0})
        return out


class CameraProcessor:
    """Drop-in for the matching half of peopleCount.CameraProcessor (:822-896): three-way decision
    at 0.45 / 0.35 over ALL tenants, and the same stats dict."""

    def __init__(self, store, recognition_threshold: float = CAMPUS_THRESHOLD,
                 unknown_threshold: float = CAMPUS_UNKNOWN, matcher=None):
        self.store = store
        self.matcher = matcher if matcher is not None else Matcher(store)
        self._match = getattr(self.matcher, "match_host", None) or self.matcher.match
        self.recognition_threshold = recognition_threshold
        self.unknown_threshold = unknown_threshold

    def process(self, embeddings: np.ndarray):
        """Returns (events, stats): events[f] is ('recognized', id, float score) |
        ('unknown', None, None) | ('ignored', None, None)."""
        stats = {"faces": 0, "recognized": 0, "unknown": 0}
        store = self.store
        if len(store) == 0:                               # peopleCount.py:850-851
            return [], stats
        stats["faces"] = len(embeddings)
        if len(embeddings) == 0:
            return [], stats
        events = []
        with _reading(store):        # match + id lookup are one read section (compact() renumbers rows)
            r = self._match(embeddings, 1, self.recognition_threshold, with_ids=False)
            # three-way decision in fp32 (peopleCount.py:876-887); ids are looked up only for recognised faces
            accept = r.accept.tolist()
            below = (r.scores[:, 0] < np.float32(self.unknown_threshold)).tolist()   # best_score stays -1 when nothing matched
            rows, scores = r.rows[:, 0].tolist(), r.scores[:, 0].tolist()
            for f in range(len(accept)):
                if accept[f]:
                    events.append(("recognized", store.id_of(rows[f]), scores[f]))
                elif below[f]:
                    events.append(("unknown", None, None))
                else:
                    events.append(("ignored", None, None))
        stats["recognized"] = sum(accept)
        stats["unknown"] = sum(1 for a, b in zip(accept, below) if b and not a)
        return events, stats
