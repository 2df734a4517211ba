"""Online clustering of unknown faces on the device (BASELINE config 5's
"online enrol/update").  Mirrors peopleCount.UnknownPerson (:52-91) and the matching rule of
CampusPeopleManager.process_unknown_detection (:432-500):

  * a cluster keeps its last 10 embeddings; its representative is their plain mean, NOT re-normalised
    (:63-74) -> the clusters live in a FRG_STORE_RAW gallery (rows stored as given);
  * a face is compared by RAW dot product with each representative in creation order, the running
    best is tracked, and the first cluster with dot >= 0.65 is joined (:446-452).  Every earlier
    cluster scored < 0.65, so this is "first row with dot >= 0.65": frg_first_match, one launch.
  * otherwise the face founds a new cluster (unknown_{campus}_{n+1}, :475).

The gallery is tiny and mutates at every observation; each observation is one first_match call plus
one single-row upsert, both enqueued on the store's stream order.
"""
from __future__ import annotations

from collections import deque
from typing import Deque, Dict, List, Tuple

import numpy as np

from .gallery import GalleryStore
from .matcher import Matcher

UNKNOWN_SIMILARITY_THRESHOLD = 0.65     # peopleCount.py:232


class UnknownClusterer:
    def __init__(self, dim: int = 512, device: int = 0, threshold: float = UNKNOWN_SIMILARITY_THRESHOLD,
                 window: int = 10, campus_id: str = "campus"):
        self.store = GalleryStore(dim=dim, capacity=256, device=device, bf16_plane=False, raw=True)
        self.matcher = Matcher(self.store)
        self.threshold = threshold
        self.window = window
        self.campus_id = campus_id
        self.members: List[Deque[np.ndarray]] = []
        self.detection_count: List[int] = []

    def __len__(self):
        return len(self.members)

    def unknown_id(self, cluster: int) -> str:
        return "unknown_%s_%d" % (self.campus_id, cluster + 1)         # peopleCount.py:475

    def observe(self, face_embedding: np.ndarray) -> Tuple[int, bool]:
        """face_embedding: the normalised query the reference passes on (peopleCount.py:884).  Returns
        (cluster ordinal, created)."""
        q = np.ascontiguousarray(face_embedding, dtype=np.float32)
        hit = -1
        if self.members:
            rows, _ = self.matcher.first_above(q[None, :], self.threshold, strict=False, query_prenormalised=True)
            hit = int(rows[0])
        if hit >= 0:
            self.members[hit].append(q)
            avg = np.mean(list(self.members[hit]), axis=0)             # peopleCount.py:74, not re-normalised
            self.store.overwrite_rows([hit], avg[None, :], np.zeros(1, np.int32))
            self.detection_count[hit] += 1
            return hit, False
        d: Deque[np.ndarray] = deque(maxlen=self.window)
        d.append(q)
        self.members.append(d)
        self.detection_count.append(1)
        self.store.append_rows(q[None, :], np.zeros(1, np.int32))
        return len(self.members) - 1, True

    def representatives(self) -> np.ndarray:
        vecs, _ = self.store.read_rows()
        return vecs

    def close(self):
        self.store.close()
