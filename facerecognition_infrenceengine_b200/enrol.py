"""Enrolment-time checks of the reference's embedding worker, on the device gallery
(trainingServer.py:170-214).

  * duplicate face:   first stored template of the company, in cursor order, with cos > 0.4
                      (trainingServer.py:170-200) -> one exact scan over the tenant's rows instead of
                      one GridFS fetch + unpickle + np.dot per stored person.
  * same person:      every pair of the <= 3 pose embeddings must have cos >= 0.4
                      (trainingServer.py:202-214).
  * stored template:  plain fp32 mean of the pose embeddings (trainingServer.py:355); the gallery
                      normalises it on ingest.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from .gallery import GalleryStore
from .matcher import Matcher

SIMILARITY_THRESHOLD = 0.4     # trainingServer.py:70
DUPLICATE_THRESHOLD = 0.4      # trainingServer.py:71


class EnrolmentChecker:
    def __init__(self, store, duplicate_threshold: float = DUPLICATE_THRESHOLD,
                 similarity_threshold: float = SIMILARITY_THRESHOLD, matcher=None):
        """store: a ``GalleryStore``, or a ``ShardedGallery`` with ``matcher=ShardedMatcher(store)`` (the
        duplicate scan then runs on every rank's block and the lowest global row wins; collective)."""
        self.store = store
        self.matcher = matcher if matcher is not None else Matcher(store)
        self.duplicate_threshold = duplicate_threshold
        self.similarity_threshold = similarity_threshold
        self._scratch: Optional[GalleryStore] = None

    def check_duplicate_face(self, new_embedding: np.ndarray, company_id: Optional[str] = None) -> Tuple[bool, Optional[str]]:
        """(is_duplicate, id of the FIRST stored person with cos > threshold) - trainingServer.py:170-200."""
        rows, _ = self.matcher.first_above(np.asarray(new_embedding, np.float32)[None, :], self.duplicate_threshold,
                                           strict=True, company_id=company_id)
        if rows[0] < 0:
            return False, None
        return True, self.store.id_of(int(rows[0]))

    def check_image_similarity(self, embeddings: Sequence[np.ndarray]) -> Tuple[bool, Optional[Tuple[int, int]]]:
        """(all poses show the same person, first offending pair) - trainingServer.py:202-214.  The pose
        embeddings go through the same device path: a scratch gallery of the poses, every pose matched
        against it, pairs read off in the reference's (i, j > i) order."""
        n = len(embeddings)
        if n < 2:
            return True, None
        if self._scratch is None:
            dev = self.store.device if getattr(self.store, "device", None) is not None else self.store.store.device
            self._scratch = GalleryStore(dim=self.store.dim, capacity=16, device=dev, bf16_plane=False)
        sc = self._scratch
        if sc.rows:
            sc.remove_rows(list(range(sc.rows)))
            sc.compact()
        E = np.stack([np.asarray(e, np.float32) for e in embeddings])
        sc.append_rows(E)
        r = Matcher(sc).match(E, k=n, threshold=self.similarity_threshold, variant="scan_f32", with_ids=False)
        S = np.full((n, n), np.nan, np.float32)
        for i in range(n):
            for j in range(n):
                if r.rows[i, j] >= 0:
                    S[i, r.rows[i, j]] = r.scores[i, j]
        thr = np.float32(self.similarity_threshold)
        for i in range(n):
            for j in range(i + 1, n):
                if S[i, j] < thr:
                    return False, (i, j)
        return True, None

    @staticmethod
    def template_from_poses(pose_embeddings: Sequence[np.ndarray]) -> np.ndarray:
        """trainingServer.py:355."""
        return np.mean(pose_embeddings, axis=0)
