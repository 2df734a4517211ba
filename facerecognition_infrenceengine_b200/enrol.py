"""Enrolment-time checks of the reference's embedding worker, on the device gallery
(trainingServer.py:170-214).

  * duplicate face:   first stored template of the company, in cursor order, with cos > 0.4
                      (trainingServer.py:170-200) -> one exact scan over the tenant's rows instead of
                      one GridFS fetch + unpickle + np.dot per stored person.
  * same person:      every pair of the <= 3 pose embeddings must have cos >= 0.4
                      (trainingServer.py:202-214).
  * stored template:  plain fp32 mean of the pose embeddings (trainingServer.py:355); the gallery
                      normalises it on ingest.

Which gallery to scan.  The reference's duplicate scan reads the Mongo collection, not the live cache: EVERY
document of the company that has an embeddingId - whatever its status, blacklisted or not - and only the
collection being enrolled into (employees OR visitors, chosen by `id_field`).  The live matching gallery
(`EmbeddingManager.store`) is the wrong thing to scan: it has evicted exactly the inactive / blacklisted people
(infrenceServer.py:234-258) a re-enrolment must still be caught against, and it mixes employees and visitors
under one tenant tag.  Use an :class:`EnrolmentGallery` - every enrolled template, tagged (company, kind),
never evicted - and pass `kind=` to ``check_duplicate_face``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .gallery import GalleryStore
from .matcher import Matcher, _reading


def enrolment_tenant(company_id: Optional[str], kind: Optional[str]) -> Optional[str]:
    """Tenant key of an enrolment gallery: one tag per (company, collection) - trainingServer.py:174-178 scans
    `collection.find({'companyId': company_id, ...})` of ONE collection."""
    if kind is None:
        return company_id
    return "%s|%s" % ("" if company_id is None else company_id, kind)


class EnrolmentGallery:
    """Every enrolled template of every company, whatever the person's status: what trainingServer.py:170-200
    scans.  A thin id-level wrapper over a ``GalleryStore`` whose tenant tag is (company, kind) and whose
    metadata keeps `doc[id_field]`, the value the reference's check returns (trainingServer.py:191)."""

    def __init__(self, dim: int = 512, capacity: int = 1024, device: int = 0, store: Optional[GalleryStore] = None):
        self.store = store if store is not None else GalleryStore(dim, capacity, device, bf16_plane=False)

    def add(self, doc_ids: Sequence[str], embeddings: np.ndarray, company_ids: Sequence[Optional[str]],
            kinds: Sequence[str], ref_ids: Optional[Sequence] = None):
        """Cursor order = insertion order; an existing document is overwritten in place."""
        meta: List[Dict] = [{"kind": k, "ref_id": (None if ref_ids is None else ref_ids[i])} for i, k in enumerate(kinds)]
        self.store.upsert(list(doc_ids), embeddings, [enrolment_tenant(c, k) for c, k in zip(company_ids, kinds)], meta)

    def remove(self, doc_ids: Sequence[str]) -> int:
        """Only when the DOCUMENT is deleted - never for a status change."""
        return self.store.remove(doc_ids)

    def close(self):
        self.store.close()

SIMILARITY_THRESHOLD = 0.4     # trainingServer.py:70
DUPLICATE_THRESHOLD = 0.4      # trainingServer.py:71


class EnrolmentChecker:
    def __init__(self, store, duplicate_threshold: float = DUPLICATE_THRESHOLD,
                 similarity_threshold: float = SIMILARITY_THRESHOLD, matcher=None):
        """store: a ``GalleryStore``, or a ``ShardedGallery`` with ``matcher=ShardedMatcher(store)`` (the
        duplicate scan then runs on every rank's block and the lowest global row wins; collective)."""
        self.store = store
        self.matcher = matcher if matcher is not None else Matcher(store.store if isinstance(store, EnrolmentGallery) else store)
        self.duplicate_threshold = duplicate_threshold
        self.similarity_threshold = similarity_threshold
        self._scratch: Optional[GalleryStore] = None

    def check_duplicate_face(self, new_embedding: np.ndarray, company_id: Optional[str] = None,
                             kind: Optional[str] = None) -> Tuple[bool, Optional[str]]:
        """(is_duplicate, id of the FIRST stored person with cos > threshold) - trainingServer.py:170-200.
        kind: 'employee' / 'visitor' - scan only that collection's templates of the company (an
        :class:`EnrolmentGallery` store).  The id returned is the stored `ref_id` (`doc[id_field]`,
        trainingServer.py:191) when the gallery carries one, else the row's document id."""
        store = self.store.store if isinstance(self.store, EnrolmentGallery) else self.store
        with _reading(store):            # scan + row -> id translation are one read section (compact() renumbers)
            rows, _ = self.matcher.first_above(np.asarray(new_embedding, np.float32)[None, :], self.duplicate_threshold,
                                               strict=True, company_id=enrolment_tenant(company_id, kind))
            if rows[0] < 0:
                return False, None
            pid = store.id_of(int(rows[0]))
            meta = store.metadata(pid) if pid is not None else None
        ref = (meta or {}).get("ref_id")
        return True, (ref if ref is not None else pid)

    def check_image_similarity(self, embeddings: Sequence[np.ndarray]) -> Tuple[bool, Optional[Tuple[int, int]]]:
        """(all poses show the same person, first offending pair) - trainingServer.py:202-214.  The pose
        embeddings go through the same device path: a scratch gallery of the poses, every pose matched
        against it, pairs read off in the reference's (i, j > i) order."""
        n = len(embeddings)
        if n < 2:
            return True, None
        if self._scratch is None:
            base = self.store.store if isinstance(self.store, EnrolmentGallery) else self.store
            dev = base.device if getattr(base, "device", None) is not None else base.store.device
            self._scratch = GalleryStore(dim=base.dim, capacity=16, device=dev, bf16_plane=False)
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
