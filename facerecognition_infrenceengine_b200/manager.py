"""``EmbeddingManager`` - same name, methods and sync semantics as the reference's two managers
(infrenceServer.py:36-398 "live", peopleCount.py:695-819 "campus"), over the GPU-resident store.

What changes: the per-frame cost.  ``get_embeddings_for_company`` no longer runs two Mongo
queries and rebuilds a dict per frame (infrenceServer.py:343-380) and ``get_all`` no longer copies
the dict (peopleCount.py:816-819); both return a :class:`GalleryView` - (store, tenant) - that the
matcher turns into a per-row tag comparison inside the kernel.

Where records come from is pluggable (``source``): the reference reads Mongo + GridFS; this
environment has neither, so a source is any object with the three methods of :class:`ListSource`.
Records are plain dicts shaped like the reference's documents, with the unpickled vector inline
under ``'embedding'``.
"""
from __future__ import annotations

import logging
import threading
import time
from datetime import datetime, timezone
from typing import Dict, Iterable, List, Optional

import numpy as np

from .gallery import GalleryStore

logger = logging.getLogger(__name__)


# ---- eligibility filters: the Mongo queries of the loaders, restated on dict records ----------
def _utcnow() -> datetime:
    """Naive UTC timestamp, what the reference's `datetime.utcnow()` returns (infrenceServer.py:112,226)."""
    return datetime.now(timezone.utc).replace(tzinfo=None)


def eligible_employee(doc: Dict) -> bool:
    """infrenceServer.py:95-99 == peopleCount.py:738-742."""
    return (doc.get("status") == "active" and doc.get("blacklisted") is False
            and doc.get("embedding_status", "done") == "done" and doc.get("embedding") is not None)


def eligible_visitor(doc: Dict) -> bool:
    """infrenceServer.py:126 == peopleCount.py:747."""
    return doc.get("embedding_status", "done") == "done" and doc.get("embedding") is not None


def inactive_employee(doc: Dict) -> bool:
    """infrenceServer.py:238-243: status != 'active' OR blacklisted."""
    return doc.get("status") != "active" or doc.get("blacklisted") is True


class ListSource:
    """In-memory record source (tests, examples).  A Mongo-backed source implements the same three
    methods with the reference's queries."""

    def __init__(self, employees: Optional[List[Dict]] = None, visitors: Optional[List[Dict]] = None):
        self.employees = employees or []
        self.visitors = visitors or []

    def employee_docs(self, since: Optional[datetime] = None) -> List[Dict]:
        return [d for d in self.employees if eligible_employee(d) and (since is None or d.get("lastUpdated") >= since)]

    def visitor_docs(self, since: Optional[datetime] = None) -> List[Dict]:
        return [d for d in self.visitors if eligible_visitor(d) and (since is None or d.get("lastUpdated") >= since)]

    def inactive_employee_ids(self) -> List[str]:
        return [str(d["_id"]) for d in self.employees if inactive_employee(d)]


class BroadcastSource:
    """Document source for a row-sharded gallery (one process per GPU): rank 0 asks the real source, every
    rank receives the same documents, so the replicated id / tenant tables of ``ShardedGallery`` stay identical
    and every rank applies the same upserts and removals in the same order.  Pass a CPU (gloo) process group
    for this control traffic - ``dist.new_group(backend="gloo")`` - so that it never interleaves with the
    collectives of the data path when the sync thread runs beside the matcher."""

    def __init__(self, source, group=None):
        import torch.distributed as dist
        self.source = source
        self.group = group
        self.rank = dist.get_rank(group)

    def _bcast(self, fetch):
        import torch.distributed as dist
        box = [fetch() if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                   group=self.group)
        return box[0]

    def employee_docs(self, since: Optional[datetime] = None) -> List[Dict]:
        return self._bcast(lambda: list(self.source.employee_docs(since)))

    def visitor_docs(self, since: Optional[datetime] = None) -> List[Dict]:
        return self._bcast(lambda: list(self.source.visitor_docs(since)))

    def inactive_employee_ids(self) -> List[str]:
        return self._bcast(lambda: list(self.source.inactive_employee_ids()))


class GalleryView:
    """What the matcher needs instead of a dict copy: the store and an optional tenant filter."""

    def __init__(self, store: GalleryStore, company_id: Optional[str] = None):
        self.store = store
        self.company_id = company_id

    def __len__(self):
        if self.company_id is None:
            return len(self.store)
        if hasattr(self.store, "count_tenant"):          # sharded gallery: host-side table, no device read
            return self.store.count_tenant(self.company_id)
        return self.store.count_tenant(self.company_id)     # host-side table: no device read

    def __bool__(self):
        return len(self.store) > 0


class EmbeddingManager:
    """mode='live'  : incremental sync by lastUpdated every 30 s + eviction (infrenceServer.py:175-258)
    mode='campus': full reload every 60 s, never evicts (peopleCount.py:766-776)."""

    def __init__(self, source, dim: int = 512, device: int = 0, mode: str = "live",
                 sync_interval: Optional[float] = None, capacity: int = 1024, bf16_plane: bool = True,
                 store=None):
        """store: an existing ``GalleryStore``, or a ``sharded.ShardedGallery`` (one process per GPU; give every
        rank the same documents, e.g. through ``BroadcastSource``).  Default: a new store on `device`."""
        assert mode in ("live", "campus")
        self.source = source
        self.mode = mode
        self.store = store if store is not None else GalleryStore(dim, capacity, device, bf16_plane)
        self.embeddings_lock = threading.Lock()
        self.last_sync_time: Optional[datetime] = None
        self.is_initial_load = True
        self.sync_interval = sync_interval if sync_interval is not None else (30 if mode == "live" else 60)
        self.running = False
        self.sync_thread: Optional[threading.Thread] = None
        self._evicted_since_compaction = 0
        self._compact_not_before = 0.0       # monotonic time before which a failed compaction is not retried
        self.rejected_records = 0            # malformed documents skipped so far (logged, as the reference does)
        self._initial_load()

    # ---- loading ----------------------------------------------------------------------------
    def _initial_load(self):
        """infrenceServer.py:62-91 / peopleCount.py:716-734 - errors are logged, not raised (:89-91)."""
        try:
            self._load_updated_embeddings(self.source.employee_docs(), self.source.visitor_docs())
            self.last_sync_time = _utcnow()
            self.is_initial_load = False
        except Exception as ex:                      # noqa: BLE001
            logger.error("Error in initial load: %s", ex)

    def _load_updated_embeddings(self, employees: Iterable[Dict], visitors: Iterable[Dict]):
        """infrenceServer.py:260-341 / peopleCount.py:778-814: employees first, then visitors; each
        vector is divided by its norm on ingest (done on the device)."""
        ids, vecs, comps, meta = [], [], [], []
        dim = getattr(self.store, "dim", None)

        def vector_of(doc):
            # the reference wraps every person in its own try/except and skips the bad ones
            # (infrenceServer.py:264-341): one malformed document must not fail the batch, let alone every
            # later sync.  What it would have stored is a float vector of the model's size.
            v = np.asarray(doc["embedding"], dtype=np.float32)
            if v.ndim != 1 or (dim is not None and v.shape[0] != dim):
                raise ValueError("embedding of shape %s, expected (%s,)" % (v.shape, dim))
            return v

        for e in employees:
            try:
                v = vector_of(e)
                pid = str(e["_id"])
            except Exception as ex:                  # noqa: BLE001 - skip and log, as the reference does
                self.rejected_records += 1
                logger.warning("Error loading embedding for employee %s: %s", e.get("_id") if isinstance(e, dict) else e, ex)
                continue
            ids.append(pid)
            vecs.append(v)
            comps.append(None if e.get("companyId") is None else str(e["companyId"]))
            meta.append({"name": e.get("employeeName", "Unknown"), "employeeId": e.get("employeeId", "Unknown"),
                         "email": e.get("employeeEmail", ""), "mobile": e.get("employeeMobile", ""),
                         "type": "employee", "lastUpdated": e.get("lastUpdated")})
        for d in visitors:
            try:
                v = vector_of(d)
                pid = str(d["_id"])
            except Exception as ex:                  # noqa: BLE001
                self.rejected_records += 1
                logger.warning("Error loading embedding for visitor %s: %s", d.get("_id") if isinstance(d, dict) else d, ex)
                continue
            ids.append(pid)
            vecs.append(v)
            comps.append(None if d.get("companyId") is None else str(d["companyId"]))
            meta.append({"name": d.get("visitorName", "Unknown"), "type": "visitor",
                         "lastUpdated": d.get("lastUpdated")})
        if ids:
            with self.embeddings_lock:
                self.store.upsert(ids, np.stack(vecs), comps, meta)

    def _remove_inactive_employees(self):
        """infrenceServer.py:234-258."""
        with self.embeddings_lock:
            gone = self.store.remove(self.source.inactive_employee_ids())
            self._evicted_since_compaction += gone
            self._maybe_compact()
            return gone

    # the reference's `del self.embeddings[id]` frees the entry; here an evicted row stays behind as a tombstone
    # (and keeps its company's row window wide) until the store is compacted.  Do that once tombstones are
    # both many and a sizeable share of the gallery - never for the odd eviction.
    COMPACT_MIN_DEAD = 1024
    COMPACT_DEAD_SHARE = 0.25
    COMPACT_RETRY_S = 600

    def _maybe_compact(self):
        if self._evicted_since_compaction < self.COMPACT_MIN_DEAD or not hasattr(self.store, "compact"):
            return False
        st = self.store.stats()
        dead = int(st.rows) - int(st.live)
        if dead >= self.COMPACT_MIN_DEAD and dead >= self.COMPACT_DEAD_SHARE * int(st.rows):
            if time.monotonic() < self._compact_not_before:
                return False
            try:
                self.store.compact()
            except Exception as ex:                  # noqa: BLE001 - e.g. no device memory for the bounce chunk
                # compaction is an optimisation: the gallery is intact (frg_store_compact changes nothing
                # unless it succeeds), the sync goes on, and it is not retried every cycle
                self._compact_not_before = time.monotonic() + self.COMPACT_RETRY_S
                logger.warning("gallery compaction failed, next attempt in %d s: %s", self.COMPACT_RETRY_S, ex)
                return False
            self._evicted_since_compaction = 0
            return True
        return False

    def _sync_embeddings(self):
        if self.mode == "live":
            if self.last_sync_time is None:
                return
            since = self.last_sync_time
            employees = self.source.employee_docs(since)
            visitors = self.source.visitor_docs(since)
            self._remove_inactive_employees()
            if employees or visitors:
                self._load_updated_embeddings(employees, visitors)
            self.last_sync_time = _utcnow()
        else:
            self._load_updated_embeddings(self.source.employee_docs(), self.source.visitor_docs())
            self.last_sync_time = _utcnow()

    def force_sync(self):
        """infrenceServer.py:382-384."""
        self._sync_embeddings()

    def start_sync(self):
        """infrenceServer.py:158-166 / peopleCount.py:750-758."""
        if self.running:
            return
        self.running = True
        self.sync_thread = threading.Thread(target=self._sync_loop, daemon=True)
        self.sync_thread.start()

    def stop_sync(self):
        self.running = False
        if self.sync_thread:
            self.sync_thread.join(timeout=5)

    def _sync_loop(self):
        while self.running:
            try:
                if self.mode == "campus":
                    self._sleep(self.sync_interval)
                    if not self.running:
                        break
                    self._sync_embeddings()
                else:
                    self._sync_embeddings()
                    self._sleep(self.sync_interval)
            except Exception:                       # log-and-retry, infrenceServer.py:181-183
                self._sleep(5)

    def _sleep(self, seconds: float):
        end = time.time() + seconds
        while self.running and time.time() < end:
            time.sleep(min(0.05, max(0.0, end - time.time())))

    # ---- what the matchers ask for -------------------------------------------------------------
    def get_embeddings_for_company(self, company_id: str) -> GalleryView:
        """infrenceServer.py:343-380 -> a tenant filter, not a dict rebuild."""
        return GalleryView(self.store, str(company_id))

    def get_all(self) -> GalleryView:
        """peopleCount.py:816-819 -> a handle, not a copy."""
        return GalleryView(self.store, None)

    def get_stats(self) -> Dict:
        """infrenceServer.py:386-398 (same keys) + device-side figures."""
        with self.embeddings_lock:
            metas = [self.store.metadata(i) or {} for i in self.store.ids()]
            st = self.store.stats()
            return {
                "total_embeddings": len(metas),
                "employees": sum(1 for m in metas if m.get("type") == "employee"),
                "visitors": sum(1 for m in metas if m.get("type") == "visitor"),
                "last_sync": self.last_sync_time.isoformat() if self.last_sync_time else None,
                "initial_load_complete": not self.is_initial_load,
                "device_rows": int(st.rows), "device_live_rows": int(st.live),
                "device_capacity": int(st.capacity), "device_bytes": int(st.bytes), "version": int(st.version),
            }
