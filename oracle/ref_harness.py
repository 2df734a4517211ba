"""Runs the reference's OWN matching / gallery code, verbatim, under module stubs.
TEST INFRASTRUCTURE - build container only (needs /root/reference, which does not exist on
the GPU box; nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this module).

The reference cannot be imported as shipped: it needs insightface, pymongo, gridfs, bson,
flask and flask_cors, and ``infrenceServer.py:682`` opens a Mongo connection at import time
(SURVEY.md section 8c).  Here those modules are replaced in ``sys.modules`` by small in-memory
fakes, after which

  * ``peopleCount.CameraProcessor.process_frame``        (peopleCount.py:843-896)
  * ``infrenceServer.FaceRecognitionProcessor.recognize_faces`` (infrenceServer.py:515-563)
  * both ``EmbeddingManager`` loaders / sync / eviction / tenant subset
    (infrenceServer.py:62-91,185-258,260-380; peopleCount.py:720-819)
  * ``CampusPeopleManager.process_unknown_detection``    (peopleCount.py:432-500)

execute unmodified.  The fakes only supply data (documents, pickled vectors, detected
"faces") and capture results; no arithmetic is done here.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import logging
import os
import pickle
import sys
import tempfile
import types
from datetime import datetime
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("FRG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "peopleCount.py"))


# ----------------------------------------------------------------------------- fakes
class ObjectId:
    """bson.ObjectId stand-in: an opaque, hashable 24-hex id."""

    def __init__(self, oid=None):
        if isinstance(oid, ObjectId):
            oid = oid._s
        if oid is None:
            ObjectId._n = getattr(ObjectId, "_n", 0) + 1
            oid = "%024x" % ObjectId._n
        self._s = str(oid)

    def __str__(self):
        return self._s

    __repr__ = __str__

    def __eq__(self, o):
        return isinstance(o, ObjectId) and o._s == self._s

    def __hash__(self):
        return hash(self._s)


def _get_path(doc: Dict, dotted: str):
    cur: Any = doc
    for part in dotted.split("."):
        if not isinstance(cur, dict) or part not in cur:
            return _MISSING
        cur = cur[part]
    return cur


_MISSING = object()


def _match(doc: Dict, query: Dict) -> bool:
    """The handful of Mongo operators the reference's queries use."""
    for key, cond in query.items():
        if key == "$or":
            if not any(_match(doc, q) for q in cond):
                return False
            continue
        val = _get_path(doc, key)
        if isinstance(cond, dict) and any(k.startswith("$") for k in cond):
            for op, arg in cond.items():
                if op == "$exists":
                    if (val is not _MISSING) != bool(arg):
                        return False
                elif op == "$ne":
                    if val is not _MISSING and val == arg:
                        return False
                elif op == "$gte":
                    if val is _MISSING or not (val >= arg):
                        return False
                elif op == "$gt":
                    if val is _MISSING or not (val > arg):
                        return False
                elif op == "$lt":
                    if val is _MISSING or not (val < arg):
                        return False
                elif op == "$in":
                    if val is _MISSING or val not in arg:
                        return False
                else:
                    raise NotImplementedError(op)
        else:
            if val is _MISSING or val != cond:
                return False
    return True


class FakeCollection:
    def __init__(self):
        self.docs: List[Dict] = []

    def find(self, query=None, projection=None):
        return [d for d in self.docs if _match(d, query or {})]   # cursor order = insertion order

    def find_one(self, query=None):
        r = self.find(query)
        return r[0] if r else None

    def count_documents(self, query=None):
        return len(self.find(query))

    # write paths used by background threads of CampusPeopleManager: accepted and dropped
    def create_index(self, *a, **k):
        return None

    def bulk_write(self, *a, **k):
        return types.SimpleNamespace(upserted_count=0, modified_count=0)

    def insert_many(self, *a, **k):
        return None

    def insert_one(self, *a, **k):
        return None

    def update_one(self, *a, **k):
        return None


class FakeDB(dict):
    def __missing__(self, name):
        self[name] = FakeCollection()
        return self[name]


class FakeMongoClient:
    _dbs: Dict[str, FakeDB] = {}

    def __init__(self, *a, **k):
        pass

    def __getitem__(self, name):
        return FakeMongoClient._dbs.setdefault(name, FakeDB())

    def close(self):
        pass


class _GridOut:
    def __init__(self, blob: bytes):
        self._b = blob

    def read(self):
        return self._b


class FakeGridFS:
    _buckets: Dict[str, Dict[Any, bytes]] = {}

    def __init__(self, db, collection="fs"):
        self.files = FakeGridFS._buckets.setdefault(collection, {})

    def get(self, file_id):
        return _GridOut(self.files[file_id])

    def put(self, blob: bytes, **k):
        fid = ObjectId()
        self.files[fid] = blob
        return fid


class FakeFace:
    def __init__(self, emb: np.ndarray, i: int = 0):
        self.normed_embedding = emb
        self.bbox = np.array([10.0 + i, 20.0, 110.0 + i, 140.0], dtype=np.float32)
        self.det_score = 0.99


class FakeDetector:
    """Stands in for insightface FaceAnalysis: ``get(frame)`` hands back pre-made faces."""

    def __init__(self, *a, **k):
        self.faces: List[FakeFace] = []

    def prepare(self, *a, **k):
        pass

    def get(self, frame):
        return list(self.faces)


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("insightface")
    mod("insightface.app", FaceAnalysis=FakeDetector)
    mod("pymongo", MongoClient=FakeMongoClient, UpdateOne=lambda *a, **k: ("UpdateOne", a, k))
    mod("pymongo.collection", Collection=FakeCollection)
    mod("pymongo.errors", ConnectionFailure=type("ConnectionFailure", (Exception,), {}),
        OperationFailure=type("OperationFailure", (Exception,), {}))
    mod("gridfs", GridFS=FakeGridFS)
    mod("bson", ObjectId=ObjectId)

    class _Flask:
        def __init__(self, *a, **k):
            pass

        def route(self, *a, **k):
            return lambda fn: fn

        def run(self, *a, **k):
            pass

    mod("flask", Flask=_Flask, request=types.SimpleNamespace(json=None, args={}),
        jsonify=lambda *a, **k: (a, k))
    mod("flask_cors", CORS=lambda *a, **k: None)


_loaded: Dict[str, types.ModuleType] = {}


def load(module_name: str) -> types.ModuleType:
    """Import ``infrenceServer``, ``peopleCount`` or ``trainingServer`` from the reference tree
    under the stubs."""
    if module_name in _loaded:
        return _loaded[module_name]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    os.environ["MONGODB_URI"] = "stub://in-memory"      # never let the shipped default URI be used
    scratch = tempfile.mkdtemp(prefix="frg_ref_")       # the modules create *.log files in cwd
    cwd = os.getcwd()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.chdir(scratch)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            m = importlib.import_module(module_name)
    finally:
        os.chdir(cwd)
    # silence the reference's per-row logging (it logs every loaded embedding at INFO)
    logging.getLogger().setLevel(logging.CRITICAL)
    for h in list(logging.getLogger().handlers):
        logging.getLogger().removeHandler(h)
    logging.getLogger(m.__name__).setLevel(logging.CRITICAL)
    _loaded[module_name] = m
    return m


def reset_database():
    FakeMongoClient._dbs.clear()
    FakeGridFS._buckets.clear()


# ----------------------------------------------------------------------------- data seeding
def seed_people(employees: Iterable[Dict], visitors: Iterable[Dict], db_name: str = "factorylyticsDB"):
    """Each dict: {'id': 24-hex str, 'vec': ndarray (raw, as the enrol worker pickles it),
    'company': 24-hex str, optional 'status', 'blacklisted', 'emb_status', 'lastUpdated', 'name'}."""
    db = FakeMongoClient()[db_name]
    efs = FakeGridFS(db, collection="employee_embeddings")
    vfs = FakeGridFS(db, collection="visitor_embeddings")
    for e in employees:
        fid = efs.put(pickle.dumps(e["vec"]))
        db["employeeInfo"].docs.append({
            "_id": ObjectId(e["id"]), "companyId": ObjectId(e.get("company", "c" * 24)),
            "status": e.get("status", "active"), "blacklisted": e.get("blacklisted", False),
            "employeeName": e.get("name", "emp"), "employeeId": e.get("id"),
            "lastUpdated": e.get("lastUpdated", datetime.utcnow()),
            "employeeEmbeddings": {"buffalo_l": {"embeddingId": fid, "status": e.get("emb_status", "done")}},
        })
    for v in visitors:
        fid = vfs.put(pickle.dumps(v["vec"]))
        db["visitors"].docs.append({
            "_id": ObjectId(v["id"]), "companyId": ObjectId(v.get("company", "c" * 24)),
            "visitorName": v.get("name", "vis"),
            "lastUpdated": v.get("lastUpdated", datetime.utcnow()),
            # the visitor loader wraps embeddingId in ObjectId(...) (infrenceServer.py:313)
            "visitorEmbeddings": {"buffalo_l": {"embeddingId": fid, "status": v.get("emb_status", "done")}},
        })
    return db


# ----------------------------------------------------------------------------- drivers
class _CapturingManager:
    """Stands in for CampusPeopleManager as the consumer of match results."""

    def __init__(self):
        self.events: List[Dict] = []

    def process_detection(self, person_id, person_info, camera_id, timestamp, score):
        self.events.append({"kind": "recognized", "id": person_id, "score": score})

    def process_unknown_detection(self, camera_id, timestamp, face_embedding, bbox):
        self.events.append({"kind": "unknown", "embedding": np.array(face_embedding, copy=True)})


class _DictManager:
    """Stands in for EmbeddingManager where only the dict snapshot matters."""

    def __init__(self, embeddings: Dict[str, np.ndarray], metadata: Dict[str, Dict]):
        self.embeddings, self.metadata = embeddings, metadata

    def get_all(self):
        return self.embeddings.copy(), self.metadata.copy()

    def get_embeddings_for_company(self, company_id):
        return self.embeddings, self.metadata


def run_campus_frame(ids: List[str], gallery: np.ndarray, faces: np.ndarray):
    """peopleCount.CameraProcessor.process_frame (peopleCount.py:843-896), verbatim.

    ``gallery`` rows are the ALREADY-LOADED unit vectors (what EmbeddingManager holds);
    ``faces`` are raw ``normed_embedding`` values.  Returns per face
    (kind 'recognized'|'unknown'|'ignored', id or None, score or None) + the stats dict.
    """
    pc = load("peopleCount")
    em = _DictManager({i: g for i, g in zip(ids, gallery)}, {i: {"name": i, "type": "employee"} for i in ids})
    out: List[Dict] = []
    stats_total = {"faces": 0, "recognized": 0, "unknown": 0}
    # one face per frame so that captured events map 1:1 to faces (ignored faces emit no event)
    for fi, e in enumerate(faces):
        mgr = _CapturingManager()
        proc = pc.CameraProcessor(em, mgr)
        proc.face_detector = FakeDetector()
        proc.face_detector.faces = [FakeFace(e, fi)]
        stats = proc.process_frame(np.zeros((4, 4, 3), np.uint8), "cam0")
        for k in stats_total:
            stats_total[k] += stats[k]
        if mgr.events:
            ev = mgr.events[0]
            out.append({"kind": ev["kind"], "id": ev.get("id"), "score": ev.get("score")})
        else:
            out.append({"kind": "ignored", "id": None, "score": None})
    return out, stats_total


def run_campus_frame_batch(ids: List[str], gallery: np.ndarray, faces: np.ndarray) -> Dict:
    """Same, all faces in ONE frame (the shape the CPU baseline times)."""
    pc = load("peopleCount")
    em = _DictManager({i: g for i, g in zip(ids, gallery)}, {i: {"name": i, "type": "employee"} for i in ids})
    mgr = _CapturingManager()
    proc = pc.CameraProcessor(em, mgr)
    proc.face_detector = FakeDetector()
    proc.face_detector.faces = [FakeFace(e, i) for i, e in enumerate(faces)]
    return proc.process_frame(np.zeros((4, 4, 3), np.uint8), "cam0")


def run_live_frame(ids: List[str], gallery: np.ndarray, faces: np.ndarray):
    """infrenceServer.FaceRecognitionProcessor.recognize_faces (infrenceServer.py:515-563),
    verbatim; results captured at the draw call (:555).  Returns per face
    (name 'Unknown' or the id, type, reported score)."""
    srv = load("infrenceServer")
    em = _DictManager({i: g for i, g in zip(ids, gallery)}, {i: {"name": i, "type": "employee"} for i in ids})
    proc = srv.FaceRecognitionProcessor(em)
    proc.face_detector = FakeDetector()
    proc.face_detector.faces = [FakeFace(e, i) for i, e in enumerate(faces)]
    captured: List[Dict] = []

    def capture(frame, bbox, color, person_info, det_score, recognition_score):
        captured.append({"name": person_info["name"], "type": person_info["type"],
                         "score": recognition_score})
        return frame

    proc.draw_enhanced_bounding_box = capture
    proc.recognize_faces(np.zeros((4, 4, 3), np.uint8), "c" * 24)
    return captured


def run_unknown_clustering(embeddings: np.ndarray):
    """peopleCount.CampusPeopleManager.process_unknown_detection (:432-500) verbatim on a
    stream of unit embeddings.  Returns per observation (cluster ordinal, created?)."""
    pc = load("peopleCount")
    mgr = pc.CampusPeopleManager("stub://", "db_unknown")
    mgr.running = False
    mgr.camera_configs["cam0"] = {"campus_id": "campus0", "type": pc.CameraType.ENTRY, "name": "cam0"}
    out = []
    for e in embeddings:
        before = len(mgr.unknown_people["campus0"])
        counts = {k: u.detection_count for k, u in mgr.unknown_people["campus0"].items()}
        mgr.process_unknown_detection("cam0", datetime.utcnow(), e, [0, 0, 1, 1])
        after = mgr.unknown_people["campus0"]
        if len(after) > before:
            out.append((len(after) - 1, True))
        else:
            hit = [i for i, (k, u) in enumerate(after.items()) if u.detection_count != counts[k]]
            out.append((hit[0], False))
    avgs = [np.array(u.avg_embedding, copy=True) for u in mgr.unknown_people["campus0"].values()]
    return out, avgs


def live_manager():
    """A fresh infrenceServer.EmbeddingManager over the current fake database."""
    srv = load("infrenceServer")
    return srv.EmbeddingManager("stub://", "factorylyticsDB")


def campus_manager():
    pc = load("peopleCount")
    return pc.EmbeddingManager("stub://", "factorylyticsDB")


def run_enrol_checks(new_embedding: np.ndarray, existing_raw: List[np.ndarray], poses: List[np.ndarray]):
    """trainingServer.FaceEmbeddingWorker._check_duplicate_face (:170-200) and
    ._check_image_similarity (:202-214), verbatim, called unbound on a bare namespace that
    carries only ``config`` (constructing the worker would start the detector and signal
    handlers).  Returns (duplicate position or -1, same_person ok, offending pair)."""
    ts = load("trainingServer")
    company = ObjectId("c" * 24)
    coll = FakeCollection()
    fs = ts.employee_embedding_fs
    for j, v in enumerate(existing_raw):
        fid = fs.put(pickle.dumps(v))
        coll.docs.append({"_id": ObjectId(), "companyId": company, "employee": j,
                          "employeeEmbeddings": {"buffalo_l": {"embeddingId": fid}}})
    me = types.SimpleNamespace(config=ts.WorkerConfig())
    is_dup, dup = ts.FaceEmbeddingWorker._check_duplicate_face(me, new_embedding, company, coll, "employee")
    ok, pair = ts.FaceEmbeddingWorker._check_image_similarity(me, poses)
    return (int(dup) if is_dup else -1), bool(ok), pair
