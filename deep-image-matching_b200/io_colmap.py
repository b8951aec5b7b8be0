"""COLMAP database writer at the end of the hot path: mirror of the reference's ``io/h5_to_db.py:44-113`` (``export_to_colmap``)
on top of the in-memory feature store and match tables instead of features.h5 / matches.h5 (h5py is absent from this image; the
reference re-reads both files here once more).

Same database layout as the reference's ``utils/database.py`` (COLMAP's own schema): ``cameras`` (model id, width, height, float64
params), ``images``, ``keypoints`` (float32 (N,2) blobs), ``matches`` (raw matches, uint32 (S,2)) and ``two_view_geometries``
(verified matches, config 2, identity F / E / H like the reference's ``add_two_view_geometry`` defaults, or the F estimated by
``geometric_verification``); ``pair_id = id1 * (2**31 - 1) + id2`` with ``id1 < id2`` (columns swapped otherwise).  Cameras: one per
image (``single_camera=False``) or one shared, model ``simple-radial`` with the reference's focal prior ``1.2 * max(w, h)`` when no
EXIF focal length is known (h5_to_db.py:342-384).  Reading EXIF (PIL) is left to the caller: pass ``focal`` per image if known.
"""
from __future__ import annotations

import sqlite3
from pathlib import Path

import numpy as np

MAX_IMAGE_ID = 2 ** 31 - 1
CAMERA_MODELS = {"simple-pinhole": 0, "pinhole": 1, "simple-radial": 2, "opencv": 4}

SCHEMA = """
CREATE TABLE IF NOT EXISTS cameras (camera_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, model INTEGER NOT NULL, width INTEGER NOT NULL,
    height INTEGER NOT NULL, params BLOB, prior_focal_length INTEGER NOT NULL);
CREATE TABLE IF NOT EXISTS images (image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE, camera_id INTEGER NOT NULL,
    prior_qw REAL, prior_qx REAL, prior_qy REAL, prior_qz REAL, prior_tx REAL, prior_ty REAL, prior_tz REAL,
    CONSTRAINT image_id_check CHECK(image_id >= 0 and image_id < 2147483647), FOREIGN KEY(camera_id) REFERENCES cameras(camera_id));
CREATE TABLE IF NOT EXISTS keypoints (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
    FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS descriptors (image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
    FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE);
CREATE TABLE IF NOT EXISTS matches (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);
CREATE TABLE IF NOT EXISTS two_view_geometries (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB,
    config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB, qvec BLOB, tvec BLOB);
CREATE UNIQUE INDEX IF NOT EXISTS index_name ON images(name);
"""


def image_ids_to_pair_id(id1: int, id2: int) -> int:
    if id1 > id2:
        id1, id2 = id2, id1
    return id1 * MAX_IMAGE_ID + id2


def camera_params(model: str, width: int, height: int, focal: float | None = None) -> np.ndarray:
    """create_camera (h5_to_db.py:116-147): default intrinsics of each model from the image size and the focal prior."""
    f = float(focal) if focal else 1.2 * max(width, height)
    if model == "simple-pinhole":
        return np.array([f, width / 2, height / 2])
    if model == "pinhole":
        return np.array([f, f, width / 2, height / 2])
    if model == "simple-radial":
        return np.array([f, width / 2, height / 2, 0.1])
    if model == "opencv":
        return np.array([f, f, width / 2, height / 2, 0.0, 0.0, 0.0, 0.0])
    raise RuntimeError(f"Invalid camera model {model}")


def export_to_colmap(features: dict, matches: dict, database_path="database.db", raw_matches: dict | None = None, camera_model: str = "simple-radial",
                     single_camera: bool = False, focal: dict | None = None, fundamental: dict | None = None) -> dict:
    """features: {image name: FeaturesDict with ``keypoints`` (N,2) and ``image_size`` [H,W]} (e.g. FeatureStore.get_features);
    matches / raw_matches: {(name0, name1): int (S,2)} verified / raw match tables; fundamental: optional {(name0, name1): (3,3)}.
    Returns {image name: image_id}.  An existing database file is deleted first, like the reference does (:84-90)."""
    path = Path(database_path)
    if path.exists():
        path.unlink()
    db = sqlite3.connect(str(path))
    db.executescript(SCHEMA)
    ids, shared_cam = {}, None
    for name in features:  # h5 group order = insertion order
        hw = np.asarray(features[name]["image_size"]).astype(int).ravel()
        height, width = int(hw[0]), int(hw[1])
        if single_camera and shared_cam is not None:
            cam = shared_cam
        else:
            params = camera_params(camera_model, width, height, (focal or {}).get(name))
            cam = db.execute("INSERT INTO cameras VALUES (?, ?, ?, ?, ?, ?)",
                             (None, str(CAMERA_MODELS[camera_model]), width, height, np.asarray(params, np.float64).tobytes(), False)).lastrowid
            shared_cam = cam
        ids[name] = db.execute("INSERT INTO images VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)", (None, name, cam, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)).lastrowid
        kp = np.asarray(features[name]["keypoints"], np.float32)
        if kp.ndim >= 2:
            db.execute("INSERT INTO keypoints VALUES (?, ?, ?, ?)", (ids[name],) + kp.shape + (kp.tobytes(),))

    def rows(table):
        for (n0, n1), m in (table or {}).items():
            i0, i1 = ids[n0], ids[n1]
            m = np.asarray(m).reshape(-1, 2)
            if i0 > i1:
                m = m[:, ::-1]
            yield (n0, n1), image_ids_to_pair_id(i0, i1), np.ascontiguousarray(m, np.uint32)

    seen = set()
    for _, pid, m in rows(raw_matches):
        if pid not in seen:
            seen.add(pid)
            db.execute("INSERT INTO matches VALUES (?, ?, ?, ?)", (pid,) + m.shape + (m.tobytes(),))
    seen = set()
    eye, q, t = np.eye(3).tobytes(), np.array([1.0, 0.0, 0.0, 0.0]).tobytes(), np.zeros(3).tobytes()
    for key, pid, m in rows(matches):
        if pid in seen:
            continue
        seen.add(pid)
        F = (fundamental or {}).get(key)
        Fb = np.asarray(F, np.float64).tobytes() if F is not None else eye
        db.execute("INSERT INTO two_view_geometries VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)", (pid,) + m.shape + (m.tobytes(), 2, Fb, eye, eye, q, t))
    db.commit()
    db.close()
    return ids
