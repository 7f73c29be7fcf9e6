"""Drop-in for the verify paths of deep_insight_face/predictions.py: `TripletPrediction.verify` (:104-150) and
`SiamesePrediction.verify` (:52-89).

The CNN that turns an image into an embedding (`_embedding`, predictions.py:152-156) is out of scope: pass any
callable `emd_model(image) -> [D] or [1, D]` (or hand `verify` the embedding itself).  The identity database is
the reference's python dict; `build_gallery` turns it into a device-resident `Gallery` for 1:N `identify`.
"""
from __future__ import annotations

import numpy as np

from . import _ffi
from .api import face_distance
from .gallery import Gallery
from .networks.siamese import euclidean_distance


class TripletPrediction:
    def __init__(self, emd_model=None, img_size=(96, 96)):
        assert len(img_size) == 2, "Invalid Image size format"   # predictions.py:100
        self.emd_model = emd_model
        self.img_size = img_size
        self._gallery = None
        self._names = []

    def _embedding(self, image) -> np.ndarray:
        if self.emd_model is None:
            return np.asarray(image, dtype=np.float32).reshape(1, -1)
        out = self.emd_model(image)
        return _ffi.host_array(out, np.float32).reshape(1, -1)

    def verify(self, image_path, identity, database, threshold=0.7):
        """predictions.py:104-150: dist = ||encoding - database[identity]||_2, valid iff dist < threshold."""
        encoding = self._embedding(image_path)
        dist = float(face_distance(encoding.reshape(-1), np.asarray(database[identity], dtype=np.float32).reshape(-1)))
        if dist < threshold:
            print("It's " + str(identity))
            is_valid = True
        else:
            print("It's not " + str(identity))
            is_valid = False
        return dist, is_valid

    # ---- 1:N extension (BASELINE.json: "1:N gallery match ... top-k identity lookup") and the persistent
    # identity index of SURVEY 8f row 4: the python dict of predictions.py:112 becomes a device-resident Gallery
    # with incremental enrol / forget; ids are positions in the name table, so a forgotten id is never reused.
    def build_gallery(self, database: dict, metric="l2", precision="tf32x3", capacity=None):
        self._names = list(database.keys())
        rows = np.stack([np.asarray(database[n], dtype=np.float32).reshape(-1) for n in self._names])
        self._gallery = Gallery(max(int(capacity or 0), len(self._names)), rows.shape[1], metric, precision)
        self._gallery.add(rows, np.arange(len(self._names), dtype=np.int64))
        return self._gallery

    def enroll(self, name, encoding) -> int:
        """Add one identity to the live index; returns its id."""
        if self._gallery is None:
            raise RuntimeError("call build_gallery(database) first")
        ident = len(self._names)
        self._gallery.add(np.asarray(encoding, dtype=np.float32).reshape(1, -1), np.array([ident], dtype=np.int64))
        self._names.append(name)
        return ident

    def forget(self, name) -> int:
        """Remove every enrolled row of `name` from the live index; returns the number of rows removed."""
        if self._gallery is None:
            raise RuntimeError("call build_gallery(database) first")
        ids = [i for i, n in enumerate(self._names) if n == name]
        return self._gallery.remove(ids=ids) if ids else 0

    def identify(self, image, k=1):
        """Top-k identities of one image: [(name, score)], best first (squared L2 ascending / cosine descending)."""
        if self._gallery is None:
            raise RuntimeError("call build_gallery(database) first")
        scores, ids = self._gallery.search(self._embedding(image), k)
        return [(self._names[i], float(s)) for s, i in zip(scores[0], ids[0]) if i >= 0]

class SiamesePrediction:
    """predictions.py:46-97.  The reference feeds (encoding, stored encoding) pairs to the siamese model, whose
    head is the `euclidean_distance` Lambda of networks/siamese.py:22-24; here that head runs on the GPU over the
    pair embeddings.  `np.average(pred, axis=-1)[0]` of the reference (:79) averages over the size-1 last axis and
    keeps the FIRST stored encoding's distance (the method is marked "TODO: Fix this method"); that behaviour is
    kept by default, `average_all=True` gives the mean over every stored encoding the comment describes."""

    def __init__(self, emd_model=None, img_size=(112, 112), average_all=False):
        assert len(img_size) == 2, "Invalid Image size format"   # predictions.py:48
        self.emd_model = emd_model
        self.img_size = img_size
        self.average_all = bool(average_all)

    def _embedding(self, image) -> np.ndarray:
        if self.emd_model is None:
            return np.asarray(image, dtype=np.float32).reshape(1, -1)
        return _ffi.host_array(self.emd_model(image), np.float32).reshape(1, -1)

    def verify(self, image_path, identity, database, threshold=0.3):
        encoding = self._embedding(image_path)
        stored = np.asarray(database[identity], dtype=np.float32)
        stored = stored.reshape(-1, encoding.shape[1])
        pred = euclidean_distance([np.repeat(encoding, stored.shape[0], axis=0), stored])   # [n, 1]
        per_pair = np.average(pred, axis=-1)
        dist = float(per_pair.mean() if self.average_all else per_pair[0])
        if dist < threshold:
            print("It's " + str(identity))
            is_valid = True
        else:
            print("It's not " + str(identity))
            is_valid = False
        return dist, is_valid
