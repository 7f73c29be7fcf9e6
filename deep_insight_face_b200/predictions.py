"""Drop-in for the verify path of deep_insight_face/predictions.py:104-150 (`TripletPrediction.verify`).

The CNN that turns an image into an embedding (`_embedding`, predictions.py:152-156) is out of scope: pass any
callable `emd_model(image) -> [D] or [1, D]` (or hand `verify` the embedding itself).  The identity database is
the reference's python dict; `build_gallery` turns it into a device-resident `Gallery` for 1:N `identify`.
"""
from __future__ import annotations

import numpy as np

from . import _ffi
from .api import face_distance
from .gallery import Gallery


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

    # ---- 1:N extension (BASELINE.json: "1:N gallery match ... top-k identity lookup")
    def build_gallery(self, database: dict, metric="l2", precision="tf32x3"):
        self._names = list(database.keys())
        rows = np.stack([np.asarray(database[n], dtype=np.float32).reshape(-1) for n in self._names])
        self._gallery = Gallery(len(self._names), rows.shape[1], metric, precision)
        self._gallery.add(rows)
        return self._gallery

    def identify(self, image, k=1):
        """Top-k identities of one image: [(name, score)], best first (squared L2 ascending / cosine descending)."""
        if self._gallery is None:
            raise RuntimeError("call build_gallery(database) first")
        scores, ids = self._gallery.search(self._embedding(image), k)
        return [(self._names[i], float(s)) for s, i in zip(scores[0], ids[0]) if i >= 0]
