"""Drop-in for the distance functions of deep_insight_face/api.py:94-104,242-256 (the image pipeline around
them - landmarks, alignment, thumbnails - is out of scope; the reference module itself does not import)."""
from __future__ import annotations

import numpy as np

from . import _ffi
from .networks.siamese import euclidean_distance
from .networks.utils import distance_to_proba, gaussian_kernel_dist_to_prob


def face_distance(face_encodings, face_to_compare):
    """api.py:94-104: Euclidean distance of one encoding pair (`np.linalg.norm(a - b, axis=0)` - correct for
    1-D inputs, which is how `compare_faces` calls it); empty list -> np.empty((0))."""
    if len(face_encodings) == 0:
        return np.empty((0))
    a = np.atleast_2d(_ffi.host_array(face_encodings, np.float32))
    b = np.atleast_2d(_ffi.host_array(face_to_compare, np.float32))
    d = euclidean_distance([a, np.broadcast_to(b, a.shape).copy()])[:, 0]
    # the kernel clamps sum((a-b)^2) at K.epsilon() (siamese.py:24); undo it for an exact zero
    d = np.where(d * d <= 1.0000001e-7, np.sqrt(np.sum(np.square(a - b), axis=1)), d)
    return d[0] if np.ndim(face_encodings) == 1 else d


def compare_faces(known_face_encodings, face_encoding_to_check, tolerance=0.6):
    """api.py:242-256: distance of the first known encoding to the first candidate and a match score - the
    gaussian-kernel probability inside the tolerance, 1 / (1 + d) outside it."""
    d = face_distance(known_face_encodings[0], face_encoding_to_check[0])
    to_score = gaussian_kernel_dist_to_prob if d <= tolerance else distance_to_proba
    return d, to_score(d)
