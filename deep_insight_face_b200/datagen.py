"""Host-side callers either side of the loss: the PK batch sampler of
deep_insight_face/datagen/generator.py:15-41 and the LFW-style pairs.txt writer of
scripts/generate_pairs.py:60-76 and the pickled verification `.bin` of scripts/raw_img_tf.py:77-86.
Pure index / text logic - nothing here touches the GPU."""
from __future__ import annotations

import io
import pickle
from typing import List, Sequence, Tuple

import numpy as np


def _paths_of(cls) -> Sequence:
    return cls.image_paths if hasattr(cls, "image_paths") else cls


def sample_people(dataset, people_per_batch: int, images_per_person: int, rng=None):
    """generator.py:15-41: shuffle the classes, take up to `images_per_person` shuffled images from each until
    `people_per_batch * images_per_person` images are collected.  Returns (image_paths, num_per_class) as the
    reference does; `rng` (np.random.Generator) replaces the reference's global np.random for reproducibility."""
    rng = np.random.default_rng() if rng is None else rng
    nrof_images = people_per_batch * images_per_person
    class_indices = np.arange(len(dataset))
    rng.shuffle(class_indices)
    i = 0
    image_paths: List = []
    num_per_class: List[int] = []
    while len(image_paths) < nrof_images:
        if i >= len(class_indices):
            raise ValueError("dataset has too few images for the requested batch")
        paths = _paths_of(dataset[class_indices[i]])
        image_indices = np.arange(len(paths))
        rng.shuffle(image_indices)
        n = min(len(paths), images_per_person, nrof_images - len(image_paths))
        image_paths += [paths[j] for j in image_indices[:n]]
        num_per_class.append(n)
        i += 1
    return image_paths, num_per_class


def pk_labels(num_per_class: Sequence[int], one_hot: bool = True) -> np.ndarray:
    """Label layout the losses ingest (common/losses.py:35 argmax's a one-hot [B, C] matrix)."""
    lab = np.repeat(np.arange(len(num_per_class)), num_per_class)
    return np.eye(len(num_per_class), dtype=np.float32)[lab] if one_hot else lab.astype(np.int32)


Match = Tuple[str, int, int]
Mismatch = Tuple[str, int, str, int]


def write_pairs_to_file(fname: str, match_folds: List[List[Match]], mismatch_folds: List[List[Mismatch]],
                        num_folds: int, num_matches_mismatches: int) -> None:
    """scripts/generate_pairs.py:60-76: header `folds\\tN`, then per fold `name\\ti\\tj` matches followed by
    `name1\\ti\\tname2\\tj` mismatches."""
    with io.open(fname, "w", io.DEFAULT_BUFFER_SIZE, encoding="utf-8") as f:
        f.write("{}\t{}\n".format(num_folds, num_matches_mismatches))
        for match_fold, mismatch_fold in zip(match_folds, mismatch_folds):
            for m in match_fold:
                f.write("{}\t{}\t{}\n".format(m[0], m[1], m[2]))
            for mm in mismatch_fold:
                f.write("{}\t{}\t{}\t{}\n".format(mm[0], mm[1], mm[2], mm[3]))
        f.flush()


def pairs_issame(pairs) -> np.ndarray:
    """issame flag of each pairs.txt row (3 fields = same person, 4 = different; evaluation/utility.py:228-236)."""
    return np.array([len(p) == 3 for p in pairs], dtype=bool)


def write_test_bin(fname: str, encoded_images: Sequence[bytes], issame_list: Sequence[bool]) -> None:
    """scripts/raw_img_tf.py:77-86: the verification `.bin` is `pickle.dump([encoded_jpegs, issame_list])` with
    two images per pair in pairs.txt order (get_paths order).  Encoding the JPEGs is the caller's business."""
    if len(encoded_images) != 2 * len(issame_list):
        raise ValueError("%d images for %d pairs: expected two per pair" % (len(encoded_images), len(issame_list)))
    with open(fname, "wb") as f:
        pickle.dump([list(encoded_images), [bool(s) for s in issame_list]], f)


def read_test_bin(fname: str):
    """Inverse of write_test_bin -> (encoded_images, issame bool array).  The embeddings of image 2i and 2i+1 are
    what evaluation.utility.evaluate expects at rows 2i and 2i+1 (evaluation/utility.py:18-19)."""
    with open(fname, "rb") as f:
        images, issame = pickle.load(f, encoding="bytes")
    if len(images) != 2 * len(issame):
        raise ValueError("%s: %d images for %d pairs" % (fname, len(images), len(issame)))
    return list(images), np.asarray(issame, dtype=bool)
