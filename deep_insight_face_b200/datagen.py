"""Host-side callers either side of the loss: the PK batch sampler of
deep_insight_face/datagen/generator.py:15-41 and the LFW-style pairs.txt writer of
scripts/generate_pairs.py:60-76 and the pickled verification `.bin` of scripts/raw_img_tf.py:77-86.
Pure index / text logic - nothing here touches the GPU."""
from __future__ import annotations

import pickle
from typing import List, Sequence, Tuple

import numpy as np


def _paths_of(cls) -> Sequence:
    return cls.image_paths if hasattr(cls, "image_paths") else cls


def sample_people(dataset, people_per_batch: int, images_per_person: int, rng=None):
    """PK sampler with the contract of generator.py:15-41: classes are visited in one random order; every visited
    class contributes min(its size, images_per_person, images still missing) randomly chosen images, until
    `people_per_batch * images_per_person` images are drawn.  Returns (image_paths, num_per_class).

    The draws are one shuffle of the class order, then one shuffle of the member order per visited class.  With
    `rng=None` they come from numpy's global generator, as in the reference - `np.random.seed(s)` therefore
    reproduces the reference's batches exactly (tests/golden/host_reference.json holds batches produced by the
    reference's own function); pass an `np.random.Generator` for a private stream."""
    shuffle = np.random.shuffle if rng is None else rng.shuffle
    missing = people_per_batch * images_per_person
    order = np.arange(len(dataset))
    shuffle(order)
    picked: List = []
    per_class: List[int] = []
    for cls in order:
        if missing == 0:
            break
        members = _paths_of(dataset[cls])
        perm = np.arange(len(members))
        shuffle(perm)
        take = min(len(members), images_per_person, missing)
        picked.extend(members[j] for j in perm[:take])
        per_class.append(take)
        missing -= take
    if missing:
        raise ValueError("dataset has too few images for the requested batch")
    return picked, per_class


def pk_labels(num_per_class: Sequence[int], one_hot: bool = True) -> np.ndarray:
    """Label layout the losses ingest (common/losses.py:35 argmax's a one-hot [B, C] matrix)."""
    lab = np.repeat(np.arange(len(num_per_class)), num_per_class)
    return np.eye(len(num_per_class), dtype=np.float32)[lab] if one_hot else lab.astype(np.int32)


Match = Tuple[str, int, int]
Mismatch = Tuple[str, int, str, int]


def write_pairs_to_file(fname: str, match_folds: List[List[Match]], mismatch_folds: List[List[Mismatch]],
                        num_folds: int, num_matches_mismatches: int) -> None:
    """The pairs.txt layout of scripts/generate_pairs.py:60-76 (byte-identical, see tests/golden): a header
    `folds<TAB>N`, then per fold its matches `name<TAB>i<TAB>j` followed by its mismatches
    `name1<TAB>i<TAB>name2<TAB>j`, one record per line."""
    records = [(num_folds, num_matches_mismatches)]
    for same, different in zip(match_folds, mismatch_folds):
        records.extend(tuple(m[:3]) for m in same)
        records.extend(tuple(mm[:4]) for mm in different)
    text = "".join("\t".join(str(field) for field in rec) + "\n" for rec in records)
    with open(fname, "w", encoding="utf-8") as f:
        f.write(text)


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
