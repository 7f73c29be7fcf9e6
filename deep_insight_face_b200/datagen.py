"""Host-side callers either side of the loss: the PK batch sampler of
deep_insight_face/datagen/generator.py:15-41, the pairs.txt expansion and one-hot label construction of
generator.py:43-124 (triplet_image_pairs / facematch_image_pairs / create_pairs), the LFW-style pairs.txt writer of
scripts/generate_pairs.py:60-76 and the pickled verification `.bin` of scripts/raw_img_tf.py:77-86.
Pure index / text logic - nothing here touches the GPU."""
from __future__ import annotations

import os
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


def _stem(img_dir_path: str, name: str, idx) -> str:
    return os.path.join(img_dir_path, name, "%s_%04d" % (name, int(idx)))


def _find(stem: str):
    """The image behind a stem, or None (the reference's add_extension raises RuntimeError here, which its
    `except InvalidPairsError` does not catch - a missing image aborts its whole listing; rows are skipped and
    counted instead, the policy evaluation.utility.get_paths already follows)."""
    for ext in (".jpg", ".png"):
        if os.path.exists(stem + ext):
            return stem + ext
    return None


def facematch_image_pairs(img_dir_path: str, pairs):
    """generator.py:79-109: pairs.txt rows -> [(path0, path1, issame)] plus the class names of the kept rows.  As in
    the reference the names are pair[0] and pair[2] of every kept row - for a 3-field (match) row pair[2] is the
    second image NUMBER, so image numbers appear among the names; kept as it is (pinned in tests/golden).  Rows of
    any other length are dropped silently, as in the reference."""
    out, names, skipped = [], [], 0
    for pair in pairs:
        if len(pair) == 3:
            stems, same = (_stem(img_dir_path, pair[0], pair[1]), _stem(img_dir_path, pair[0], pair[2])), True
        elif len(pair) == 4:
            stems, same = (_stem(img_dir_path, pair[0], pair[1]), _stem(img_dir_path, pair[2], pair[3])), False
        else:
            continue
        found = [_find(s) for s in stems]
        if None in found:
            skipped += 1
            continue
        out.append((found[0], found[1], same))
        for n in (pair[0], pair[2]):
            if n not in names:
                names.append(n)
    if skipped > 0:
        print('Skipped %d image pairs' % skipped)
    return out, names


def triplet_image_pairs(img_dir_path: str, pairs, rng=None):
    """generator.py:43-76: every 4-field row (name1, i, name2, j) becomes (anchor, positive, negative) - anchor and
    negative are the two images of the row, the positive is another image of name1 drawn at random (the first entry
    of a shuffled listing of name1's directory that is not the anchor; hidden files excluded).  3-field rows and
    identities with a single image give no triplet.  The listing is SORTED before the shuffle, so that
    `np.random.seed(s)` (rng=None: numpy's global generator, as in the reference) fixes the triplets on every
    filesystem; the reference shuffles os.listdir's arbitrary order."""
    shuffle = np.random.shuffle if rng is None else rng.shuffle
    out, names, skipped = [], [], 0
    for pair in pairs:
        if len(pair) != 4:
            continue
        anchor, negative = _find(_stem(img_dir_path, pair[0], pair[1])), _find(_stem(img_dir_path, pair[2], pair[3]))
        if anchor is None or negative is None:
            skipped += 1
            continue
        listing = np.array(sorted(f for f in os.listdir(os.path.join(img_dir_path, pair[0])) if not f.startswith('.')))
        shuffle(listing)
        positive = next((os.path.join(img_dir_path, pair[0], str(f)) for f in listing
                         if str(f) != os.path.basename(anchor)), None)
        if positive is None:
            continue
        out.append((anchor, positive, negative))
        for n in (pair[0], pair[2]):
            if n not in names:
                names.append(n)
    if skipped > 0:
        print('Skipped %d image pairs' % skipped)
    return out, names


def create_pairs(img_dir_path: str, func=None, pairs_txt: str = 'pairs.txt'):
    """generator.py:112-124: read pairs.txt, expand it with `func` (facematch_image_pairs / triplet_image_pairs) and
    build the one-hot label matrix: row i is the one-hot of the rank of nb_classes[i] among the sorted names.
    Returns (pairs, nb_classes, one_hot [len(nb_classes), len(nb_classes)] float32).  nb_classes is in first-seen
    order (the reference returns list(set(...)), an arbitrary order; the name -> one-hot row mapping is the same)."""
    assert func is not None, "func should be of type Callable"
    from .evaluation.utility import read_pairs

    pairs, nb_classes = func(img_dir_path, read_pairs(os.path.expanduser(pairs_txt)))
    rank = {name: i for i, name in enumerate(sorted(nb_classes))}
    one_hot = np.zeros((len(nb_classes), len(nb_classes)), dtype=np.float32)
    for i, name in enumerate(nb_classes):
        one_hot[i, rank[name]] = 1.0
    return pairs, nb_classes, one_hot


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
