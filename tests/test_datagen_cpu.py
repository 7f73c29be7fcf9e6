"""Host-side helpers around the hot path: PK sampler, label layout, pairs.txt round trip."""
import numpy as np
import pytest

from deep_insight_face_b200 import datagen


def test_sample_people_gives_p_times_k():
    dataset = [[f"p{c}_{i}.jpg" for i in range(2 + c % 5)] for c in range(40)]
    paths, per_class = datagen.sample_people(dataset, 18, 4, rng=np.random.default_rng(0))
    assert len(paths) == 72 and sum(per_class) == 72 and max(per_class) <= 4
    assert len(set(paths)) == 72
    onehot = datagen.pk_labels(per_class)
    assert onehot.shape == (72, len(per_class)) and (onehot.sum(1) == 1).all()
    assert (np.argmax(onehot, 1) == datagen.pk_labels(per_class, one_hot=False)).all()


def test_pairs_file_round_trip(tmp_path):
    matches = [[("Ann", 1, 2), ("Bob", 3, 1)], [("Cy", 1, 4)]]
    mism = [[("Ann", 1, "Bob", 2)], [("Cy", 2, "Ann", 3), ("Bob", 1, "Cy", 1)]]
    f = tmp_path / "pairs.txt"
    datagen.write_pairs_to_file(str(f), matches, mism, 2, 2)
    lines = f.read_text().splitlines()
    assert lines[0] == "2\t2" and lines[1] == "Ann\t1\t2" and lines[3] == "Ann\t1\tBob\t2"
    # the reader half lives with the evaluators (reference: evaluation/utility.py:256-262); no GPU needed for it
    pairs = [ln.split("\t") for ln in lines[1:]]
    assert datagen.pairs_issame(pairs).tolist() == [True, True, False, True, False, False]


def test_get_paths_and_test_bin_round_trip(tmp_path):
    """pairs.txt rows -> image paths (evaluation/utility.py:222-253) -> pickled .bin (scripts/raw_img_tf.py:77-86)."""
    from deep_insight_face_b200 import datagen
    from deep_insight_face_b200.evaluation import utility as U

    for name, idx, ext in (("ann", 1, ".jpg"), ("ann", 2, ".png"), ("bob", 1, ".jpg")):
        d = tmp_path / name
        d.mkdir(exist_ok=True)
        (d / ("%s_%04d%s" % (name, idx, ext))).write_bytes(b"img-%s-%d" % (name.encode(), idx))
    pairs = [["ann", "1", "2"], ["ann", "2", "bob", "1"], ["ann", "1", "9"], ["bob", "1", "cat", "1"]]
    paths, issame = U.get_paths(str(tmp_path), pairs)
    assert issame == [True, False] and len(paths) == 4
    assert paths[0].endswith("ann_0001.jpg") and paths[1].endswith("ann_0002.png") and paths[3].endswith("bob_0001.jpg")
    with pytest.raises(RuntimeError):
        U.add_extension(str(tmp_path / "ann" / "ann_0009"))
    blobs = [open(p, "rb").read() for p in paths]
    fname = str(tmp_path / "test.bin")
    datagen.write_test_bin(fname, blobs, issame)
    got, same = datagen.read_test_bin(fname)
    assert got == blobs and same.tolist() == issame
    with pytest.raises(ValueError):
        datagen.write_test_bin(fname, blobs[:3], issame)
