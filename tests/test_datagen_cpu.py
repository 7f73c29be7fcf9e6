"""Host-side helpers around the hot path: PK sampler, label layout, pairs.txt round trip."""
import numpy as np

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
