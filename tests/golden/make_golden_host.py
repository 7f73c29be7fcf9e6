"""Generates tests/golden/host_reference.json from the REFERENCE's own host-side code, in the build container:

    python tests/golden/make_golden_host.py

 * `sample_people` (deep_insight_face/datagen/generator.py:15-41): the module does not import (keras / tensorflow /
   broken imports), so the function's source is cut out of the file with `ast` and executed as is
   under fixed `np.random.seed`s;
 * `write_pairs_to_file` (scripts/generate_pairs.py:60-76): the module imports here; its output text is recorded;
 * `facematch_image_pairs`, `triplet_image_pairs`, `create_pairs` (generator.py:43-124): cut out with `ast` like
   `sample_people` and executed as they are on an image tree built by `image_tree()` below, with the names the
   module's (broken) imports would have bound: `add_extension` is the reference's own (evaluation/utility.py:247,
   imported), `read_pairs` the reference's loop without its final np.array() (which raises on mixed 3/4-field rows
   under numpy >= 1.24), `to_categorical` = np.eye(num_classes)[idx], `InvalidPairsError` an Exception subclass
   (common/utils.py defines none of the four); `os.listdir` is wrapped to return sorted names so that the shuffle of
   triplet_image_pairs is reproducible across filesystems.  Paths are recorded relative to the tree;
 * `TripletPrediction.verify` / `SiamesePrediction.verify` (predictions.py:104-150, :52-89): cut out and executed on a stub
   `self` (see _verify_goldens);
 * default arguments of the functions / constructors the drop-in mirrors, read from the reference's AST.
/root/reference does not exist on the GPU box; only the .json travels.
"""
import ast
import importlib.util
import json
import os
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def function_source(path, name, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    scope = tree.body
    if cls:
        scope = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in scope if isinstance(n, ast.FunctionDef) and n.name == name)
    return ast.get_source_segment(src, node), node


def defaults_of(path, name, cls=None):
    _, node = function_source(path, name, cls)
    args = [a.arg for a in node.args.args]
    vals = []
    for dflt in node.args.defaults:
        try:
            vals.append(ast.literal_eval(dflt))
        except ValueError:
            vals.append("expr:" + ast.unparse(dflt))   # e.g. np.arange(0, 4, 0.01): recorded as source text
    return dict(zip(args[len(args) - len(vals):], vals))


class PersonClass:
    """The shape `sample_people` expects of a dataset entry: len() and .image_paths."""

    def __init__(self, name, n):
        self.name = name
        self.image_paths = [f"{name}/{name}_{i:04d}.jpg" for i in range(n)]

    def __len__(self):
        return len(self.image_paths)


def dataset(seed, n_people):
    rng = np.random.default_rng(seed)
    return [PersonClass(f"person{p:03d}", int(rng.integers(1, 9))) for p in range(n_people)]


TREE = {"ann": [1, 2, 3, 5], "bob": [1, 2], "cat": [1], "dan": [2, 4, 6], "eve": [1, 3]}      # name -> image numbers
PAIR_ROWS = [["ann", "1", "2"], ["ann", "2", "bob", "1"], ["dan", "2", "6"], ["bob", "1", "cat", "1"], ["cat", "1", "ann", "5"],
             ["eve", "1", "3"], ["dan", "4", "eve", "3"], ["bob", "2", "dan", "6"], ["ann", "3", "5"], ["ann", "1"]]


def image_tree(root):
    """<root>/<name>/<name>_%04d.(jpg|png) (odd numbers .jpg, even .png), a hidden file per directory, pairs.txt."""
    for name, numbers in TREE.items():
        os.makedirs(os.path.join(root, name))
        open(os.path.join(root, name, ".hidden"), "w").close()
        for n in numbers:
            open(os.path.join(root, name, "%s_%04d.%s" % (name, n, "jpg" if n % 2 else "png")), "w").close()
    with open(os.path.join(root, "pairs.txt"), "w") as f:
        f.write("1\t%d\n" % len(PAIR_ROWS))
        for row in PAIR_ROWS:
            f.write("\t".join(row) + "\n")
    return os.path.join(root, "pairs.txt")


class _SortedListdirOs:
    """`os` as the cut-out functions see it: listdir returns sorted names."""
    path = os.path

    @staticmethod
    def listdir(d):
        return sorted(os.listdir(d))


def _pair_listing_goldens():
    import sys

    sys.path.insert(0, REF)
    from deep_insight_face.evaluation import utility as ref_util

    def read_pairs(fn):                       # utility.py:256-262 without the final np.array()
        with open(fn) as f:
            return [line.strip().split("\t") for line in f.readlines()[1:]]

    gen = os.path.join(REF, "deep_insight_face/datagen/generator.py")
    ns = {"np": np, "os": _SortedListdirOs, "add_extension": ref_util.add_extension, "read_pairs": read_pairs,
          "InvalidPairsError": type("InvalidPairsError", (Exception,), {}),
          "to_categorical": lambda idx, num_classes: np.eye(num_classes, dtype=np.float32)[np.asarray(idx)]}
    for name in ("facematch_image_pairs", "triplet_image_pairs", "create_pairs"):
        exec(function_source(gen, name)[0], ns)   # the reference's functions, verbatim
    out = {}
    with tempfile.TemporaryDirectory() as td:
        pairs_txt = image_tree(td)
        rel = lambda p: os.path.relpath(p, td)
        pairs, names = ns["facematch_image_pairs"](td, PAIR_ROWS)
        out["facematch"] = {"pairs": [[rel(a), rel(b), bool(s)] for a, b, s in pairs], "names": sorted(names)}
        out["triplet"] = []
        for seed in (0, 1, 2):
            np.random.seed(seed)
            trip, names = ns["triplet_image_pairs"](td, PAIR_ROWS)
            out["triplet"].append({"seed": seed, "triplets": [[rel(x) for x in t] for t in trip], "names": sorted(names)})
        for key, func in (("create_facematch", "facematch_image_pairs"), ("create_triplet", "triplet_image_pairs")):
            np.random.seed(5)
            pairs, names, one_hot = ns["create_pairs"](td, func=ns[func], pairs_txt=pairs_txt)
            out[key] = {"n_pairs": len(pairs), "one_hot_of": {n: [float(v) for v in one_hot[i]] for i, n in enumerate(names)}}
    return out


def verify_cases():
    """(encoding [1, D], stored encodings [n, 1, D], threshold) triples around both default thresholds; regenerated by
    the tests from the same seed."""
    rng = np.random.default_rng(77)
    cases = []
    for i, (gap, thr) in enumerate([(0.02, 0.7), (0.06, 0.7), (0.2, 0.7), (0.01, 0.3), (0.04, 0.3), (0.0, 0.3), (0.05, 0.5)]):
        D = 128
        enc = rng.standard_normal((1, D)).astype(np.float32)
        enc /= np.linalg.norm(enc)
        stored = np.stack([enc + gap * (j + 1) * rng.standard_normal((1, D)).astype(np.float32) for j in range(3)])
        cases.append((f"case{i}", enc, stored.astype(np.float32), thr))
    return cases


def _verify_goldens():
    """predictions.py:104-150 (TripletPrediction.verify) and :52-89 (SiamesePrediction.verify), cut out with `ast` and
    executed as they are on a stub `self`: `_embedding` returns the case's encoding, the siamese `model.predict` is the
    euclidean_distance head of networks/siamese.py:22-24 written out in numpy (the CNN is out of scope)."""
    import contextlib
    import io
    import types

    pred = os.path.join(REF, "deep_insight_face/predictions.py")
    ns = {"np": np}
    exec(function_source(pred, "verify", "TripletPrediction")[0].replace("def verify", "def triplet_verify"), ns)
    exec(function_source(pred, "verify", "SiamesePrediction")[0].replace("def verify", "def siamese_verify"), ns)

    def head(pair):
        a, b = (np.asarray(p, dtype=np.float32).reshape(len(p), -1) for p in pair)
        return np.sqrt(np.maximum(np.sum(np.square(a - b), axis=1, keepdims=True), 1e-7))

    out = []
    for name, enc, stored, thr in verify_cases():
        me = types.SimpleNamespace(_embedding=lambda _p, e=enc: e, model=types.SimpleNamespace(predict=head))
        rec = {"name": name, "threshold": thr}
        for key, fn, db in (("triplet", ns["triplet_verify"], {"who": stored[0]}), ("siamese", ns["siamese_verify"], {"who": stored})):
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                dist, ok = fn(me, "img.jpg", "who", db, threshold=thr)
            rec[key] = {"dist": float(dist), "is_valid": bool(ok), "printed": buf.getvalue()}
        out.append(rec)
    return out


def main():
    out = {"sample_people": [], "defaults": {}}
    out["pair_listing"] = _pair_listing_goldens()
    out["verify"] = _verify_goldens()
    src, _ = function_source(os.path.join(REF, "deep_insight_face/datagen/generator.py"), "sample_people")
    ns = {"np": np}
    exec(src, ns)  # the reference's function, verbatim
    for seed, n_people, P, K in ((0, 40, 18, 4), (1, 25, 6, 3), (2, 60, 45, 2), (3, 12, 4, 8)):
        ds = dataset(100 + seed, n_people)
        np.random.seed(seed)
        paths, counts = ns["sample_people"](ds, P, K)
        out["sample_people"].append({"seed": seed, "dataset_seed": 100 + seed, "n_people": n_people, "P": P, "K": K,
                                     "image_paths": paths, "num_per_class": [int(c) for c in counts]})
    spec = importlib.util.spec_from_file_location("ref_generate_pairs", os.path.join(REF, "scripts/generate_pairs.py"))
    gp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gp)
    matches = [[(f"p{f}_{i}", i + 1, i + 2) for i in range(3)] for f in range(2)]
    mismatches = [[(f"p{f}_{i}", i + 1, f"q{f}_{i}", 7 - i) for i in range(3)] for f in range(2)]
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "pairs.txt")
        gp.write_pairs_to_file(fn, matches, mismatches, 2, 3)
        out["pairs_txt"] = {"matches": matches, "mismatches": mismatches, "num_folds": 2, "n": 3, "text": open(fn).read()}
    d = out["defaults"]
    losses = os.path.join(REF, "deep_insight_face/common/losses.py")
    d["TripletLossWapper.__init__"] = defaults_of(losses, "__init__", "TripletLossWapper")
    d["BatchHardTripletLossEuclideanAutoAlpha.__init__"] = defaults_of(losses, "__init__", "BatchHardTripletLossEuclideanAutoAlpha")
    d["triplet_loss"] = defaults_of(os.path.join(REF, "deep_insight_face/networks/triplet.py"), "triplet_loss")
    d["_accuracy"] = defaults_of(os.path.join(REF, "deep_insight_face/networks/siamese.py"), "_accuracy")
    pred = os.path.join(REF, "deep_insight_face/predictions.py")
    d["TripletPrediction.verify"] = defaults_of(pred, "verify", "TripletPrediction")
    d["SiamesePrediction.verify"] = defaults_of(pred, "verify", "SiamesePrediction")
    d["compare_faces"] = defaults_of(os.path.join(REF, "deep_insight_face/api.py"), "compare_faces")
    util = os.path.join(REF, "deep_insight_face/evaluation/utility.py")
    d["evaluate"] = defaults_of(util, "evaluate")
    d["distance"] = defaults_of(util, "distance")
    with open(os.path.join(HERE, "host_reference.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote host_reference.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
