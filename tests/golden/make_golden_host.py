"""Generates tests/golden/host_reference.json from the REFERENCE's own host-side code, in the build container:

    python tests/golden/make_golden_host.py

 * `sample_people` (deep_insight_face/datagen/generator.py:15-41): the module does not import (keras / tensorflow /
   broken imports), so the function's source is cut out of the file with `ast` and executed as is
   under fixed `np.random.seed`s;
 * `write_pairs_to_file` (scripts/generate_pairs.py:60-76): the module imports here; its output text is recorded;
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


def main():
    out = {"sample_people": [], "defaults": {}}
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
