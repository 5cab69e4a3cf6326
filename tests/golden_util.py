"""Loader / replayer of the committed fixtures under tests/golden/ -- vectors produced by the reference itself
(oracle/_ref, see tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

from groundgrid_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAYER_NAMES = ("points", "ground", "groundpatch", "minGroundHeight", "maxGroundHeight", "groundCandidates", "planeDist",
               "m2", "meanVariance", "pointsRaw", "variance")


def case_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_case(path):
    z = np.load(path)
    case = {"name": os.path.basename(path), "dimension": float(z["dimension"]), "resolution": float(z["resolution"]),
            "cells": int(z["cells"]), "expected": z["expected"],
            "config": {str(k): float(v) for k, v in zip(z["config_keys"], z["config_values"])},
            "ground_0": z["ground_0"], "groundpatch_0": z["groundpatch_0"], "scans": [],
            "final": {name: z["layer_" + name] for name in LAYER_NAMES}}
    for k in range(int(z["n_scans"])):
        pts = z[f"points_{k}"].view(synth.POINT_DTYPE)
        case["scans"].append({"points": pts, "origin": z[f"origin_{k}"], "base_z": float(z[f"base_z_{k}"]),
                              "labels": z[f"labels_{k}"], "order": z[f"order_{k}"], "pose": z[f"pose_{k}"], "q": z[f"q_{k}"],
                              "t": z[f"t_{k}"], "T": z[f"T_{k}"], "moved": int(z[f"moved_{k}"]), "position": z[f"position_{k}"],
                              "prior_ground": z[f"prior_ground_{k}"], "prior_groundpatch": z[f"prior_groundpatch_{k}"]})
    return case


def int_config(cfg):
    """Integer-typed configuration fields come back from the fixture as floats."""
    ints = {"point_count_cell_variance_threshold", "max_ring", "thread_count"}
    return {k: (int(v) if k in ints else v) for k, v in cfg.items()}


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def replay(case, impl, update, filter_cloud, layers=LAYER_NAMES):
    """Runs one fixture through an implementation and asserts every recorded number bit for bit.
    impl: object with set_config / init_map / position / layer;  update(impl, x, y, T) -> moved;
    filter_cloud(impl, points, origin, base_z) -> (labels, order)."""
    name = case["name"]
    if case["config"]:
        impl.set_config(**int_config(case["config"]))
    impl.init_map(0.0, 0.0, 0.0)
    assert impl.n == case["cells"], name
    assert same(impl.layer("ground"), case["ground_0"]) and same(impl.layer("groundpatch"), case["groundpatch_0"]), name
    for k, s in enumerate(case["scans"]):
        if k:
            moved = update(impl, float(s["pose"][0]), float(s["pose"][1]), s["T"])
            assert int(bool(moved)) == s["moved"], (name, k)
        assert np.array_equal(np.asarray(impl.position()), s["position"]), (name, k, impl.position(), s["position"])
        assert same(impl.layer("ground"), s["prior_ground"]), (name, k, "rolled ground")
        assert same(impl.layer("groundpatch"), s["prior_groundpatch"]), (name, k, "rolled groundpatch")
        labels, order = filter_cloud(impl, s["points"], s["origin"], s["base_z"])
        assert np.array_equal(labels, s["labels"]), (name, k, int((labels != s["labels"]).sum()))
        assert np.array_equal(order, s["order"]), (name, k)
    for lname in layers:
        assert same(impl.layer(lname), case["final"][lname]), (name, lname)
