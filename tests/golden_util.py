"""Loader of the committed fixtures under tests/golden/ (see tests/golden/make_golden.py for where they come from)."""
import glob
import os

import numpy as np

from groundgrid_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_case(path):
    z = np.load(path)
    case = {"dimension": float(z["dimension"]), "resolution": float(z["resolution"]),
            "config": {str(k): float(v) for k, v in zip(z["config_keys"], z["config_values"])},
            "ground_0": z["ground_0"], "groundpatch_0": z["groundpatch_0"], "scans": [],
            "final": {"ground": z["ground"], "groundpatch": z["groundpatch"], "variance": z["variance"],
                      "points": z["points_layer"], "minGroundHeight": z["minGroundHeight"]}}
    for k in range(int(z["n_scans"])):
        pts = z[f"points_{k}"].view(synth.POINT_DTYPE)
        case["scans"].append({"points": pts, "origin": z[f"origin_{k}"], "base_z": float(z[f"base_z_{k}"]),
                              "labels": z[f"labels_{k}"], "order": z[f"order_{k}"]})
    return case


def int_config(cfg):
    """Integer-typed configuration fields come back from the fixture as floats."""
    ints = {"point_count_cell_variance_threshold", "max_ring", "thread_count"}
    return {k: (int(v) if k in ints else v) for k, v in cfg.items()}
