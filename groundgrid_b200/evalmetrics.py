"""Metric definitions of the reference's evaluator (scripts/eval_groundpoint_classifier.py:135-195),
computed from the per-label tallies of gg_eval_accumulate (counts[id] = (predicted ground, predicted
non-ground) for ground-truth label id; the label travels in the `ring` field)."""

# cfg/semantic-kitti-all.yaml:2-36
LABELS = {0: "unlabeled", 1: "outlier", 10: "car", 11: "bicycle", 13: "bus", 15: "motorcycle", 16: "on-rails", 18: "truck",
          20: "other-vehicle", 30: "person", 31: "bicyclist", 32: "motorcyclist", 40: "road", 44: "parking", 48: "sidewalk",
          49: "other-ground", 50: "building", 51: "fence", 52: "other-structure", 60: "lane-marking", 70: "vegetation", 71: "trunk",
          72: "terrain", 80: "pole", 81: "traffic-sign", 99: "other-object", 252: "moving-car", 253: "moving-bicyclist",
          254: "moving-person", 255: "moving-motorcyclist", 256: "moving-on-rails", 257: "moving-bus", 258: "moving-truck",
          259: "moving-other-vehicle"}
# eval_groundpoint_classifier.py:74-79
GROUND = ["road", "sidewalk", "parking", "lane-marking"]
ADDITIONAL_GROUND = ["other-ground", "terrain"]
NON_GROUND = ["bicycle", "moving-bicyclist", "motorcycle", "moving-motorcyclist", "person", "moving-person", "traffic-sign", "car",
              "moving-car", "motorcyclist", "bicyclist", "truck", "moving-truck", "building", "fence", "trunk", "pole", "bus", "on-rails",
              "other-vehicle", "other-structure", "other-object", "moving-on-rails", "moving-bus", "moving-other-vehicle"]


def metrics(counts):
    """counts: array [id][2] (ground, non-ground).  Returns the evaluator's summary numbers."""
    by_name = {name: (int(counts[i][0]), int(counts[i][1])) for i, name in LABELS.items()}
    tp = sum(by_name[n][0] for n in GROUND + ADDITIONAL_GROUND)              # ground predicted ground
    fn = sum(by_name[n][1] for n in GROUND + ADDITIONAL_GROUND)              # ground predicted non-ground
    fp = sum(by_name[n][0] for n in NON_GROUND)                              # obstacle predicted ground
    tn = sum(by_name[n][1] for n in NON_GROUND)
    gt_ground = sum(sum(by_name[n]) for n in GROUND + ADDITIONAL_GROUND)
    out = dict(tp=tp, fn=fn, fp=fp, tn=tn)
    out["precision"] = tp / (fp + tp) if fp + tp else float("nan")
    out["recall"] = tp / (fn + tp) if fn + tp else float("nan")
    out["f1"] = 2 * tp / (2 * tp + fp + fn) if tp + fp + fn else float("nan")
    out["accuracy"] = (tp + tn) / (tp + tn + fp + fn) if tp + tn + fp + fn else float("nan")
    out["iou_ground"] = tp / (fp + gt_ground) if fp + gt_ground else float("nan")
    return out
