"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

from oracle.make_golden import checksum, synth_image_u8, synth_state_dict  # noqa: F401  (pure functions, no reference import)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def layouts():
    out = {}
    for fn in ("state_dict_layouts.json", "state_dict_layouts_p6.json", "state_dict_layouts_v4.json"):
        with open(os.path.join(GOLDEN, fn)) as f:
            out.update(json.load(f))
    return out


P6_STRIDES = [8, 16, 32, 64]
P6_ANCHORS = [[19, 27, 44, 40, 38, 94], [96, 68, 86, 152, 180, 137], [140, 301, 303, 264, 238, 542],
              [436, 615, 739, 380, 925, 792]]
GAIN_N6 = 2.22   # oracle/make_golden_p6.py
GAINS_V4 = {"s_r40": 2.1, "s_r31": 2.05}   # oracle/make_golden_v4.py


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def dets_from_npz(z, n):
    return [{k: z[f"det{i}_{k}"] for k in ("scores", "labels", "boxes")} for i in range(n)]


def assert_dets_close(got, ref, box_atol, score_atol, allow_tie_swaps=False):
    """got/ref: dicts of numpy arrays. Labels and order exact unless equal scores swap places."""
    assert len(got["scores"]) == len(ref["scores"]), (len(got["scores"]), len(ref["scores"]))
    if len(ref["scores"]) == 0:
        return
    gl, rl = np.asarray(got["labels"]), np.asarray(ref["labels"])
    gb, rb = np.asarray(got["boxes"], dtype=np.float64), np.asarray(ref["boxes"], dtype=np.float64)
    gs, rs = np.asarray(got["scores"], dtype=np.float64), np.asarray(ref["scores"], dtype=np.float64)
    np.testing.assert_allclose(gs, rs, atol=score_atol, rtol=0)
    if allow_tie_swaps and not np.array_equal(gl, rl):
        # compare as multisets within groups of (nearly) equal score
        order_g = np.lexsort((gb[:, 0], gb[:, 1], gl, -np.round(gs, 6)))
        order_r = np.lexsort((rb[:, 0], rb[:, 1], rl, -np.round(rs, 6)))
        gl, rl, gb, rb = gl[order_g], rl[order_r], gb[order_g], rb[order_r]
    assert np.array_equal(gl, rl)
    np.testing.assert_allclose(gb, rb, atol=box_atol, rtol=0)


def to_np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def _iou_matrix(a, b):
    x1 = np.maximum(a[:, None, 0], b[None, :, 0]); y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2]); y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]); ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter + 1e-12)


def match_fraction(got, ref, iou_thr=0.9):
    """Fraction of reference detections that have a same-label detection with IoU > iou_thr."""
    if len(ref["scores"]) == 0:
        return 1.0 if len(got["scores"]) == 0 else 0.0
    if len(got["scores"]) == 0:
        return 0.0
    iou = _iou_matrix(np.asarray(ref["boxes"], dtype=np.float64), np.asarray(got["boxes"], dtype=np.float64))
    same = np.asarray(ref["labels"])[:, None] == np.asarray(got["labels"])[None, :]
    return float(((iou > iou_thr) & same).any(axis=1).mean())
