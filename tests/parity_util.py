"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

from oracle.make_golden import checksum, synth_image_u8, synth_state_dict  # noqa: F401  (pure functions, no reference import)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def layouts():
    with open(os.path.join(GOLDEN, "state_dict_layouts.json")) as f:
        return json.load(f)


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
