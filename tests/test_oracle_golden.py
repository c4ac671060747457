"""Pins oracle/restate.py against fixtures produced by the unmodified reference (oracle/make_golden.py).
CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import restate as R
import parity_util as util


def test_resize_shape_and_batch_geometry(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "letterbox_geometry.json")))
    for h, w, nh, nw in g["sizes"]:
        assert R.resize_shape(h, w) == (nh, nw), (h, w)
    # the 639 trap
    assert R.resize_shape(800, 600) == (639, 479)
    assert R.resize_shape(417, 523) == (510, 639)
    for b in g["batches"]:
        sizes = [(g["sizes"][i][2], g["sizes"][i][3]) for i in b["idx"]]
        assert R.batch_shape(sizes) == (b["Hb"], b["Wb"])
        for (nh, nw), off in zip(sizes, b["offsets"]):
            assert R.pad_offsets(b["Hb"], b["Wb"], nh, nw) == tuple(off)
        probe = np.array([[10.0, 20.0, 300.5, 400.25], [0.0, 0.0, b["Wb"], b["Hb"]]], dtype=np.float32)
        for i, ref in zip(b["idx"], b["scaled"]):
            got = R.scale_coords(probe, b["Hb"], b["Wb"], g["sizes"][i][0], g["sizes"][i][1])
            assert np.array_equal(got, np.array(ref, dtype=np.float32)), (i, got, ref)


def test_letterbox_pixels(golden_dir):
    z = util.load_npz("letterbox_pixels.npz")
    ims = [torch.from_numpy(z[f"img{i}"]) for i in range(4)]
    batch, sizes, _ = R.letterbox(ims, 96.0, 96.0)
    assert tuple(batch.shape) == z["batch"].shape
    assert [tuple(s) for s in sizes] == [tuple(s) for s in z["sizes"]]
    err = np.abs(batch.numpy() - z["batch"]).max()
    assert err <= 5e-5, err  # fp32 restatement vs ATen CPU kernel
    # identity resize (96x96 image) must be bit exact
    assert np.array_equal(batch.numpy()[2], z["batch"][2])


def test_network_features_and_heads(golden_dir):
    z = util.load_npz("network_n.npz")
    sd = util.synth_state_dict(util.layouts()["n"], knob_obj=7.0, knob_cls=4.5, seed=0)
    assert util.checksum(sd) == pytest.approx(float(z["checksum"]), rel=1e-12), "synthetic weights not reproducible here"
    net = R.Net(sd)
    with torch.no_grad():
        feats = net.backbone(torch.from_numpy(z["x"]))
        heads = net.head(feats)
    for got, key in zip(feats, ("p3", "p4", "p5")):
        np.testing.assert_allclose(got.numpy(), z[key], atol=2e-5, rtol=1e-5)
    for got, key in zip(heads, ("h0", "h1", "h2")):
        np.testing.assert_allclose(got.numpy(), z[key], atol=2e-5, rtol=1e-5)
    dets = R.postprocess(heads, 0.15, 0.45, 300)
    util.assert_dets_close(dets[0], util.dets_from_npz(z, 1)[0], box_atol=1e-3, score_atol=1e-5, allow_tie_swaps=True)


@pytest.mark.parametrize("case", ["few", "trick", "vanilla", "empty"])
def test_postprocess_bit_exact(case, golden_dir):
    z = util.load_npz(f"postprocess_{case}.npz")
    heads = [torch.from_numpy(z[f"h{i}"]) for i in range(3)]
    dets = R.postprocess(heads, float(z["thr"]), 0.45, 300)
    for got, ref in zip(dets, util.dets_from_npz(z, 2)):
        assert np.array_equal(got["labels"], ref["labels"])
        assert np.array_equal(got["scores"], ref["scores"])
        assert np.array_equal(got["boxes"], ref["boxes"])
    if case == "vanilla":
        assert all(d["n_candidates"] > 1000 for d in dets)   # exercises the per-class branch
    if case in ("few", "trick"):
        assert all(0 < d["n_candidates"] <= 1000 for d in dets)  # exercises the offset-trick branch


def test_end_to_end(golden_dir):
    z = util.load_npz("e2e_n.npz")
    sd = util.synth_state_dict(util.layouts()["n"], knob_obj=7.0, knob_cls=4.5, seed=0)
    ims = [torch.from_numpy(z["img0"]), torch.from_numpy(z["img1"])]
    dets = R.detect(sd, ims, score_thresh=0.15, size=(128, 128))
    for got, ref in zip(dets, util.dets_from_npz(z, 2)):
        util.assert_dets_close(got, ref, box_atol=2e-2, score_atol=2e-5, allow_tie_swaps=True)


def test_nms_semantics_small_cases():
    # strict '>' : IoU exactly 0.5 at thr 0.5 is kept; ties keep index order; degenerate boxes are kept
    b = np.array([[0, 0, 2, 1], [1, 0, 3, 1], [10, 10, 9, 9]], dtype=np.float32)
    s = np.array([0.9, 0.9, 0.1], dtype=np.float32)
    l = np.zeros(3, dtype=np.int64)
    keep = R.batched_nms(b, s, l, 1.0 / 3.0, R.EXACT_PER_CLASS)
    assert keep.tolist() == [0, 1, 2]
    keep = R.batched_nms(b, s, l, 0.33, R.EXACT_PER_CLASS)
    assert keep.tolist() == [0, 2]
    assert R.batched_nms(b[:0], s[:0], l[:0], 0.5).tolist() == []
