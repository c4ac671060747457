"""The other model sizes of BASELINE.json's configs (m / l / x: channel widths 48..1280, deeper C3 stacks, bf16,
mixed input sizes, 1280-pixel canvas) against the CPU oracle (B200)."""
import pytest
import torch

import parity_util as util
from oracle import restate as R
from yolort_b200.models import yolov5l, yolov5m, yolov5x

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(ctor, name, **kw):
    sd = util.synth_state_dict(util.layouts()[name], knob_obj=7.0, knob_cls=4.5, seed=1)
    m = ctor(**kw).eval()
    m.load_state_dict(sd)
    return m.to(DEV), sd


def _check(m, sd, ims, min_frac, iou=0.9, **kw):
    ref = R.detect(sd, ims, **kw)
    out = m([im.to(DEV) for im in ims])
    for got, want, im in zip(out, ref, ims):
        frac = util.match_fraction(util.to_np(got), want, iou_thr=iou)
        print(tuple(im.shape[1:]), "matched", round(frac, 3), len(got["scores"]), len(want["scores"]))
        assert frac >= min_frac


def test_yolov5m_fp16_widths_48_to_768():
    m, sd = _build(yolov5m, "m", size=(160, 160), score_thresh=0.2)
    ims = [util.synth_image_u8(120, 160, 1), util.synth_image_u8(160, 96, 2)]
    _check(m, sd, ims, 0.8, score_thresh=0.2, size=(160, 160))


def test_yolov5m_bf16():
    m, sd = _build(yolov5m, "m", size=(160, 160), score_thresh=0.2)
    m = m.to(torch.bfloat16)
    ims = [util.synth_image_u8(160, 160, 3)]
    _check(m, sd, ims, 0.4, iou=0.8, score_thresh=0.2, size=(160, 160))


def test_yolov5l_mixed_sizes():
    m, sd = _build(yolov5l, "l", size=(192, 192), score_thresh=0.2)
    ims = [util.synth_image_u8(h, w, 10 + i) for i, (h, w) in enumerate([(150, 192), (192, 100), (97, 131)])]
    _check(m, sd, ims, 0.8, score_thresh=0.2, size=(192, 192))


def test_yolov5x_widths_80_to_1280():
    m, sd = _build(yolov5x, "x", size=(128, 128), score_thresh=0.2)
    ims = [util.synth_image_u8(128, 128, 5)]
    _check(m, sd, ims, 0.8, score_thresh=0.2, size=(128, 128))


def test_yolov5x_1280_canvas_runs():
    """BASELINE.json configs[4] shape (1280x1280, 100 800 anchors): plumbing + output contract (no oracle: the CPU
    path needs ~0.8 TFLOP per image)."""
    m, _ = _build(yolov5x, "x", size=(1280, 1280), score_thresh=0.3)
    out = m([torch.randint(0, 256, (3, 1280, 1280), dtype=torch.uint8, device=DEV)])
    assert len(out) == 1 and out[0]["boxes"].shape[1] == 4 and out[0]["boxes"].shape[0] <= 300
