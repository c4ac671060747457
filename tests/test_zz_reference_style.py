"""Mirrors of the reference's own unit tests for the hot path (test/test_models.py:98-300), so that a user of the
reference finds the same contracts here.  CPU: shapes of backbone+PAN outputs (through the lowering) and of the
anchor generator for every version / P6.  GPU: PostProcess called the reference way (head_outputs, grids, shifts)
for 3 and 4 levels, checked against the oracle instead of only for types."""
import numpy as np
import pytest
import torch

import parity_util as util
from oracle import restate as R
from yolort_b200.models._checkpoint import get_yolov5_size
from yolort_b200.models.anchor_utils import AnchorGenerator
from yolort_b200.models.backbone_utils import darknet_pan_backbone
from yolort_b200.models.box_head import PostProcess, YOLOHead


def _in_channels(width_multiple, use_p6):
    return [int(g * width_multiple) for g in ([256, 512, 768, 1024] if use_p6 else [256, 512, 1024])]


def _strides(use_p6):
    return [8, 16, 32, 64] if use_p6 else [8, 16, 32]


def _anchor_grids(use_p6):
    return util.P6_ANCHORS if use_p6 else R.DEFAULT_ANCHORS


@pytest.mark.parametrize("depth_multiple,width_multiple,version,use_p6", [
    (0.33, 0.5, "r3.1", False), (0.33, 0.5, "r4.0", False), (0.33, 0.5, "r6.0", False), (0.33, 0.5, "r6.0", True),
    (0.67, 0.75, "r6.0", False)])
def test_backbone_with_pan_shapes(depth_multiple, width_multiple, version, use_p6):
    """test_models.py:188-222: one output per level with (C_l, H / s_l, W / s_l); here read off the lowered plan."""
    from yolort_b200.engine import lower_yolo
    from yolort_b200.models.yolo import YOLO

    size = get_yolov5_size(depth_multiple, width_multiple)
    bb = darknet_pan_backbone(f"darknet_{size}_{version.replace('.', '_')}", depth_multiple, width_multiple,
                              version=version, use_p6=use_p6)
    assert bb.out_channels == _in_channels(width_multiple, use_p6)
    model = YOLO(bb, 80, strides=_strides(use_p6), anchor_grids=_anchor_grids(use_p6)).eval()
    L, x0, heads, feats = lower_yolo(model, torch.float16, torch.device("cpu"))
    assert len(feats) == (4 if use_p6 else 3)
    for (name, view), c, s in zip(feats.items(), _in_channels(width_multiple, use_p6), _strides(use_p6)):
        assert (view.C, view.buf.div) == (c, s), name
    for hb, s in zip(heads, _strides(use_p6)):
        assert hb.div == s and hb.C == 256          # 255 logits padded to 256 channels


@pytest.mark.parametrize("use_p6", [False, True])
@pytest.mark.parametrize("batch_size,height,width", [(4, 448, 320), (2, 384, 640)])
def test_anchor_generator_shapes(use_p6, batch_size, height, width):
    """test_models.py:232-249."""
    strides = _strides(use_p6)
    fmaps = [torch.rand(batch_size, c, height // s, width // s) for c, s in zip(_in_channels(0.5, use_p6), strides)]
    anchors = AnchorGenerator(strides, _anchor_grids(use_p6))(fmaps)
    assert len(anchors) == 2 and len(anchors[0]) == len(anchors[1]) == len(strides)
    for i, s in enumerate(strides):
        assert tuple(anchors[0][i].shape) == (1, 3, height // s, width // s, 2)
        assert tuple(anchors[1][i].shape) == (1, 3, height // s, width // s, 2)


def test_yolo_head_parameters():
    """test_models.py:253-273 (shapes of the head): one 1x1 conv per level with A * (nc + 5) outputs."""
    head = YOLOHead(_in_channels(0.5, False), 3, _strides(False), 80)
    assert [tuple(c.weight.shape) for c in head.head] == [(255, 128, 1, 1), (255, 256, 1, 1), (255, 512, 1, 1)]


@pytest.mark.gpu
@pytest.mark.parametrize("use_p6", [False, True])
def test_postprocessors_reference_call(use_p6):
    """test_models.py:283-300: PostProcess(strides, 0.5, 0.45, 100)(head_outputs, grids, shifts) on random heads."""
    N, H, W = 4, 416, 352
    strides, grids_cfg = _strides(use_p6), _anchor_grids(use_p6)
    g = torch.Generator().manual_seed(3 + use_p6)
    fmaps = [torch.zeros(N, 1, H // s, W // s) for s in strides]
    # logits in [-5, 1): ~500 candidates per image above 0.5, the 100 best survive (detections_per_img)
    heads = [torch.rand(N, 3, H // s, W // s, 85, generator=g) * 6.0 - 5.0 for s in strides]
    grids, shifts = AnchorGenerator(strides, grids_cfg)(fmaps)
    out = PostProcess(strides, 0.5, 0.45, 100)([h.to("cuda:0") for h in heads], grids, shifts)
    assert len(out) == N and isinstance(out[0], dict)
    ref = R.postprocess(heads, 0.5, 0.45, 100, strides=strides, anchor_grids=grids_cfg)
    for got, want in zip(out, ref):
        assert all(isinstance(got[k], torch.Tensor) for k in ("boxes", "labels", "scores"))
        # exact parity of this kernel is asserted on the committed fixtures (test_gpu_postprocess.py); here the scores
        # sit in a narrow band (0.50-0.53), so allow the cut at 100 to fall differently for near-equal scores
        assert len(got["scores"]) == len(want["scores"]) == 100
        assert util.match_fraction(util.to_np(got), want, iou_thr=0.99) >= 0.97
