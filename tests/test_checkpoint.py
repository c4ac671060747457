"""Upstream-checkpoint loader vs the reference converter (CPU; needs /root/reference, skipped on the GPU box)."""
import os
import sys

import pytest
import torch

from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present on this machine")


def _make_upstream(tmp_path_factory, yaml_rel: str) -> str:
    ref_import.import_reference()
    from yolort.v5 import add_yolov5_context

    path = str(tmp_path_factory.mktemp("ckpt") / (os.path.basename(yaml_rel).replace(".yaml", "") + "_synth.pt"))
    with add_yolov5_context():
        from models.yolo import Model  # noqa: resolved inside the reference's yolort/v5 tree

        cfg = os.path.join(ref_import.REFERENCE_ROOT, "yolort", "v5", "models", yaml_rel)
        torch.manual_seed(0)
        m = Model(cfg)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
        torch.save({"model": m.half(), "ema": None}, path)
    # the loader under test must not need the upstream tree on sys.path
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    return path


@pytest.fixture(scope="module")
def upstream_ckpt(tmp_path_factory):
    """A synthetic upstream yolov5n checkpoint, pickled exactly like ultralytics does (classes models.yolo.* /
    models.common.*), built from the reference's vendored upstream tree (SURVEY.md section 8c recipe)."""
    return _make_upstream(tmp_path_factory, "yolov5n.yaml")


@pytest.fixture(scope="module")
def upstream_ckpt_p6(tmp_path_factory):
    return _make_upstream(tmp_path_factory, os.path.join("hub", "yolov5n6.yaml"))


def test_converted_state_dict_equals_reference_converter(upstream_ckpt):
    from yolort.models._checkpoint import load_from_ultralytics as ref_load

    from yolort_b200.models._checkpoint import load_from_ultralytics

    mine = load_from_ultralytics(upstream_ckpt)
    ref = ref_load(upstream_ckpt)
    assert mine["num_classes"] == ref["num_classes"] == 80
    assert (mine["depth_multiple"], mine["width_multiple"], mine["size"]) == (0.33, 0.25, "n")
    assert mine["strides"] == [8, 16, 32] and isinstance(mine["strides"][0], int)
    assert mine["anchor_grids"] == ref["anchor_grids"]
    assert list(mine["state_dict"].keys()) == list(ref["state_dict"].keys())
    for k, v in ref["state_dict"].items():
        assert mine["state_dict"][k].dtype == v.dtype and torch.equal(mine["state_dict"][k], v), k


def test_load_from_yolov5_builds_a_model(upstream_ckpt):
    from yolort_b200.models import YOLOv5

    m = YOLOv5.load_from_yolov5(upstream_ckpt, score_thresh=0.3)
    assert m.model.post_process.score_thresh == 0.3
    assert m.model.anchor_generator.strides == [8, 16, 32]
    assert len(m.state_dict()) == 348


def test_p6_checkpoint_equals_reference_converter(upstream_ckpt_p6):
    """_checkpoint.py:49-58: 4 strides -> use_p6, module index maps of the 34-module upstream graph."""
    from yolort.models._checkpoint import load_from_ultralytics as ref_load

    from yolort_b200.models import YOLOv5
    from yolort_b200.models._checkpoint import load_from_ultralytics

    mine = load_from_ultralytics(upstream_ckpt_p6)
    ref = ref_load(upstream_ckpt_p6)
    assert mine["use_p6"] is True and ref["use_p6"] is True
    assert mine["strides"] == [8, 16, 32, 64]
    assert mine["anchor_grids"] == ref["anchor_grids"]
    assert list(mine["state_dict"].keys()) == list(ref["state_dict"].keys())
    for k, v in ref["state_dict"].items():
        assert mine["state_dict"][k].dtype == v.dtype and torch.equal(mine["state_dict"][k], v), k
    m = YOLOv5.load_from_yolov5(upstream_ckpt_p6, size_divisible=64)
    assert m.model.anchor_generator.strides == [8, 16, 32, 64] and len(m.model.head.head) == 4
