"""Plan infrastructure on the device (B200): shared weights across shapes, arena liveness reuse, LRU, invalidation when
weights change, non-current devices, callable sub-modules / forward hooks, training-mode head outputs."""
import time

import numpy as np
import pytest
import torch

import parity_util as util
from yolort_b200 import _C
from yolort_b200.models import yolov5n, yolov5s
from yolort_b200.models.yolo import YOLO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model_n(seed=0, **kw):
    sd = util.synth_state_dict(util.layouts()["n"], knob_obj=7.0, knob_cls=4.5, seed=seed)
    m = yolov5n(size=(128, 128), score_thresh=0.15, **kw).eval()
    m.load_state_dict(sd)
    return m.to(DEV), sd


def _same(a, b):
    return len(a) == len(b) and all(torch.equal(x[k], y[k]) for x, y in zip(a, b) for k in ("scores", "labels", "boxes"))


def test_weights_lowered_once_and_new_shapes_are_cheap():
    m = yolov5s().eval().to(DEV)
    eng = m.model.engine()
    m.model.get_plan(2, 320, 320)
    torch.cuda.synchronize()
    t = []
    for hw in ((256, 320), (320, 256), (384, 640), (640, 640), (96, 128)):
        t0 = time.perf_counter()
        m.model.get_plan(3, *hw)
        torch.cuda.synchronize()
        t.append((time.perf_counter() - t0) * 1e3)
    print("plan creation ms for 5 new shapes:", [round(x, 1) for x in t])
    assert eng.lowerings == 1
    assert sorted(t)[len(t) // 2] < 50.0          # second-shape plan creation < 50 ms
    assert len({id(p._low) for p in eng._plans.values()}) == 1


def test_arena_reuse_equals_unshared_arena_and_is_smaller():
    m, sd = _model_n()
    ims = [util.synth_image_u8(90, 128, 21).to(DEV), util.synth_image_u8(100, 75, 22).to(DEV)]
    out = m(ims)
    plan = m.model.get_plan(2, 128, 128)
    full = m.model.get_plan(2, 128, 128, keep_intermediates=True)
    assert plan is not full and plan.arena_bytes < 0.5 * full.arena_bytes
    geoms, (Hb, Wb) = m.transform.geometry(ims)
    m.transform.letterbox_into(ims, geoms, Hb, Wb, full.input, _C.YB_LAYOUT_S2D16)
    full.run()
    plan.run()
    torch.cuda.synchronize()
    for a, b in zip(plan.heads, full.heads):
        assert torch.equal(a, b)
    assert _same(out, m(ims))


def test_plan_cache_is_bounded():
    m, sd = _model_n()
    eng = m.model.engine()
    eng.MAX_PLANS = 4
    for k in range(7):
        m.model.get_plan(1, 64 + 32 * k, 64)
    assert len(eng._plans) == 4
    first = m.model.get_plan(1, 64, 64)           # evicted -> rebuilt, weights still lowered once
    assert eng.lowerings == 1 and first is m.model.get_plan(1, 64, 64)


def test_new_weights_through_the_wrapper_after_a_forward():
    m, sd = _model_n(seed=0)
    ims = [util.synth_image_u8(128, 128, 5).to(DEV)]
    out0 = m(ims)
    sd1 = util.synth_state_dict(util.layouts()["n"], knob_obj=7.0, knob_cls=4.5, seed=3)
    m.load_state_dict(sd1)                         # nn.Module.load_state_dict on the PARENT
    out1 = m(ims)
    fresh, _ = _model_n(seed=3)
    assert _same(out1, fresh(ims)) and not _same(out0, out1)
    # in-place edit of a parameter is picked up as well (parameter-version fingerprint)
    with torch.no_grad():
        m.model.head.head[0].bias.add_(1.0)
    out2 = m(ims)
    assert not _same(out1, out2) and m.model.engine().lowerings == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two devices")
def test_model_on_a_non_current_device():
    m, sd = _model_n()
    m1 = m.to("cuda:1")
    ims = [util.synth_image_u8(90, 128, 21), util.synth_image_u8(100, 75, 22)]
    assert torch.cuda.current_device() == 0
    out1 = m1([im.to("cuda:1") for im in ims])
    assert torch.cuda.current_device() == 0 and out1[0]["boxes"].device.index == 1
    m0, _ = _model_n()
    out0 = m0([im.to("cuda:0") for im in ims])
    for a, b in zip(out0, out1):
        assert torch.equal(a["labels"].cpu(), b["labels"].cpu()) and torch.equal(a["boxes"].cpu(), b["boxes"].cpu())


def test_backbone_and_head_are_callable_and_hooks_fire():
    """yolort/utils/hooks.py:7-26 (FeatureExtractor): forward hooks on backbone / head see the stage outputs; the staged
    forward returns the detections of the fast path."""
    m, sd = _model_n()
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 96, 128, generator=g).to(DEV)
    fast = m.model(x)
    feats = m.model.backbone(x)
    assert [tuple(f.shape) for f in feats] == [(2, 64, 12, 16), (2, 128, 6, 8), (2, 256, 3, 4)]
    heads = m.model.head(feats)
    assert [tuple(h.shape) for h in heads] == [(2, 3, 12, 16, 85), (2, 3, 6, 8, 85), (2, 3, 3, 4, 85)]
    dets = m.model.post_process(heads)
    assert _same(fast, dets)
    seen = {}
    h1 = m.model.backbone.register_forward_hook(lambda mod, inp, out: seen.__setitem__("backbone", out))
    h2 = m.model.head.register_forward_hook(lambda mod, inp, out: seen.__setitem__("head", out))
    staged = m.model(x)
    assert set(seen) == {"backbone", "head"} and _same(fast, staged)
    for a, b in zip(seen["head"], heads):
        assert torch.equal(a, b)
    # through the YOLOv5 wrapper: transform -> model -> rescale, hooks still fire, same detections as the fused path
    ims = [util.synth_image_u8(90, 128, 21).to(DEV), util.synth_image_u8(100, 75, 22).to(DEV)]
    seen.clear()
    staged_w = m(ims)
    h1.remove()
    h2.remove()
    fast_w = m(ims)
    assert set(seen) == {"backbone", "head"}
    for a, b in zip(staged_w, fast_w):
        assert torch.equal(a["labels"], b["labels"])
        np.testing.assert_allclose(a["boxes"].cpu().numpy(), b["boxes"].cpu().numpy(), rtol=0, atol=2e-3)
        np.testing.assert_allclose(a["scores"].cpu().numpy(), b["scores"].cpu().numpy(), rtol=0, atol=1e-6)


def test_training_mode_hands_the_head_outputs_to_the_criterion():
    """yolort/models/yolo.py:168-171: in training mode the detector returns criterion(targets, head_outputs), the head
    outputs being the raw per-level [N, A, H, W, nc+5] list (box_head.py:68-82)."""
    m, sd = _model_n()
    g = torch.Generator().manual_seed(2)
    x = torch.rand(1, 3, 64, 96, generator=g).to(DEV)
    eval_heads = m.model.head(m.model.backbone(x))
    got = {}

    def criterion(targets, head_outputs):
        got["targets"], got["outs"] = targets, head_outputs
        return {"n_levels": len(head_outputs)}

    m.model.compute_loss = criterion
    m.model.train()
    res = m.model(x, targets="T")
    m.model.eval()
    assert res == {"n_levels": 3} and got["targets"] == "T"
    for a, b in zip(got["outs"], eval_heads):
        assert torch.equal(a, b) and a.shape[1] == 3 and a.shape[-1] == 85
    m.model.compute_loss = None
    m.model.train()
    with pytest.raises(NotImplementedError):
        m.model(x, targets="T")
    m.model.eval()


def test_plan_replayed_as_a_cuda_graph_gives_the_same_logits():
    m, sd = _model_n()
    ims = [util.synth_image_u8(90, 128, 21).to(DEV), util.synth_image_u8(100, 75, 22).to(DEV)]
    want = m(ims)
    plan = m.model.get_plan(2, 128, 128)
    heads = [h.clone() for h in plan.heads]
    for h in plan.heads:
        h.zero_()
    plan.run_graph()
    plan.run_graph()
    torch.cuda.synchronize()
    for a, b in zip(plan.heads, heads):
        assert torch.equal(a, b)
    m.model.engine().graphs = True
    m.model.engine()._plans.clear()
    assert _same(want, m(ims)) and m.model.get_plan(2, 128, 128).use_graph
