"""`predict(paths)`: threaded CPU decode -> pinned staging -> one H2D -> HWC letterbox."""
import os

import pytest
import torch

import parity_util as util
from oracle import restate as R
from yolort_b200.models import yolov5n

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _write(tmp_path):
    from torchvision.io import write_jpeg, write_png

    paths = []
    for i, (h, w) in enumerate([(90, 128), (100, 75), (128, 128), (61, 117)]):
        im = util.synth_image_u8(h, w, 40 + i)
        # smooth the noise so that JPEG keeps some structure
        im = torch.nn.functional.avg_pool2d(im.float().unsqueeze(0), 5, 1, 2).squeeze(0).round().to(torch.uint8)
        p = str(tmp_path / f"im{i}.{'png' if i % 2 else 'jpg'}")
        (write_png if i % 2 else write_jpeg)(im, p)
        paths.append(p)
    return paths


def test_predict_paths_equals_predict_tensors_and_oracle(tmp_path):
    from torchvision.io import ImageReadMode, read_image

    paths = _write(tmp_path)
    sd = util.synth_state_dict(util.layouts()["n"], knob_obj=7.0, knob_cls=4.5, seed=0)
    m = yolov5n(size=(128, 128), score_thresh=0.15).eval()
    m.load_state_dict(sd)
    m = m.to(DEV)
    decoded = [read_image(p, mode=ImageReadMode.RGB) for p in paths]
    for rep in range(3):                       # exercises both pinned staging slots and their reuse
        got = m.predict(paths)
        want = m.predict([d.contiguous() for d in decoded])      # planar tensors, per-image placement
        assert len(got) == len(want) == 4
        for a, b in zip(got, want):
            assert torch.equal(a["labels"], b["labels"]) and torch.equal(a["scores"], b["scores"])
            assert torch.equal(a["boxes"], b["boxes"])
    one = m.predict(paths[0])                  # a single path string (yolov5.py:236)
    assert len(one) == 1 and one[0]["boxes"].shape[1] == 4
    ref = R.detect(sd, [d.contiguous() for d in decoded], score_thresh=0.15, size=(128, 128))
    for a, r, d in zip(got, ref, decoded):
        frac = util.match_fraction(util.to_np(a), r, iou_thr=0.9, side=max(int(v) for v in d.shape[1:]))
        print("ingest matched", round(frac, 3), len(a["scores"]), len(r["scores"]))
        assert frac >= 0.96      # measured 0.987 .. 1.0


def test_custom_loader_is_respected(tmp_path):
    paths = _write(tmp_path)[:2]
    m = yolov5n(size=(128, 128), score_thresh=0.3).eval().to(DEV)
    seen = []

    def loader(p):
        seen.append(os.path.basename(p))
        return torch.zeros(3, 64, 64, dtype=torch.uint8)

    out = m.predict(paths, image_loader=loader)
    assert seen == ["im0.jpg", "im1.png"] and len(out) == 2
