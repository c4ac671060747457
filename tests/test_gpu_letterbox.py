"""Letterbox kernel vs the reference fixtures / oracle (B200)."""
import numpy as np
import pytest
import torch

import parity_util as util
from oracle import restate as R
from yolort_b200 import _C
from yolort_b200.models.transform import YOLOTransform

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_pixels_vs_reference_fixture_uint8():
    z = util.load_npz("letterbox_pixels.npz")
    ims = [torch.from_numpy(z[f"img{i}"]).to(DEV) for i in range(4)]
    tr = YOLOTransform(96, 96)
    nt, _ = tr(ims)
    got = nt.tensors.cpu().numpy()
    assert got.shape == z["batch"].shape and got.dtype == np.float32
    assert [tuple(s) for s in nt.image_sizes] == [tuple(s) for s in z["sizes"]]
    err = np.abs(got - z["batch"]).max()
    print("letterbox max abs err vs reference:", err)
    assert err <= 5e-5          # tolerance of SURVEY.md appendix A.3 (fp32 source-coordinate rounding)
    assert np.array_equal(got[2], z["batch"][2])   # identity resize: exact copy


def test_float_inputs_and_s2d_layout_agree_with_nchw():
    g = torch.Generator().manual_seed(5)
    ims = [torch.rand(3, 70, 101, generator=g), torch.rand(3, 128, 128, generator=g), torch.rand(3, 55, 40, generator=g)]
    tr = YOLOTransform(128, 128)
    ref, sizes, _ = R.letterbox(ims, 128.0, 128.0)
    dims = [im.to(DEV) for im in ims]
    nt, _ = tr(dims)
    assert np.abs(nt.tensors.cpu().numpy() - ref.numpy()).max() <= 5e-5
    geoms, (Hb, Wb) = tr.geometry(dims)
    for dt in (torch.float16, torch.bfloat16):
        s2d = torch.empty((3, Hb // 2, Wb // 2, 16), dtype=dt, device=DEV)
        tr.letterbox_into(dims, geoms, Hb, Wb, s2d, _C.YB_LAYOUT_S2D16)
        nchw = torch.empty((3, 3, Hb, Wb), dtype=dt, device=DEV)
        tr.letterbox_into(dims, geoms, Hb, Wb, nchw, _C.YB_LAYOUT_NCHW)
        # s2d[n, Y, X, (dy*2+dx)*4 + c] == nchw[n, c, 2Y+dy, 2X+dx]; channel 3 of every quad is zero
        v = s2d.view(3, Hb // 2, Wb // 2, 2, 2, 4)
        assert torch.all(v[..., 3] == 0)
        back = v[..., :3].permute(0, 5, 1, 3, 2, 4).reshape(3, 3, Hb, Wb)
        assert torch.equal(back, nchw)
        assert (nchw.float().cpu() - ref).abs().max() <= (2e-3 if dt == torch.float16 else 8e-3)


def test_mixed_batch_geometry_639_trap():
    # sizes whose long side resizes to 639 in one batch with a 640 one
    ims = [torch.randint(0, 256, (3, 800, 600), dtype=torch.uint8), torch.randint(0, 256, (3, 480, 640), dtype=torch.uint8)]
    ref, sizes, geo = R.letterbox(ims)
    nt, _ = YOLOTransform(640, 640)([im.to(DEV) for im in ims])
    assert nt.image_sizes == [tuple(s) for s in sizes] and nt.image_sizes[0] == (639, 479)
    assert np.abs(nt.tensors.cpu().numpy() - ref.numpy()).max() <= 5e-5


def test_rejects_bad_inputs():
    tr = YOLOTransform(64, 64)
    with pytest.raises(ValueError):
        tr([torch.rand(1, 3, 8, 8, device=DEV)])
    with pytest.raises(_C.NativeLibraryError):
        tr([torch.rand(3, 8, 8)])   # CPU tensor: no fallback
