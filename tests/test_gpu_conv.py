"""tcgen05 implicit-GEMM convolution vs a plain fp32 PyTorch conv on the same fp16/bf16-rounded operands (B200)."""
import pytest
import torch
import torch.nn.functional as F

from yolort_b200 import _C
from yolort_b200.engine import pack_bias, pack_weight

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def run_conv(N, H, W, Cin, Cout, k, s, p, dtype=torch.float16, act=True, residual=False, in_pad=0, out_pad=0, seed=0,
             bias_scale=0.5):
    g = torch.Generator().manual_seed(seed)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    in_cs, out_cs = Cin + in_pad, Cout + out_pad
    x_full = (torch.randn(N, H, W, in_cs, generator=g)).to(dtype).to(DEV)
    in_off = in_pad // 2 // 8 * 8
    out_off = out_pad // 2 // 8 * 8
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5).to(dtype)
    b = torch.randn(Cout, generator=g) * bias_scale
    wp, ci_pad, co_pad = pack_weight(w.double(), dtype, DEV)
    bp = pack_bias(b.double(), co_pad, DEV)
    out_full = torch.full((N, Ho, Wo, out_cs), 7.0, dtype=dtype, device=DEV)
    res = (torch.randn(N, Ho, Wo, Cout, generator=g)).to(dtype).to(DEV) if residual else None
    d = _C.OpDesc()
    d.kind, d.dtype = _C.YB_OP_CONV, _C.dtype_code(dtype)
    d.N, d.H, d.W, d.Cin, d.in_cstride = N, H, W, Cin, in_cs
    d.in_ = x_full.data_ptr() + in_off * 2
    d.Ho, d.Wo, d.Cout, d.out_cstride = Ho, Wo, Cout, out_cs
    d.out = out_full.data_ptr() + out_off * 2
    act_code = {True: _C.YB_ACT_SILU, False: _C.YB_ACT_NONE, "hardswish": _C.YB_ACT_HARDSWISH,
                "leaky": _C.YB_ACT_LEAKY01}[act]
    d.ksize, d.stride, d.pad, d.act = k, s, p, act_code
    d.weight, d.Cin_pad, d.Cout_pad, d.bias = wp.data_ptr(), ci_pad, co_pad, bp.data_ptr()
    if residual:
        d.residual, d.res_cstride = res.data_ptr(), Cout
    plan = _C.Plan([d], DEV)
    plan.run()
    torch.cuda.synchronize()
    x = x_full[..., in_off:in_off + Cin].float().permute(0, 3, 1, 2)
    ref = F.conv2d(x, w.float().to(DEV), b.to(DEV), s, p)
    if act == "hardswish":
        ref = F.hardswish(ref)
    elif act == "leaky":
        ref = F.leaky_relu(ref, 0.1)
    elif act:
        ref = F.silu(ref)
    if residual:
        ref = ref + res.float().permute(0, 3, 1, 2)
    got = out_full[..., out_off:out_off + Cout].float().permute(0, 3, 1, 2)
    # channels outside the destination window must be untouched
    if out_pad:
        mask = torch.ones(out_cs, dtype=torch.bool)
        mask[out_off:out_off + Cout] = False
        assert torch.all(out_full[..., mask.to(DEV)] == 7.0)
    err = (got - ref).abs()
    tol = (2.0 ** -9 if dtype == torch.float16 else 2.0 ** -6)   # SURVEY.md section 8c stage-wise bound
    bound = tol * (1.0 + ref.abs())
    bad = (err > bound).sum().item()
    print(f"conv N{N} {H}x{W} {Cin}->{Cout} k{k}s{s} {dtype}: max_abs_err={err.max().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.2f} violations={bad}/{err.numel()}")
    if bad:
        idx = (err > bound).nonzero()[:8]
        print("first violations (n,c,y,x):", idx.tolist())
        print("got", got[tuple(idx[0])].item(), "ref", ref[tuple(idx[0])].item())
    assert bad == 0


@pytest.mark.parametrize("cin,cout", [(64, 128), (128, 64), (32, 32), (16, 32), (256, 256), (512, 256), (48, 96), (80, 160)])
def test_conv1x1(cin, cout):
    run_conv(2, 20, 20, cin, cout, 1, 1, 0)


def test_conv1x1_ragged_m_and_channel_windows():
    run_conv(3, 13, 11, 64, 64, 1, 1, 0, in_pad=64, out_pad=32)       # M = 429, not a tile multiple; sliced views


@pytest.mark.parametrize("cin,cout,s", [(64, 64, 1), (32, 64, 2), (16, 32, 1), (128, 128, 2), (256, 256, 1), (48, 48, 1)])
def test_conv3x3(cin, cout, s):
    run_conv(2, 24, 40, cin, cout, 3, s, 1)


def test_conv3x3_crosses_image_boundaries():
    run_conv(5, 6, 10, 64, 64, 3, 1, 1)      # 60 pixels/image: every 128-pixel tile spans 2-3 images
    run_conv(4, 10, 6, 32, 64, 3, 2, 1)      # stride 2, Wo=3


def test_bottleneck_residual_and_inplace_window():
    run_conv(2, 20, 20, 64, 64, 3, 1, 1, residual=True, out_pad=64)


def test_head_conv_bias_no_activation():
    run_conv(2, 20, 20, 128, 256, 1, 1, 0, act=False)


def test_wide_output_splits_into_n_tiles():
    run_conv(1, 16, 16, 64, 320, 1, 1, 0)    # block_n = 160 x 2 tiles
    run_conv(1, 16, 16, 128, 512, 3, 1, 1)   # block_n = 256 x 2 tiles


def test_bf16():
    run_conv(2, 20, 20, 64, 128, 1, 1, 0, dtype=torch.bfloat16)
    run_conv(2, 20, 20, 64, 64, 3, 2, 1, dtype=torch.bfloat16)


def test_large_m_many_tiles():
    run_conv(8, 80, 80, 64, 64, 3, 1, 1)     # 51 200 pixels -> 400 CTAs


def test_rejects_unsupported():
    d = _C.OpDesc()
    d.kind = 99
    with pytest.raises(_C.NativeLibraryError):
        _C.Plan([d], DEV)


# ---- halo-patch 3x3 kernel (conv3x3_patch_sm100.cu) -------------------------------------------------------
import os


@pytest.mark.parametrize("mode", ["0", "2"])
@pytest.mark.parametrize("cin,cout", [(64, 64), (16, 32), (32, 32), (128, 128), (256, 256)])
def test_patch_conv_view_modes(cin, cout, mode, monkeypatch):
    """All three ways of addressing the taps inside the staged patch must give the same convolution."""
    monkeypatch.setenv("YB_PATCH_MODE", mode)
    run_conv(2, 32, 40, cin, cout, 3, 1, 1, seed=3)


def test_patch_conv_ragged_edges_residual_and_windows():
    run_conv(3, 40, 44, 64, 64, 3, 1, 1, residual=True, in_pad=32, out_pad=64)   # W=44: last column tile half empty
    run_conv(1, 48, 24, 128, 128, 3, 1, 1, dtype=torch.bfloat16)


def test_patch_conv_matches_im2col_kernel(monkeypatch):
    """Same layer through both kernels (the generic im2col path is forced with YB_DISABLE_PATCH_CONV=1)."""
    monkeypatch.setenv("YB_DISABLE_PATCH_CONV", "1")
    run_conv(2, 32, 32, 64, 64, 3, 1, 1, seed=11)
    monkeypatch.delenv("YB_DISABLE_PATCH_CONV")
    run_conv(2, 32, 32, 64, 64, 3, 1, 1, seed=11)


@pytest.mark.parametrize("act", ["hardswish", "leaky"])
def test_r31_activations(act):
    """Hardswish (r3.1 Conv) and LeakyReLU(0.1) (BottleneckCSP) epilogues, on both kernels (1x1 generic, 3x3 patch)."""
    run_conv(2, 24, 40, 64, 64, 1, 1, 0, act=act, bias_scale=2.0)
    run_conv(2, 32, 32, 32, 64, 3, 1, 1, act=act, residual=(act == "hardswish"), bias_scale=2.0)
    run_conv(1, 20, 28, 48, 96, 3, 2, 1, act=act, dtype=torch.bfloat16)
