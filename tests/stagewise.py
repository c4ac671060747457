"""Stage-wise parity at REAL shapes: every launch of a plan against a plain fp32 PyTorch op applied to the launch's
OWN (fp16/bf16) input buffer -- "each conv block vs fp32 on the same rounded inputs/weights: |err| <= 2^-9 (fp16) /
2^-6 (bf16) x (1 + |ref|)" (SURVEY.md 8c.1).  The fp32 reference runs on the GPU with TF32 disabled."""
import torch
import torch.nn.functional as F

from yolort_b200 import _C
from yolort_b200.engine import fold_conv_bn

TOL = {torch.float16: 2.0 ** -9, torch.bfloat16: 2.0 ** -6}


def _nchw(t):
    return t.float().permute(0, 3, 1, 2)


def _act(y, code):
    if code == _C.YB_ACT_SILU:
        return F.silu(y)
    if code == _C.YB_ACT_HARDSWISH:
        return F.hardswish(y)
    if code == _C.YB_ACT_LEAKY01:
        return F.leaky_relu(y, 0.1)
    return y


def check_plan_stagewise(model_yolo, plan, verbose=True):
    """`plan` must have been created with keep_intermediates=True and its input canvas written.  Launches the plan ONE
    op at a time and checks each launch right after it ran (C3 blocks overwrite channel windows in place -- the last
    bottleneck writes over cv1's half of the concat buffer, its own residual -- so inputs are only valid at that
    moment; the residual is snapshotted before the launch).  Returns [(name, violations, max_err)]."""
    assert plan.keep_intermediates
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    L = plan._low.L
    tol = TOL[plan.dtype]
    out = []
    # A launch covers one op, or two when a 1x1 convolution rides as the chained tail of its predecessor
    # (PlanInstance.launch_ops).  keep_intermediates plans store the first output of a fused launch too, so the tail is
    # checked against fp32 applied to exactly the fp16 tile it consumed on chip.
    for li, grp in enumerate(plan.launch_ops):
        snaps = {}
        for i in grp:
            op = L.ops[i]
            if op.residual is not None:
                snaps[i] = plan.buffers[op.residual.buf.name][..., op.residual.ch0: op.residual.ch0 + op.residual.C].clone()
        plan.run(li, 1)
        torch.cuda.synchronize()
        for i in grp:
            _check_op(model_yolo, plan, L.ops[i], snaps.get(i), tol, verbose, out, fused=len(grp) > 1)
    return out


def _check_op(model_yolo, plan, op, res_snapshot, tol, verbose, out, fused=False):
    if True:
        src = plan.buffers[op.src.buf.name][..., op.src.ch0: op.src.ch0 + op.src.C]
        dst = plan.buffers[op.dst.buf.name][..., op.dst.ch0: op.dst.ch0 + op.dst.C]
        got = _nchw(dst)
        if op.kind == _C.YB_OP_SPP_POOL:
            x = _nchw(src)
            p1 = F.max_pool2d(x, 5, 1, 2)
            p2 = F.max_pool2d(x, 9, 1, 4)
            p3 = F.max_pool2d(x, 13, 1, 6)
            ref = torch.cat([p1, p2, p3], 1)
        elif op.kind == _C.YB_OP_UPSAMPLE2X:
            ref = F.interpolate(_nchw(src), scale_factor=2.0, mode="nearest")
        elif op.pack > 1:
            # the stem: compare with the module's own 6x6/s2/p2 convolution over the un-space-to-depth'ed canvas
            stem = model_yolo.backbone.body["0"]
            w, b = fold_conv_bn(stem)
            s2d = plan.input.float()                                     # [N, H/2, W/2, 16], channel (dy*2+dx)*4 + c
            n, h2, w2, _ = s2d.shape
            x = s2d.view(n, h2, w2, 2, 2, 4)[..., :3].permute(0, 5, 1, 3, 2, 4).reshape(n, 3, 2 * h2, 2 * w2)
            ref = _act(F.conv2d(x, w.to(plan.dtype).float(), b.float(), 2, 2), op.act)
        else:
            co, ci, k = op.dst.C, op.src.C, op.ksize
            w = op.weight[:co, :, :ci].float().view(co, k, k, ci).permute(0, 3, 1, 2).contiguous()
            ref = _act(F.conv2d(_nchw(src), w, op.bias[:co], op.stride, op.pad), op.act)
            if res_snapshot is not None:
                ref = ref + _nchw(res_snapshot)
        err = (got - ref).abs()
        bad = int((err > tol * (1.0 + ref.abs())).sum().item())
        mx = float(err.max().item())
        if verbose and (bad or mx > 0.05):
            print(f"  stage {op.name}: violations {bad}/{err.numel()} max_abs_err {mx:.3e} ref_absmax {float(ref.abs().max()):.2f}")
        out.append((op.name + (" [fused launch]" if fused else ""), bad, mx))
        del ref, err, got
