"""Times single convolution launches (bs32 yolov5s shapes) with the kernels' ablation knobs (YB_CONV_DBG):
0 = full kernel, 1 = no epilogue math/stores, 2 = no MMA, 4 = no TMA stores, 3 = loads only.  Run on the GPU box."""
import os, sys
sys.path.insert(0, ".")
import torch
from yolort_b200 import _C
from yolort_b200.engine import pack_bias, pack_weight

DEV = torch.device("cuda:0")

def build(N, H, W, Cin, Cout, k, s, p, residual, nbuf):
    g = torch.Generator().manual_seed(0)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    xs = [torch.randn(N, H, W, Cin, generator=g).half().to(DEV) for _ in range(nbuf)]
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5)
    wp, ci_pad, co_pad = pack_weight(w.double(), torch.float16, DEV)
    bp = pack_bias(torch.zeros(Cout).double(), co_pad, DEV)
    outs = [torch.empty(N, Ho, Wo, Cout, dtype=torch.float16, device=DEV) for _ in range(nbuf)]
    res = torch.randn(N, Ho, Wo, Cout, generator=g).half().to(DEV) if residual else None
    plans = []
    for x, o in zip(xs, outs):
        d = _C.OpDesc()
        d.kind, d.dtype = _C.YB_OP_CONV, _C.YB_F16
        d.N, d.H, d.W, d.Cin, d.in_cstride, d.in_ = N, H, W, Cin, Cin, x.data_ptr()
        d.Ho, d.Wo, d.Cout, d.out_cstride, d.out = Ho, Wo, Cout, Cout, o.data_ptr()
        d.ksize, d.stride, d.pad, d.act = k, s, p, _C.YB_ACT_SILU
        d.weight, d.Cin_pad, d.Cout_pad, d.bias = wp.data_ptr(), ci_pad, co_pad, bp.data_ptr()
        if residual:
            d.residual, d.res_cstride = res.data_ptr(), Cout
        plans.append(_C.Plan([d], DEV))
    return plans, (xs, outs, wp, bp, res)

def timeit(plans, reps=30):
    for p in plans: p.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): plans[i % len(plans)].run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

SHAPES = [  # name, H, W, Cin, Cout, k, s, p, residual
    ("stem sp4 64->128 k3 @320x80", 320, 80, 64, 128, 3, 1, 1, False),
    ("1x1 64->64 @160", 160, 160, 64, 64, 1, 1, 0, False),
    ("3x3 32->32 @160 +res", 160, 160, 32, 32, 3, 1, 1, True),
    ("1x1 128->128 @80", 80, 80, 128, 128, 1, 1, 0, False),
    ("3x3 64->64 @80 +res", 80, 80, 64, 64, 3, 1, 1, True),
    ("3x3 64->128 s2 @160", 160, 160, 64, 128, 3, 2, 1, False),
    ("3x3 128->128 @40 +res", 40, 40, 128, 128, 3, 1, 1, True),
    ("1x1 256->256 @40", 40, 40, 256, 256, 1, 1, 0, False),
    ("3x3 256->256 @20", 20, 20, 256, 256, 3, 1, 1, False),
]
if __name__ == "__main__":
    print(f"{'layer':26s} {'variant':10s} " + " ".join(f"dbg{d:<2d}" .rjust(9) for d in (0, 1, 3, 11, 27)) + "   (us, 4 rotating buffers | same buffer for dbg0)")
    for name, H, W, Cin, Cout, k, s, p, res in SHAPES:
        variants = [("default", {})]
        if k == 3 and s == 1:
            variants = [("patch/tma", {"YB_PATCH_LOADER": "0"}), ("im2col", {"YB_DISABLE_PATCH_CONV": "1"})]
        for vname, env in variants:
            row = []
            for dbg in (0, 1, 3, 11, 27):
                os.environ["YB_CONV_DBG"] = str(dbg)
                for kk, vv in env.items(): os.environ[kk] = vv
                plans, keep = build(32, H, W, Cin, Cout, k, s, p, res, 4)
                row.append(timeit(plans))
                if dbg == 0:
                    same = timeit(plans[:1])
                for kk in env: os.environ.pop(kk)
                del plans, keep
            print(f"{name:26s} {vname:10s} " + " ".join(f"{t:9.1f}" for t in row) + f"   | {same:9.1f}")
    os.environ.pop("YB_CONV_DBG", None)
