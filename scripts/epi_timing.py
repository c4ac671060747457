"""Epilogue phase clocks of the 1x1 / im2col kernel (instrumented build: make -C yolort_b200/csrc variant NAME=epitime
EXTRA=-DYB_EPI_TIMING; run with YB_LIB_PATH=scratch/lib_epitime.so).  Per selected plan op: clock64 totals of CTA 0 /
epilogue group 0 over the op's tiles: loop | set-up | accumulator wait | box compute | fences | store-read wait | barrier |
store issue."""
import ctypes, os, sys
sys.path.insert(0, ".")
import torch
import yolort_b200.models as M
from yolort_b200 import _C

dev = torch.device("cuda:0")
m = M.yolov5s(score_thresh=0.25).eval().to(dev)
os.environ["YB_NO_CHAIN"] = os.environ.get("YB_NO_CHAIN", "1")
m.model.engine().fuse_chains = os.environ["YB_NO_CHAIN"] != "1"
plan = m.model.get_plan(32, 640, 640)
for _ in range(5):
    plan.run()
torch.cuda.synchronize()
lib = _C.lib()
buf = (ctypes.c_ulonglong * 16)()
names = ("loop", "setup", "acc_wait", "box", "fence", "st_read", "barrier", "st_issue")
for i, nm in enumerate(plan.op_names):
    if plan.op_flops[i] == 0:
        continue
    lib.yb_debug_epi_ticks(None, 1)
    plan.run(i, 1)
    torch.cuda.synchronize()
    lib.yb_debug_epi_ticks(buf, 0)
    t = [int(v) for v in buf[:8]]
    tot = sum(t)
    if tot == 0:
        print(f"{i:3d} {nm[:44]:44s} (halo-patch kernel: not instrumented)")
        continue
    print(f"{i:3d} {nm[:44]:44s} total {tot:8d} clk | " + " ".join(f"{n} {100 * v / tot:4.1f}%" for n, v in zip(names, t)))
