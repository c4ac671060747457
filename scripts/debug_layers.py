"""Find the first backbone layer whose output departs from the fp32 oracle (run on the GPU box):
python scripts/debug_layers.py <n|s|m|l|x> [H W]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import parity_util as util
from oracle import restate as R
from yolort_b200 import models

name = sys.argv[1] if len(sys.argv) > 1 else "l"
H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (192, 192)
dev = "cuda:0"
sd = util.synth_state_dict(util.layouts()[name], knob_obj=7.0, knob_cls=4.5, seed=1, gain={"m":1.7,"l":1.7,"x":1.5}.get(name,2.0))
m = getattr(models, f"yolov5{name}")(size=(H, W), score_thresh=0.2).eval(); m.load_state_dict(sd); m = m.to(dev)
g = torch.Generator().manual_seed(0)
x = torch.rand(2, 3, H, W, generator=g)
m.model(x.to(dev)); plan = m.model.get_plan(2, H, W); m.model.run_plan(plan); torch.cuda.synchronize()
net = R.Net(sd)
c3, c4, c5 = m.model.backbone.out_channels
with torch.no_grad():
    t = x.half().float(); outs = {}
    for i in range(9):
        t = net.conv(t, f"backbone.body.{i}") if i in (0, 1, 3, 5, 7) else net.c3(t, f"backbone.body.{i}", True)
        outs[i] = t
def got(i):
    if i == 4: return plan.buffers["pan.cat2[up(lat2)|f4]"][..., c3:]
    if i == 6: return plan.buffers["pan.cat1[up(lat1)|f6]"][..., c4:]
    return plan.buffers[f"body.{i}"]
for i in range(9):
    a = got(i).float().permute(0, 3, 1, 2).cpu().numpy(); b = outs[i].numpy()
    err = np.abs(a - b); rr = np.sqrt((err ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-12)
    print(f"body.{i}: shape {b.shape} rel_rms {rr:.3e} max {err.max():.4f}")
for op_i, nm in enumerate(plan.op_names): pass
print("ops:", len(plan.op_names))
