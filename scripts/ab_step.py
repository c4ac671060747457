"""Device time of one full yolov5s batch-32 640x640 step (letterbox -> plan -> decode+NMS) and of the plan alone,
for the library named by YB_LIB_PATH.  python scripts/ab_step.py [reps]"""
import os, sys, time
sys.path.insert(0, ".")
import torch
from yolort_b200.models import yolov5s

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = yolov5s(score_thresh=0.25).eval().to(dev)
ims = [[torch.randint(0, 256, (3, 640, 640), dtype=torch.uint8, device=dev) for _ in range(32)] for _ in range(4)]
t_end = time.time() + 1.5
while time.time() < t_end:          # clock ramp + caches
    m.forward_padded(ims[0])
torch.cuda.synchronize()
plan = m.model.get_plan(32, 640, 640)
def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = []
for rnd in range(3):
    step = timed(lambda i: m.forward_padded(ims[i % 4]), reps)
    pl = timed(lambda i: plan.run(), reps)
    res.append((step, pl))
tag = os.environ.get("YB_LIB_PATH", "default")
print(f"{os.path.basename(tag):34s} step ms {' '.join(f'{a:.3f}' for a, _ in res)} | plan ms {' '.join(f'{b:.3f}' for _, b in res)}")
