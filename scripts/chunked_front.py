"""Does running the front of the plan (stem .. first tapped C3) per image chunk help on device-resident batches?
Smaller tensors stay in the 126 MB L2 between layers, at the price of 4x the launches of the front."""
import sys
sys.path.insert(0, ".")
import torch
from yolort_b200.models import yolov5s
dev = torch.device("cuda:0")
m = yolov5s(score_thresh=0.25).eval().to(dev)
full = m.model.get_plan(32, 640, 640)
ch = m.model.get_plan(32, 640, 640, chunked=True)
print("front launches", ch.front_ops, "chunks", ch.front_chunks, "launches", ch.plan.n_ops)
def timed(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def run_chunked():
    for k in range(ch.front_chunks):
        ch.run_front_chunk(k)
    ch.run_rest()
for rnd in range(3):
    print(f"plan ms: whole batch {timed(full.run):.3f} | front in {ch.front_chunks} chunks {timed(run_chunked):.3f} | "
          f"front only: whole {timed(lambda: full.plan.run(0, full.front_ops if full.front_ops else ch.front_ops)):.3f} "
          f"chunked {timed(lambda: [ch.run_front_chunk(k) for k in range(ch.front_chunks)]):.3f}")
