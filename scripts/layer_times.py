"""Per-op CUDA-event times of one plan (default yolov5s batch 32 640x640 fp16), full-plan passes with an event pair
around every op (so each op sees the cache state the real step gives it).  python scripts/layer_times.py [model] [batch] [size] [reps] [dtype]"""
import os, sys, time
sys.path.insert(0, ".")
import torch
import yolort_b200.models as M

name = sys.argv[1] if len(sys.argv) > 1 else "yolov5s"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
size = int(sys.argv[3]) if len(sys.argv) > 3 else 640
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dt = sys.argv[5] if len(sys.argv) > 5 else "f16"
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = getattr(M, name)(score_thresh=0.25, size=(size, size)).eval().to(dev)
if dt == "bf16":
    m = m.to(torch.bfloat16)
plan = m.model.get_plan(batch, size, size)
n = plan.plan.n_ops
t_end = time.time() + 1.0
while time.time() < t_end:
    plan.run()
torch.cuda.synchronize()
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(reps)]
for r in range(reps):
    ev[r][0].record()
    for i in range(n):
        plan.run(i, 1)
        ev[r][i + 1].record()
torch.cuda.synchronize()
tot = 0.0
print(f"# {name} batch {batch} {size}x{size} {dt}: per-op us (median of {reps} passes, events between ops)")
for i in range(n):
    ts = sorted(ev[r][i].elapsed_time(ev[r][i + 1]) * 1e3 for r in range(reps))
    t = ts[len(ts) // 2]
    tot += t
    fl = plan.op_flops[i]
    print(f"{i:3d} {t:8.1f} us  {fl / t / 1e6 if fl else 0:7.1f} TF/s  {plan.op_names[i]}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.run()
e1.record()
torch.cuda.synchronize()
print(f"# sum of ops {tot:.1f} us; plan back-to-back {e0.elapsed_time(e1) / reps * 1e3:.1f} us; total GFLOP {sum(plan.op_flops) / 1e9:.1f}")
