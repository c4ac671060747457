"""CUDA-event times of the stages of one yolov5s batch-32 640x640 step (letterbox | plan | begin | decode | nms), the GPU
kept busy while the host enqueues; plus wall-clock of predict() on pinned host images (chunk-pipelined vs plain)."""
import sys, time
sys.path.insert(0, ".")
import torch
import bench
from yolort_b200 import _C
from yolort_b200.models import yolov5s

dev = torch.device("cuda:0")
m = yolov5s(score_thresh=0.25).eval()
m.load_state_dict(bench.make_state_dict(m))
m = m.to(dev)
host = torch.stack(bench.make_images(32, 1234)).pin_memory()
ims = [t for t in host.to(dev)]
for _ in range(30):
    m.forward_padded(ims)
torch.cuda.synchronize()
names = ("letterbox", "plan", "begin", "decode", "nms")
acc = {k: [] for k in names}
marks = []
def mark(_=None):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(e)
pc = m.model.post_config()
for i in range(22):
    marks.clear()
    torch.cuda._sleep(4_000_000)
    mark()
    plan, rescale = m._prepare(ims)
    mark()
    plan.run()
    mark()
    _C.decode_nms_padded(plan.heads, "nhwc", pc["strides"], pc["anchors_px"], pc["num_classes"], pc["score_thresh"],
                         pc["nms_thresh"], pc["detections_per_img"], pc["semantics"], rescale, stage_hook=mark)
    torch.cuda.synchronize()
    if i >= 2:
        for k, n in enumerate(names):
            acc[n].append(marks[k].elapsed_time(marks[k + 1]) * 1e3)
print("stage us:", {n: round(sorted(v)[len(v) // 2], 1) for n, v in acc.items()})
# predict() wall clock
hl = [host[j] for j in range(32)]
def wall(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
t_pipe = wall(lambda: m.predict(hl))
orig = m._predict_pipelined
m._predict_pipelined = lambda x: None
t_plain = wall(lambda: m.predict(hl))
m._predict_pipelined = orig
t_dev = wall(lambda: m(ims))
def enqueue_only():
    t0 = time.perf_counter(); m.forward_padded(ims); return (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize(); enq = sorted(enqueue_only() for _ in range(20))[10]; torch.cuda.synchronize()
print(f"predict(host) ms: pipelined {t_pipe:.3f}  plain {t_plain:.3f}  | forward(device list) {t_dev:.3f} | host enqueue time of forward_padded {enq:.3f}")
