import sys; sys.path.insert(0, ".")
import torch, bench
from yolort_b200 import _C
from yolort_b200.models import yolov5s
dev = torch.device("cuda:0")
m = yolov5s(score_thresh=0.25).eval(); m.load_state_dict(bench.make_state_dict(m)); m = m.to(dev)
ims = [im.to(dev) for im in bench.make_images(32, 1234)]
for _ in range(3):
    out = m.forward_padded(ims)
torch.cuda.synchronize()
print("nms phase clocks (sort, phaseA, compact, bitmatrix, resolve, rest):", _C.nms_phase_clocks(dev), "status", out[4].tolist())
