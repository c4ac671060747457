"""N device-resident steps of yolov5s batch 32 640x640 (for ncu captures).  python scripts/one_step.py [steps]"""
import sys
sys.path.insert(0, ".")
import torch
from bench import make_state_dict, make_images
from yolort_b200.models import yolov5s

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
m = yolov5s(score_thresh=0.25).eval()
m.load_state_dict(make_state_dict(m))
m = m.to(dev)
b = torch.stack(make_images(32, 1234)).to(dev)
for i in range(steps):
    out = m.forward_padded([b[j] for j in range(32)])
torch.cuda.synchronize()
print("status", out[4].cpu().tolist())
