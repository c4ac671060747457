"""Per-launch breakdown of an ncu `gpu__time_duration` CSV of one bench step (yolov5s bs32 640)."""
import csv, re, sys
sys.path.insert(0, '.')
import torch
from yolort_b200.engine import lower_yolo
from yolort_b200.models import yolov5s

path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, r in enumerate(rows) if 'letterbox' in r['Kernel Name']]
a, b = idx[0], idx[1]
L, _, _, _ = lower_yolo(yolov5s().model, torch.float16, torch.device('cpu'))
ops = L.ops
k = 0
tot = 0
alltot = 0
for r in rows[a:b]:
    name = re.sub(r'\(.*', '', r['Kernel Name']).split('::')[-1][:28]
    t = float(r['Metric Value'].replace(',', '')) / 1e3
    alltot += t
    extra = ''
    if 'conv_umma' in name or 'conv3x3_patch' in name or 'spp_pool' in name or 'upsample' in name:
        op = ops[k]; k += 1
        hw = 640 // op.dst.buf.div
        if op.kind == 0:
            fl = 32 * hw * hw * op.flops_per_pixel // op.pack
            byts = 32 * ((640 // op.src.buf.div) ** 2 * op.src.C + hw * hw * op.dst.C) * 2
            extra = f"{op.name:32s} {op.src.C:4d}->{op.dst.C:4d} k{op.ksize}s{op.stride} out{hw:3d}  {fl/t/1e6:7.1f} TF/s {byts/t/1e3:7.1f} GB/s"
            tot += t
        else:
            extra = op.name
    print(f"{t:8.1f} us  grid {r['Grid Size']:>14s} {name:28s} {extra}")
print('conv total us', round(tot, 1), ' step total us', round(alltot, 1))
