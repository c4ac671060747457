"""Per-launch table of one step from an ncu CSV (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`,
kernels of namespace yb only, one whole step: letterbox .. nms).  python scripts/analyze_launches.py <csv> [out.json]
Prints time, achieved TFLOP/s (conv launches) and DRAM bytes per launch; writes the conv DRAM traffic of the step."""
import csv, json, re, sys
sys.path.insert(0, '.')
import torch
from yolort_b200.engine import lower_yolo
from yolort_b200.models import yolov5s

path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
launches = {}
for r in rows:
    d = launches.setdefault(int(r['ID']), {'name': re.sub(r'\(.*', '', r['Kernel Name']).split('::')[-1][:30], 'grid': r['Grid Size']})
    v = float(r['Metric Value'].replace(',', ''))
    u = r['Metric Unit']
    if 'time' in r['Metric Name']:
        d['us'] = v * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1e-3)
    else:
        d['rd' if 'read' in r['Metric Name'] else 'wr'] = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
order = [launches[k] for k in sorted(launches)]
start = next(i for i, d in enumerate(order) if 'letterbox' in d['name'])
order = order[start:]
L, _, _, _ = lower_yolo(yolov5s().model, torch.float16, torch.device('cpu'))
ops, k = L.ops, 0
tot = conv_t = conv_rd = conv_wr = 0.0
for d in order:
    extra = ''
    if any(s in d['name'] for s in ('conv_umma', 'conv3x3_patch', 'spp_pool', 'upsample2x')) and k < len(ops):
        op = ops[k]; k += 1
        hw = 640 // op.dst.buf.div
        if op.kind == 0:
            fl = 32 * hw * hw * op.flops_per_pixel // op.pack
            extra = f"{op.name[:34]:34s} {op.src.C:4d}->{op.dst.C:4d} k{op.ksize}s{op.stride} out{hw:3d} {fl / d['us'] / 1e6:7.1f} TF/s"
            conv_t += d['us']; conv_rd += d.get('rd', 0); conv_wr += d.get('wr', 0)
        else:
            extra = op.name
    tot += d['us']
    print(f"{d['us']:8.1f} us  rd {d.get('rd', 0) / 1e6:7.1f} MB wr {d.get('wr', 0) / 1e6:7.1f} MB  {d['name']:30s} {extra}")
    if 'nms_image' in d['name']:
        break
print(f"conv launches: {conv_t:.1f} us, DRAM read {conv_rd / 1e9:.3f} GB + write {conv_wr / 1e9:.3f} GB; whole step {tot:.1f} us (serialised, cold caches)")
if len(sys.argv) > 2:
    json.dump({"dram_bytes_per_step": conv_rd + conv_wr, "dram_read": conv_rd, "dram_write": conv_wr, "conv_us_serialised": conv_t,
               "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, one step of yolov5s batch 32 640x640"},
              open(sys.argv[2], 'w'), indent=1)
