"""Generate tests/golden/* by running the UNMODIFIED reference (imported from /root/reference).

Run in the development container only:  python -m oracle.make_golden
The fixtures pin oracle/restate.py (tests/test_oracle_golden.py) and, through it, the CUDA path.
TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import zlib

import numpy as np
import torch

from . import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ---- deterministic, construction-order-independent parameters ------------------------------------------
def synth_state_dict(shapes: dict, knob_obj: float = 0.0, knob_cls: float = 0.0, seed: int = 0,
                     gain: float = 2.0) -> dict:
    """Weights as a pure function of (key name, shape, seed): conv ~ N(0, gain/fan_in) (the deeper
    m/l/x residual chains need gain < 2 to keep activations inside fp16 range, as trained weights do); BN statistics
    randomised so that folding is exercised; head bias = the constructor's
    prior (box_head.py:40-46) plus a "load knob" that raises objectness / class logits."""
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) & 0x7FFFFFFF)
        if k.endswith("num_batches_tracked"):
            t = torch.zeros(shp, dtype=torch.int64)
        elif k.endswith("conv.weight") or (k.startswith("model.head") and k.endswith(".weight")) or \
                (len(shp) == 4 and k.endswith(".weight")):      # bare cv2 / cv3 of the r3.1 BottleneckCSP
            fan_in = shp[1] * shp[2] * shp[3]
            t = torch.randn(shp, generator=g) * (gain / fan_in) ** 0.5
        elif k.endswith("bn.weight") or k.endswith("running_var"):
            t = torch.rand(shp, generator=g) + 0.5
        elif k.endswith("bn.bias") or k.endswith("running_mean"):
            t = torch.randn(shp, generator=g) * 0.1
        elif k.startswith("model.head") and k.endswith(".bias"):
            lvl = int(k.split(".")[3])
            stride = (8, 16, 32, 64)[lvl]
            na = 3
            b = torch.randn(shp, generator=g).view(na, -1) * 0.1
            b[:, 4] += np.log(8 / (640 / stride) ** 2) + knob_obj
            b[:, 5:] += np.log(0.6 / (b.shape[1] - 5 - 0.999999)) + knob_cls
            t = b.reshape(-1)
        else:
            raise KeyError(k)
        sd[k] = t
    return sd


def checksum(sd: dict) -> float:
    return float(sum(v.double().abs().sum().item() * ((i % 7) + 1) for i, (k, v) in enumerate(sorted(sd.items()))))


def synth_image_u8(h: int, w: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)


def main():
    yolort = ref_import.import_reference()
    from yolort.models import YOLOv5, yolov5l, yolov5m, yolov5n, yolov5s
    from yolort.models.anchor_utils import AnchorGenerator
    from yolort.models.box_head import PostProcess
    from yolort.models.transform import scale_coords, YOLOTransform

    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)

    # 1. state-dict layouts ---------------------------------------------------------------------------
    layouts = {}
    for name, ctor in (("n", yolov5n), ("s", yolov5s), ("m", yolov5m), ("l", yolov5l)):
        layouts[name] = {k: list(v.shape) for k, v in ctor().state_dict().items()}
    layouts["x"] = {k: list(v.shape) for k, v in YOLOv5(arch="yolov5_darknet_pan_x_r60").state_dict().items()}
    with open(os.path.join(GOLDEN, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f, separators=(",", ":"))

    # 2. letterbox geometry over many sizes + scale_coords -------------------------------------------------
    tr = YOLOTransform(640, 640)
    rng = np.random.RandomState(4321)
    sizes = [(500, 375), (768, 1000), (800, 600), (417, 523), (640, 640), (1280, 1280), (416, 416), (1080, 1920),
             (950, 950), (720, 1280), (333, 1000), (37, 53)]
    sizes += [(int(a), int(b)) for a, b in rng.randint(416, 1281, size=(120, 2))]
    geo = []
    for h, w in sizes:
        out, _ = tr.resize(torch.zeros(3, h, w))
        geo.append((h, w, int(out.shape[1]), int(out.shape[2])))
    batches = []
    for bi in range(12):
        idx = rng.choice(len(sizes), size=rng.randint(1, 6), replace=False)
        ims = [torch.zeros(3, geo[i][2], geo[i][3]) for i in idx]
        for j, im in enumerate(ims):
            im[:, 0, 0] = 1.0 + j  # marker at the top-left corner of each pasted image
        bt = tr.batch_images(ims)
        offs = []
        for j in range(len(ims)):
            pos = (bt[j, 0] == 1.0 + j).nonzero()[0]
            offs.append((int(pos[0]), int(pos[1])))
        Hb, Wb = int(bt.shape[2]), int(bt.shape[3])
        boxes = torch.tensor([[10.0, 20.0, 300.5, 400.25], [0.0, 0.0, float(Wb), float(Hb)]])
        sc = [scale_coords(boxes, torch.tensor([Hb, Wb]), (geo[i][0], geo[i][1])).numpy() for i in idx]
        batches.append({"idx": [int(i) for i in idx], "Hb": Hb, "Wb": Wb, "offsets": offs,
                        "scaled": [s.tolist() for s in sc]})
    with open(os.path.join(GOLDEN, "letterbox_geometry.json"), "w") as f:
        json.dump({"sizes": geo, "batches": batches, "probe_boxes": [[10.0, 20.0, 300.5, 400.25], "full"]}, f,
                  separators=(",", ":"))

    # 3. letterbox pixels on small images (size=(96,96)) -----------------------------------------------------
    tr_small = YOLOTransform(96, 96)
    small = [synth_image_u8(40, 61, 11), synth_image_u8(75, 50, 12), synth_image_u8(96, 96, 13), synth_image_u8(131, 97, 14)]
    nt, _ = tr_small([im / 255.0 for im in small])
    np.savez_compressed(os.path.join(GOLDEN, "letterbox_pixels.npz"),
                        batch=nt.tensors.numpy(), sizes=np.array(nt.image_sizes),
                        **{f"img{i}": im.numpy() for i, im in enumerate(small)})

    # 4. network: yolov5n, synthetic weights, one 96x128 input -------------------------------------------------
    sd = synth_state_dict(layouts["n"], knob_obj=7.0, knob_cls=4.5, seed=0)
    m = yolov5n(size=(128, 128), score_thresh=0.15).eval()
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(99)
    x = torch.rand(1, 3, 96, 128, generator=g)
    with torch.no_grad():
        feats = m.model.backbone(x)
        heads = m.model.head(feats)
        dets = m.model(x)
    np.savez_compressed(os.path.join(GOLDEN, "network_n.npz"), x=x.numpy(), checksum=np.float64(checksum(sd)),
                        p3=feats[0].numpy(), p4=feats[1].numpy(), p5=feats[2].numpy(),
                        h0=heads[0].numpy(), h1=heads[1].numpy(), h2=heads[2].numpy(),
                        **{f"det{i}_{k}": v.numpy() for i, d in enumerate(dets) for k, v in d.items()})

    # 5. post-process on synthetic head logits: both torchvision branches ------------------------------------------
    pp_cases = {}
    ag = AnchorGenerator([8, 16, 32], [[10, 13, 16, 30, 33, 23], [30, 61, 62, 45, 59, 119], [116, 90, 156, 198, 373, 326]])
    for name, (hw, mu, thr, seed) in {"few": ((64, 96), -3.0, 0.25, 1), "trick": ((96, 96), -2.2, 0.25, 2),
                                     "vanilla": ((96, 128), -2.0, 0.1, 3), "empty": ((64, 64), -9.0, 0.25, 4)}.items():
        g = torch.Generator().manual_seed(seed)
        H, W = hw
        hs = [torch.randn(2, 3, H // s, W // s, 85, generator=g) * 1.5 + mu for s in (8, 16, 32)]
        grids, shifts = ag([torch.zeros(1, 1, H // s, W // s) for s in (8, 16, 32)])
        out = PostProcess([8, 16, 32], thr, 0.45, 300)(hs, grids, shifts)
        case = {f"h{i}": h.numpy() for i, h in enumerate(hs)}
        case["thr"] = np.float32(thr)
        for i, d in enumerate(out):
            for k, v in d.items():
                case[f"det{i}_{k}"] = v.numpy()
        pp_cases[name] = case
        print(name, "detections", [len(d["scores"]) for d in out])
        np.savez_compressed(os.path.join(GOLDEN, f"postprocess_{name}.npz"), **case)

    # 6. end to end: yolov5n, two uint8-derived images of different sizes, size=(128,128) ---------------------------------
    ims = [synth_image_u8(90, 128, 21), synth_image_u8(100, 75, 22)]
    with torch.no_grad():
        out = m([im / 255.0 for im in ims])
    np.savez_compressed(os.path.join(GOLDEN, "e2e_n.npz"), img0=ims[0].numpy(), img1=ims[1].numpy(),
                        **{f"det{i}_{k}": v.numpy() for i, d in enumerate(out) for k, v in d.items()})
    print("golden fixtures written to", GOLDEN)
    for fn in sorted(os.listdir(GOLDEN)):
        print(f"  {fn}: {os.path.getsize(os.path.join(GOLDEN, fn)) / 1024:.1f} KiB")
    print("e2e detections:", [len(d["scores"]) for d in out], "network dets:", [len(d["scores"]) for d in dets])


if __name__ == "__main__":
    main()
