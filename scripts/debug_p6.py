"""Localise an end-to-end mismatch of the P6 fixture: model output vs (a) the oracle's post-process applied to the
GPU's OWN head logits + its own rescale, (b) the reference fixture."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import parity_util as util
from oracle import restate as R
from yolort_b200.models import yolov5n6
DEV = "cuda:0"
sd = util.synth_state_dict(util.layouts()["n6"], knob_obj=7.0, knob_cls=4.5, seed=0, gain=util.GAIN_N6)
m = yolov5n6(size=(192, 192), score_thresh=0.15).eval(); m.load_state_dict(sd); m = m.to(DEV)
z = util.load_npz("e2e_n6.npz")
ims = [torch.from_numpy(z["img0"]), torch.from_numpy(z["img1"])]
out = m([im.to(DEV) for im in ims])
geoms, (Hb, Wb) = m.transform.geometry(ims)
plan = m.model.get_plan(2, Hb, Wb)
print("canvas", Hb, Wb, "sizes", [tuple(i.shape) for i in ims])
heads = []
for i, h in enumerate(plan.heads):
    hh = h[..., :255].float().cpu()
    heads.append(hh.view(*hh.shape[:3], 3, 85).permute(0, 3, 1, 2, 4).contiguous())
ref_own = R.postprocess(heads, 0.15, 0.45, 300, R.TV_AUTO, util.P6_STRIDES, util.P6_ANCHORS)
for d, im in zip(ref_own, ims):
    d["boxes"] = R.scale_coords(d["boxes"], Hb, Wb, int(im.shape[-2]), int(im.shape[-1]))
for k, (g, r, f) in enumerate(zip(out, ref_own, util.dets_from_npz(z, 2))):
    g = util.to_np(g)
    st_own = util.pair_stats(g, r, 192.0)
    st_fix = util.pair_stats(g, f, 192.0)
    print(f"img{k}: vs oracle-on-own-logits {st_own}")
    print(f"img{k}: vs fixture {st_fix}")
    n = min(len(g['scores']), len(r['scores']), 6)
    print("  got boxes", np.round(g['boxes'][:n], 2).tolist())
    print("  own boxes", np.round(r['boxes'][:n], 2).tolist())
    print("  got scores", g['scores'][:n].tolist(), "labels", g['labels'][:n].tolist())
    print("  own scores", r['scores'][:n].tolist(), "labels", r['labels'][:n].tolist(), "cands", r["n_candidates"])
