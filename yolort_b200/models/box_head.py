"""Detection head and post-processing front-ends.

`YOLOHead` keeps the reference parameter layout (yolort/models/box_head.py:14-82: one 1x1 conv with
bias per level, 3*(nc+5) outputs, bias initialised as `:40-46`).  `PostProcess` has the reference
constructor (box_head.py:363-386) and runs anchor-decode + multi-label threshold + batched NMS +
top-k as ONE native call (`yb_decode_nms`, csrc/decode_nms.cu) instead of the per-image Python loop
at box_head.py:414-427.
"""
import math
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn, Tensor

from .. import _C
from .common import _PlanOnly

# NMS semantics selector (see include/yolort_b200.h).  torchvision.ops.batched_nms switches between the
# coordinate-offset trick and exact per-class NMS on numel.
NMS_TV_AUTO = 0
NMS_EXACT_PER_CLASS = 1
NMS_OFFSET_TRICK = 2


class YOLOHead(_PlanOnly):
    def __init__(self, in_channels: List[int], num_anchors: int, strides: List[int], num_classes: int):
        super().__init__()
        if not isinstance(in_channels, list):
            in_channels = [in_channels] * len(strides)
        self.num_anchors = num_anchors
        self.num_classes = num_classes
        self.num_outputs = num_classes + 5
        self.strides = strides
        blocks = nn.ModuleList(nn.Conv2d(ch, self.num_outputs * num_anchors, 1) for ch in in_channels)
        for conv, s in zip(blocks, strides):
            with torch.no_grad():
                b = conv.bias.view(num_anchors, -1)
                b[:, 4] += math.log(8 / (640 / float(s)) ** 2)  # ~8 objects per 640 image
                b[:, 5:] += math.log(0.6 / (num_classes - 0.999999))
        self.head = blocks

    def forward(self, x: List[Tensor]) -> List[Tensor]:
        """`head(features)` (box_head.py:68-82): per level [N, A, H, W, nc+5] raw logits -- the training-mode output
        of the detector as well -- executed as the head launch range of the owning YOLO's plan."""
        owner = self.__dict__.get("_yb_owner")
        if not owner:
            return super().forward(x)
        return owner[0].run_head(list(x))


class PostProcess(nn.Module):
    """Decode + threshold + batched NMS + top-k on the device.

    forward() accepts the reference argument list (head_outputs [N,A,H,W,K] per level, grids, shifts);
    grids/shifts are accepted for signature compatibility -- the kernel recomputes them from
    `strides`/`anchors_px` (they are pure functions of the level shape).
    """

    def __init__(self, strides: List[int], score_thresh: float, nms_thresh: float, detections_per_img: int,
                 anchors_px: Optional[Sequence[Sequence[float]]] = None, nms_semantics: int = NMS_TV_AUTO):
        super().__init__()
        self.strides = strides
        self.score_thresh = score_thresh
        self.nms_thresh = nms_thresh
        self.detections_per_img = detections_per_img
        self.anchors_px = anchors_px
        self.nms_semantics = nms_semantics

    def forward(self, head_outputs: List[Tensor], grids: Optional[List[Tensor]] = None,
                shifts: Optional[List[Tensor]] = None) -> List[Dict[str, Tensor]]:
        if self.anchors_px is None:
            if shifts is None:
                raise ValueError("PostProcess needs anchors_px (or reference-style shifts)")
            anchors_px = [s[0, :, 0, 0, :].reshape(-1).float().tolist() for s in shifts]
        else:
            anchors_px = self.anchors_px
        return _C.decode_nms(
            head_outputs, layout="nahwk", strides=self.strides, anchors_px=anchors_px,
            score_thresh=self.score_thresh, nms_thresh=self.nms_thresh,
            detections_per_img=self.detections_per_img, semantics=self.nms_semantics,
        )
