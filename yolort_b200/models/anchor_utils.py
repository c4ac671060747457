"""Anchor constants of the detection head.

The reference regenerates dense grid/shift tensors on every forward
(yolort/models/anchor_utils.py:19-67).  The decode kernel derives the grid from the thread's
(x, y) and only needs the per-level strides and the per-anchor pixel sizes, so this class keeps
those constants; `forward` still returns reference-shaped grids/shifts for callers (and tests) that
want them, computed on the host side with plain tensor indexing.
"""
from typing import List, Tuple

import torch
from torch import nn, Tensor


class AnchorGenerator(nn.Module):
    def __init__(self, strides: List[int], anchor_grids: List[List[float]]):
        super().__init__()
        if len(strides) != len(anchor_grids):
            raise ValueError("strides and anchor_grids must have one entry per level")
        self.strides = [int(s) for s in strides]
        self.anchor_grids = [list(map(float, a)) for a in anchor_grids]
        self.num_layers = len(anchor_grids)
        self.num_anchors = len(anchor_grids[0]) // 2

    def anchors_px(self) -> List[List[float]]:
        """Per level, [aw0, ah0, aw1, ah1, ...] in pixels, evaluated the way the reference does:
        fp32(anchor / stride) * stride (anchor_utils.py:46-57); exact for the default anchors."""
        out = []
        for lvl in range(self.num_layers):
            a = torch.tensor(self.anchor_grids[lvl], dtype=torch.float32)
            s = torch.tensor(float(self.strides[lvl]), dtype=torch.float32)
            out.append(((a / s) * self.strides[lvl]).tolist())
        return out

    def forward(self, feature_maps: List[Tensor]) -> Tuple[List[Tensor], List[Tensor]]:
        dtype, device = feature_maps[0].dtype, feature_maps[0].device
        grids, shifts = [], []
        px = self.anchors_px()
        for lvl, fm in enumerate(feature_maps):
            h, w = int(fm.shape[-2]), int(fm.shape[-1])
            xs = torch.arange(w, device=device).to(dtype).view(1, 1, 1, w).expand(1, self.num_anchors, h, w)
            ys = torch.arange(h, device=device).to(dtype).view(1, 1, h, 1).expand(1, self.num_anchors, h, w)
            grids.append(torch.stack((xs, ys), dim=-1))
            a = torch.tensor(px[lvl], dtype=dtype, device=device).view(1, self.num_anchors, 1, 1, 2)
            shifts.append(a.expand(1, self.num_anchors, h, w, 2).contiguous())
        return grids, shifts
