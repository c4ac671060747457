"""PANet neck (r6.0) -- parameter container.

Block order and indices follow the reference (yolort/models/path_aggregation_network.py:77-165):
inner_blocks = [SPP, Conv1x1, Upsample, C3, Conv1x1, Upsample], layer_blocks = [C3, Conv3x3s2, C3,
Conv3x3s2, C3].  Data flow (`:199-239`) is lowered by yolort_b200/engine.py.
"""
from typing import List

from torch import nn

from ._utils import depth_gain
from .common import C3, Conv, SPP, _PlanOnly


class PathAggregationNetwork(_PlanOnly):
    def __init__(self, in_channels: List[int], depth_multiple: float, version: str = "r6.0", use_p6: bool = False):
        super().__init__()
        if version != "r6.0":
            raise NotImplementedError(f"only upstream version 'r6.0' is built here, got {version!r}")
        if use_p6:
            raise NotImplementedError("P6 variants are listed as 'next' in SURVEY.md section 8(f)")
        if len(in_channels) != 3:
            raise ValueError("Length of in channels should be 3.")
        c3, c4, c5 = in_channels
        n = depth_gain(3, depth_multiple)
        self.intermediate_blocks = None
        self.inner_blocks = nn.ModuleList(
            [
                SPP(c5, c5, k=(5, 9, 13)),
                Conv(c5, c4, 1, 1),
                nn.Upsample(scale_factor=2),
                C3(c5, c4, n=n, shortcut=False),
                Conv(c4, c3, 1, 1),
                nn.Upsample(scale_factor=2),
            ]
        )
        self.layer_blocks = nn.ModuleList(
            [
                C3(c4, c3, n=n, shortcut=False),
                Conv(c3, c3, 3, 2),
                C3(c4, c4, n=n, shortcut=False),
                Conv(c4, c4, 3, 2),
                C3(c5, c5, n=n, shortcut=False),
            ]
        )
