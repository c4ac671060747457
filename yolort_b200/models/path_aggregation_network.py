"""PANet neck (r6.0 / r4.0 / r3.1) -- parameter container.

Block order and indices follow the reference (yolort/models/path_aggregation_network.py:77-165):
3 levels: inner_blocks = [SPP, Conv1x1, Upsample, C3, Conv1x1, Upsample], layer_blocks = [C3, Conv3x3s2, C3,
Conv3x3s2, C3].  P6 (4 levels, `:10-41,119-126,148-154`): an intermediate [Conv3x3s2, C3] block builds the
stride-64 map, inner_blocks gain [Conv1x1, Upsample, C3] after the SPP and layer_blocks gain [Conv3x3s2, C3].
r3.1 / r4.0 (`:104-111`): the first inner block is a BottleneckCSP / C3 without shortcut instead of the SPP
(their SPP is the last module of the body) and r3.1 uses BottleneckCSP + Hardswish everywhere.
Data flow (`:199-239`) is lowered by yolort_b200/engine.py.
"""
from typing import List

from torch import nn

from ._utils import depth_gain
from .common import BottleneckCSP, C3, Conv, SPP, _PlanOnly

_BLOCK = {"r3.1": BottleneckCSP, "r4.0": C3}   # path_aggregation_network.py:242-245


class IntermediateLevelP6(_PlanOnly):
    """Stride-64 level appended to the body taps (path_aggregation_network.py:10-41); the Sequential keeps the
    reference's parameter names `intermediate_blocks.p6.{0,1}`."""

    def __init__(self, depth_multiple: float, in_channel: int, out_channel: int, version: str = "r4.0"):
        super().__init__()
        self.p6 = nn.Sequential(
            Conv(in_channel, out_channel, k=3, s=2, version=version),
            _BLOCK[version](out_channel, out_channel, n=depth_gain(3, depth_multiple)),
        )


class PathAggregationNetwork(_PlanOnly):
    def __init__(self, in_channels: List[int], depth_multiple: float, version: str = "r6.0", use_p6: bool = False):
        super().__init__()
        if version not in ("r3.1", "r4.0", "r6.0"):
            raise NotImplementedError(f"Version {version} is not implemented yet.")
        mv = "r4.0" if version == "r6.0" else version      # module version (`:87`)
        block = _BLOCK[mv]
        n = depth_gain(3, depth_multiple)
        ch = list(in_channels)
        if use_p6:
            if len(ch) != 4:
                raise ValueError("Length of in channels should be 4.")
            self.intermediate_blocks = IntermediateLevelP6(depth_multiple, ch[2], ch[3], version=mv)
        else:
            if len(ch) != 3:
                raise ValueError("Length of in channels should be 3.")
            self.intermediate_blocks = None
        if version == "r6.0":
            inner: List[nn.Module] = [SPP(ch[-1], ch[-1], k=(5, 9, 13))]
        else:
            inner = [block(ch[-1], ch[-1], n=n, shortcut=False)]
        if use_p6:
            inner += [Conv(ch[-1], ch[2], 1, 1, version=mv), nn.Upsample(scale_factor=2),
                      block(ch[1] + ch[-1], ch[2], n=n, shortcut=False)]
        inner += [
            Conv(ch[2], ch[1], 1, 1, version=mv),
            nn.Upsample(scale_factor=2),
            block(ch[-1], ch[1], n=n, shortcut=False),
            Conv(ch[1], ch[0], 1, 1, version=mv),
            nn.Upsample(scale_factor=2),
        ]
        self.inner_blocks = nn.ModuleList(inner)
        layer: List[nn.Module] = [
            block(ch[1], ch[0], n=n, shortcut=False),
            Conv(ch[0], ch[0], 3, 2, version=mv),
            block(ch[1], ch[1], n=n, shortcut=False),
            Conv(ch[1], ch[1], 3, 2, version=mv),
            block(ch[-1], ch[2], n=n, shortcut=False),
        ]
        if use_p6:
            layer += [Conv(ch[2], ch[2], 3, 2, version=mv), block(ch[1] + ch[-1], ch[-1], n=n, shortcut=False)]
        self.layer_blocks = nn.ModuleList(layer)
