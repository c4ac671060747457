"""CSPDarknet (r6.0) body: 6x6/s2 stem followed by four [3x3/s2 Conv, C3] stages.

Layout and channel rules follow the reference (yolort/models/darknetv6.py:76-98); module indices
0..8 are what the PAN taps (4, 6, 8) and what the state-dict keys are built from.
"""
from typing import List

from torch import nn

from ._utils import depth_gain, make_divisible
from .common import C3, Conv


def darknet_v6_features(depth_multiple: float, width_multiple: float, last_channel: int = 1024) -> nn.Sequential:
    widths = [make_divisible(c * width_multiple, 8) for c in (64, 128, 256, 512)]
    last = make_divisible(last_channel * width_multiple, 8)
    repeats = [depth_gain(n, depth_multiple) for n in (3, 6, 9)] + [depth_gain(3, depth_multiple)]
    outs = widths[1:] + [last]

    layers: List[nn.Module] = [Conv(3, widths[0], k=6, s=2, p=2)]
    c_in = widths[0]
    for n, c_out in zip(repeats, outs):
        layers.append(Conv(c_in, c_out, k=3, s=2))
        layers.append(C3(c_out, c_out, n=n))
        c_in = c_out
    return nn.Sequential(*layers)
