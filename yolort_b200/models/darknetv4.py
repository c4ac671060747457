"""CSPDarknet (r3.1 / r4.0) body: Focus stem, three [3x3/s2 Conv, block] stages, a 3x3/s2 Conv and the SPP.

Layout and channel rules follow the reference (yolort/models/darknetv4.py:50-102): stage repeats [3, 9, 9],
stage widths [128, 256, 512], block = BottleneckCSP (r3.1) or C3 (r4.0); module indices 0..8 are what the PAN
taps (4, 6, 8 -- the last one is the SPP output) and what the state-dict keys are built from.
"""
from typing import List

from torch import nn

from ._utils import depth_gain, make_divisible
from .common import BottleneckCSP, C3, Conv, Focus, SPP

BLOCKS = {"r3.1": BottleneckCSP, "r4.0": C3}


def darknet_v4_features(depth_multiple: float, width_multiple: float, version: str = "r4.0",
                        last_channel: int = 1024) -> nn.Sequential:
    if version not in BLOCKS:
        raise NotImplementedError("Currently the module version used in DarkNetV4 is r3.1 or r4.0")
    block = BLOCKS[version]
    c_in = make_divisible(64 * width_multiple, 8)
    layers: List[nn.Module] = [Focus(3, c_in, k=3, version=version)]
    for n, c in zip((3, 9, 9), (128, 256, 512)):
        c_out = make_divisible(c * width_multiple, 8)
        layers.append(Conv(c_in, c_out, k=3, s=2, version=version))
        layers.append(block(c_out, c_out, n=depth_gain(n, depth_multiple)))
        c_in = c_out
    last = make_divisible(last_channel * width_multiple, 8)
    layers.append(Conv(c_in, last, k=3, s=2, version=version))
    layers.append(SPP(last, last, k=(5, 9, 13), version=version))
    return nn.Sequential(*layers)
