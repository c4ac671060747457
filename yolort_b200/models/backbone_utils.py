"""`darknet_pan_backbone`: CSPDarknet body + PAN, same constructor contract as the reference
(yolort/models/backbone_utils.py:60-122).  `body` keeps the module names "0".."8" so that
state-dict keys read `backbone.body.<i>...`, as produced by torchvision's IntermediateLayerGetter
in the reference (`backbone_utils.py:45`)."""
from typing import List, Optional

from torch import nn

from .common import _PlanOnly
from .darknetv4 import darknet_v4_features
from .darknetv6 import darknet_v6_features
from .path_aggregation_network import PathAggregationNetwork


class BackboneWithPAN(_PlanOnly):
    def __init__(self, body: nn.Sequential, returned_layers: List[int], in_channels_list: List[int],
                 depth_multiple: float, version: str, use_p6: bool = False):
        super().__init__()
        last = max(returned_layers)
        self.body = nn.ModuleDict({str(i): m for i, m in enumerate(body) if i <= last})
        self.returned_layers = list(returned_layers)
        self.pan = PathAggregationNetwork(in_channels_list, depth_multiple, version=version, use_p6=use_p6)
        self.out_channels = in_channels_list

    def forward(self, x):
        """`backbone(x)` (backbone_utils.py:54-57): [N,3,H,W] -> PAN feature maps [N,C_l,H/s_l,W/s_l], executed as the
        backbone launch range of the owning YOLO's plan (forward hooks fire: yolort/utils/hooks.py:7-26)."""
        owner = self.__dict__.get("_yb_owner")
        if not owner:
            return super().forward(x)
        return owner[0].run_backbone(x)


def darknet_pan_backbone(
    backbone_name: str,
    depth_multiple: float,
    width_multiple: float,
    pretrained: Optional[bool] = False,
    returned_layers: Optional[List[int]] = None,
    version: str = "r6.0",
    use_p6: bool = False,
) -> BackboneWithPAN:
    if version not in ("r3.1", "r4.0", "r6.0"):
        raise NotImplementedError("Currently only supports version 'r3.1', 'r4.0' and 'r6.0'.")
    if pretrained:
        raise ValueError("no backbone checkpoints exist offline")
    last_channel = 768 if use_p6 else 1024
    if version == "r6.0":
        body = darknet_v6_features(depth_multiple, width_multiple, last_channel=last_channel)
    else:
        body = darknet_v4_features(depth_multiple, width_multiple, version=version, last_channel=last_channel)
    if returned_layers is None:
        returned_layers = [4, 6, 8]
    grow_widths = [256, 512, 768, 1024] if use_p6 else [256, 512, 1024]
    in_channels_list = [int(gw * width_multiple) for gw in grow_widths]
    return BackboneWithPAN(body, returned_layers, in_channels_list, depth_multiple, version, use_p6=use_p6)
