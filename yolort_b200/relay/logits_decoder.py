"""`LogitsDecoder`: the post-process without NMS (yolort/relay/logits_decoder.py:10-61).

The reference swaps it in for `PostProcess` (relay/trt_inference.py:43) so that the graph ends in dense
`(boxes [N, A, 4], scores [N, A, num_classes])` -- the input of TensorRT's EfficientNMS plugin
(relay/trt_graphsurgeon.py:212-246).  Here it is one launch of `yb_decode_dense` over the head logits.
"""
from typing import List, Optional, Sequence, Tuple

from torch import nn, Tensor

from .. import _C


class LogitsDecoder(nn.Module):
    def __init__(self, strides: List[int], anchors_px: Optional[Sequence[Sequence[float]]] = None) -> None:
        """
        Args:
            strides (List[int]): Strides of the AnchorGenerator.
            anchors_px: per level [aw0, ah0, aw1, ah1, ...] in pixels; when None they are read from the
                reference-style `shifts` argument of forward (AnchorGenerator output).
        """
        super().__init__()
        self.strides = [int(s) for s in strides]
        self.anchors_px = anchors_px

    def forward(self, head_outputs: List[Tensor], grids: Optional[List[Tensor]] = None,
                shifts: Optional[List[Tensor]] = None) -> Tuple[Tensor, Tensor]:
        """head_outputs: per level [N, A, H, W, K] (reference layout).  grids are implied by the level shapes."""
        anchors_px = self.anchors_px
        if anchors_px is None:
            if shifts is None:
                raise ValueError("LogitsDecoder needs anchors_px (or reference-style shifts)")
            anchors_px = [s[0, :, 0, 0, :].reshape(-1).float().tolist() for s in shifts]
        num_classes = int(head_outputs[0].shape[-1]) - 5
        return _C.decode_dense(head_outputs, "nahwk", self.strides, anchors_px, num_classes)

    def decode_plan_heads(self, heads: List[Tensor], anchors_px, num_classes: int) -> Tuple[Tensor, Tensor]:
        """Same on the plan's NHWC head buffers (channel a*K + k): what YOLO.forward calls."""
        return _C.decode_dense(heads, "nhwc", self.strides, anchors_px, num_classes)
