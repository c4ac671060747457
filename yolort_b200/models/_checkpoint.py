"""Upstream (ultralytics/yolov5) checkpoint -> yolort state-dict layout.

Same contract as the reference converter (yolort/models/_checkpoint.py:16-94 `load_from_ultralytics`, module
index maps `:53-64`, anchor recovery `:34-47`) with two differences that fix the latent problems listed in
SURVEY.md section 0.10:

  * the upstream pickle references classes `models.yolo.*` / `models.common.*`; instead of putting a copy of
    the upstream source tree on `sys.path` (yolort/v5/helper.py:15-29) the pickle is read with an Unpickler
    that materialises those names as bare `nn.Module` stubs -- only their parameters/buffers/attributes are
    needed, none of their code;
  * `strides` is returned as `List[int]` (the reference hands a Tensor to `SetCriterion`, which breaks).
"""
import pickle
import types
from typing import Any, Dict, List

import torch
from torch import nn

__all__ = ["load_from_ultralytics", "load_upstream_model", "get_yolov5_size"]

_STUB_PREFIXES = ("models.", "utils.", "models", "utils")


class _Stub(nn.Module):
    """Attribute container standing in for an upstream class."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("upstream modules are loaded as parameter containers only")


_stub_cache: Dict[str, type] = {}


def _stub_class(module: str, name: str) -> type:
    key = f"{module}.{name}"
    cls = _stub_cache.get(key)
    if cls is None:
        cls = type(name, (_Stub,), {"__module__": module})
        _stub_cache[key] = cls
    return cls


class _UpstreamUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        if module.split(".")[0] in ("models", "utils"):
            return _stub_class(module, name)
        return super().find_class(module, name)


def _pickle_module() -> types.ModuleType:
    m = types.ModuleType("yolort_b200_upstream_pickle")
    m.__dict__.update({k: getattr(pickle, k) for k in dir(pickle) if not k.startswith("__")})
    m.Unpickler = _UpstreamUnpickler

    def load(f, **kw):
        return _UpstreamUnpickler(f, **kw).load()

    m.load = load
    return m


def load_upstream_model(checkpoint_path: str) -> nn.Module:
    """Returns the fp32 upstream model object (stub classes), `ema` preferred over `model` as in
    yolort/v5/helper.py:67-72."""
    ckpt = torch.load(checkpoint_path, map_location="cpu", pickle_module=_pickle_module(), weights_only=False)
    if isinstance(ckpt, dict):
        model = ckpt["ema"] if ckpt.get("ema") is not None else ckpt["model"]
    else:
        model = ckpt
    return model.float().eval()


def get_yolov5_size(depth_multiple: float, width_multiple: float) -> str:
    table = {(0.33, 0.25): "n", (0.33, 0.5): "s", (0.67, 0.75): "m", (1.0, 1.0): "l", (1.33, 1.25): "x"}
    try:
        return table[(depth_multiple, width_multiple)]
    except KeyError:
        raise NotImplementedError(
            f"Currently does't support architecture with depth: {depth_multiple} and width: {width_multiple}") from None


def _sequential(model: nn.Module) -> nn.Module:
    while not isinstance(model, nn.Sequential):
        model = model.model
    return model


def load_from_ultralytics(checkpoint_path: str, version: str = "r6.0") -> Dict[str, Any]:
    if version not in ("r3.1", "r4.0", "r6.0"):     # _checkpoint.py:26-30; the module index maps are shared (`:60-64`)
        raise NotImplementedError(f"Currently does not support version: {version}.")
    up = load_upstream_model(checkpoint_path)
    seq = _sequential(up)
    detect = seq[-1]
    num_classes = int(up.yaml["nc"])
    depth_multiple, width_multiple = up.yaml["depth_multiple"], up.yaml["width_multiple"]
    strides: List[int] = [int(s) for s in torch.as_tensor(up.stride).tolist()]
    num_anchors = int(detect.anchors.shape[1])
    # anchors in pixels = Detect.anchors (grid units) * stride  (_checkpoint.py:38-43)
    anchor_grids = (detect.anchors.float() * torch.as_tensor(detect.stride).float().view(-1, 1, 1)).reshape(
        len(strides), 2 * num_anchors).tolist()
    use_p6 = len(strides) == 4                                # _checkpoint.py:49-51
    if len(strides) not in (3, 4):
        raise NotImplementedError(f"checkpoints with {len(strides)} detection levels are not supported")
    if use_p6:                                                # _checkpoint.py:53-58
        inner_map = {"0": 11, "1": 12, "3": 15, "4": 16, "6": 19, "7": 20}
        layer_map = {"0": 23, "1": 24, "2": 26, "3": 27, "4": 29, "5": 30, "6": 32}
        p6_map = {"0": 9, "1": 10}
        head_ind = 33
    else:
        inner_map = {"0": 9, "1": 10, "3": 13, "4": 14}          # _checkpoint.py:60
        layer_map = {"0": 17, "1": 18, "2": 20, "3": 21, "4": 23}  # _checkpoint.py:61
        p6_map = {}
        head_ind = 24

    sd: Dict[str, torch.Tensor] = {}

    def take(prefix: str, src: nn.Module) -> None:
        for k, v in src.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    for i in range(9):
        take(f"backbone.body.{i}", seq[i])
    for ours, theirs in p6_map.items():
        take(f"backbone.pan.intermediate_blocks.p6.{ours}", seq[theirs])
    for ours, theirs in inner_map.items():
        take(f"backbone.pan.inner_blocks.{ours}", seq[theirs])
    for ours, theirs in layer_map.items():
        take(f"backbone.pan.layer_blocks.{ours}", seq[theirs])
    for i, conv in enumerate(seq[head_ind].m):
        take(f"head.head.{i}", conv)
    # the reference returns a half-precision state dict (_checkpoint.py:81)
    sd = {k: (v.half() if v.is_floating_point() else v) for k, v in sd.items()}
    return {
        "num_classes": num_classes,
        "depth_multiple": depth_multiple,
        "width_multiple": width_multiple,
        "strides": strides,
        "anchor_grids": anchor_grids,
        "use_p6": use_p6,
        "size": get_yolov5_size(depth_multiple, width_multiple),
        "state_dict": sd,
    }
