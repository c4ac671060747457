"""Wire formats for downstream deployments."""
from .logits_decoder import LogitsDecoder

__all__ = ["LogitsDecoder"]
