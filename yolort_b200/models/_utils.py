"""Channel/depth arithmetic shared by the model zoo.

Mirrors the integer rules of the reference (yolort/models/_utils.py:10-23 `_make_divisible`,
yolort/models/darknetv6.py:70-96 repeat rule).
"""
from typing import Optional


def make_divisible(v: float, divisor: int = 8, min_value: Optional[int] = None) -> int:
    """Round `v` to the nearest multiple of `divisor`, never dropping more than 10 %."""
    floor = divisor if min_value is None else min_value
    rounded = (int(v + divisor / 2) // divisor) * divisor
    out = rounded if rounded > floor else floor
    if out < 0.9 * v:
        out += divisor
    return out


def depth_gain(n: int, depth_multiple: float) -> int:
    """Number of bottlenecks in a stage: max(round(n * depth), 1) (Python banker's round)."""
    g = round(n * depth_multiple)
    return g if g > 1 else 1
