"""Import stub so the reference (yolort) can be imported offline as a parity oracle.

Only the plotting helpers of the reference touch matplotlib; none of them is on the
inference path. Test infrastructure only.
"""


def use(*_a, **_k):
    return None


def rc(*_a, **_k):
    return None
