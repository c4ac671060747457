"""Import the UNMODIFIED reference (zhiqwang/yolort at /root/reference) as a parity oracle.

Works only in the development container (the GPU box has no /root/reference): used by
oracle/make_golden.py to generate tests/golden/* and by tests that are skipped when the tree is absent.
The only missing hard import of the reference is matplotlib (yolort/utils/image_utils.py:8), provided by
the 2-file stub in oracle/_stub.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

REFERENCE_ROOT = "/root/reference"
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stub")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "yolort"))


def import_reference():
    """Returns the `yolort` package of the reference tree."""
    if not available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present on this machine")
    for p in (_STUB, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
    import contextlib
    import io

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with contextlib.redirect_stdout(io.StringIO()):  # silences the font-download message of plots.py
            import yolort  # noqa: F401
            import yolort.models  # noqa: F401
    return yolort
