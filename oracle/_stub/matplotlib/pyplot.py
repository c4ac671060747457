"""Stub of matplotlib.pyplot (see package docstring)."""


class _Dummy:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Dummy()
