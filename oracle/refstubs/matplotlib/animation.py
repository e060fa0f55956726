"""matplotlib stub submodule (test tooling only)."""


class _Anything:
    def __init__(self, *a, **k):
        raise RuntimeError("matplotlib stub: plotting is unavailable in this image")


def __getattr__(name):
    return _Anything
