"""Import-time stand-in for matplotlib (absent from this image); test tooling only."""
rcParams = {}


def use(*a, **k):
    pass
