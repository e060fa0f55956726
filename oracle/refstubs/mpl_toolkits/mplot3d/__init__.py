"""mpl_toolkits.mplot3d stub (test tooling only)."""


class Axes3D:  # pragma: no cover
    pass
