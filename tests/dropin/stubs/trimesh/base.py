class Trimesh:  # noqa: D101 - see package docstring
    pass
