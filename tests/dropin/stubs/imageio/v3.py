from . import imread, imwrite  # noqa: F401
