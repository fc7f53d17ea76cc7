"""TEST INFRASTRUCTURE: stand-in for `imageio` (absent from this image): imread / imwrite through PIL."""
import numpy as np
from PIL import Image


def imread(path):
    return np.asarray(Image.open(path))


def imwrite(path, array):
    a = np.asarray(array)
    if a.dtype != np.uint8:
        a = (np.clip(a, 0, 1) * 255).astype(np.uint8) if a.dtype.kind == "f" else a.astype(np.uint8)
    if a.ndim == 3 and a.shape[2] == 1:
        a = a[:, :, 0]
    Image.fromarray(a).save(path)
