import numpy as np
from scipy import ndimage as ndi


def rectangle(nrows, ncols, dtype=np.uint8):
    return np.ones((nrows, ncols), dtype=dtype)


def disk(radius, dtype=np.uint8):
    L = np.arange(-radius, radius + 1)
    X, Y = np.meshgrid(L, L)
    return np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)


def dilation(image, footprint=None, out=None):
    # scikit-image mirrors the footprint because scipy's grey_dilation mirrors it back.
    fp = np.array(footprint)[::-1, ::-1]
    if out is None:
        out = np.empty_like(image)
    ndi.grey_dilation(image, footprint=fp, output=out)
    return out


def erosion(image, footprint=None, out=None):
    if out is None:
        out = np.empty_like(image)
    ndi.grey_erosion(image, footprint=np.array(footprint), output=out)
    return out
