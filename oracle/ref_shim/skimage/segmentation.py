import numpy as np
from scipy import ndimage as ndi
from .morphology import dilation, erosion


def find_boundaries(label_img, connectivity=1, mode='thick', background=0):
    # defaults only: thick boundaries = dilation != erosion over the cross footprint
    if label_img.dtype == bool:
        label_img = label_img.astype(np.uint8)
    fp = ndi.generate_binary_structure(label_img.ndim, connectivity)
    return dilation(label_img, fp) != erosion(label_img, fp)


def slic(*a, **k):
    raise NotImplementedError("slic is not available in the stand-in")
