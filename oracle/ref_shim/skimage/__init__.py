from . import io, color, segmentation, morphology, measure  # noqa: F401
