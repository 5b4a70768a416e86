"""drop-in for cuda_functions/roi_align_3D/roi_align/crop_and_resize.py:10-69"""
from ....native_ops import CropAndResizeFunction, CropAndResize  # noqa: F401
