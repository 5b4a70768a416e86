"""drop-in for cuda_functions/roi_align_2D/roi_align/crop_and_resize.py"""
from ....native_ops import CropAndResizeFunction2D as CropAndResizeFunction, CropAndResize2D as CropAndResize  # noqa: F401
