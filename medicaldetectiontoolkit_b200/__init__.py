"""medicaldetectiontoolkit_b200 — B200 (sm_100a) native hot path behind the Medical Detection Toolkit's operator surface.

See DESIGN.md for the scope and INTEGRATION.md for how the reference binds to it.
"""
import sys

__version__ = "0.1.0"


def install_dropin():
    """Register this package's `cuda_functions` under the reference's top-level module name, so that the reference's unmodified
    `from cuda_functions.nms_3D.pth_nms import nms_gpu as nms_3D` (mrcnn.py:24-27, retina_unet.py:26-27) resolves to libmdt_b200."""
    import importlib
    base = __name__ + ".cuda_functions"
    names = ["", ".nms_2D", ".nms_2D.pth_nms", ".nms_3D", ".nms_3D.pth_nms", ".roi_align_2D", ".roi_align_2D.roi_align",
             ".roi_align_2D.roi_align.crop_and_resize", ".roi_align_3D", ".roi_align_3D.roi_align", ".roi_align_3D.roi_align.crop_and_resize"]
    for n in names:
        sys.modules["cuda_functions" + n] = importlib.import_module(base + n)
