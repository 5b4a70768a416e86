"""ctypes binding of libmdt_b200.so (the C-ABI declared in include/mdt_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.  PyTorch is used only for device
memory and the current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmdt_b200.so")

c_i64p = ctypes.POINTER(ctypes.c_int64)


class MdtError(RuntimeError):
    pass


class Conv3dDesc(ctypes.Structure):
    """mirror of struct mdt_conv3d_desc (include/mdt_b200.h)"""
    _fields_ = [(n, ctypes.c_int) for n in (
        "n", "d", "h", "w", "cin", "cout", "kd", "kh", "kw", "sd", "sh", "sw", "pd", "ph", "pw", "relu", "precision", "algo")]


_lib = None

_VP = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_D = ctypes.c_double
_SZ = ctypes.c_size_t
_I64 = ctypes.c_int64

# name -> (restype, argtypes); every symbol declared in include/mdt_b200.h must appear here (tests/test_abi.py checks both directions)
SIGNATURES = {
    "mdt_version": (_I, []),
    "mdt_error_string": (ctypes.c_char_p, [_I]),
    "mdt_launch_count": (ctypes.c_ulonglong, []),
    "mdt_nms_mask_3d": (_I, [_I, _VP, _VP, _F, _VP]),
    "mdt_nms_mask_2d": (_I, [_I, _VP, _VP, _F, _VP]),
    "mdt_nms_workspace_bytes": (_SZ, [_I]),
    "mdt_nms_3d": (_I, [_VP, _I, _F, _VP, _SZ, _VP, _VP, _VP]),
    "mdt_nms_2d": (_I, [_VP, _I, _F, _VP, _SZ, _VP, _VP, _VP]),
    "mdt_crop_and_resize_3d_forward": (_I, [_VP, c_i64p, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _VP, c_i64p, _VP]),
    "mdt_crop_and_resize_3d_backward": (_I, [_VP, c_i64p, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _I, _VP, c_i64p, _I, _I64, _VP]),
    "mdt_crop_and_resize_2d_forward": (_I, [_VP, c_i64p, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _F, _VP, c_i64p, _VP]),
    "mdt_crop_and_resize_2d_backward": (_I, [_VP, c_i64p, _VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP, c_i64p, _I, _I64, _VP]),
    "mdt_pyramid_roi_align_forward": (_I, [_I, ctypes.POINTER(_VP), c_i64p, ctypes.POINTER(_I), _I, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, _VP, c_i64p, _VP]),
    "mdt_pyramid_roi_align_backward": (_I, [_I, _VP, c_i64p, _VP, _VP, _VP, _I, _I, _I, _I, _I, _I, ctypes.POINTER(_VP), c_i64p, ctypes.POINTER(_I), _I, _I,
                                            c_i64p, _VP]),
    "mdt_anchor_match_workspace_bytes": (_SZ, [_I]),
    "mdt_anchor_match": (_I, [_I, _VP, _I, _VP, _VP, _I, _D, _D, _VP, _SZ, _VP, _VP, _VP, _VP]),
    "mdt_anchor_delta_targets": (_I, [_I, _VP, _VP, _VP, _VP, _I, _I, ctypes.POINTER(_D), _VP, _VP]),
    "mdt_conv3d_workspace_bytes": (_SZ, [ctypes.POINTER(Conv3dDesc), _I]),
    "mdt_conv3d_fprop": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_conv3d_dgrad": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_conv3d_wgrad": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_conv3d_algo": (_I, [ctypes.POINTER(Conv3dDesc), _I]),
    "mdt_conv3d_variant": (_I, [ctypes.POINTER(Conv3dDesc), _I]),
    "mdt_debug_conv_tcw_prof": (_I, [ctypes.POINTER(ctypes.c_ulonglong)]),
    "mdt_conv3d_backward_fused": (_I, [ctypes.POINTER(Conv3dDesc), _I]),
    "mdt_conv3d_backward_workspace_bytes": (_SZ, [ctypes.POINTER(Conv3dDesc), _I]),
    "mdt_conv3d_backward": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_conv3d_split_bytes": (_SZ, [ctypes.POINTER(Conv3dDesc)]),
    "mdt_conv3d_split": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP]),
    "mdt_conv3d_fprop_presplit": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_conv3d_out_split_bytes": (_SZ, [ctypes.POINTER(Conv3dDesc)]),
    "mdt_conv3d_fprop_presplit_out": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_conv3d_fprop_out": (_I, [ctypes.POINTER(Conv3dDesc), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_upsample221_forward": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "mdt_upsample221_backward": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "mdt_maxpool3d_forward": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I), _VP]),
    "mdt_maxpool3d_backward": (_I, [_VP, _VP, _VP, _I, _I, _I, _I, _I, ctypes.POINTER(_I), ctypes.POINTER(_I), ctypes.POINTER(_I), _VP]),
    "mdt_upsample_nearest_forward": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mdt_upsample_nearest_backward": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _I, _VP]),
    "mdt_seg_loss_workspace_bytes": (_SZ, [_I]),
    "mdt_seg_loss_forward": (_I, [_VP, c_i64p, _VP, _I, _I64, _I, _F, _F, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_seg_loss_backward": (_I, [_VP, c_i64p, _VP, _I, _I64, _I, _F, _F, _VP, _VP, _VP, _VP]),
    "mdt_shem_workspace_bytes": (_SZ, [_I, _I]),
    "mdt_shem_class_loss_forward": (_I, [_VP, _VP, _I, _I, _VP, _I, _I, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_shem_class_loss_backward": (_I, [_VP, _I, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "mdt_wbc_workspace_bytes": (_SZ, [_I, _I]),
    "mdt_wbc": (_I, [_VP, _VP, _VP, _I, _I, _I, _D, _D, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "mdt_nms_2to3d_workspace_bytes": (_SZ, [_I, _I]),
    "mdt_nms_2to3d": (_I, [_VP, _VP, _I, _D, _I, _VP, _VP, _VP, _VP, _SZ, _VP]),
}


def load():
    """Load libmdt_b200.so once; raises MdtError if it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MdtError("libmdt_b200.so not found at %s - build it with `make -C medicaldetectiontoolkit_b200/csrc` "
                       "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise MdtError("libmdt_b200: %s (code %d)" % (load().mdt_error_string(code).decode(), code))


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def i64arr(vals):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise MdtError("libmdt_b200 ops take CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)
