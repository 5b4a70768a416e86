"""conv3d / conv2d modules and the NDConvGenerator factory on libmdt_b200's conv kernels.

Reference surface mirrored: utils/model_utils.py:732-781 `NDConvGenerator(dim)(c_in, c_out, ks, pad=0, stride=1, norm=None, relu='relu')`
returns nested nn.Sequential wrappers whose state-dict keys (`...0.weight`, `...0.bias`, bare `weight` when relu is None) and parameter
shapes ([Cout, Cin, kd, kh, kw]) are kept, so reference checkpoints / initialisations load unchanged (SURVEY.md §5 checkpoint row).

Activations are NDHWC in memory (torch.channels_last_3d; the logical shape stays [N, C, Y, X, Z] so the operator surface is unchanged).
bias + ReLU (+ residual) are fused into the conv epilogue; the ReLU module that the reference appends stays in the Sequential as a
parameter-free marker so the key structure is identical.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L

_CL3 = torch.channels_last_3d

# precision: 0 = fp32-faithful (3-pass split-bf16 on tcgen05 / fp32 SIMT), 1 = single-pass bf16 on tcgen05
DEFAULT_PRECISION = 0
# algo: 0 auto, 1 force SIMT, 2 force tcgen05, 4 force the pointwise fp32 streaming kernels (1x1x1 stride 1)
DEFAULT_ALGO = 0

_ws_cache = {}

# conv modules whose output feeds another conv emit the bf16 split planes of their result from the epilogue (module attribute emit_split)
EMIT_SPLIT = True

# when a list, every conv kernel call appends a (start, end) CUDA-event pair recorded on the launching stream (bench.py roofline)
EVENT_LOG = None


def _workspace(nbytes, device):
    """one grow-only scratch buffer per device/stream-less use: conv calls on a stream are ordered, so reuse is safe"""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _ev_start():
    if EVENT_LOG is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _ev_end(e0, tag=None):
    if e0 is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        EVENT_LOG.append((e0, e1, tag))


_desc_cache = {}


def _desc(x_shape, w_shape, stride, padding, relu, precision, algo):
    key = (tuple(x_shape), tuple(w_shape), tuple(stride), tuple(padding), bool(relu), int(precision), int(algo))
    d = _desc_cache.get(key)
    if d is None:
        n, cin, dd, h, w = x_shape
        cout, _, kd, kh, kw = w_shape
        d = L.Conv3dDesc(n, dd, h, w, cin, cout, kd, kh, kw, stride[0], stride[1], stride[2], padding[0], padding[1], padding[2],
                         int(bool(relu)), int(precision), int(algo))
        _desc_cache[key] = d
    return d


_plan_cache = {}


def _plan(lib, d, ps):
    """(workspace bytes, algorithm) of a descriptor/pass, cached: two ctypes calls saved per conv call"""
    key = (id(d), ps)
    v = _plan_cache.get(key)
    if v is None:
        v = (lib.mdt_conv3d_workspace_bytes(d, ps), lib.mdt_conv3d_algo(d, ps))
        _plan_cache[key] = v
    return v


def _out_shape(x_shape, w_shape, stride, padding):
    n, _, d, h, w = x_shape
    cout, _, kd, kh, kw = w_shape
    return (n, cout, (d + 2 * padding[0] - kd) // stride[0] + 1, (h + 2 * padding[1] - kh) // stride[1] + 1,
            (w + 2 * padding[2] - kw) // stride[2] + 1)


def _split_of(lib, x, d, precision):
    """canonical split form of x (mdt_conv3d_split), cached ON the tensor: sibling convs reading the same activation and the later weight
    gradient reuse it instead of re-splitting (the form depends on the tensor alone: channel count and W)"""
    hit = getattr(x, "_mdt_split", None)
    if hit is not None and hit[1] == x._version and hit[2] == precision:
        return hit[0]
    xs = torch.empty(lib.mdt_conv3d_split_bytes(d), dtype=torch.uint8, device=x.device)
    L.check(lib.mdt_conv3d_split(d, L.ptr(x), L.ptr(xs), L.stream_ptr()))
    try:
        x._mdt_split = (xs, x._version, precision)
    except Exception:
        pass
    return xs


def conv3d_forward(x, weight, bias, stride, padding, relu=False, residual=None, precision=None, algo=None, want_split=False, emit_split=False):
    """y = relu?(conv3d(x, weight) + bias (+ residual)); x logical [N, C, D, H, W]; returns a channels_last_3d tensor. No autograd.
    want_split: also return the split form of x used by the tcgen05 path (or None) so the caller can hand it to the backward pass.
    emit_split: on the tcgen05 path the conv epilogue also writes y as (hi, lo) bf16 planes (mdt_conv3d_fprop_presplit_out) and caches them
    on y, so that a conv consuming y skips its operand-split pass (split_rows_kernel was 8 % of the round-1 step)."""
    lib = L.load()
    L.require_cuda(x, weight, bias, residual)
    precision = DEFAULT_PRECISION if precision is None else precision
    algo = DEFAULT_ALGO if algo is None else algo
    x = x.contiguous(memory_format=_CL3)
    w = weight.contiguous()
    y = torch.empty(_out_shape(x.shape, w.shape, stride, padding), dtype=torch.float32, device=x.device, memory_format=_CL3)
    if residual is not None:
        residual = residual.contiguous(memory_format=_CL3)
    d = _desc(x.shape, w.shape, stride, padding, relu, precision, algo)
    nbytes, which = _plan(lib, d, 0)
    ws = _workspace(nbytes, x.device)
    ev = _ev_start()
    xs = None
    if which == 2:
        xs = _split_of(lib, x, d, precision)
        if emit_split and EMIT_SPLIT:
            ys = torch.empty(lib.mdt_conv3d_out_split_bytes(d), dtype=torch.uint8, device=x.device)
            L.check(lib.mdt_conv3d_fprop_presplit_out(d, L.ptr(xs), L.ptr(w), L.ptr(bias), L.ptr(residual), L.ptr(y), L.ptr(ys), L.ptr(ws), ws.numel(),
                                                      L.stream_ptr()))
            y._mdt_split = (ys, y._version, precision)
        else:
            L.check(lib.mdt_conv3d_fprop_presplit(d, L.ptr(xs), L.ptr(w), L.ptr(bias), L.ptr(residual), L.ptr(y), L.ptr(ws), ws.numel(), L.stream_ptr()))
    elif which == 4 and emit_split and EMIT_SPLIT:
        # pointwise conv feeding a tcgen05 conv: fp32 rows in, fp32 result + its split planes out
        ys = torch.empty(lib.mdt_conv3d_out_split_bytes(d), dtype=torch.uint8, device=x.device)
        L.check(lib.mdt_conv3d_fprop_out(d, L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(residual), L.ptr(y), L.ptr(ys), L.ptr(ws), ws.numel(), L.stream_ptr()))
        y._mdt_split = (ys, y._version, precision)
    else:
        L.check(lib.mdt_conv3d_fprop(d, L.ptr(x), L.ptr(w), L.ptr(bias), L.ptr(residual), L.ptr(y), L.ptr(ws), ws.numel(), L.stream_ptr()))
    _ev_end(ev, (0, tuple(x.shape), tuple(w.shape), tuple(stride), which))
    return (y, xs) if want_split else y


def conv3d_dgrad(dy, weight, x_shape, stride, padding, precision=None, algo=None):
    lib = L.load()
    precision = DEFAULT_PRECISION if precision is None else precision
    algo = DEFAULT_ALGO if algo is None else algo
    dy = dy.contiguous(memory_format=_CL3)
    w = weight.contiguous()
    dx = torch.empty(tuple(x_shape), dtype=torch.float32, device=dy.device, memory_format=_CL3)
    d = _desc(x_shape, w.shape, stride, padding, False, precision, algo)
    nbytes, which = _plan(lib, d, 1)
    ws = _workspace(nbytes, dy.device)
    ev = _ev_start()
    L.check(lib.mdt_conv3d_dgrad(d, L.ptr(dy), L.ptr(w), L.ptr(dx), L.ptr(ws), ws.numel(), L.stream_ptr()))
    _ev_end(ev, (1, tuple(x_shape), tuple(w.shape), tuple(stride), which))
    return dx


def conv3d_wgrad(x, dy, w_shape, stride, padding, want_bias, precision=None, algo=None):
    lib = L.load()
    precision = DEFAULT_PRECISION if precision is None else precision
    algo = DEFAULT_ALGO if algo is None else algo
    x = x.contiguous(memory_format=_CL3)
    dy = dy.contiguous(memory_format=_CL3)
    dw = torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device)
    db = torch.empty(w_shape[0], dtype=torch.float32, device=x.device) if want_bias else None
    d = _desc(x.shape, w_shape, stride, padding, False, precision, algo)
    nbytes, which = _plan(lib, d, 2)
    ws = _workspace(nbytes, x.device)
    ev = _ev_start()
    L.check(lib.mdt_conv3d_wgrad(d, L.ptr(x), L.ptr(dy), L.ptr(dw), L.ptr(db), L.ptr(ws), ws.numel(), L.stream_ptr()))
    _ev_end(ev, (2, tuple(x.shape), tuple(w_shape), tuple(stride), which))
    return dw, db


_bwd_plan_cache = {}


def conv3d_backward(x, gy, y_relu, weight, stride, padding, need_dx, want_bias, want_masked, precision=None, algo=None, x_split=None):
    """Fused backward of one conv (see mdt_conv3d_backward in include/mdt_b200.h).  Returns (dx|None, dw, db|None, gy_masked|None), or None when
    the fused tcgen05 path does not apply to this shape (caller then runs mask + dgrad + wgrad separately)."""
    lib = L.load()
    precision = DEFAULT_PRECISION if precision is None else precision
    algo = DEFAULT_ALGO if algo is None else algo
    d = _desc(x.shape, weight.shape, stride, padding, False, precision, algo)
    key = (id(d), bool(need_dx))
    plan = _bwd_plan_cache.get(key)
    if plan is None:
        plan = (lib.mdt_conv3d_backward_fused(d, int(need_dx)), lib.mdt_conv3d_backward_workspace_bytes(d, int(need_dx)))
        _bwd_plan_cache[key] = plan
    if not plan[0]:
        return None
    x = x.contiguous(memory_format=_CL3)
    gy = gy.contiguous(memory_format=_CL3)
    w = weight.contiguous()
    dev = x.device
    dx = torch.empty(tuple(x.shape), dtype=torch.float32, device=dev, memory_format=_CL3) if need_dx else None
    dw = torch.empty(tuple(w.shape), dtype=torch.float32, device=dev)
    db = torch.empty(w.shape[0], dtype=torch.float32, device=dev) if want_bias else None
    gm = torch.empty_like(gy) if want_masked else None
    ws = _workspace(plan[1], dev)
    ev = _ev_start()
    L.check(lib.mdt_conv3d_backward(d, L.ptr(x), L.ptr(x_split), L.ptr(gy), L.ptr(y_relu), L.ptr(w), L.ptr(dx), L.ptr(dw), L.ptr(db), L.ptr(gm),
                                    L.ptr(ws), ws.numel(), L.stream_ptr()))
    _ev_end(ev, (3, tuple(x.shape), tuple(w.shape), tuple(stride), _plan(lib, d, 2)[1]))
    return dx, dw, db, gm


class _Conv3dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, stride, padding, relu, precision, algo, emit_split=False):
        prec = DEFAULT_PRECISION if precision is None else precision
        y, xs = conv3d_forward(x, weight, bias, stride, padding, relu, residual, prec, algo, want_split=True, emit_split=emit_split)
        ctx.cfg = (stride, padding, relu, prec, algo, tuple(x.shape), bias is not None, residual is not None)
        ctx.xs = xs   # split form of x: the weight gradient reads it instead of splitting x again
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        stride, padding, relu, precision, algo, x_shape, has_bias, has_res = ctx.cfg
        gy = gy.contiguous(memory_format=_CL3)
        need_w = ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2])
        if need_w:
            want_masked = has_res and ctx.needs_input_grad[3] and relu
            fused = conv3d_backward(x, gy, y if relu else None, weight, stride, padding, ctx.needs_input_grad[0], has_bias, want_masked, precision, algo,
                                    x_split=getattr(ctx, 'xs', None))
            if fused is not None:
                gx, gw, gb, gm = fused
                gres = (gm if relu else gy) if (has_res and ctx.needs_input_grad[3]) else None
                return gx, gw, gb, gres, None, None, None, None, None, None
        if relu:
            gy = torch.ops.aten.threshold_backward(gy, y, 0.0)  # ReLU mask (elementwise, HBM-bound)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = conv3d_dgrad(gy, weight, x_shape, stride, padding, precision, algo)
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            gw, gb = conv3d_wgrad(x, gy, tuple(weight.shape), stride, padding, has_bias, precision, algo)
        gres = gy if (has_res and ctx.needs_input_grad[3]) else None
        return gx, gw, gb, gres, None, None, None, None, None, None


def _triple(v):
    return tuple(int(i) for i in v) if isinstance(v, (tuple, list)) else (int(v),) * 3


def _pair(v):
    return tuple(int(i) for i in v) if isinstance(v, (tuple, list)) else (int(v),) * 2


class Conv3d(nn.Module):
    """nn.Conv3d-compatible module (same parameter names, shapes and default initialisation) running on libmdt_b200."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, fused_relu=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _triple(kernel_size), _triple(stride), _triple(padding)
        self.fused_relu = fused_relu
        self.precision = None
        self.algo = None
        self.emit_split = True      # set False on convs whose output never feeds another conv (final layers, laterals feeding adds)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # identical to torch.nn.modules.conv._ConvNd.reset_parameters, so a seeded build draws the same numbers as the reference's nn.Conv3d
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x, residual=None):
        return _Conv3dFn.apply(x, self.weight, self.bias, residual, self.stride, self.padding, self.fused_relu, self.precision, self.algo, self.emit_split)

    def extra_repr(self):
        return "{}, {}, kernel_size={}, stride={}, padding={}, fused_relu={}".format(
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.fused_relu)


class Conv2d(nn.Module):
    """nn.Conv2d-compatible module: runs as a depth-1 conv3d on the same kernels ([N, C, H, W] <-> [N, C, 1, H, W])"""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, fused_relu=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = _pair(kernel_size), _pair(stride), _pair(padding)
        self.fused_relu = fused_relu
        self.precision = None
        self.algo = None
        self.emit_split = False     # the split cache is keyed on the 5-D tensor; 2-D convs re-wrap their activations (unsqueeze) per call
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(in_channels * self.kernel_size[0] * self.kernel_size[1])
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x, residual=None):
        res = residual.unsqueeze(2) if residual is not None else None
        y = _Conv3dFn.apply(x.unsqueeze(2), self.weight.unsqueeze(2), self.bias, res, (1,) + self.stride, (0,) + self.padding,
                            self.fused_relu, self.precision, self.algo)
        return y.squeeze(2)


def pointwise_eligible(c_in, c_out, ks=1, stride=1, pad=0):
    """mirror of conv_pw_supported (csrc/conv3d_pw.cu): 1x1x1 stride-1 convs with cin * cout <= 2592 run on the fp32 streaming kernels, which read
    fp32 rows: a conv feeding ONLY such convs need not emit split planes"""
    one = lambda v, t: all(int(i) == t for i in (v if isinstance(v, (tuple, list)) else (v,)))
    return one(ks, 1) and one(stride, 1) and one(pad, 0) and c_in * c_out <= 2592


def no_split_consumer(module):
    """mark the conv inside an NDConvGenerator result (bare conv or Sequential) as NOT feeding another conv: its epilogue then skips the
    bf16 split planes of the result (final layers of heads, laterals that only feed adds / up-sampling)"""
    for m in module.modules():
        if isinstance(m, (Conv3d, Conv2d)):
            m.emit_split = False
    return module


class FusedReLU(nn.Module):
    """Parameter-free marker standing where the reference puts nn.ReLU(inplace=True): the ReLU already ran in the conv epilogue."""

    def forward(self, x):
        return x


class NDConvGenerator(object):
    """generic conv(+norm)(+relu) factory, 2D or 3D — same call shape and module nesting as utils/model_utils.py:732-781"""

    def __init__(self, dim):
        self.dim = dim

    def __call__(self, c_in, c_out, ks, pad=0, stride=1, norm=None, relu='relu'):
        fuse = (relu == 'relu') and norm is None
        conv_cls = Conv2d if self.dim == 2 else Conv3d
        conv = conv_cls(c_in, c_out, kernel_size=ks, padding=pad, stride=stride, fused_relu=fuse)
        if norm is not None:
            if norm == 'instance_norm':
                norm_layer = nn.InstanceNorm2d(c_out) if self.dim == 2 else nn.InstanceNorm3d(c_out)
            elif norm == 'batch_norm':
                norm_layer = nn.BatchNorm2d(c_out) if self.dim == 2 else nn.BatchNorm3d(c_out)
            else:
                raise ValueError('norm type as specified in configs is not implemented... {}'.format(norm))
            conv = nn.Sequential(conv, norm_layer)
        if relu is not None:
            if relu == 'relu':
                relu_layer = FusedReLU() if fuse else nn.ReLU(inplace=True)
            elif relu == 'leaky_relu':
                relu_layer = nn.LeakyReLU(inplace=True)
            else:
                raise ValueError('relu type as specified in configs is not implemented...')
            conv = nn.Sequential(conv, relu_layer)
        return conv
