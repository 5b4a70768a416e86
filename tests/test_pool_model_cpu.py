"""Executable model of the max-pool kernels' index arithmetic (csrc/resample.cu: maxpool_fwd_kernel / maxpool_bwd_kernel) against torch's CPU
max_pool3d forward + autograd backward for every (kernel, stride, padding) family the C entry point accepts (pad <= kernel / 2, floor mode).
The GPU tests run the BASELINE shapes (k3 s(2,2,1) p1, 2D k3 s2 p1); this pins the GENERAL window formulas of the backward gather:
    windows containing input index i along one axis:  o in [max(0, (i + p - k + s) // s), min(O - 1, (i + p) // s)]."""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def model_forward(x, k, s, p):
    N, C, D, H, W = x.shape
    O = [(n + 2 * pp - kk) // ss + 1 for n, kk, ss, pp in zip((D, H, W), k, s, p)]
    y = np.full((N, C, *O), -np.inf, dtype=x.dtype)
    arg = np.full((N, C, *O), -1, dtype=np.int64)
    for od, oh, ow in itertools.product(*[range(o) for o in O]):
        for a, b, e in itertools.product(range(k[0]), range(k[1]), range(k[2])):
            d, h, w = od * s[0] - p[0] + a, oh * s[1] - p[1] + b, ow * s[2] - p[2] + e
            if not (0 <= d < D and 0 <= h < H and 0 <= w < W):
                continue
            v = x[:, :, d, h, w]
            cur = y[:, :, od, oh, ow]
            upd = (v > cur) | np.isnan(v) | (arg[:, :, od, oh, ow] < 0)          # ATen's rule: first maximum wins, NaN propagates
            y[:, :, od, oh, ow] = np.where(upd, v, cur)
            arg[:, :, od, oh, ow] = np.where(upd, (a * k[1] + b) * k[2] + e, arg[:, :, od, oh, ow])
    return y, arg, O


def model_backward(gy, arg, xshape, k, s, p, O):
    N, C, D, H, W = xshape
    gx = np.zeros(xshape, dtype=gy.dtype)
    rng = lambda i, kk, ss, pp, o: range(max(0, (i + pp - kk + ss) // ss), min(o - 1, (i + pp) // ss) + 1)
    for d, h, w in itertools.product(range(D), range(H), range(W)):
        for od in rng(d, k[0], s[0], p[0], O[0]):
            for oh in rng(h, k[1], s[1], p[1], O[1]):
                for ow in rng(w, k[2], s[2], p[2], O[2]):
                    off = ((d - (od * s[0] - p[0])) * k[1] + (h - (oh * s[1] - p[1]))) * k[2] + (w - (ow * s[2] - p[2]))
                    assert 0 <= off < k[0] * k[1] * k[2]                      # the window really contains (d, h, w)
                    gx[:, :, d, h, w] += np.where(arg[:, :, od, oh, ow] == off, gy[:, :, od, oh, ow], 0)
    return gx


@pytest.mark.parametrize("k,s,p", [((3, 3, 3), (2, 2, 1), (1, 1, 1)), ((1, 3, 3), (1, 2, 2), (0, 1, 1)), ((2, 2, 2), (2, 2, 2), (0, 0, 0)),
                                   ((3, 3, 3), (1, 1, 1), (1, 1, 1)), ((3, 2, 1), (2, 1, 1), (1, 0, 0)), ((3, 3, 3), (3, 3, 3), (0, 1, 1)),
                                   ((2, 3, 2), (1, 2, 2), (1, 1, 0))])
def test_pool_window_arithmetic(k, s, p):
    rs = np.random.RandomState(sum(k) * 7 + sum(s))
    x = rs.randn(1, 2, 7, 6, 5)
    xt = torch.from_numpy(x).requires_grad_(True)
    yt = F.max_pool3d(xt, k, s, p)
    y, arg, O = model_forward(x, k, s, p)
    assert tuple(yt.shape[2:]) == tuple(O) and np.array_equal(y, yt.detach().numpy())
    gy = rs.randn(*yt.shape)
    yt.backward(torch.from_numpy(gy))
    gx = model_backward(gy, arg, x.shape, k, s, p, O)
    assert np.allclose(gx, xt.grad.numpy(), rtol=0, atol=1e-12)


def test_pool_ties_take_the_first_maximum():
    x = np.zeros((1, 1, 4, 4, 4))
    xt = torch.from_numpy(x).requires_grad_(True)
    k, s, p = (3, 3, 3), (2, 2, 1), (1, 1, 1)
    yt = F.max_pool3d(xt, k, s, p)
    yt.sum().backward()
    y, arg, O = model_forward(x, k, s, p)
    gx = model_backward(np.ones_like(y), arg, x.shape, k, s, p, O)
    assert np.array_equal(gx, xt.grad.numpy())
