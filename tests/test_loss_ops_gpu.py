"""Loss-side and resampling kernels (csrc/loss_ops.cu, csrc/resample.cu) against plain PyTorch restatements of the reference formulas
(floating-point kernels => torch fp64 reference, tolerances written at each check):

  * max pooling k3 / s(2,2,1) / p1 and nearest x2 up-sampling of models/backbone.py:63-64,147-153: forward bit-exact, backward
    up to the order of the fp32 sums (<= 12 terms; ATen itself uses atomics): 1e-5;
  * segmentation loss = batch_dice(softmax, one_hot) + cross_entropy (utils/model_utils.py:833-858, retina_unet.py:446-448): values 1e-6,
    gradient 1e-5 relative to its max;
  * SHEM class loss (retina_unet.py:126-164, model_utils.py:674-691): the fused kernel chain must select the SAME pool and the SAME sample as
    the torch-op formulation fed with the same uniform keys (that formulation is pinned to the reference in tests/test_model_golden.py).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from medicaldetectiontoolkit_b200 import backbone as B
from medicaldetectiontoolkit_b200 import native_ops
from medicaldetectiontoolkit_b200 import retina_unet as RU

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CL3 = torch.channels_last_3d


@pytest.mark.parametrize("shape", [(2, 18, 16, 16, 24), (1, 36, 10, 14, 7), (2, 5, 8, 8, 8)])
def test_maxpool3d_matches_torch(shape):
    torch.manual_seed(0)
    x = torch.randn(shape, device=DEV).contiguous(memory_format=CL3).requires_grad_(True)
    pool = B.MaxPool(3, 3, (2, 2, 1), 1)
    y = pool(x)
    xr = x.detach().clone().requires_grad_(True)
    yr = F.max_pool3d(xr, 3, (2, 2, 1), 1)
    assert y.shape == yr.shape
    assert torch.equal(y, yr)                                       # selection: bit-exact
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-5, atol=1e-5)   # fp32 sums of <= 12 gradients in a different order (ATen: atomics)
    # ties and -inf / NaN follow ATen's update rule (first maximum in (d, h, w) order; NaN propagates)
    t = torch.zeros(1, 4, 6, 6, 4, device=DEV).contiguous(memory_format=CL3).requires_grad_(True)
    tr = t.detach().clone().requires_grad_(True)
    pool(t).sum().backward()
    F.max_pool3d(tr, 3, (2, 2, 1), 1).sum().backward()
    assert torch.equal(t.grad, tr.grad)
    n = torch.randn(1, 4, 6, 6, 4, device=DEV)
    n[0, 1, 2, 3, 1] = float('nan')
    n[0, 2, 0, 0, 0] = float('-inf')
    a, b = pool(n.contiguous(memory_format=CL3)), F.max_pool3d(n, 3, (2, 2, 1), 1)
    assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.nan_to_num(a, 7.0), torch.nan_to_num(b, 7.0))


def test_maxpool2d_matches_torch():
    torch.manual_seed(1)
    x = torch.randn(2, 12, 20, 28, device=DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = B.MaxPool(2, 3, 2, 1)(x), F.max_pool2d(xr, 3, 2, 1)
    assert torch.equal(y, yr)
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 36, 4, 4, 8), (1, 18, 3, 5, 7), (2, 12, 6, 10)])
def test_nearest_up2_matches_torch(shape):
    torch.manual_seed(2)
    fmt = CL3 if len(shape) == 5 else torch.channels_last
    x = torch.randn(shape, device=DEV).contiguous(memory_format=fmt).requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    y, yr = B.nearest_up2(x), F.interpolate(xr, scale_factor=2)
    assert torch.equal(y, yr)
    g = torch.randn_like(yr)
    y.backward(g)
    yr.backward(g)
    assert torch.allclose(x.grad, xr.grad, rtol=0, atol=1e-5)


def _seg_reference(logits, seg, fpw=1.0, smooth=1e-6):
    """the reference's formulas in fp64: batch_dice(F.softmax(l, 1), one_hot) (model_utils.py:833-858) and F.cross_entropy(l, seg)"""
    l = logits.detach().double().requires_grad_(True)
    c = l.shape[1]
    p = F.softmax(l, dim=1)
    y = F.one_hot(seg.long(), c).movedim(-1, 1).double()
    axes = (0,) + tuple(range(2, l.dim()))
    inter = (p * y).sum(axes)
    den = (fpw * p + y).sum(axes)
    dice = torch.mean(((2 * inter + smooth) / (den + smooth))[1:])
    ce = F.cross_entropy(l, seg.long())
    return l, dice, ce


@pytest.mark.parametrize("shape,fmt", [((2, 2, 32, 32, 24), "cl"), ((2, 3, 16, 16, 8), "nc"), ((1, 4, 40, 24), "nc"), ((2, 2, 128, 128, 128), "cl")])
def test_seg_loss_matches_reference_formulas(shape, fmt):
    torch.manual_seed(4)
    logits = torch.randn(shape, device=DEV) * 2
    if fmt == "cl":
        logits = logits.contiguous(memory_format=CL3 if len(shape) == 5 else torch.channels_last)
    logits.requires_grad_(True)
    seg = (torch.rand((shape[0],) + shape[2:], device=DEV) < 0.1).to(torch.uint8) * torch.randint(1, shape[1], (shape[0],) + shape[2:], device=DEV,
                                                                                                  dtype=torch.uint8)
    dice, ce = native_ops.seg_loss(logits, seg.contiguous())
    l64, dice_r, ce_r = _seg_reference(logits, seg)
    assert abs(dice.item() - dice_r.item()) < 1e-6 * max(1.0, abs(dice_r.item()))
    assert abs(ce.item() - ce_r.item()) < 2e-6 * max(1.0, abs(ce_r.item()))
    loss = (1 - dice) * 0.5 + ce * 0.5                                # how retina_unet.py:448 combines them
    loss.backward()
    ((1 - dice_r) * 0.5 + ce_r * 0.5).backward()
    gr = l64.grad
    err = (logits.grad.double() - gr).abs().max().item() / gr.abs().max().item()
    assert err < 1e-5, err
    # determinism: fixed-order reduction => bit-identical repeats
    d2, c2 = native_ops.seg_loss(logits.detach(), seg.contiguous())
    assert d2.item() == dice.item() and c2.item() == ce.item()


def _shem_case(A, n_cls, n_pos, seed, neutral=0.2):
    g = torch.Generator(device=DEV)
    g.manual_seed(seed)
    logits = torch.randn(A, n_cls, device=DEV, generator=g) * 1.5
    match = torch.full((A,), -1, dtype=torch.int32, device=DEV)
    if neutral > 0:
        match[torch.randperm(A, device=DEV, generator=g)[: int(A * neutral)]] = 0
    pos = torch.sort(torch.randperm(A, device=DEV, generator=g)[:n_pos])[0]
    if n_pos:
        match[pos] = torch.randint(1, n_cls, (n_pos,), device=DEV, generator=g, dtype=torch.int32)
    return logits, match, pos


@pytest.mark.parametrize("A,n_cls,n_pos,max_pos,poolsize", [
    (50000, 3, 3, 3, 20),            # Retina U-Net shape of the loss (cfg2: max_pos 3 -> pool 60)
    (50001, 3, 0, 3, 20),            # no positive: one negative is still drawn (np.max((1, n_pos)))
    (4097, 2, 16, 16, 20),           # RPN shape (mrcnn.py:176-213): pool 320, two chunks, second almost empty
    (3000, 3, 2, 8, 20),             # fewer positives than the cap: pool_size < k_pool
    (1347840, 3, 3, 3, 20),          # all anchors of cfg2
    (900, 2, 40, 64, 20),            # pool larger than the number of negatives
])
def test_shem_fused_equals_torch_formulation(A, n_cls, n_pos, max_pos, poolsize):
    logits, match, pos = _shem_case(A, n_cls, n_pos, seed=A % 97)
    res = {}
    for fused in (True, False):
        RU.FUSED_LOSSES = fused
        try:
            l = logits.clone().requires_grad_(True)
            gen = torch.Generator(device=DEV)
            gen.manual_seed(11)
            loss, neg_ix = RU.compute_class_loss(match, l, shem_poolsize=poolsize, max_pos=max_pos, generator=gen, pos_ids=pos)
            (loss * 1.7).backward()
            res[fused] = (loss.item(), neg_ix.clone(), l.grad.clone())
        finally:
            RU.FUSED_LOSSES = True
    (lf, nf, gf), (lt, nt, gt) = res[True], res[False]
    assert abs(lf - lt) < 2e-6 * max(1.0, abs(lt)), (lf, lt)
    assert sorted(nf[nf >= 0].tolist()) == sorted(nt[nt >= 0].tolist())          # the same sampled negatives (ranks inside the negative subset)
    assert (nf >= 0).sum().item() == min(max(1, n_pos), int((match == -1).sum().item()), nf.numel())
    assert torch.allclose(gf, gt, rtol=1e-5, atol=1e-7)
    assert int((gf.abs().sum(1) > 0).sum().item()) <= 2 * max_pos


def test_shem_pool_is_the_top_of_the_negatives():
    """independent of the torch formulation: the sampled negatives come from the poolsize * n_pos best-scoring negatives"""
    A, n_pos = 200000, 3
    logits, match, pos = _shem_case(A, 3, n_pos, seed=5)
    loss, neg_ix = RU.compute_class_loss(match, logits, shem_poolsize=20, max_pos=3, pos_ids=pos)
    neg_all = torch.nonzero(match == -1).squeeze(1)
    picked = neg_all[neg_ix[neg_ix >= 0]]
    assert picked.numel() == n_pos and picked.unique().numel() == n_pos
    probs = F.softmax(logits[neg_all], 1)[:, 1:].max(1)[0]
    pool = neg_all[probs.sort(descending=True)[1][:60]]
    assert set(picked.tolist()) <= set(pool.tolist())
    want = (F.cross_entropy(logits[pos], match[pos].long()) + F.cross_entropy(logits[picked], torch.zeros(n_pos, dtype=torch.long, device=DEV))) / 2
    assert abs(loss.item() - want.item()) < 1e-5


def test_seg_loss_single_term_backward():
    """only one of the two outputs enters the loss: autograd hands None for the other (dice-only / CE-only configurations)"""
    torch.manual_seed(9)
    logits = torch.randn(1, 3, 8, 8, 8, device=DEV).requires_grad_(True)
    seg = torch.randint(0, 3, (1, 8, 8, 8), device=DEV, dtype=torch.uint8)
    dice, ce = native_ops.seg_loss(logits, seg)
    (1 - dice).backward()
    l64, dice_r, ce_r = _seg_reference(logits, seg)
    (1 - dice_r).backward()
    assert float((logits.grad.double() - l64.grad).abs().max() / l64.grad.abs().max()) < 1e-5
    logits.grad = None
    dice, ce = native_ops.seg_loss(logits, seg)
    ce.backward()
    l64, dice_r, ce_r = _seg_reference(logits, seg)
    ce_r.backward()
    assert float((logits.grad.double() - l64.grad).abs().max() / l64.grad.abs().max()) < 1e-5
