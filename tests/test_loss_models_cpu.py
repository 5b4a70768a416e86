"""CPU statements of the ALGORITHMS inside csrc/loss_ops.cu (the kernels themselves run on the GPU: tests/test_loss_ops_gpu.py):

  * the analytic gradient the segmentation-loss backward kernel evaluates per voxel,
        dL/dz_k = p_k (q_k - sum_c q_c p_c) + g_ce / N (p_k - y_k),   q_c = g_dice / (C-1) (2 y_c / D_c - (2 I_c + s) fpw / D_c^2)  (c >= 1),
    against torch autograd of the reference's formulas (utils/model_utils.py:833-858, retina_unet.py:446-448) in fp64;
  * the SHEM selection chain (keys = 1 + bits(max fg probability) for negatives, top-k per 4096-anchor chunk, top-k of the candidates, pool of
    shem_poolsize * max(1, n_pos), sample = the negative_count smallest uniform keys of the pool) against the torch-op formulation of
    retina_unet.compute_class_loss, which tests/test_model_golden.py pins to the reference (retina_unet.py:126-164, model_utils.py:674-691).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from medicaldetectiontoolkit_b200 import model_utils as mutils
from medicaldetectiontoolkit_b200 import retina_unet as RU


@pytest.mark.parametrize("C,fpw", [(2, 1.0), (4, 0.7)])
def test_seg_loss_backward_formula(C, fpw):
    torch.manual_seed(0)
    smooth = 1e-6
    z = torch.randn(2, C, 6, 5, 4, dtype=torch.float64, requires_grad=True)
    t = torch.randint(0, C, (2, 6, 5, 4))
    p = F.softmax(z, 1)
    y = F.one_hot(t, C).movedim(-1, 1).double()
    dice = mutils.batch_dice(p, y, false_positive_weight=fpw, smooth=smooth)
    ce = F.cross_entropy(z, t)
    g_dice, g_ce = -0.5, 0.5                                    # L = (1 - dice) / 2 + ce / 2
    (g_dice * dice + g_ce * ce).backward()
    # the kernel's per-voxel formula from the per-class sums
    pd, yd = p.detach(), y
    axes = (0, 2, 3, 4)
    I, P, T = (pd * yd).sum(axes), pd.sum(axes), yd.sum(axes)
    D = fpw * P + T + smooth
    qa = torch.zeros(C, dtype=torch.float64)
    qb = torch.zeros(C, dtype=torch.float64)
    qa[1:] = g_dice * 2.0 / (D[1:] * (C - 1))
    qb[1:] = g_dice * (2.0 * I[1:] + smooth) * fpw / (D[1:] ** 2 * (C - 1))
    q = qa.view(1, C, 1, 1, 1) * yd - qb.view(1, C, 1, 1, 1)
    dot = (q * pd).sum(1, keepdim=True)
    n_vox = t.numel()
    grad = pd * (q - dot) + g_ce / n_vox * (pd - yd)
    assert float((grad - z.grad).abs().max() / z.grad.abs().max()) < 1e-12


def _shem_model(logits, match, pos_ids, rand, k_pos, poolsize, chunk=4096):
    """numpy model of shem_level1 / shem_reduce / shem_final: returns (sampled anchor ids in sample order, pool anchor ids in pool order)"""
    A = logits.shape[0]
    z = logits.astype(np.float32)
    e = np.exp(z - z.max(1, keepdims=True)).astype(np.float32)
    score = (e[:, 1:].max(1) / e.sum(1)).astype(np.float32)
    key = np.where(match == -1, score.view(np.uint32).astype(np.uint64) + 1, 0)
    k_pool = min(A, poolsize * k_pos)
    cands = []
    for c0 in range(0, A, chunk):                               # level 1: chunk top-k (stable: ties keep ascending anchor order)
        ids = np.arange(c0, min(A, c0 + chunk))
        o = np.argsort(-key[ids].astype(np.int64), kind="stable")[:k_pool]
        cands.append(ids[o][key[ids[o]] > 0])
    cand = np.concatenate(cands)
    while cand.size > chunk:                                    # further levels
        nxt = []
        for c0 in range(0, cand.size, chunk):
            ids = cand[c0:c0 + chunk]
            nxt.append(ids[np.argsort(-key[ids].astype(np.int64), kind="stable")[:k_pool]])
        cand = np.concatenate(nxt)
    pool = cand[np.argsort(-key[cand].astype(np.int64), kind="stable")[:k_pool]]
    n_pos, n_neg = int((match > 0).sum()), int((match == -1).sum())
    negative_count = max(1, n_pos)
    pool_size = min(poolsize * negative_count, n_neg)
    keys = np.full(k_pool, 2.0, dtype=np.float32)
    m = min(pool.size, pool_size, k_pool)
    keys[:m] = rand[:m]
    order = np.argsort(keys, kind="stable")[: min(k_pool, k_pos)]
    sel = [int(pool[j]) for r, j in enumerate(order) if r < negative_count and keys[j] < 1.5]
    return sel, pool


@pytest.mark.parametrize("A,n_cls,n_pos,max_pos", [(20000, 3, 3, 3), (9000, 2, 0, 3), (5000, 3, 2, 8), (700, 2, 20, 32)])
def test_shem_selection_chain_equals_the_torch_formulation(A, n_cls, n_pos, max_pos):
    rs = np.random.RandomState(A)
    logits = (rs.randn(A, n_cls) * 0.5).astype(np.float32)
    match = np.full(A, -1, dtype=np.int32)
    match[rs.permutation(A)[: A // 5]] = 0
    pos = np.sort(rs.permutation(A)[:n_pos])
    match[pos] = rs.randint(1, n_cls, size=n_pos)
    k_pos = min(A, max_pos)
    k_pool = min(A, 20 * k_pos)
    gen = torch.Generator().manual_seed(3)
    rand = torch.rand(k_pool, generator=gen).numpy()
    sel, pool = _shem_model(logits, match, pos, rand, k_pos, 20)
    gen = torch.Generator().manual_seed(3)                       # the torch formulation draws the same k_pool keys
    loss, neg_ix = RU.compute_class_loss(torch.from_numpy(match), torch.from_numpy(logits), shem_poolsize=20, max_pos=max_pos, generator=gen,
                                         pos_ids=torch.from_numpy(pos))
    neg_all = np.nonzero(match == -1)[0]
    picked = neg_all[neg_ix[neg_ix >= 0].numpy()]
    assert sorted(picked.tolist()) == sorted(sel)
    # and the loss the final kernel evaluates on those rows
    lt = torch.from_numpy(logits)
    ce_pos = F.cross_entropy(lt[pos], torch.from_numpy(match[pos]).long(), reduction='sum') / max(1, len(pos)) if len(pos) else torch.zeros(())
    ce_neg = F.cross_entropy(lt[sel], torch.zeros(len(sel), dtype=torch.long), reduction='sum') / max(1, len(sel))
    assert abs(float(loss) - float((ce_pos + ce_neg) / 2)) < 1e-6
