"""Model surface pinned to the REFERENCE: every comparison here is against fixtures produced by the reference's own, unmodified
models/retina_unet.py, models/mrcnn.py, models/retina_net.py and utils/model_utils.py (tests/golden/make_model_golden.py, run in the
build container on CPU with the shims of tests/golden/ref_shims.py).  Inputs come from tests/golden/golden_inputs.py (identical bits
on both sides); sampling is neutralised on both sides (model_utils.SAMPLING = "identity" == torch.randperm -> arange in the generator).

CPU tests (-m "not gpu"): the pure-torch functions (losses, SHEM, box IoU, dice, ...).  GPU tests: everything that runs libmdt_b200
(NMS, RoIAlign, matching, convs) — function level with exact discrete results, whole models at the small and the BASELINE sizes
(cfg2 = Retina U-Net 2x128^3, cfg3 = Mask R-CNN 2x128^3 / 512 proposals).

Tolerances: logits/deltas 1e-4 of max|ref| (north_star), losses 1e-4 relative (2e-3 where a loss is a mean over <= 6 samples of
O(1e-2) values), gradients relative L2 <= 1e-2, same norm and direction (a pre-activation within rounding distance of zero flips its ReLU mask
under any re-association; measured 2e-5..7e-3 from the heads down to the stem, see test_model_gpu._rel_l2), boxes/indices exact.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import detweights  # noqa: E402
import golden_inputs as GI  # noqa: E402

from medicaldetectiontoolkit_b200 import model_utils as mutils  # noqa: E402

DEV = "cuda:0"
gpu = pytest.mark.gpu
T = torch.from_numpy
sub = detweights.subsample


@pytest.fixture(scope="module")
def funcs():
    return np.load(os.path.join(GOLD, "model_funcs.npz"))


@pytest.fixture(autouse=True)
def identity_sampling():
    old = mutils.SAMPLING
    mutils.SAMPLING = "identity"
    yield
    mutils.SAMPLING = old


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)) if a.size else 0.0


def _rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _rows_sorted(a):
    a = np.asarray(a, dtype=np.float64)
    return a[np.lexsort(a.T[::-1])] if a.size else a


def _np(x):
    return x.detach().cpu().numpy()


# ================================================================================================== CPU: pure-torch functions
def test_class_loss_vs_reference(funcs):
    """retina_unet.compute_class_loss / mrcnn.compute_rpn_class_loss incl. SHEM pool + sampled negative indices
    (models/retina_unet.py:126-164, models/mrcnn.py:176-213, utils/model_utils.py:674-691); 'many_pos' has 100 positives with the
    reference signature (no max_pos) — ADVICE r1: must not be capped at 64"""
    from medicaldetectiontoolkit_b200 import mrcnn, retina_unet
    for name, (m, lg, pool) in GI.class_loss_cases().items():
        if lg.shape[1] == 2:
            loss, neg = mrcnn.compute_rpn_class_loss(T(m), T(lg), pool)
        else:
            loss, neg = retina_unet.compute_class_loss(T(m), T(lg), pool)
        want = float(funcs["class_loss__" + name][0])
        assert abs(float(loss) - want) <= 1e-5 * max(1.0, abs(want)), (name, float(loss), want)
        got_neg = _np(neg)
        got_neg = got_neg[got_neg >= 0]
        assert got_neg.tolist() == funcs["class_neg__" + name].tolist(), name
        # the fixed-shape fast path (max_pos = the matching's cap) must agree whenever the cap holds
        n_pos = int((m > 0).sum())
        if lg.shape[1] == 3 and n_pos > 0:
            loss2, _ = retina_unet.compute_class_loss(T(m), T(lg), pool, max_pos=n_pos + 2)
            assert abs(float(loss2) - want) <= 1e-5 * max(1.0, abs(want)), name
            # ... and so must the path fed with the matching's own list of positives (no search over the anchor array)
            loss3, neg3 = retina_unet.compute_class_loss(T(m), T(lg), pool, max_pos=n_pos, pos_ids=torch.nonzero(T(m) > 0).squeeze(1))
            assert abs(float(loss3) - want) <= 1e-5 * max(1.0, abs(want)) and _np(neg3)[_np(neg3) >= 0].tolist() == got_neg.tolist(), name


def test_bbox_loss_vs_reference(funcs):
    from medicaldetectiontoolkit_b200 import mrcnn, retina_unet
    for name, (t, p, m) in GI.bbox_loss_cases().items():
        a = float(retina_unet.compute_bbox_loss(T(t), T(p), T(m)))
        b = float(mrcnn.compute_rpn_bbox_loss(T(t), T(p), T(m)))
        assert abs(a - float(funcs["bbox_loss__" + name][0])) <= 1e-5, name
        assert abs(b - float(funcs["rpn_bbox_loss__" + name][0])) <= 1e-5, name


def test_mrcnn_head_losses_vs_reference(funcs):
    from medicaldetectiontoolkit_b200 import mrcnn
    t_cls, logits, t_del, p_del, t_m, p_m = GI.mrcnn_loss_inputs()
    z = np.zeros_like(t_cls)
    got = {
        "mrcnn_class_loss": mrcnn.compute_mrcnn_class_loss(T(t_cls), T(logits)),
        "mrcnn_bbox_loss": mrcnn.compute_mrcnn_bbox_loss(T(t_del), T(p_del), T(t_cls)),
        "mrcnn_mask_loss": mrcnn.compute_mrcnn_mask_loss(T(t_m), T(p_m), T(t_cls)),
        "mrcnn_bbox_loss_nopos": mrcnn.compute_mrcnn_bbox_loss(T(t_del), T(p_del), T(z)),
        "mrcnn_mask_loss_nopos": mrcnn.compute_mrcnn_mask_loss(T(t_m), T(p_m), T(z)),
    }
    for k, v in got.items():
        assert abs(float(v) - float(funcs[k][0])) <= 1e-5 * max(1.0, abs(float(funcs[k][0]))), k


def test_utils_vs_reference(funcs):
    """bbox_overlaps_{2D,3D}, unique1d, batch_dice, get_one_hot_encoding, shem, log2 (utils/model_utils.py:430-501,645-663,674-691,785-858)"""
    u = GI.utils_inputs()
    assert _rel(_np(mutils.bbox_overlaps_3D(T(u["b3a"]), T(u["b3b"]))), funcs["overlaps3"]) <= 1e-6
    assert _rel(_np(mutils.bbox_overlaps_2D(T(u["b2a"]), T(u["b2b"]))), funcs["overlaps2"]) <= 1e-6
    assert _np(mutils.unique1d(T(u["uniq"]))).tolist() == funcs["unique1d"].tolist()
    ohe = mutils.get_one_hot_encoding(u["dice_seg"], 3)
    assert ohe.sum(axis=(0, 2, 3, 4)).tolist() == funcs["one_hot_sum"].tolist()
    ohe_dev = torch.nn.functional.one_hot(T(u["dice_seg"]).long()[:, 0], 3).movedim(-1, 1)      # the on-device form train_forward uses
    assert np.array_equal(_np(ohe_dev), ohe)
    assert abs(float(mutils.batch_dice(T(u["dice_pred"]), T(ohe).float())) - float(funcs["batch_dice"][0])) <= 1e-6
    assert abs(float(mutils.batch_dice(T(u["dice_pred"]), T(ohe).float(), false_positive_weight=2.0)) - float(funcs["batch_dice_fpw"][0])) <= 1e-6
    assert _np(mutils.shem(T(u["shem_probs"]), 7, 10)).tolist() == funcs["shem"].tolist()
    assert _np(mutils.shem(T(u["shem_probs"][:30]), 7, 10)).tolist() == funcs["shem_small_pool"].tolist()
    assert _rel(_np(mutils.log2(T(u["log2_x"]))), funcs["log2"]) <= 1e-6


# ================================================================================================== GPU: function level
def _assert_detections(got, want, dim, score_tol=1e-5):
    got, want = _rows_sorted(got), _rows_sorted(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.array_equal(got[:, :2 * dim + 2], want[:, :2 * dim + 2])        # rounded pixel boxes, batch index, class id: exact
    assert np.abs(got[:, -1] - want[:, -1]).max() <= score_tol


@gpu
def test_retina_refine_detections_vs_reference(funcs):
    """models/retina_unet.py:194-271: global top-k, decode, clip, round, per (element, class) NMS at 1e-5, top-30 per element"""
    from medicaldetectiontoolkit_b200 import retina_unet
    cf, probs, deltas, bix = GI.retina_refine_inputs()
    anchors = T(mutils.generate_pyramid_anchors(None, cf)).float().to(DEV)
    det = retina_unet.refine_detections(anchors, T(probs).to(DEV), T(deltas).to(DEV), T(bix).to(DEV), cf)
    _assert_detections(_np(det), funcs["retina_refine"], 3)


@gpu
def test_proposal_layer_vs_reference(funcs):
    """models/mrcnn.py:297-369"""
    from medicaldetectiontoolkit_b200 import mrcnn
    cf, probs, deltas, count = GI.proposal_inputs()
    anchors = T(mutils.generate_pyramid_anchors(None, cf)).float().to(DEV)
    boxes, props = mrcnn.proposal_layer(T(probs).to(DEV), T(deltas).to(DEV), count, anchors, cf)
    assert _rel(_np(boxes), funcs["proposal_boxes"]) <= 1e-5
    assert _rel(_np(props), funcs["proposal_props"]) <= 1e-5


@gpu
@pytest.mark.parametrize("size,pool", [("small", (7, 7, 3)), ("small", (14, 14, 5)), ("cfg3", (7, 7, 3))])
def test_pyramid_roi_align_vs_reference(funcs, size, pool):
    """models/mrcnn.py:373-457 on the four FPN levels: level assignment, per-level RoIAlign, original roi order; values and the gradient
    w.r.t. every feature map.  'cfg3' = BASELINE config 3 shapes (2x36x{32x32x128 ... 4x4x16}, 2 x 512 proposals)."""
    from medicaldetectiontoolkit_b200 import mrcnn
    fm, rois = GI.pyramid_inputs(size)
    fmt = [T(f).to(DEV).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True) for f in fm]
    y = mrcnn.pyramid_roi_align(fmt, T(rois).to(DEV), pool, [0, 1, 2, 3], 3)
    tag = "pyr_%s_%d" % (size, pool[0])
    assert list(y.shape) == funcs[tag + "_shape"].tolist()
    assert _rel(sub(_np(y), 65536), funcs[tag]) <= 1e-5
    g = T(np.random.RandomState(7).randn(*y.shape).astype(np.float32)).to(DEV)
    y.backward(g)
    for i, f in enumerate(fmt):
        assert _rel(sub(_np(f.grad), 32768), funcs[tag + "_g%d" % i]) <= 1e-4, (tag, i)


@gpu
def test_detection_target_layer_vs_reference(funcs):
    """models/mrcnn.py:461-613: IoU thresholds 0.3 / 0.01, positive sub-sampling, GT assignment, box_refinement / bbox_std_dev targets,
    RoIAlign(28,28,10) of the GT masks + round, SHEM negatives"""
    from medicaldetectiontoolkit_b200 import mrcnn
    cf, bp, sc, gcls, gbox, gmask = GI.detection_target_inputs()
    six, tcls, tdel, tmask = mrcnn.detection_target_layer(T(bp).to(DEV), T(sc).to(DEV), gcls, gbox, gmask, cf)
    assert _np(six).tolist() == funcs["dtl_ix"].tolist()
    assert _np(tcls).tolist() == funcs["dtl_cls"].tolist()
    assert _rel(_np(tdel), funcs["dtl_deltas"]) <= 1e-5
    m = _np(tmask)
    assert m.shape == funcs["dtl_masks"].shape
    assert float((m != funcs["dtl_masks"]).mean()) <= 1e-4      # rounded RoIAlign of a binary cuboid: a sample within 1e-7 of 0.5 may flip


@gpu
def test_mrcnn_refine_detections_vs_reference(funcs):
    """models/mrcnn.py:620-714"""
    from medicaldetectiontoolkit_b200 import mrcnn
    cf, rois, probs, deltas, bix = GI.mrcnn_refine_inputs()
    det = mrcnn.refine_detections(T(rois).to(DEV), T(probs).to(DEV), T(deltas).to(DEV), T(bix).to(DEV), cf)
    _assert_detections(_np(det), funcs["mrcnn_refine"], 3)


# ================================================================================================== GPU: whole models
def _build(case):
    from medicaldetectiontoolkit_b200 import mrcnn, retina_net, retina_unet
    cf, model, B = GI.model_case(case)
    mod = {"retina_unet": retina_unet, "retina_net": retina_net, "mrcnn": mrcnn}[model]
    net = mod.net(cf, None)
    GI.tame_(detweights.fill_(net), model)
    return cf, model, B, net.to(DEV)


def _batch_from_golden(cf, g, B, with_masks, case):
    data = GI.synthetic_batch(cf, B, seed=GI.case_seed(case))['data']
    bt, lab = g["bb_target"], g["roi_labels"]
    boxes = [bt[bt[:, -1] == b][:, :-1].astype(np.int64) for b in range(B)]
    labels = [lab[bt[:, -1] == b] for b in range(B)]
    seg = np.zeros((B, 1) + tuple(cf.patch_size), dtype=np.uint8)
    masks = []
    for b in range(B):
        ms = []
        for bx in boxes[b]:
            sl = (slice(int(bx[0]), int(bx[2])), slice(int(bx[1]), int(bx[3]))) + ((slice(int(bx[4]), int(bx[5])),) if cf.dim == 3 else ())
            seg[(b, 0) + sl] = 1
            m = np.zeros((1,) + tuple(cf.patch_size), dtype=np.uint8)
            m[(0,) + sl] = 1
            ms.append(m)
        masks.append(np.array(ms))
    batch = {'data': data, 'seg': seg, 'bb_target': boxes, 'roi_labels': labels, 'pid': ['g%d' % i for i in range(B)]}
    if with_masks:
        batch['roi_masks'] = masks
    return batch


def _check_grads(net, g, model, tol):
    params = dict(net.named_parameters())
    assert [k for k, _ in net.named_parameters()] == list(g["keys"])                     # state-dict keys: the reference's, in order
    assert sorted(k for k, p in params.items() if p.grad is None) == sorted(g["nograd"])      # parameters autograd never reaches (Fpn.P1_*)
    errs = {}
    for k in GI.GRAD_KEYS[model]:
        a, b = sub(_np(params[k].grad), 8192).astype(np.float64), g["grad__" + k].astype(np.float64)
        errs[k] = _rel_l2(a, b)
        nrm = float(params[k].grad.norm())
        assert abs(nrm - float(g["gradnorm__" + k][0])) <= tol * float(g["gradnorm__" + k][0]), (k, nrm, float(g["gradnorm__" + k][0]))
        if np.linalg.norm(b) > 0:
            assert float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300)) >= 1 - tol * tol, k      # same direction
        else:
            assert np.linalg.norm(a) == 0, k
    assert max(errs.values()) <= tol, errs
    return errs


def _match_fraction(got, want, dim):
    """fraction of the reference's detections found in ours (same rounded box, element, class; score within 1e-4)"""
    if want.shape[0] == 0:
        return 1.0
    hit = 0
    for w in want:
        same = np.all(got[:, :2 * dim + 2] == w[:2 * dim + 2], axis=1) & (np.abs(got[:, -1] - w[-1]) <= 1e-4)
        hit += bool(same.any())
    return hit / want.shape[0]


@gpu
@pytest.mark.parametrize("case", ["retina_unet_small", "retina_unet_cfg2", "retina_net_cfg1"])
def test_retina_models_vs_reference(case):
    """forward (class logits, box deltas, seg logits, detections), every loss term of train_forward and parameter gradients against the
    unmodified models/retina_unet.py / retina_net.py (models/retina_unet.py:381-456,477-513); cfg2 = BASELINE config 2 (2 x 1 x 128^3)"""
    g = np.load(os.path.join(GOLD, "model_%s.npz" % case))
    cf, model, B, net = _build(case)
    batch = _batch_from_golden(cf, g, B, False, case)
    img = T(batch['data']).to(DEV)
    with torch.no_grad():
        det, cl, bb, seg = net.forward(img)
    assert list(cl.shape) == g["class_logits_shape"].tolist()
    assert _rel(sub(_np(cl)), g["class_logits"]) <= 1e-4
    assert _rel(sub(_np(bb)), g["bb_outputs"]) <= 1e-4
    if model == 'retina_unet':
        assert _rel(sub(_np(seg)), g["seg_logits"]) <= 1e-4
    assert _match_fraction(_np(det), g["detections"], cf.dim) >= 0.9
    assert det.shape[0] == g["detections"].shape[0]

    from medicaldetectiontoolkit_b200 import retina_unet as RU
    log = {"class": [], "bbox": []}
    orig_c, orig_b = RU.compute_class_loss, RU.compute_bbox_loss

    def rec_c(*a, **k):
        out = orig_c(*a, **k)
        log["class"].append(out)
        return out

    def rec_b(*a, **k):
        out = orig_b(*a, **k)
        log["bbox"].append(out)
        return out

    RU.compute_class_loss, RU.compute_bbox_loss = rec_c, rec_b
    try:
        np.random.seed(0)          # the generator seeds numpy the same way: replays the reference's np.random.choice sub-sampling of positives
        res = net.train_forward(batch)
    finally:
        RU.compute_class_loss, RU.compute_bbox_loss = orig_c, orig_b
    res['torch_loss'].backward()
    got_c = np.array([float(l[0]) for l in log["class"]])
    got_b = np.array([float(l) for l in log["bbox"]])
    assert np.abs(got_c - g["class_loss"]).max() <= 1e-4 * np.abs(g["class_loss"]).max()
    assert np.abs(got_b - g["bbox_loss"]).max() <= 1e-4 * max(1.0, np.abs(g["bbox_loss"]).max())
    negs = np.concatenate([_np(l[1])[_np(l[1]) >= 0] for l in log["class"]])
    assert negs.tolist() == g["neg_ix"].tolist()                                      # the SHEM picks (indices into the negative subset)
    assert abs(float(res['torch_loss']) - float(g["loss"][0])) <= 1e-4 * abs(float(g["loss"][0]))
    assert [len(b) for b in res['boxes']] == g["n_boxes"].tolist()
    assert sorted(set(bx['box_type'] for b in res['boxes'] for bx in b)) == list(g["box_types"])
    # argmax of the seg logits: voxels whose two logits agree to ~1e-6 may fall on either side
    assert abs(float(np.asarray(res['seg_preds']).sum()) - float(g["seg_preds_sum"][0])) <= 1e-4 * np.asarray(res['seg_preds']).size
    _check_grads(net, g, model, 1e-2)   # encoder weights sit behind ~50 ReLU layers: mask flips of near-zero pre-activations (see module docstring)


@gpu
@pytest.mark.parametrize("case", ["mrcnn_small", "mrcnn_cfg3"])
def test_mrcnn_vs_reference(case):
    """RPN outputs, proposals, detections, detection masks, detection targets, all five loss terms and parameter gradients against the
    unmodified models/mrcnn.py (:853-966, :987-1083); cfg3 = BASELINE config 3 (2 x 128^3, 512 proposals, RoIAlign 7x7x3)"""
    from medicaldetectiontoolkit_b200 import mrcnn as MR
    g = np.load(os.path.join(GOLD, "model_%s.npz" % case))
    cf, model, B, net = _build(case)
    batch = _batch_from_golden(cf, g, B, True, case)
    img = T(batch['data']).to(DEV)
    with torch.no_grad():
        rl, rd, props, det, dm = net.forward(img)
    assert list(rl.shape) == g["rpn_logits_shape"].tolist()
    assert _rel(sub(_np(rl)), g["rpn_logits"]) <= 1e-4
    assert _rel(sub(_np(rd)), g["rpn_deltas"]) <= 1e-4
    # proposals: same boxes in the same order (scores well separated); tolerate a few near-tie swaps at the tail
    p_got, p_want = _np(props), g["proposals"]
    assert p_got.shape == p_want.shape
    close = np.all(np.abs(p_got - p_want) <= 1e-3 * np.maximum(1.0, np.abs(p_want)), axis=2)
    assert close.mean() >= 0.95, close.mean()
    assert _match_fraction(_np(det), g["detections"], cf.dim) >= 0.9
    if _match_fraction(_np(det), g["detections"], cf.dim) == 1.0:
        assert list(dm.shape) == g["detection_masks_shape"].tolist()    # mask-head values are pinned through mrcnn_mask_loss and the mask.* gradients

    log = {}
    names = ("compute_rpn_class_loss", "compute_rpn_bbox_loss", "compute_mrcnn_class_loss", "compute_mrcnn_bbox_loss", "compute_mrcnn_mask_loss",
             "detection_target_layer")
    orig = {n: getattr(MR, n) for n in names}

    def rec(n):
        def f(*a, **k):
            out = orig[n](*a, **k)
            log.setdefault(n, []).append(out)
            return out
        return f

    # second-stage class scores of all proposals (what SHEM ranks the negatives by)
    if close.all():
        assert _rel(_np(net.batch_mrcnn_class_scores), g["class_scores"]) <= 1e-4
    else:
        rows = np.repeat(close.reshape(-1), 1)
        assert _rel(_np(net.batch_mrcnn_class_scores)[rows], g["class_scores"][rows]) <= 1e-4
    for n in names:
        setattr(MR, n, rec(n))
    try:
        np.random.seed(0)
        res = net.train_forward(batch)
    finally:
        for n in names:
            setattr(MR, n, orig[n])
    res['torch_loss'].backward()
    six, tcls, tdel, tmask = log["detection_target_layer"][0]
    assert _np(tcls).tolist() == g["dtl_cls"].tolist()
    got_ix, want_ix = _np(six), g["dtl_ix"]
    n_pos = int((g["dtl_cls"] > 0).sum())
    assert got_ix[:n_pos].tolist() == want_ix[:n_pos].tolist()                          # positive samples: exact
    if got_ix.tolist() != want_ix.tolist():
        # SHEM negatives are the top-scoring ones: a pick may differ from the reference's only if the reference scores of the two are a near-tie
        ref_fg = g["class_scores"][:, 1:].max(1)
        assert sorted(got_ix[n_pos:] // (ref_fg.shape[0] // B)) == sorted(want_ix[n_pos:] // (ref_fg.shape[0] // B))
        assert np.all(ref_fg[got_ix[n_pos:]] >= ref_fg[want_ix[n_pos:]].min() - 1e-4), (got_ix, want_ix)
        pytest.skip("SHEM near-tie resolved differently than the reference (scores within 1e-4): sample-dependent losses not comparable")
    assert _rel(_np(tdel), g["dtl_deltas"]) <= 1e-3
    assert np.abs(_np(tmask).reshape(tmask.shape[0], -1).sum(1) - g["dtl_masks_sum"]).max() <= 2
    got = {
        "rpn_class_loss": np.array([float(l[0]) for l in log["compute_rpn_class_loss"]]),
        "rpn_bbox_loss": np.array([float(l) for l in log["compute_rpn_bbox_loss"]]),
        "mrcnn_class_loss": np.array([float(log["compute_mrcnn_class_loss"][0])]),
        "mrcnn_bbox_loss": np.array([float(log["compute_mrcnn_bbox_loss"][0])]),
        "mrcnn_mask_loss": np.array([float(log["compute_mrcnn_mask_loss"][0])]),
    }
    for k, v in got.items():
        assert np.abs(v - g[k]).max() <= 2e-4 * max(1.0, np.abs(g[k]).max()), (k, v, g[k])
    negs = np.concatenate([_np(l[1])[_np(l[1]) >= 0] for l in log["compute_rpn_class_loss"]])
    assert negs.tolist() == g["rpn_neg_ix"].tolist()
    assert abs(float(res['torch_loss']) - float(g["loss"][0])) <= 2e-4 * abs(float(g["loss"][0]))
    assert sorted(set(bx['box_type'] for b in res['boxes'] for bx in b)) == list(g["box_types"])
    _check_grads(net, g, model, 1e-2)


# ================================================================================================== CPU: the bench's CPU port is pinned too
def test_cpu_port_matches_reference_goldens():
    """oracle/cpu_step.py (the `--impl reference` / cpu_baseline arm: stock torch.nn.Conv3d on the host) reproduces the logits of the
    reference's own models/retina_unet.py under the same weights — the timed CPU arm computes the reference's function"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_step
    case = "retina_unet_small"
    g = np.load(os.path.join(GOLD, "model_%s.npz" % case))
    cf, model, B = GI.model_case(case)
    net = cpu_step.build_cpu_net(cf)
    assert [k for k, _ in net.named_parameters()] == list(g["keys"])
    GI.tame_(detweights.fill_(net), model)
    img = T(GI.synthetic_batch(cf, B, seed=GI.case_seed(case))['data'])
    with torch.no_grad():
        cl, bb, seg = net(img)
    assert list(cl.shape) == g["class_logits_shape"].tolist()
    assert _rel(sub(_np(cl)), g["class_logits"]) <= 1e-4
    assert _rel(sub(_np(bb)), g["bb_outputs"]) <= 1e-4
    assert _rel(sub(_np(seg)), g["seg_logits"]) <= 1e-4
