"""Model-surface goldens: outputs of the REFERENCE's unmodified models/retina_unet.py, models/mrcnn.py, models/retina_net.py and
utils/model_utils.py (imported from /root/reference through ref_shims.py; nothing copied) on the seeded inputs of golden_inputs.py.
Run once in the build container:   python tests/golden/make_model_golden.py [funcs] [<model case> ...]

  model_funcs.npz            function level: refine_detections (both models), compute_*_loss, proposal_layer, pyramid_roi_align (small and
                             BASELINE-cfg3 size, forward + gradient), detection_target_layer, mrcnn losses, bbox_overlaps, unique1d,
                             batch_dice, one-hot, shem, log2
  model_<case>.npz           whole model: forward outputs (sub-sampled logits, detections, proposals), every loss term of train_forward,
                             parameter gradients (sub-sampled) under name-keyed deterministic weights (detweights.py + golden_inputs.TAME)

Sampling is neutralised on BOTH sides the same way: torch.randperm -> identity (the tests put our sampler in the same mode through
model_utils.SAMPLING), and the synthetic batches have <= rpn_train_anchors_per_image // 2 positives so np.random.choice is never drawn.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import detweights  # noqa: E402
import golden_inputs as GI  # noqa: E402
import ref_shims as RS  # noqa: E402

T = torch.from_numpy
sub = detweights.subsample


def identity_randperm(n, *a, **k):
    return torch.arange(n)


class Patched:
    """temporarily replace attributes (restored on exit)"""

    def __init__(self, *triples):
        self.triples = triples

    def __enter__(self):
        self.saved = [(o, k, getattr(o, k)) for o, k, _ in self.triples]
        for o, k, v in self.triples:
            setattr(o, k, v)

    def __exit__(self, *a):
        for o, k, v in self.saved:
            setattr(o, k, v)


def recorder(fn, log, name):
    def wrapped(*a, **k):
        out = fn(*a, **k)
        log.setdefault(name, []).append(out)
        return out
    return wrapped


def np_(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


# ---------------------------------------------------------------------------------------------------------------- function level
def make_funcs():
    out = {}
    RS.install_import_shims()
    with RS.torch04_semantics(), Patched((torch, "randperm", identity_randperm)):
        import utils.model_utils as mutils
        ru = RS.load_ref_module("retina_unet")
        mr = RS.load_ref_module("mrcnn")
        log = RS.Logger()

        # retina refine_detections
        cf, probs, deltas, bix = GI.retina_refine_inputs()
        anchors = T(mutils.generate_pyramid_anchors(log, cf)).float()
        det = ru.refine_detections(anchors, T(probs), T(deltas), T(bix), RS.ref_cf(cf))
        out["retina_refine"] = np_(det)
        print("retina_refine", det.shape)

        # retina / rpn class + bbox losses
        for name, (m, lg, pool) in GI.class_loss_cases().items():
            if lg.shape[1] == 2:
                loss, neg = mr.compute_rpn_class_loss(T(m), T(lg), pool)
            else:
                loss, neg = ru.compute_class_loss(T(m), T(lg), pool)
            out["class_loss__" + name] = np_(loss).reshape(-1)
            out["class_neg__" + name] = np.asarray(neg).astype(np.int64)
            print("class_loss", name, float(np_(loss).reshape(-1)[0]), np.asarray(neg)[:6])
        for name, (t, p, m) in GI.bbox_loss_cases().items():
            out["bbox_loss__" + name] = np_(ru.compute_bbox_loss(T(t), T(p), T(m))).reshape(-1)
            out["rpn_bbox_loss__" + name] = np_(mr.compute_rpn_bbox_loss(T(t), T(p), T(m))).reshape(-1)

        # proposal_layer
        cf, probs, deltas, count = GI.proposal_inputs()
        anchors = T(mutils.generate_pyramid_anchors(log, cf)).float()
        nb, props = mr.proposal_layer(T(probs), T(deltas), count, anchors, RS.ref_cf(cf))
        out["proposal_boxes"], out["proposal_props"] = np_(nb), np.asarray(props)
        print("proposal_layer", nb.shape, "non-pad", int((np.asarray(props)[..., -1] > 0).sum()))

        # pyramid_roi_align: small (values + gradient), cfg3 size (values sub-sampled + gradient sub-sampled)
        for size in ("small", "cfg3"):
            fm, rois = GI.pyramid_inputs(size)
            for pool in ((7, 7, 3), (14, 14, 5)):
                if size == "cfg3" and pool[0] == 14:
                    continue
                fmt = [T(f).requires_grad_(True) for f in fm]
                y = mr.pyramid_roi_align(fmt, T(rois), pool, [0, 1, 2, 3], 3)
                g = T(np.random.RandomState(7).randn(*y.shape).astype(np.float32))
                y.backward(g)
                tag = "pyr_%s_%d" % (size, pool[0])
                out[tag] = sub(np_(y), 65536)
                out[tag + "_shape"] = np.array(y.shape)
                for i, f in enumerate(fmt):
                    out[tag + "_g%d" % i] = sub(np_(f.grad), 32768)
                print(tag, y.shape)

        # detection_target_layer
        cf, bp, sc, gcls, gbox, gmask = GI.detection_target_inputs()
        six, tcls, tdel, tmask = mr.detection_target_layer(T(bp), T(sc), gcls, gbox, gmask, RS.ref_cf(cf))
        out["dtl_ix"], out["dtl_cls"], out["dtl_deltas"], out["dtl_masks"] = np_(six), np_(tcls), np_(tdel), np_(tmask).astype(np.uint8)
        print("detection_target_layer", np_(six), np_(tcls))

        # mrcnn refine_detections
        cf, rois, probs, deltas, bix = GI.mrcnn_refine_inputs()
        det = mr.refine_detections(T(rois), T(probs), T(deltas), T(bix), RS.ref_cf(cf))
        out["mrcnn_refine"] = np_(det)
        print("mrcnn_refine", det.shape)

        # mrcnn head losses
        t_cls, logits, t_del, p_del, t_m, p_m = GI.mrcnn_loss_inputs()
        out["mrcnn_class_loss"] = np_(mr.compute_mrcnn_class_loss(T(t_cls), T(logits))).reshape(-1)
        out["mrcnn_bbox_loss"] = np_(mr.compute_mrcnn_bbox_loss(T(t_del), T(p_del), T(t_cls))).reshape(-1)
        out["mrcnn_mask_loss"] = np_(mr.compute_mrcnn_mask_loss(T(t_m), T(p_m), T(t_cls))).reshape(-1)
        z = np.zeros_like(t_cls)
        out["mrcnn_bbox_loss_nopos"] = np_(mr.compute_mrcnn_bbox_loss(T(t_del), T(p_del), T(z))).reshape(-1)
        out["mrcnn_mask_loss_nopos"] = np_(mr.compute_mrcnn_mask_loss(T(t_m), T(p_m), T(z))).reshape(-1)

        # utils
        u = GI.utils_inputs()
        out["overlaps3"] = np_(mutils.bbox_overlaps_3D(T(u["b3a"]), T(u["b3b"])))
        out["overlaps2"] = np_(mutils.bbox_overlaps_2D(T(u["b2a"]), T(u["b2b"])))
        out["unique1d"] = np_(mutils.unique1d(T(u["uniq"])))
        ohe = mutils.get_one_hot_encoding(u["dice_seg"], 3)
        out["one_hot_sum"] = ohe.sum(axis=(0, 2, 3, 4))
        out["batch_dice"] = np_(mutils.batch_dice(T(u["dice_pred"]), T(ohe).float())).reshape(-1)
        out["batch_dice_fpw"] = np_(mutils.batch_dice(T(u["dice_pred"]), T(ohe).float(), false_positive_weight=2.0)).reshape(-1)
        out["shem"] = np_(mutils.shem(T(u["shem_probs"]), 7, 10))
        out["shem_small_pool"] = np_(mutils.shem(T(u["shem_probs"][:30]), 7, 10))
        out["log2"] = np_(mutils.log2(T(u["log2_x"])))
    np.savez_compressed(os.path.join(HERE, "model_funcs.npz"), **out)
    print("wrote model_funcs.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


# ---------------------------------------------------------------------------------------------------------------- whole models
def cuboid_batch(cf, data, boxes_per_elem, labels_per_elem, with_masks):
    """batch dict from explicit GT boxes: seg / roi_masks are the boxes' cuboids (format of configs.synthetic_batch)"""
    B = data.shape[0]
    seg = np.zeros((B, 1) + tuple(cf.patch_size), dtype=np.uint8)
    masks = []
    for b in range(B):
        ms = []
        for bx in boxes_per_elem[b]:
            sl = (slice(int(bx[0]), int(bx[2])), slice(int(bx[1]), int(bx[3]))) + ((slice(int(bx[4]), int(bx[5])),) if cf.dim == 3 else ())
            seg[(b, 0) + sl] = 1
            m = np.zeros((1,) + tuple(cf.patch_size), dtype=np.uint8)
            m[(0,) + sl] = 1
            ms.append(m)
        masks.append(np.array(ms))
    batch = {'data': data, 'seg': seg, 'bb_target': [np.asarray(b) for b in boxes_per_elem], 'roi_labels': [np.asarray(l) for l in labels_per_elem],
             'pid': ['g%d' % i for i in range(B)]}
    if with_masks:
        batch['roi_masks'] = masks
    return batch


def make_model(case):
    cf, model, B = GI.model_case(case)
    RS.install_import_shims()
    out = {}
    t0 = time.time()
    with RS.torch04_semantics(), Patched((torch, "randperm", identity_randperm)):
        m = RS.load_ref_module(model)
        rcf = RS.ref_cf(cf)
        rcf.operate_stride1 = cf.operate_stride1
        net = m.net(rcf, RS.Logger())
        GI.tame_(detweights.fill_(net), model)
        batch = GI.synthetic_batch(cf, B, seed=GI.case_seed(case), with_masks=(model == 'mrcnn'))
        img = T(batch['data']).float()

        if model == 'mrcnn':
            # pass 1: GT boxes := two of the reference's own proposals per element (so that detection_target_layer has positives)
            with torch.no_grad():
                _, _, props, _, _ = net.forward(img)
            boxes, labels = [], []
            for b in range(B):
                p = np.round(np.asarray(props[b])[:, :6])
                ext = np.stack([p[:, 2] - p[:, 0], p[:, 3] - p[:, 1], p[:, 5] - p[:, 4]], 1)
                ok = np.where((ext.min(1) >= 4) & (np.asarray(props[b])[:, 6] > 0))[0]
                pick = ok[[0, len(ok) // 2]] if len(ok) > 1 else ok[:1]
                boxes.append(p[pick].astype(np.int64))
                labels.append(np.array([1 + (i + b) % 2 for i in range(len(pick))]))
            batch = cuboid_batch(cf, batch['data'], boxes, labels, True)
        out["bb_target"] = np.array([np.asarray(b) for b in batch['bb_target']], dtype=object) if False else np.concatenate(
            [np.concatenate([np.asarray(b), np.full((len(b), 1), i)], 1) for i, b in enumerate(batch['bb_target'])], 0)
        out["roi_labels"] = np.concatenate([np.asarray(l) for l in batch['roi_labels']])

        log = {}
        patches = [(F, "cross_entropy", recorder(F.cross_entropy, log, "ce"))]
        if model == 'mrcnn':
            for fn in ("compute_rpn_class_loss", "compute_rpn_bbox_loss", "compute_mrcnn_class_loss", "compute_mrcnn_bbox_loss",
                       "compute_mrcnn_mask_loss", "detection_target_layer", "proposal_layer", "refine_detections"):
                patches.append((m, fn, recorder(getattr(m, fn), log, fn)))
        else:
            for fn in ("compute_class_loss", "compute_bbox_loss", "refine_detections"):
                patches.append((m, fn, recorder(getattr(m, fn), log, fn)))
            patches.append((m.mutils, "batch_dice", recorder(m.mutils.batch_dice, log, "batch_dice")))
        with Patched(*patches):
            np.random.seed(0)
            if model == 'retina_net':
                fwd = net.forward(img)
                res = net.train_forward(batch)
            elif model == 'retina_unet':
                res = net.train_forward(batch)
            else:
                res = net.train_forward(batch)
        res['torch_loss'].backward()

        out["loss"] = np_(res['torch_loss']).reshape(-1)
        if model == 'mrcnn':
            out["rpn_class_loss"] = np.array([float(np_(l[0]).reshape(-1)[0]) for l in log["compute_rpn_class_loss"]])
            out["rpn_neg_ix"] = np.concatenate([np.asarray(l[1]).astype(np.int64) for l in log["compute_rpn_class_loss"]])
            out["rpn_bbox_loss"] = np.array([float(np_(l).reshape(-1)[0]) for l in log["compute_rpn_bbox_loss"]])
            for k in ("compute_mrcnn_class_loss", "compute_mrcnn_bbox_loss", "compute_mrcnn_mask_loss"):
                out[k[8:]] = np_(log[k][0]).reshape(-1)
            six, tcls, tdel, tmask = log["detection_target_layer"][0]
            out["dtl_ix"], out["dtl_cls"], out["dtl_deltas"] = np_(six), np_(tcls), np_(tdel)
            out["dtl_masks_sum"] = np_(tmask).reshape(tmask.shape[0], -1).sum(1)
            nb, props = log["proposal_layer"][-1]
            out["proposals"] = np.asarray(props)
            out["detections"] = np_(log["refine_detections"][-1])
            # continuous outputs of a fresh forward (same weights; train_forward does not change them)
            with torch.no_grad():
                rl, rd, _, det, dm = net.forward(img)
            out["rpn_logits"], out["rpn_deltas"] = sub(np_(rl)), sub(np_(rd))
            out["class_scores"] = np_(net.batch_mrcnn_class_scores)
            # margin of the SHEM picks: lowest selected negative vs best unselected negative of the same element (printed; want >> 1e-4)
            fg = out["class_scores"][:, 1:].max(1)
            P_ = fg.shape[0] // B
            neg_ix = out["dtl_ix"][out["dtl_cls"] == 0]
            for b in range(B):
                sel = neg_ix[(neg_ix >= b * P_) & (neg_ix < (b + 1) * P_)]
                rest = np.setdiff1d(np.arange(b * P_, (b + 1) * P_), out["dtl_ix"])
                print("    shem margin elem", b, float(fg[sel].min() - np.sort(fg[rest])[-1]), "sel", sel)
            out["rpn_logits_shape"] = np.array(rl.shape)
            out["detection_masks"] = sub(np_(dm))
            out["detection_masks_shape"] = np.array(dm.shape)
        else:
            cl = log["compute_class_loss"]
            out["class_loss"] = np.array([float(np_(l[0]).reshape(-1)[0]) for l in cl])
            out["neg_ix"] = np.concatenate([np.asarray(l[1]).astype(np.int64) for l in cl])
            out["bbox_loss"] = np.array([float(np_(l).reshape(-1)[0]) for l in log["compute_bbox_loss"]])
            out["detections"] = np_(log["refine_detections"][-1])
            if model == 'retina_unet':
                out["dice"] = np_(log["batch_dice"][-1]).reshape(-1)
                out["seg_ce"] = np_(log["ce"][-1]).reshape(-1)
            with torch.no_grad():
                det, cl_, bb_, seg_ = net.forward(img)
            out["class_logits"], out["bb_outputs"] = sub(np_(cl_)), sub(np_(bb_))
            out["class_logits_shape"] = np.array(cl_.shape)
            out["logit_std"] = np.array([float(cl_.std()), float(bb_.std())])
            if model == 'retina_unet':
                out["seg_logits"] = sub(np_(seg_))
            out["seg_preds_sum"] = np.array([float(np.asarray(res['seg_preds']).sum())])
        out["n_boxes"] = np.array([len(b) for b in res['boxes']])
        out["box_types"] = np.array(sorted(set(bx['box_type'] for b in res['boxes'] for bx in b)))
        grads = dict(net.named_parameters())
        for k in GI.GRAD_KEYS[model]:
            g = grads[k].grad
            out["grad__" + k] = sub(np_(g), 8192)
            out["gradnorm__" + k] = np.array([float(g.norm())])
        out["nograd"] = np.array([k for k, p in net.named_parameters() if p.grad is None])
        out["keys"] = np.array([k for k, _ in net.named_parameters()])
        print(case, res['logger_string'], "| %.1fs" % (time.time() - t0))
        for k in ("class_loss", "bbox_loss", "rpn_class_loss", "rpn_bbox_loss", "mrcnn_class_loss", "mrcnn_bbox_loss", "mrcnn_mask_loss", "dtl_cls",
                  "logit_std"):
            if k in out:
                print("   ", k, out[k])
        print("    detections", out["detections"].shape, "nograd", list(out["nograd"]))
    np.savez_compressed(os.path.join(HERE, "model_%s.npz" % case), **out)
    print("wrote model_%s.npz" % case, sum(v.nbytes for v in out.values()) // 1024, "KiB raw")


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("THREADS", "8")))
    what = sys.argv[1:] or ["funcs"] + list(GI.MODEL_CASES)
    for w in what:
        if w == "funcs":
            make_funcs()
        else:
            make_model(w)
