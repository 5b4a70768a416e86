"""CPU port of the Retina U-Net training step — the `cpu_baseline` / `--impl reference` leg.  TEST INFRASTRUCTURE ONLY.

The reference (torch 0.4.1, TH/cffi extensions) cannot be installed or imported on the benchmark box, so the timed CPU
arm is a port (`cpu_baseline.kind = "port"`): the same network graph as models/retina_unet.py + models/backbone.py built from stock
torch.nn.Conv3d on the host cores (what the reference's NDConvGenerator builds, utils/model_utils.py:739-781), numpy fp64 anchor
matching (oracle/matching_oracle.py, pinned to the reference's outputs) and the C restatement of its NMS (oracle/mdt_oracle.c).
The network topology classes (FPN, towers) are generic in the `conv` factory exactly like the reference's, so they are shared with the
product package; every op that the product runs in libmdt_b200 is replaced here by its CPU oracle.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (_ROOT, os.path.join(_ROOT, "tests"), os.path.join(_ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import matching_oracle as MO  # noqa: E402
import _oracle as O  # noqa: E402


class TorchConvGenerator(object):
    """conv(+relu) factory on stock torch modules — restates utils/model_utils.py:732-781 for norm=None"""

    def __init__(self, dim):
        self.dim = dim

    def __call__(self, c_in, c_out, ks, pad=0, stride=1, norm=None, relu='relu'):
        conv = (nn.Conv2d if self.dim == 2 else nn.Conv3d)(c_in, c_out, kernel_size=ks, padding=pad, stride=stride)
        if norm is not None:
            raise ValueError("cpu port covers norm=None (all BASELINE configs)")
        if relu is not None:
            conv = nn.Sequential(conv, nn.ReLU(inplace=True) if relu == 'relu' else nn.LeakyReLU(inplace=True))
        return conv


def build_cpu_net(cf):
    from medicaldetectiontoolkit_b200 import model_utils as mutils
    from medicaldetectiontoolkit_b200.backbone import FPN
    from medicaldetectiontoolkit_b200.retina_unet import BBRegressor, Classifier

    class CpuNet(nn.Module):
        def __init__(self):
            super().__init__()
            conv = TorchConvGenerator(cf.dim)
            self.np_anchors = mutils.generate_pyramid_anchors(None, cf)
            self.anchors = torch.from_numpy(self.np_anchors).float()
            self.Fpn = FPN(cf, conv, operate_stride1=cf.operate_stride1)
            self.Classifier = Classifier(cf, conv)
            self.BBRegressor = BBRegressor(cf, conv)
            self.final_conv = conv(cf.end_filts, cf.num_seg_classes, ks=1, pad=0, norm=None, relu=None)

        def forward(self, img):
            outs = self.Fpn(img)
            seg_logits = self.final_conv(outs[0])
            fmaps = [outs[i + 1] for i in cf.pyramid_levels]
            cl = torch.cat([self.Classifier(p)[0] for p in fmaps], 1)
            bb = torch.cat([self.BBRegressor(p)[0] for p in fmaps], 1)
            return cl, bb, seg_logits

    return CpuNet()


def refine_detections_cpu(cf, anchors, probs, deltas, batch_ixs):
    """retina_unet.py:194-271 on the host: global sort, top pre_nms_limit, decode, per (element, class) NMS (C oracle), top-k"""
    from medicaldetectiontoolkit_b200 import model_utils as mutils
    dim = cf.dim
    fg = probs[:, 1:].contiguous()
    flat, order = fg.view(-1).sort(descending=True)
    keep_ix = order[:cf.pre_nms_limit]
    row = torch.div(keep_ix, fg.shape[1], rounding_mode='floor')
    cls = keep_ix % fg.shape[1] + 1
    scores = flat[:cf.pre_nms_limit]
    b_ix = batch_ixs[row]
    std = torch.from_numpy(np.reshape(cf.rpn_bbox_std_dev, [1, dim * 2])).float()
    scale = torch.from_numpy(np.asarray(cf.scale)).float()
    apply = mutils.apply_box_deltas_2D if dim == 2 else mutils.apply_box_deltas_3D
    rois = torch.round(mutils.clip_to_window(cf.window, apply(anchors[row % anchors.shape[0]] / scale, deltas[row] * std) * scale))
    out = []
    for b in torch.unique(b_ix).tolist():
        sel_b = torch.nonzero(b_ix == b)[:, 0]
        kept_b = []
        for c in torch.unique(cls[sel_b]).tolist():
            ixs = sel_b[cls[sel_b] == c]
            dets = torch.cat((rois[ixs], scores[ixs].unsqueeze(1)), 1).numpy()   # already in descending score order
            kept_b.append(ixs[torch.from_numpy(O.nms(dets, cf.detection_nms_threshold, dim))])
        kept_b = torch.cat(kept_b)
        top = scores[kept_b].sort(descending=True)[1][:cf.model_max_instances_per_batch_element]
        out.append(kept_b[top])
    keep = torch.cat(out).sort()[0] if out else torch.zeros(0, dtype=torch.long)
    return torch.cat((rois[keep], b_ix[keep].unsqueeze(1).float(), cls[keep].unsqueeze(1).float(), scores[keep].unsqueeze(1)), 1)


def class_loss_cpu(match, logits, shem_poolsize=20):
    pos = torch.nonzero(match > 0).squeeze(1)
    neg = torch.nonzero(match == -1).squeeze(1)
    pos_loss = F.cross_entropy(logits[pos], match[pos].long()) if pos.numel() else logits.new_zeros(())
    if neg.numel():
        n_neg = max(1, pos.numel())
        probs = F.softmax(logits[neg], 1)
        order = probs[:, 1:].max(1)[0].sort(descending=True)[1]
        pool = order[:min(shem_poolsize * n_neg, order.numel())]
        pick = pool[torch.randperm(pool.numel())[:n_neg]]
        neg_loss = F.cross_entropy(logits[neg][pick], torch.zeros(pick.numel(), dtype=torch.long))
    else:
        neg_loss = logits.new_zeros(())
    return (pos_loss + neg_loss) / 2


def train_step_cpu(net, opt, cf, batch):
    """one full step: forward, detections, matching, losses, backward, Adam — returns the loss value"""
    from medicaldetectiontoolkit_b200.retina_unet import batch_dice
    img = torch.from_numpy(batch['data']).float()
    cl, bb, seg_logits = net(img)
    B, A = cl.shape[0], cl.shape[1]
    with torch.no_grad():
        batch_ixs = torch.arange(B).unsqueeze(1).repeat(1, A).view(-1)
        refine_detections_cpu(cf, net.anchors, F.softmax(cl.view(-1, cl.shape[-1]), 1), bb.view(-1, bb.shape[-1]), batch_ixs)
    class_loss = bbox_loss = 0.
    for b in range(B):
        gt, ids = np.asarray(batch['bb_target'][b], dtype=np.float64), np.asarray(batch['roi_labels'][b])
        labels, row_arg = MO.match_labels(net.np_anchors, gt, ids, cf.anchor_matching_iou, cf.dim)
        pos_ids = np.where(labels > 0)[0]
        extra = len(pos_ids) - cf.rpn_train_anchors_per_image // 2
        if extra > 0:
            labels[np.random.choice(pos_ids, extra, replace=False)] = 0
            pos_ids = np.where(labels > 0)[0]
        tgt = MO.delta_targets(net.np_anchors, gt, row_arg, pos_ids, cf.rpn_train_anchors_per_image, cf.rpn_bbox_std_dev, cf.dim)
        match = torch.from_numpy(labels)
        class_loss = class_loss + class_loss_cpu(match, cl[b]) / B
        if len(pos_ids):
            bbox_loss = bbox_loss + F.smooth_l1_loss(bb[b][torch.from_numpy(pos_ids)], torch.from_numpy(tgt[:len(pos_ids)]).float()) / B
    seg = torch.from_numpy(batch['seg']).long()
    ohe = F.one_hot(seg[:, 0], cf.num_seg_classes).movedim(-1, 1).float()
    loss = class_loss + bbox_loss + ((1 - batch_dice(F.softmax(seg_logits, 1), ohe)) + F.cross_entropy(seg_logits, seg[:, 0])) / 2
    opt.zero_grad()
    loss.backward()
    opt.step()
    return float(loss.item())


def calibrate_threads(max_threads):
    """torch's CPU conv3d does not scale to every hardware thread of a big host (128 threads ran 6x SLOWER than 8 on the same step):
    time one small training step (64x64x32 patch) at a few thread counts and return the fastest — 'all the host threads it can use'."""
    from medicaldetectiontoolkit_b200.configs import make_cf, synthetic_batch
    cf = make_cf('retina_unet', 3, (64, 64, 32))
    batch = synthetic_batch(cf, 1, seed=0)
    best, best_t = None, None
    cands = sorted(set(t for t in (8, 16, 32, 64, max_threads) if t <= max_threads))
    for t in cands:
        times, _ = time_cpu_steps(cf, batch, steps=1, warmup=0, threads=t)
        if best_t is None or times[0] < best_t:
            best, best_t = t, times[0]
    return best


def time_cpu_steps(cf, batch, steps=1, warmup=0, threads=None):
    """returns (seconds per step list, cores used)"""
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = build_cpu_net(cf)
    opt = torch.optim.Adam(net.parameters(), lr=cf.learning_rate[0], weight_decay=cf.weight_decay)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        train_step_cpu(net, opt, cf, batch)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, torch.get_num_threads()
