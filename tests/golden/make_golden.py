"""Generates the golden fixtures under tests/golden/ by running the REFERENCE's own Python code (imported from /root/reference,
read-only, nothing copied).  Run once in the build container:  python tests/golden/make_golden.py

Fixtures (all small, seeded):
  anchors_*.npz      utils.model_utils.generate_pyramid_anchors for a 3D and a 2D toy config  (+ sha256 of the full cfg2 grid, A = 1 347 840)
  matching_*.npz     utils.model_utils.gt_anchor_matching inputs/outputs, 3D and 2D, sub-sampling branch inactive and active (seeded)
  boxcoding.npz      apply_box_deltas_{2D,3D}, clip_boxes_3D, box_refinement
  backbone3d_*.npz   models/backbone.py FPN forward + input gradient on a tiny 3D patch with stock nn.Conv3d (CPU fp32), weights included
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

# utils/exp_utils.py imports plotting -> matplotlib (absent); only model_utils / backbone are needed, they import cleanly
import utils.model_utils as mutils  # noqa: E402


class Logger:
    def info(self, *a, **k):
        pass


from golden_cfg import cf2d, cf3d, rand_gt  # noqa: E402


def main():
    log = Logger()
    # ------------------------------------------------------------------ anchors
    c3 = cf3d()
    a3 = mutils.generate_pyramid_anchors(log, c3)
    c2 = cf2d()
    a2 = mutils.generate_pyramid_anchors(log, c2)
    full = cf3d((128, 128, 128))
    afull = mutils.generate_pyramid_anchors(log, full)
    np.savez_compressed(os.path.join(HERE, "anchors.npz"), a3=a3, a2=a2, full_shape=np.array(afull.shape),
                        full_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(afull).tobytes()).digest(), dtype=np.uint8),
                        full_head=afull[:64], full_tail=afull[-64:])
    print("anchors", a3.shape, a2.shape, afull.shape)

    # ------------------------------------------------------------------ matching
    rs = np.random.RandomState(1234)
    cases = {}
    # 3D, few positives allowed -> sub-sampling branch active (np.random seeded right before the call)
    for name, cf, anc, G, ext, tpi, seed in [("m3_sub", c3, a3, 3, (64, 64, 32), 6, 11), ("m3_nosub", c3, a3, 8, (64, 64, 32), 100000, 12),
                                             ("m3_g1", c3, a3, 1, (64, 64, 32), 100000, 13), ("m2_nosub", c2, a2, 5, (128, 128), 100000, 14),
                                             ("m2_sub", c2, a2, 4, (128, 128), 8, 15)]:
        cf.rpn_train_anchors_per_image = tpi
        gt = rand_gt(rs, G, ext, cf.dim, 6, 30 if cf.dim == 3 else 60)
        cls = rs.randint(1, 3, size=G)
        np.random.seed(seed)
        m, t = mutils.gt_anchor_matching(cf, anc, gt, cls)
        cases[name + "_gt"] = gt
        cases[name + "_cls"] = cls
        cases[name + "_matches"] = m
        cases[name + "_targets"] = t if tpi < 1000 else t[:64]
        cases[name + "_cfg"] = np.array([cf.dim, tpi, seed, cf.anchor_matching_iou])
        print(name, "pos", (m > 0).sum(), "neg", (m == -1).sum())
    # class-agnostic (RPN) call: gt_class_ids None
    c3.rpn_train_anchors_per_image = 100000
    gt = rand_gt(rs, 4, (64, 64, 32), 3, 6, 30)
    m, t = mutils.gt_anchor_matching(c3, a3, gt)
    cases["m3_rpn_gt"], cases["m3_rpn_matches"] = gt, m
    np.savez_compressed(os.path.join(HERE, "matching.npz"), **cases)

    # ------------------------------------------------------------------ box coding
    torch.manual_seed(0)
    b3 = torch.rand(200, 6) * 50
    b3 = torch.stack([b3[:, 0], b3[:, 1], b3[:, 0] + 4 + b3[:, 2], b3[:, 1] + 4 + b3[:, 3], b3[:, 4], b3[:, 4] + 2 + b3[:, 5]], 1)
    d3 = torch.randn(200, 6) * 0.3
    g3 = b3 + torch.randn(200, 6).abs()
    b2, d2 = b3[:, :4].clone(), d3[:, :4].clone()
    np.savez_compressed(os.path.join(HERE, "boxcoding.npz"), b3=b3.numpy(), d3=d3.numpy(), g3=g3.numpy(),
                        apply3=mutils.apply_box_deltas_3D(b3.clone(), d3).numpy(), apply2=mutils.apply_box_deltas_2D(b2.clone(), d2).numpy(),
                        clip3=mutils.clip_boxes_3D(b3.clone(), [0, 0, 40, 40, 0, 20]).numpy(),
                        refine3=mutils.box_refinement(b3.clone(), g3).numpy(), refine2=mutils.box_refinement(b2.clone(), g3[:, :4]).numpy())

    # ------------------------------------------------------------------ backbone (tiny)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_backbone", os.path.join(REF, "models", "backbone.py"))
    bb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bb)
    sys.path.insert(0, HERE)
    import detweights
    for tag, op1 in [("unet", True), ("mrcnn", False)]:
        cfb = types.SimpleNamespace(start_filts=18, end_filts=36, res_architecture='resnet50', sixth_pooling=False, n_channels=1, norm=None,
                                    relu='relu', n_latent_dims=0)
        conv = mutils.NDConvGenerator(3)
        fpn = detweights.fill_(bb.FPN(cfb, conv, operate_stride1=op1).float())
        x = torch.from_numpy(np.random.RandomState(3).rand(1, 1, 32, 32, 16).astype(np.float32)).requires_grad_(True)
        outs = fpn(x)
        loss = sum((o * o).mean() for o in outs)
        loss.backward()
        keys = [k for k, _ in fpn.named_parameters()]
        shapes = {k: tuple(v.shape) for k, v in fpn.state_dict().items()}
        gsel = ("C1.0.weight", "C1.0.bias", "P2_conv2.weight", "P5_conv1.bias", "C3.0.conv2.0.weight", "C5.2.conv3.weight", "C2.1.conv1.0.bias")
        grads = {"grad__" + k: detweights.subsample(p.grad.numpy()) for k, p in fpn.named_parameters() if p.grad is not None and k in gsel}
        nograd = [k for k, p in fpn.named_parameters() if p.grad is None]
        np.savez_compressed(os.path.join(HERE, "backbone3d_%s.npz" % tag), x_grad=detweights.subsample(x.grad.numpy()),
                            loss=np.array(loss.item()), keys=np.array(keys), key_shapes=np.array([str(shapes[k]) for k in keys]),
                            nograd=np.array(nograd),
                            **{"out%d" % i: detweights.subsample(o.detach().numpy()) for i, o in enumerate(outs)},
                            **{"outshape%d" % i: np.array(o.shape) for i, o in enumerate(outs)}, **grads)
        print("backbone", tag, [tuple(o.shape) for o in outs], "params", sum(int(np.prod(s)) for s in shapes.values()), "nograd", nograd)


if __name__ == "__main__":
    main()
