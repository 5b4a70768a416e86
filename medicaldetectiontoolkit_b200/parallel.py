"""Data-parallel gradient exchange: one process per GPU, ONE all-reduce per step over a flat fp32 gradient buffer.

The reference has no multi-GPU code (exec.py:38 is single-device).  The path shards by patch with no data-path collective; the only
exchange is the gradient sum.  All parameter .grad tensors are views into one contiguous buffer, so backward accumulates straight into
the message and the averaged result is what the (fused) Adam reads — no flatten/unflatten copies.  Parameters that never receive a
gradient (Fpn.P1_conv2.*, backbone.py:175) still sit in the buffer as zeros, which keeps the bucket list static across ranks.
"""
import torch
import torch.distributed as dist


class FlatGradAllReduce(object):
    def __init__(self, module, world_size, group=None):
        self.world = world_size
        self.group = group
        params = [p for p in module.parameters() if p.requires_grad]
        n = sum(p.numel() for p in params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.params = params

    def zero_grad(self):
        self.flat.zero_()

    def _check_aliasing(self):
        """`optimizer.zero_grad()` (set_to_none=True by default) or any `p.grad = None` silently detaches a parameter from the flat buffer; the
        all-reduce would then average a stale buffer and the replicas diverge.  Re-bind (and carry over) such gradients."""
        base = self.flat.data_ptr()
        off = 0
        for p in self.params:
            want = base + off * 4
            if p.grad is None:                                  # released by zero_grad(set_to_none=True) and not produced since: contributes 0
                p.grad = self.flat[off:off + p.numel()].view_as(p)
                p.grad.zero_()
            elif p.grad.data_ptr() != want:
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.grad)
                p.grad = view
            off += p.numel()

    def all_reduce(self):
        self._check_aliasing()
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.mul_(1.0 / self.world)
