"""N > 1 host logic on CPU: world_size-2 gloo run of the flat-gradient all-reduce (the only exchange step of the data-parallel path)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from medicaldetectiontoolkit_b200.parallel import FlatGradAllReduce
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.ReLU(), torch.nn.Linear(3, 2))
    unused = torch.nn.Parameter(torch.ones(5))   # a parameter that never receives a gradient (like Fpn.P1_conv2.*)
    net.register_parameter("unused", unused)
    red = FlatGradAllReduce(net, world)
    x = torch.full((2, 4), float(rank + 1))
    red.zero_grad()
    net(x).sum().backward()
    local = red.flat.clone()
    red.all_reduce()
    off_unused = [o for p, o in _offsets(red) if p is unused][0]
    q.put((rank, local, red.flat.clone(), [p.grad.data_ptr() == red.flat[o:o + p.numel()].data_ptr() for p, o in _offsets(red)], off_unused))
    dist.destroy_process_group()


def _offsets(red):
    off = 0
    for p in red.params:
        yield p, off
        off += p.numel()


def _free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:   # a fixed port can still be in TIME_WAIT from an earlier run
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_flat_grad_allreduce_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mean = (out[0][1] + out[1][1]) / 2
    for rank, local, reduced, views, off_unused in out:
        assert torch.allclose(reduced, mean)          # averaged gradients on every rank
        assert all(views)                             # .grad tensors are views of the flat buffer (no flatten copies)
        assert torch.all(reduced[off_unused:off_unused + 5] == 0)   # never-used parameter stays zero: static bucket layout
    assert not torch.allclose(out[0][1], out[1][1])   # ranks really had different local gradients


def test_flat_grad_rebinds_after_set_to_none():
    """optimizer.zero_grad() (set_to_none=True) must not silently detach parameters from the flat all-reduce buffer"""
    import torch
    from medicaldetectiontoolkit_b200.parallel import FlatGradAllReduce
    m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    red = FlatGradAllReduce(m, world_size=1)
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    m(torch.ones(2, 4)).sum().backward()
    first = red.flat.clone()
    assert first.abs().sum() > 0
    opt.zero_grad()                                  # set_to_none=True: every p.grad is gone
    m[0](torch.ones(2, 4)).sum().backward()          # only the first layer gets a fresh gradient (new tensors, not views of the buffer)
    red.all_reduce()
    off = 0
    for p in m.parameters():
        assert p.grad.data_ptr() == red.flat.data_ptr() + 4 * off
        off += p.numel()
    n0 = sum(p.numel() for p in m[0].parameters())
    assert torch.equal(red.flat[:n0], torch.cat([p.grad.reshape(-1) for p in m[0].parameters()]))
    assert red.flat[n0:].abs().sum() == 0            # stale gradients of the untouched layer are not averaged
