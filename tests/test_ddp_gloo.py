"""Host-side logic of the data-parallel layer (b200vit/ddp.py) on CPU: world_size 2 over gloo.
Checks bucket layout, hook + gradient-sink delivery, the all-reduce average and param.grad aliasing."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vit-is-all-you-need_b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _SinkLinear(torch.autograd.Function):
    """A stand-in for the fused CUDA backward: writes the weight gradient straight into the bucket slot."""

    @staticmethod
    def forward(ctx, x, w):
        from b200vit import functional as Fn
        ctx.save_for_backward(x, w)
        ctx.w = w
        ctx.sink = Fn._GRAD_SINK   # captured per autograd node, exactly like functional.TransformerStackFn
        return x @ w.t()

    @staticmethod
    def backward(ctx, dy):
        from b200vit import functional as Fn
        x, w = ctx.saved_tensors
        gw = dy.t() @ x
        if getattr(_SinkLinear, "fail_next", False):
            _SinkLinear.fail_next = False
            raise RuntimeError("injected backward failure")
        slot = Fn._slot(ctx.sink, ctx.w)      # asked once; None -> the gradient goes through autograd (functional.layer_backward)
        if slot is not None:
            slot.copy_(gw)
            gw = None   # delivered through the sink: autograd gets no second copy
            Fn._ready(ctx.sink, ctx.w)
        return dy @ w, gw


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(8, 16)
        self.w = torch.nn.Parameter(torch.randn(16, 16) * 0.1)
        self.b = torch.nn.Linear(16, 4)
        self.unused = torch.nn.Parameter(torch.zeros(3))

    def forward(self, x):
        return self.b(_SinkLinear.apply(torch.relu(self.a(x)), self.w))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, PKG)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200vit import ddp
    torch.manual_seed(123 + rank)  # different initial weights per rank: broadcast must fix that
    net = _Net()
    model = ddp.DataParallel(net, bucket_mb=0.0005)  # tiny buckets -> several buckets
    assert len(model.buckets) >= 3
    w0 = [p.detach().clone() for p in net.parameters()]
    gathered = [torch.zeros_like(w0[0]) for _ in range(world)]
    dist.all_gather(gathered, w0[0])
    assert torch.equal(gathered[0], gathered[1]), "parameters must be broadcast from rank 0"

    # reference: every rank's local gradient, computed without the wrapper, then averaged
    def local_grads(r):
        g = torch.Generator().manual_seed(1000 + r)
        x = torch.randn(5, 8, generator=g)
        y = torch.randn(5, 4, generator=g)
        ref = _Net()
        ref.load_state_dict(net.state_dict())
        ((ref(x) - y) ** 2).mean().backward()
        return [None if p.grad is None else p.grad.clone() for p in ref.parameters()], x, y

    per_rank = [local_grads(r) for r in range(world)]
    expect = []
    for i in range(len(w0)):
        gs = [per_rank[r][0][i] for r in range(world)]
        expect.append(None if gs[0] is None else sum(gs) / world)

    for step in range(2):  # twice: state must reset between backward passes
        for p in net.parameters():
            p.grad = None
        _, x, y = per_rank[rank]
        ((model(x) - y) ** 2).mean().backward()
        for p, e in zip(net.parameters(), expect):
            if e is None:
                continue
            assert p.grad is not None
            torch.testing.assert_close(p.grad, e, rtol=1e-5, atol=1e-6)
            assert model.grad_slot(p) is None, "a populated .grad must not be offered for overwriting"
            assert p.grad.data_ptr() == model._slots[p][1].data_ptr(), "param.grad must alias its bucket slot"
        # an un-wrapped model trained in the same process (teacher, second network ...) is not routed into the sink
        other, _, _ = local_grads(rank)
        assert other[0] is not None

    # no_sync(): local gradients only
    for p in net.parameters():
        p.grad = None
    with model.no_sync():
        _, x, y = per_rank[rank]
        ((model(x) - y) ** 2).mean().backward()
    idx_w = [n for n, _ in net.named_parameters()].index("w")
    torch.testing.assert_close(net.w.grad, per_rank[rank][0][idx_w], rtol=1e-5, atol=1e-6)

    # gradient accumulation: a no_sync() micro-step (above) followed by a synchronised step must reduce the SUM of both
    # micro-step gradients for sink-delivered (w) and hook-delivered (a, b) parameters alike
    ((model(x) - y) ** 2).mean().backward()
    for (name, p), e in zip(net.named_parameters(), expect):
        if e is not None:
            torch.testing.assert_close(p.grad, 2 * e, rtol=1e-5, atol=1e-6, msg=lambda m, n=name: f"accumulated {n}: {m}")

    # a backward that raises must not leave the wrapper stuck: the next step reduces normally
    for p in net.parameters():
        p.grad = None
    _SinkLinear.fail_next = True
    try:
        ((model(x) - y) ** 2).mean().backward()
        raise AssertionError("the injected failure did not surface")
    except RuntimeError as err:
        assert "injected" in str(err)
    for p in net.parameters():
        p.grad = None
    ((model(x) - y) ** 2).mean().backward()
    for (name, p), e in zip(net.named_parameters(), expect):
        if e is not None:
            torch.testing.assert_close(p.grad, e, rtol=1e-5, atol=1e-6, msg=lambda m, n=name: f"after a failed backward {n}: {m}")
    dist.destroy_process_group()
    open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")


def test_ddp_two_ranks_gloo(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
