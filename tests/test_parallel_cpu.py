"""The N>1 path on CPU: world_size-2 gloo processes exercise the bucketed flat-gradient all-reduce and
the clip sharding used by bench.py / the training environment."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_frame_inpainting_b200.parallel import FlatGradAllReducer, broadcast_module, shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)  # deliberately different replicas ...
    net = torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3), torch.nn.ReLU(), torch.nn.Conv2d(4, 1, 3),
                              torch.nn.Flatten(), torch.nn.Linear(16, 3))
    unused = torch.nn.Linear(3, 3)  # a parameter that never receives a gradient
    net.add_module("unused", unused)
    broadcast_module(net)           # ... made identical here
    red = FlatGradAllReducer(net, n_buckets=3)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    torch.manual_seed(100 + rank)
    x = torch.randn(5, 2, 8, 8)     # per-rank shard of the batch
    for _ in range(2):
        red.zero_grad()
        loss = net[:5](x).pow(2).mean()
        red.arm()
        loss.backward()
        red.finish()
        opt.step()
    flat = torch.cat([p.detach().flatten() for p in net.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        # reference: one process, both shards, mean of the two per-rank losses
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3), torch.nn.ReLU(), torch.nn.Conv2d(4, 1, 3),
                                  torch.nn.Flatten(), torch.nn.Linear(16, 3))
        ref.add_module("unused", torch.nn.Linear(3, 3))
        ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
        xs = []
        for r in range(world):
            torch.manual_seed(100 + r)
            xs.append(torch.randn(5, 2, 8, 8))
        for _ in range(2):
            ropt.zero_grad()
            sum(ref[:5](xi).pow(2).mean() for xi in xs).div(world).backward()
            ropt.step()
        rflat = torch.cat([p.detach().flatten() for p in ref.parameters()])
        out.put((torch.equal(gathered[0], gathered[1]), float((gathered[0] - rflat).abs().max())))
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same, err = out.get()
    assert same, "replicas diverged"
    assert err < 1e-6, err


def test_shard_range_partitions_clips():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_reducer_is_a_noop():
    net = torch.nn.Linear(4, 2)
    red = FlatGradAllReducer(net, n_buckets=2)
    red.zero_grad()
    red.arm()
    net(torch.ones(3, 4)).sum().backward()
    red.finish()
    assert torch.allclose(net.weight.grad, torch.full((2, 4), 3.0))
    assert net.weight.grad.data_ptr() >= red.flat.data_ptr()


def test_buckets_are_reduced_in_index_order_whatever_the_hook_order(monkeypatch):
    """Every rank must issue the same sequence of collectives even when its gradient hooks complete in another order
    (a graph that differs between ranks, unused parameters): bucket i is launched only after buckets 0..i-1."""
    import video_frame_inpainting_b200.parallel as par
    net = torch.nn.Sequential(*[torch.nn.Linear(8, 8) for _ in range(4)])
    red = FlatGradAllReducer(net, n_buckets=4)
    assert len(red.buckets) >= 3
    launched = []

    class _Done(object):
        def wait(self):
            pass

    def fake_all_reduce(t, op=None, async_op=False):
        launched.append((t.data_ptr() - red.flat.data_ptr()) // 4)
        return _Done()
    monkeypatch.setattr(par.dist, "all_reduce", fake_all_reduce)
    monkeypatch.setattr(par.dist, "get_world_size", lambda: 2)
    monkeypatch.setattr(FlatGradAllReducer, "active", property(lambda self: True))
    starts = [lo for lo, _ in red._slices]
    # hooks arrive bucket-reversed: nothing may be launched until bucket 0 is complete, then all of them in index order
    red.arm()
    for bucket in reversed(red.buckets):
        for p in bucket:
            red._hook(p)
    assert launched == starts
    # a bucket whose hooks never fire (unused parameters) is flushed by finish(), still in index order
    launched.clear()
    red.arm()
    for p in red.buckets[1]:
        red._hook(p)
    assert launched == []
    red.finish()
    assert launched == starts
