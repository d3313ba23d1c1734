"""Data-parallel plumbing of the TAI training step: one process per GPU, NCCL over NVLink.

Every op on the hot path is per-clip, so inference shards clips with no
collective at all; training adds exactly one exchange, the sum of the gradients.  ``FlatGradAllReducer``
keeps all gradients of a module in ONE contiguous buffer (each ``p.grad`` is a view into it), splits the
buffer into a few buckets and launches an asynchronous all-reduce per bucket from autograd's
post-accumulate hooks, so the reduction overlaps the rest of the backward pass; ``finish()`` waits and
divides by the world size.  Works with any ``torch.distributed`` backend (``nccl`` on the GPU box,
``gloo`` in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (clips are independent units)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_module(module, src=0):
    """Make all replicas start from rank `src`'s parameters and buffers."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)


class FlatGradAllReducer(object):
    ALIGN = 32  # floats: 128 B

    def __init__(self, module, n_buckets=4):
        self.params = [p for p in module.parameters() if p.requires_grad]
        # every gradient view starts on a 128 B boundary: autograd accumulates into the views in place
        # (`grad += new`, once per use of a shared weight -- ~800 adds per KTH step), and torch's vectorised
        # elementwise kernel needs 16 B-aligned operands; unaligned views fell back to the scalar kernel
        # (785 launches, 23 ms per step)
        align = self.ALIGN

        def padded(n):
            return (n + align - 1) // align * align

        total = sum(padded(p.numel()) for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, device=ref.device, dtype=ref.dtype)
        # gradients become available roughly in reverse parameter order: bucket 0 = last parameters
        order = list(reversed(self.params))
        target = max(1, (total + n_buckets - 1) // n_buckets)
        self.buckets, cur, cur_n = [], [], 0
        for p in order:
            cur.append(p)
            cur_n += padded(p.numel())
            if cur_n >= target:
                self.buckets.append(cur)
                cur, cur_n = [], 0
        if cur:
            self.buckets.append(cur)
        offset = total
        self._slices, self._bucket_of = [], {}
        for bi, bucket in enumerate(self.buckets):
            hi = offset
            for p in bucket:
                offset -= padded(p.numel())
                p.grad = self.flat[offset:offset + p.numel()].view_as(p)
                self._bucket_of[p] = bi
            self._slices.append((offset, hi))
        self._pending = [0] * len(self.buckets)
        self._next = 0
        self._handles = []
        self._armed = False
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    @property
    def active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def zero_grad(self):
        """Replaces optimizer.zero_grad(): the views must stay attached to the flat buffer."""
        self.flat.zero_()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr():
                raise RuntimeError("a gradient was detached from the flat buffer (zero_grad(set_to_none=True)?)")

    def arm(self):
        """Call right before backward(): buckets are reduced as soon as their last gradient lands."""
        self._pending = [len(b) for b in self.buckets]
        self._handles = []
        self._next = 0       # buckets are launched strictly in index order on every rank
        self._armed = True

    def _launch_ready(self, force=False):
        """Launch bucket i only after buckets 0..i-1: every rank issues the same sequence of collectives even when
        its hooks complete in a different order (a graph that differs between ranks, unused parameters) --
        differently sized all-reduces paired across ranks would hang or sum the wrong slices."""
        while self._next < len(self.buckets) and (force or self._pending[self._next] == 0):
            lo, hi = self._slices[self._next]
            self._handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
            self._next += 1

    def _hook(self, p):
        if not self._armed:
            return
        self._pending[self._bucket_of[p]] -= 1
        if self.active:
            self._launch_ready()

    def finish(self):
        """Call after backward(): reduce buckets whose hooks did not all fire (unused parameters), wait,
        average."""
        if not self._armed:
            return
        self._armed = False
        if not self.active:
            return
        self._launch_ready(force=True)   # buckets whose hooks did not all fire, still in index order
        for h in self._handles:
            h.wait()
        self._handles = []
        self.flat.div_(dist.get_world_size())
