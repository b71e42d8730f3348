"""Data-parallel plumbing for batched renders (one process per GPU, torch.distributed over NCCL / NVLink).

The hot path shards by minibatch image: every kernel is indexed (pixel, image) and no kernel reads another image
(SURVEY.md 8(e)), so there is no data-path collective.  The only exchange is the sum of the gradients of parameters
shared by all images (mesh vertices, uv coordinates, the texture): `SharedGradReducer` keeps them in ONE flat bucket and
reduces it with at most two collectives per step -- the large tensors (the 12.6 MB texture gradient is complete as soon as
the texture backward has run) go out while interpolate / rasterize backward are still executing, the small remainder
(vertex and uv gradients, tens of KB, complete only at the very end of backward) follows in one latency-bound call.
"""
import torch


def shard_range(n_total, rank, world):
    """Contiguous [start, stop) slice of a batch of n_total images owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_shared_grads(grads, group=None):
    """Sum the given gradient tensors over all ranks with ONE all-reduce (flatten -> all_reduce -> scatter back in place)."""
    grads = [g for g in grads if g is not None]
    if not grads or not torch.distributed.is_initialized() or torch.distributed.get_world_size(group) == 1:
        return grads
    if len(grads) == 1:
        torch.distributed.all_reduce(grads[0], group=group)
        return grads
    flat = torch.cat([g.reshape(-1) for g in grads])
    torch.distributed.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return grads


class SharedGradReducer:
    """Bucketed, overlapped all-reduce of the gradients of parameters shared by every image of the batch.

        red = SharedGradReducer([("tex", tex.shape), ("uv", uv.shape), ("verts", verts.shape)], device)
        red.begin({"tex": tex, "uv": uv, "verts": verts})     # leaf tensors of this step (requires_grad)
        loss.backward()                                        # large gradients leave as soon as autograd has accumulated them
        grads = red.finish()                                   # dict name -> reduced gradient (views of the flat bucket)

    Layout: one flat fp32 buffer, large tensors (>= `early_bytes`) first, small ones behind them.  A post-accumulate hook on
    each leaf copies its gradient into its slot; for a large tensor the hook immediately launches an asynchronous all-reduce
    of that slot (NCCL runs it on its own stream: it overlaps the backward kernels that are still to come); `finish()`
    reduces the contiguous small tail with one more collective and waits.  With one rank (or no process group) it degrades
    to the copies.  `bytes_per_step` = payload a rank contributes per step.
    """

    def __init__(self, spec, device, group=None, early_bytes=1 << 20):
        self.group = group
        order = sorted(spec, key=lambda s: -int(torch.Size(s[1]).numel()))        # large first
        self.slots = {}
        off = 0
        for name, shape in order:
            n = int(torch.Size(shape).numel())
            self.slots[name] = (off, n, tuple(shape))
            off += n
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.early = {name for name, (o, n, _) in self.slots.items() if n * 4 >= early_bytes}
        self.small_start = min([o for name, (o, n, _) in self.slots.items() if name not in self.early], default=off)
        self.bytes_per_step = off * 4
        self._handles, self._works, self._seen = [], [], set()

    def _active(self):
        return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(self.group) > 1

    def view(self, name):
        o, n, shape = self.slots[name]
        return self.flat[o:o + n].view(shape)

    def begin(self, leaves):
        for h in self._handles:
            h.remove()
        self._handles, self._works, self._seen = [], [], set()
        for name, t in leaves.items():
            assert name in self.slots and tuple(t.shape) == self.slots[name][2], "unknown parameter or shape mismatch: %s" % name
            self._handles.append(t.register_post_accumulate_grad_hook(self._make_hook(name)))

    def _make_hook(self, name):
        def hook(p):
            v = self.view(name)
            v.copy_(p.grad)
            self._seen.add(name)
            if name in self.early and self._active():
                self._works.append(torch.distributed.all_reduce(v, group=self.group, async_op=True))
        return hook

    def finish(self):
        """Reduce what is still local, wait for everything.  Parameters whose hook never fired (no gradient this step)
        contribute zeros."""
        for name in self.slots:
            if name not in self._seen:
                self.view(name).zero_()
        if self._active():
            for name in self.early:
                if name not in self._seen:                      # keep the collective sequence identical on every rank
                    self._works.append(torch.distributed.all_reduce(self.view(name), group=self.group, async_op=True))
            if self.small_start < self.flat.numel():
                self._works.append(torch.distributed.all_reduce(self.flat[self.small_start:], group=self.group, async_op=True))
            for w in self._works:
                w.wait()
        for h in self._handles:
            h.remove()
        self._handles, self._works = [], []
        return {name: self.view(name) for name in self.slots}
