"""Data-parallel plumbing for the training step.

The reference is single-process / single-device; the north_star shards the batch over 1/2/4/8
B200 of one box, one process per GPU.  Every hot-path kernel is per-sequence independent, so the
only exchange is the weight-gradient all-reduce: parameters are grouped into flat buckets in
reverse registration order (the order backward produces them); a post-accumulate-grad hook copies
each finished gradient into its bucket and, when a bucket is complete, launches an asynchronous
all-reduce on the communication stream so it overlaps the rest of backward.  ``finish()`` waits
for the outstanding collectives and averages.  Inference is batch-sharded with no collective.

Buckets: one per producer (``producers``: head, block 2: 19 MB, block 1: 19 MB at H = 512), up to 24 MB.  Round 1 used
32 MB buckets in plain reverse order: the first bucket spanned the head, block 2 and part of block 1 and NOTHING was
reduced before the very last weight-gradient GEMM.  Now block 2's gradients reduce under block 1's
recurrent backward kernel (which leaves 20 SMs idle) and only block 1's own bucket is exposed -- as one call (8 MB
buckets made it three LL-protocol calls of 43 us each on 2 GPUs), started at the event its backward records after the
weight-gradient GEMMs (``mark_grads_ready``) so that it runs beside the block's input-gradient GEMM, which leaves it
16 SMs (``rcnn_reserve_sms``: a persistent GEMM otherwise holds one CTA on every SM and NCCL's kernel waits for it).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


# Events recorded by a backward function right after the kernels that produced a parameter's gradient (keyed by
# id(param)): a bucket whose gradients all carry one waits for those events instead of for everything the compute
# stream has been given so far -- block 1's input-gradient GEMM, enqueued after its weight-gradient GEMMs but before
# autograd runs the hooks, then overlaps the all-reduce of block 1's buckets (model._BiLSTMBlockFn.backward).
GRAD_READY = {}


def mark_grads_ready(params):
    """Called from a backward pass: the gradients of ``params`` are complete at this point of the current stream."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    ev = torch.cuda.Event()
    ev.record()
    for p in params:
        GRAD_READY[id(p)] = ev


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced slice of ``n_items`` for ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Done:
    """Stand-in for a collective whose completion is already ordered on the communication stream."""

    def wait(self):
        return True


class GradAllReducer:
    def __init__(self, params, bucket_bytes: int = 24 << 20, group=None, reserve_sms: int = 16, producers=None,
                 wire_dtype: torch.dtype | None = None):
        """params: the parameters to reduce, in registration (forward) order.  producers: optional list of parameter
        lists in forward order (e.g. [block 1, block 2, head]) -- buckets never span two of them.  wire_dtype:
        ``torch.bfloat16`` sends the gradients as bf16 (half the bytes of the exposed all-reduce; the sum is still formed
        by NCCL in bf16, so the averaged gradients carry bf16 rounding) -- off by default (RCNN_DP_WIRE=bf16 turns it on for
        measurements): the default keeps the fp32 sums that make the N-GPU step equal the 1-GPU step on the whole batch."""
        import os as _os
        if wire_dtype is None and _os.environ.get("RCNN_DP_WIRE", "") == "bf16":
            wire_dtype = torch.bfloat16
        self.wire_dtype = wire_dtype
        self._wire = {}
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []          # (flat tensor, [(param, offset, numel)])
        self._bucket_of = {}
        self._pending = {}
        self._works = []
        self._handles = []
        self._needs_compute = {}     # bucket -> some gradient was copied on the compute stream: wait for that stream
        if self.world == 1 or not self.params:
            return
        # one bucket per producer: the gradients of a BidirectionalLSTM block (19 MB at H = 512) are final together, at
        # the end of that block's backward, so they travel as ONE all-reduce (three 8 MB calls cost 3 x 43 us with the
        # LL protocol, measured on 2 GPUs; one call of 19 MB is shorter and is launched once).  Parameters are taken in
        # reverse registration order (the order backward produces them); a bucket closes when the next parameter
        # belongs to another producer (`owner`: the module path up to the block) or would push it past bucket_bytes.
        if producers is None:
            producers = [self.params]
        else:
            producers = [[p for p in g if p.requires_grad] for g in producers]
            covered = {id(p) for g in producers for p in g}
            rest = [p for p in self.params if id(p) not in covered]
            if rest:
                producers = [rest] + producers
        for plist in reversed(producers):
            cur, cur_bytes = [], 0
            for p in reversed(plist):
                nbytes = p.numel() * 4
                if cur and cur_bytes + nbytes > bucket_bytes:
                    self._close(cur)
                    cur, cur_bytes = [], 0
                cur.append(p)
                cur_bytes += nbytes
            if cur:
                self._close(cur)
        import os
        reserve_sms = int(os.environ.get("RCNN_RESERVE_SMS", reserve_sms))
        if reserve_sms and self.params[0].is_cuda:
            from . import _lib
            _lib.check(_lib.lib().rcnn_reserve_sms(int(reserve_sms)), "rcnn_reserve_sms")
        self._avg = dist.get_backend(group) == "nccl"
        self._stream = torch.cuda.Stream() if self.params[0].is_cuda else None
        for p in self.params:
            self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _close(self, plist):
        total = sum(p.numel() for p in plist)
        flat = torch.zeros(total, dtype=torch.float32, device=plist[0].device)
        layout, off = [], 0
        for p in plist:
            layout.append((p, off, p.numel()))
            self._bucket_of[p] = len(self.buckets)
            off += p.numel()
        self.buckets.append((flat, layout))

    def _on_grad(self, p):
        bi = self._bucket_of[p]
        flat, layout = self.buckets[bi]
        ev = GRAD_READY.pop(id(p), None) if self._stream is not None else None
        self._needs_compute.setdefault(bi, False)
        for q, off, n in layout:
            if q is p:
                # same strides as the parameter (channels_last conv weights stay channels_last: the fused optimizers
                # require parameter and gradient layouts to match); any dense layout fills its n elements exactly
                dense = p.is_contiguous() or p.numel() == 0 or \
                    sum((sz - 1) * st for sz, st in zip(p.shape, p.stride())) + 1 == p.numel()
                view = flat[off:off + n].as_strided(p.shape, p.stride()) if dense else flat[off:off + n].view_as(p)
                if ev is not None:
                    # the gradient was complete at `ev`: copy it on the communication stream, behind that event only
                    # (not behind what the compute stream was given since, e.g. the block's input-gradient GEMM)
                    g = p.grad
                    self._stream.wait_event(ev)
                    with torch.cuda.stream(self._stream):
                        view.copy_(g)
                    g.record_stream(self._stream)
                else:
                    view.copy_(p.grad)
                    self._needs_compute[bi] = True
                p.grad = view          # the optimizer reads the reduced values in place
                break
        left = self._pending.get(bi, len(layout)) - 1
        self._pending[bi] = left
        if left == 0:
            self._launch(bi)

    def _launch(self, bi):
        flat, _ = self.buckets[bi]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if self._stream is not None:
            if self._needs_compute.pop(bi, True):
                self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                if self.wire_dtype is not None:
                    wire = self._wire.get(bi)
                    if wire is None:
                        wire = self._wire[bi] = torch.empty_like(flat, dtype=self.wire_dtype)
                    wire.copy_(flat)
                    work = dist.all_reduce(wire, op=op, group=self.group, async_op=True)
                    work.wait()                      # (the communication stream waits; the host does not)
                    flat.copy_(wire)
                    self._works.append((_Done(), flat, bi))
                else:
                    self._works.append((dist.all_reduce(flat, op=op, group=self.group, async_op=True), flat, bi))
        else:
            self._works.append((dist.all_reduce(flat, op=op, group=self.group, async_op=True), flat, bi))

    def finish(self, params=None):
        """Call after backward(): waits for the outstanding all-reduces and leaves averaged gradients in p.grad.
        ``params``: wait only for the buckets that hold these parameters (their optimizer step can then run while the
        later buckets -- the first block's, produced last -- are still being reduced); a final ``finish()`` waits for the rest."""
        if self.world == 1:
            return
        only = None if params is None else {self._bucket_of[p] for p in params if p in self._bucket_of}
        for bi, (flat, layout) in enumerate(self.buckets):   # buckets whose params got no grad this step
            if only is not None and bi not in only:
                continue
            if self._pending.get(bi, len(layout)) != 0 and any(p.grad is not None for p, _, _ in layout):
                self._launch(bi)
                self._pending[bi] = 0
        rest = []
        for work, flat, bi in self._works:
            if only is not None and bi not in only:
                rest.append((work, flat, bi))
                continue
            work.wait()
            if not self._avg:
                flat.div_(self.world)
        self._works = rest
        if only is not None:
            return
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self._works.clear()
        self._pending.clear()
        self._needs_compute.clear()
        GRAD_READY.clear()

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles.clear()
