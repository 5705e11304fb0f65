"""Data-parallel gradient exchange for the trainer step (new functionality: the
reference is single-process).

One process per GPU; the batch is sharded by rank; gradients are averaged with
NCCL all-reduce over NVLink/NVSwitch.  Parameters are grouped into a few flat
buckets in REVERSE forward order (the order backward produces them); a
bucket's all-reduce is launched on a side stream from the
post-accumulate-grad hook of its last parameter, so it overlaps the remaining
dgrad/wgrad kernels.  `wait()` joins the side stream before the optimiser.

BatchNorm statistics stay rank-local (what DistributedDataParallel gives the
reference); buffers are broadcast from rank 0 at construction.
Works with the `gloo` backend on CPU tensors too (used by the world_size-2
tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_batch(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Equal contiguous shards along dim 0 (the loss is a per-shard mean, so shards must be equal:
    UG_unet.py:92)."""
    if t.shape[0] % world_size != 0:
        raise ValueError(f"global batch {t.shape[0]} is not divisible by world_size {world_size}")
    n = t.shape[0] // world_size
    return t[rank * n:(rank + 1) * n]


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.numel = sum(p.numel() for p in params)
        self.flat: Optional[torch.Tensor] = None
        self.pending = 0
        self.work = None


class _ArenaBucket:
    """A contiguous range of the packed weight-gradient arena (engine.enable_deferred_wgrad): reduced
    in place, no flatten / copy-back; the 1/world_size factor is applied by the optimiser kernel."""

    def __init__(self, packs, flat: torch.Tensor):
        self.packs = packs
        self.flat = flat
        self.pending = 0
        self.work = None


class GradBucketReducer:
    """Bucketed, overlapped gradient averaging across ranks."""

    def __init__(self, params: Sequence[torch.nn.Parameter], bucket_mb: float = 16.0,
                 process_group=None, overlap: bool = True, conv_packs=None, arena: Optional[torch.Tensor] = None,
                 average: bool = True, tail_mb: Optional[float] = None):
        """`conv_packs` / `arena`: the deferred convolutions and their flat packed-gradient arena, in
        arena order; their weights are skipped in `params` (they get no `.grad`).
        `average=False`: `.grad` is left pointing at the SUM over ranks inside the flat bucket (no
        per-tensor scale kernels after the all-reduce); the optimiser applies 1/world_size
        (`FusedRMSprop.dense_grad_scale`), as it does for the packed arena."""
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        self.average = average
        deferred = {id(pk.conv.weight) for pk in (conv_packs or [])}
        ps = [p for p in params if p.requires_grad and id(p) not in deferred]
        # reverse forward order = the order gradients become ready
        ps = list(reversed(ps))
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self.buckets: List[_Bucket] = []
        cur: List[torch.nn.Parameter] = []
        size = 0
        for p in ps:
            cur.append(p)
            size += p.numel()
            if size >= cap:
                self.buckets.append(_Bucket(cur))
                cur, size = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        self._owner = {}
        self._hooks = []
        self._stream = None
        self.arena_buckets: List[_ArenaBucket] = []
        self._tail_arena: Optional[_ArenaBucket] = None
        self._packs = list(conv_packs or [])
        if self._packs:
            cur_p, start, off = [], 0, 0
            for pk in self._packs:
                n = pk.conv.weight.numel()
                assert pk._dwp is not None and pk._dwp.data_ptr() == arena.data_ptr() + 4 * off, \
                    "conv_packs must be given in arena order"
                cur_p.append(pk)
                off += n
                if off - start >= cap:
                    self.arena_buckets.append(_ArenaBucket(cur_p, arena[start:off]))
                    cur_p, start = [], off
            if cur_p:
                self.arena_buckets.append(_ArenaBucket(cur_p, arena[start:off]))
            # Light tail: the LAST bucket cannot start before backward ends and the optimiser waits for it,
            # so its all-reduce is exposed.  Keep it to the packs that finish last (<= tail_mb of gradients;
            # the first layers of the network are small) and give the rest to an earlier bucket that still
            # overlaps the remaining backward.  It is launched from wait() in ONE coalesced NCCL call together
            # with the flat small-tensor bucket.
            tail_mb = float(os.environ.get("UGP_DP_TAIL_MB", "2.0")) if tail_mb is None else tail_mb
            tail_cap = int(tail_mb * 1024 * 1024 / 4)
            last = self.arena_buckets[-1]
            if tail_cap > 0 and len(last.packs) > 1 and last.flat.numel() > tail_cap:
                keep, n_keep = [], 0
                for pk in reversed(last.packs):
                    n = pk.conv.weight.numel()
                    if keep and n_keep + n > tail_cap:
                        break
                    keep.insert(0, pk)
                    n_keep += n
                if len(keep) < len(last.packs):
                    head = last.packs[: len(last.packs) - len(keep)]
                    n_head = last.flat.numel() - n_keep
                    self.arena_buckets[-1] = _ArenaBucket(head, last.flat[:n_head])
                    self.arena_buckets.append(_ArenaBucket(keep, last.flat[n_head:]))
            self._tail_arena = self.arena_buckets[-1] if (tail_cap > 0 and len(self.arena_buckets) > 1) else None
            for ab in self.arena_buckets:
                for pk in ab.packs:
                    self._owner[id(pk)] = ab
                    pk.on_grad_ready = self._on_pack_ready if (overlap and self.world_size > 1) else None
        for b in self.buckets:
            for p in b.params:
                self._owner[p] = b
                if overlap and self.world_size > 1:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad_ready))
        self.reset()

    def reset(self):
        for b in self.buckets:
            b.pending = len(b.params)
            b.work = None
        for ab in self.arena_buckets:
            ab.pending = len(ab.packs)
            ab.work = None

    def _side_stream_ctx(self, dev):
        if dev.type != "cuda":
            return _null()
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev, priority=_comm_priority())
        self._stream.wait_stream(torch.cuda.current_stream(dev))
        from .engine import WgradOverlap
        for st in WgradOverlap.streams():            # a bucket's packs may come from either weight-gradient stream
            self._stream.wait_stream(st)
        return torch.cuda.stream(self._stream)

    def _on_pack_ready(self, pk):
        ab = self._owner[id(pk)]
        ab.pending -= 1
        if ab.pending == 0 and ab is not self._tail_arena:       # the tail goes out from wait(), coalesced
            self._launch_arena(ab)

    def _launch_arena(self, ab: _ArenaBucket):
        with self._side_stream_ctx(ab.flat.device):
            ab.work = dist.all_reduce(ab.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    # -- hook path (during backward) --
    def _on_grad_ready(self, p):
        b = self._owner[p]
        b.pending -= 1
        if b.pending == 0 and not (b is self.buckets[-1] and self._tail_arena is not None):
            self._launch(b)

    def _launch(self, b: _Bucket):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b.params]
        dev = grads[0].device
        if dev.type == "cuda":
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev, priority=_comm_priority())
            self._stream.wait_stream(torch.cuda.current_stream(dev))
            ctx = torch.cuda.stream(self._stream)
        else:
            ctx = _null()
        with ctx:
            # the side stream forked from the current stream above and is joined again in wait()
            # before the optimizer runs, so the gradient tensors outlive their use here
            flat = torch.cat([g.reshape(-1) for g in grads])
            b.work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            b.flat = flat

    def flush_arena(self):
        """Launches the all-reduce of every arena bucket that has not gone out yet (the light tail): called
        when the last packed weight gradient is queued, so that the packed-weight optimiser can start beside
        the end of backward.  The flat small-tensor bucket then goes out alone from wait()."""
        if self.world_size == 1:
            return
        for ab in self.arena_buckets:
            if ab.work is None and ab.flat.is_cuda:
                self._launch_arena(ab)
        # async_op collectives run on NCCL's own stream: work.wait() is what orders a stream after them
        for ab in self.arena_buckets:
            if ab.work is not None:
                ab.work.wait()                       # the CURRENT stream waits; wait() may wait again later

    def _launch_tail(self, ab: _ArenaBucket, b: _Bucket):
        """The exposed end of the exchange as ONE NCCL launch: the light last arena bucket and the flat
        bucket of the small tensors (biases, BatchNorm affine, heads), coalesced."""
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b.params]
        with self._side_stream_ctx(ab.flat.device):
            flat = torch.cat([g.reshape(-1) for g in grads])
            with dist._coalescing_manager(group=self.group, device=ab.flat.device, async_ops=True) as cm:
                dist.all_reduce(ab.flat, op=dist.ReduceOp.SUM, group=self.group)
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            ab.work = b.work = cm
            b.flat = flat

    # -- after backward --
    def wait(self):
        """Finish every bucket and write the averaged gradients back into `.grad`."""
        if self.world_size == 1:
            self.reset()
            return
        tail = self._tail_arena
        if tail is not None and tail.work is None and self.buckets and self.buckets[-1].work is None \
                and tail.flat.is_cuda:
            self._launch_tail(tail, self.buckets[-1])
        for ab in self.arena_buckets:
            if ab.work is None:
                self._launch_arena(ab)
        for b in self.buckets:
            if b.work is None:          # hooks disabled, or a parameter got no gradient this step
                self._launch(b)
        for ab in self.arena_buckets:
            ctx = torch.cuda.stream(self._stream) if ab.flat.is_cuda else _null()
            with ctx:
                ab.work.wait()
        inv = 1.0 / self.world_size
        for b in self.buckets:
            dev = b.flat.device
            ctx = torch.cuda.stream(self._stream) if dev.type == "cuda" else _null()
            with ctx:
                b.work.wait()
                off = 0
                for p in b.params:
                    n = p.numel()
                    g = b.flat[off:off + n].view_as(p)
                    if not self.average:
                        p.grad = g                      # a view of the summed bucket: no kernel
                    elif p.grad is None:
                        p.grad = g * inv
                    else:
                        torch.mul(g, inv, out=p.grad)
                    off += n
        if self._stream is not None:
            torch.cuda.current_stream().wait_stream(self._stream)
        self.reset()

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
        for pk in self._packs:
            pk.on_grad_ready = None


def _comm_priority() -> int:
    """Priority of the stream the all-reduces are queued on.  High (-1, the default): a bucket's
    collective starts as soon as SMs free up, so the exchange finishes under the remaining backward;
    UGP_NCCL_STREAM_PRIORITY=0 lets the compute kernels win the SMs instead (A/B switch)."""
    return int(os.environ.get("UGP_NCCL_STREAM_PRIORITY", "-1"))


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def broadcast_module_state(module: torch.nn.Module, src: int = 0, process_group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src=src, group=process_group)
