"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in CPU tests).

The path shards by independent image pairs: inference needs NO data-path collective (weak scaling);
training adds one exchange per step, the gradient all-reduce (3.67 M / 5.22 M fp32 parameters = 14.7 / 20.9 MB), issued
per bucket from post-accumulate-grad hooks as soon as the bucket's gradients exist, so that it overlaps the rest of the
backward (``GradSync``); with `accelerator: "ddp"` the BatchNorm partial sums are all-reduced too (train_ops.set_sync_bn).  A single
high-resolution pair (BASELINE config 5) is split into row tiles with per-layer halo exchange between row neighbours;
``row_tiles`` / ``exchange_row_halo`` are the host-side pieces of that path.
"""
from __future__ import annotations

import os
from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None) -> Tuple[int, int]:
    """RANK / WORLD_SIZE / MASTER_* come from the launcher (torchrun); returns (rank, world)."""
    if not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend)
    return dist.get_rank(), dist.get_world_size()


def shard_pairs(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous block of image pairs owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_pairs, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def make_buckets(params: Sequence[torch.nn.Parameter], bucket_bytes: int = 8 << 20) -> List[List[torch.nn.Parameter]]:
    """Reverse-registration-order buckets (gradients become ready roughly in that order during backward)."""
    buckets, cur, size = [], [], 0
    for p in reversed([p for p in params if p.requires_grad]):
        cur.append(p)
        size += p.numel() * p.element_size()
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    return buckets


class GradSync:
    """Bucketed gradient all-reduce that overlaps the backward pass.

    Every bucket owns a persistent flat fp32 buffer.  A post-accumulate-grad hook per parameter copies the finished gradient
    into its slice of the bucket (no torch.cat); the hook that completes a bucket launches its asynchronous all-reduce (SUM)
    right there, i.e. WHILE autograd is still running the backward of the earlier layers (buckets are filled in reverse
    registration order, the order in which gradients become ready).  ``sync()`` -- called once after ``backward()`` -- launches
    the buckets that did not fill (parameters without a gradient this step contribute zeros), waits for all of them and writes
    the averaged gradients back with one multi-tensor copy per bucket.  ``launched_in_backward`` counts the buckets whose
    collective was issued from a hook (tests assert it is > 0).
    """

    def __init__(self, model: torch.nn.Module, bucket_bytes: int = 8 << 20, group=None):
        self.group = group
        self.buckets = make_buckets(list(model.parameters()), bucket_bytes)
        self.flat, self.views, self.where = [], [], {}
        for bi, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            flat = torch.zeros(n, device=bucket[0].device, dtype=torch.float32)
            off, views = 0, []
            for pi, p in enumerate(bucket):
                views.append(flat[off:off + p.numel()].view_as(p))
                self.where[p] = (bi, pi)
                off += p.numel()
            self.flat.append(flat)
            self.views.append(views)
        self._ready = [set() for _ in self.buckets]
        self._work = [None] * len(self.buckets)
        self.launched_in_backward = 0
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for b in self.buckets for p in b]

    def _launch(self, bi: int):
        self._work[bi] = dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _hook(self, p: torch.nn.Parameter):
        bi, pi = self.where[p]
        if self._work[bi] is not None or pi in self._ready[bi]:
            return                                       # second accumulation into the same grad: picked up by sync()'s fallback
        self.views[bi][pi].copy_(p.grad)
        self._ready[bi].add(pi)
        if len(self._ready[bi]) == len(self.buckets[bi]):
            self._launch(bi)
            self.launched_in_backward += 1

    def __call__(self, _model=None):
        world = dist.get_world_size(self.group)
        for bi, bucket in enumerate(self.buckets):       # buckets that never filled: absent gradients count as zeros
            if self._work[bi] is None:
                for pi, p in enumerate(bucket):
                    if pi not in self._ready[bi]:
                        if p.grad is None:
                            self.views[bi][pi].zero_()
                        else:
                            self.views[bi][pi].copy_(p.grad)
                self._launch(bi)
        for bi, bucket in enumerate(self.buckets):
            self._work[bi].wait()
            self.flat[bi].div_(world)
            for p in bucket:
                if p.grad is None:
                    p.grad = torch.empty_like(p)
            torch._foreach_copy_([p.grad for p in bucket], self.views[bi])
            self._work[bi] = None
            self._ready[bi].clear()

    def remove(self):
        for h in self._handles:
            h.remove()


def make_grad_sync(model: torch.nn.Module, bucket_bytes: int = 8 << 20, group=None) -> GradSync:
    """Install the overlapping bucketed gradient all-reduce on `model`; call the returned object once after backward()."""
    return GradSync(model, bucket_bytes, group)


def broadcast_module_state(model: torch.nn.Module, src: int = 0, group=None) -> None:
    """Rank `src`'s parameters and buffers to every rank (what DDP does at construction), so that replicas start identical."""
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src, group=group)


def row_tiles(height: int, world: int, unit: int = 16) -> List[Tuple[int, int]]:
    """[start, end) rows per rank for a single image split along H (the disparity axis), boundaries on multiples of
    `unit` px (two stride-2 stages at quarter resolution need 16).  2240 rows over 8 ranks -> 18,18,18,18,17,17,17,17 units."""
    if height % unit:
        raise ValueError(f"height {height} must be a multiple of {unit}")
    units = height // unit
    base, extra = divmod(units, world)
    out, start = [], 0
    for r in range(world):
        n = (base + (1 if r < extra else 0)) * unit
        out.append((start, start + n))
        start += n
    return out


def exchange_row_halo(x: torch.Tensor, halo: int, row_dim: int, wrap: bool = False, group=None) -> torch.Tensor:
    """Concatenate `halo` rows from the row-neighbour ranks above and below (zeros at the image border, or the
    opposite end's rows when `wrap` -- the circular phase sample needs that).  Blocking send/recv pairs, even ranks
    send first; works on gloo (CPU) and NCCL (CUDA)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = x.shape[row_dim]
    top = x.narrow(row_dim, 0, halo).contiguous()
    bot = x.narrow(row_dim, n - halo, halo).contiguous()
    from_up, from_dn = torch.zeros_like(top), torch.zeros_like(bot)
    up, dn = rank - 1, rank + 1
    if wrap:
        up, dn = up % world, dn % world

    def send_recv(send_buf, dst, recv_buf, src):
        ops = []
        if 0 <= dst < world and dst != rank:
            ops.append(dist.P2POp(dist.isend, send_buf, dst, group))
        if 0 <= src < world and src != rank:
            ops.append(dist.P2POp(dist.irecv, recv_buf, src, group))
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()

    send_recv(bot, dn, from_up, up)      # my bottom rows go down; I receive my upper halo from above
    send_recv(top, up, from_dn, dn)      # my top rows go up; I receive my lower halo from below
    if wrap and world == 1:
        from_up, from_dn = bot, top
    return torch.cat([from_up, x, from_dn], dim=row_dim)
