"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in CPU tests).

The path shards by independent image pairs: inference needs NO data-path collective (weak scaling);
training adds one exchange per step, the gradient all-reduce (3.67 M / 5.22 M fp32 parameters = 14.7 / 20.9 MB), issued
per bucket as soon as the bucket's gradients exist so that it overlaps the rest of the backward.  A single
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


def make_grad_sync(model: torch.nn.Module, bucket_bytes: int = 8 << 20, group=None) -> Callable[[torch.nn.Module], None]:
    """Returns f(model) that averages gradients over ranks: flatten each bucket, async all-reduce (SUM), unflatten / world.

    All buckets are launched before any is waited on, so the collectives pipeline on the NCCL stream.
    """
    buckets = make_buckets(list(model.parameters()), bucket_bytes)

    def sync(_model=None):
        world = dist.get_world_size(group)
        pending = []
        for bucket in buckets:
            grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
            flat = torch.cat([g.reshape(-1) for g in grads])
            pending.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, bucket))
        for work, flat, bucket in pending:
            work.wait()
            flat.div_(world)
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                p.grad.copy_(flat[off:off + n].view_as(p))
                off += n

    return sync


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
