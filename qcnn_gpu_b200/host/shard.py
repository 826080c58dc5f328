"""Multi-GPU partitioning of the QVRCNN pass (one process per GPU, torch.distributed for plumbing).

The reference is single-GPU (device 0 hard-wired, inference/kernel.cu:86); the path shards two ways
(SURVEY.md 8e):
  * frames are independent (forward_blu reads only I1.x, inference/qvrcnn.cu:168-242): contiguous blocks
    of frames per rank, no communication during compute;
  * inside one frame the net is a stencil of radius 2+2+1+1 = 6 rows: horizontal strips with a 6-row
    INPUT halo from each neighbour (one neighbour exchange before compute), outputs disjoint.
The only collective is the sum of the exact int64 SSE for the PSNR report (inference/yuv_data.cpp:87-97).
Everything here is backend-agnostic: the per-rank compute is a callable supplied by the caller.  On GPUs the
halo rows are NOT exchanged by this module: the fused kernel reads them from the neighbours' peer-mapped
memory (qv_strip_*, host/multi_gpu.py); exchange_halos is the message-passing statement of the same
partition, used by the gloo CPU tests and by callers without peer access.
"""
from __future__ import annotations

from typing import Callable, Tuple

HALO = 6


def split(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, start+count) of `n` items for `rank` of `world` (first n%world ranks get one more)."""
    base, rem = divmod(n, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def strip_window(h: int, rank: int, world: int) -> Tuple[int, int, int, int]:
    """(y0, y1, r0, r1): rank's output rows [y0,y1) and the input rows [r0,r1) it needs (6-row halo, clipped).
    Every strip must hold at least HALO rows: the halo of a strip then lies entirely inside its direct neighbours
    (a shorter strip would need rows of the strip after next, which neither exchange_halos nor qv_strip_* carry)."""
    if world > 1 and h // world < HALO:
        raise ValueError("%d rows over %d ranks leaves strips shorter than the %d-row halo: use fewer ranks" % (h, world, HALO))
    y0, n = split(h, rank, world)
    y1 = y0 + n
    return y0, y1, max(0, y0 - HALO), min(h, y1 + HALO)


def exchange_halos(own, y0: int, y1: int, h: int, rank: int, world: int, dist, alloc: Callable):
    """Neighbour exchange of input-luma halo rows.  `own` is a [y1-y0, W] uint8 tensor holding this rank's rows;
    returns a [r1-r0, W] tensor = [halo from rank-1 | own | halo from rank+1].  Point-to-point only
    (isend/irecv: NVLink P2P under NCCL)."""
    import torch
    _, _, r0, r1 = strip_window(h, rank, world)
    w = own.shape[1]
    out = alloc((r1 - r0, w))
    out[y0 - r0:y0 - r0 + (y1 - y0)] = own
    ops = []
    top_n, bot_n = y0 - r0, r1 - y1
    if rank > 0:
        up_y0, up_y1, _, up_r1 = strip_window(h, rank - 1, world)
        send_n = up_r1 - up_y1                       # rows the upper neighbour needs from me
        if send_n > 0:
            ops.append(dist.P2POp(dist.isend, own[:send_n].contiguous(), rank - 1))
        if top_n > 0:
            ops.append(dist.P2POp(dist.irecv, out[:top_n], rank - 1))
    if rank < world - 1:
        dn_y0, _, dn_r0, _ = strip_window(h, rank + 1, world)
        send_n = dn_y0 - dn_r0
        if send_n > 0:
            ops.append(dist.P2POp(dist.isend, own[own.shape[0] - send_n:].contiguous(), rank + 1))
        if bot_n > 0:
            ops.append(dist.P2POp(dist.irecv, out[out.shape[0] - bot_n:], rank + 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def psnr_from_global_sse(sse_local: int, n_local: int, dist=None, device=None) -> Tuple[float, int, int]:
    """All-reduce (sum) of the exact integer SSE and the sample count, then the reference's formula
    mse = SSE/n, psnr = 10 log10(65025/mse).  Bit-identical to the sequential double accumulation of
    vrcnn_data::psnr because every partial sum is an integer < 2^53."""
    import math
    import torch
    t = torch.tensor([sse_local, n_local], dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    sse, n = int(t[0].item()), int(t[1].item())
    mse = float(sse) / float(n)
    return (10 * math.log10(65025.0 / mse) if mse > 0 else float("inf")), sse, n
