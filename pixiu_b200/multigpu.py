"""Multi-GPU extended-window setitem: the two collectives between the C-ABI phases.

`setitem_sharded(ctrl, keys, vals)` runs one batch through `pixiu_mg_setitem_{begin,mid,end}` and
issues `all_reduce(MAX)` on the per-position match lengths and `all_reduce(MIN)` on the per-run
(idx << 16 | to) candidates with torch.distributed (NCCL over NVLink on the GPUs).  The device
buffers handed out by the library are wrapped as torch tensors without a copy.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _DevArray:
    """__cuda_array_interface__ view of `count` uint32 at a raw device pointer"""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<u4", "data": (ptr, False), "version": 2}


def wrap_u32(ptr: int, count: int, device) -> torch.Tensor:
    # int32 view: torch has no uint32 collectives; MAX on match lengths (< 65536) is sign-safe, MIN on
    # candidates is made sign-safe by the caller-side bias below
    return torch.as_tensor(_DevArray(ptr, count), device=device).view(torch.int32)


def setitem_sharded(c, keys, vals, device=None, group=None):
    """one replicated batch through the sharded window; returns (rc, saved) like setitem_batch"""
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    p, cnt = c.mg_setitem_begin(keys, vals)
    if cnt:
        m = wrap_u32(p, cnt, device)
        torch.cuda.synchronize(device)
        dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
        torch.cuda.synchronize(device)
    p, cnt = c.mg_setitem_mid()
    if cnt:
        cand = wrap_u32(p, cnt, device)
        torch.cuda.synchronize(device)
        # unsigned MIN through a signed collective: flip the top bit, reduce, flip back
        cand ^= -0x80000000
        dist.all_reduce(cand, op=dist.ReduceOp.MIN, group=group)
        cand ^= -0x80000000
        torch.cuda.synchronize(device)
    return c.mg_setitem_end()
