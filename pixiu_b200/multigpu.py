"""Multi-GPU extended-window setitem (BASELINE config 5): rendezvous helpers.

The data plane is inside libpixiu_b200.so: `pixiu_mg_setitem_batch` runs the three encode phases and issues
`ncclAllReduce(MAX)` on the per-position match lengths and `ncclAllReduce(MIN)` on the per-run
(idx << 16 | to) candidates on the store's own CUDA stream (pixiu_b200/csrc/mgcomm.cu).  What is left for the
host program is the usual NCCL bootstrap: rank 0 creates a 128-byte unique id and every rank receives it.  Two
ways are offered; neither touches the data path:

  init_comm_file(ctrl, rank, world, path)    a file on a shared filesystem (no other dependency)
  init_comm_torch(ctrl)                      a torch.distributed broadcast, for programs launched by torchrun
"""
from __future__ import annotations

import os
import time

from .ctrl import NCCL_UNIQUE_ID_BYTES


def init_comm_file(c, rank: int, world: int, path: str, timeout_s: float = 120.0):
    if rank == 0:
        uid = c.mg_unique_id()
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
    else:
        t0 = time.time()
        while not (os.path.exists(path) and os.path.getsize(path) == NCCL_UNIQUE_ID_BYTES):
            if time.time() - t0 > timeout_s:
                raise TimeoutError(f"no NCCL unique id at {path}")
            time.sleep(0.01)
        uid = open(path, "rb").read()
    c.mg_comm_init(rank, world, uid)


def init_comm_torch(c, group=None):
    """bootstrap only: the unique id travels through an existing torch.distributed group (any backend)"""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.zeros(NCCL_UNIQUE_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(c.mg_unique_id()), dtype=torch.uint8).clone()
    t = t.to(dev)
    dist.broadcast(t, src=0, group=group)
    c.mg_comm_init(rank, world, bytes(t.cpu().numpy().tobytes()))


def setitem_sharded(c, keys, vals):
    """one replicated batch through the sharded window; returns (rc, saved) like setitem_batch"""
    return c.mg_setitem_batch(keys, vals)
