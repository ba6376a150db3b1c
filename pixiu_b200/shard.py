"""Key partitioning for multi-GPU operation (one process and one store per GPU).

The reference window is ~12.5 MB (PiXiuCtrl.cpp:13), so a shard never needs another shard's text:
records are routed to `owner(key)`, every rank runs the unmodified single-GPU path on its records
and there is no data-path collective.  Lookups are routed the same way; results are gathered with
`torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

_FNV_OFFSET = np.uint64(0xCBF29CE484222325)
_FNV_PRIME = np.uint64(0x100000001B3)


def key_hash(key: bytes) -> int:
    """64-bit FNV-1a (stable across processes, unlike Python's hash())"""
    h = 0xCBF29CE484222325
    for b in key:
        h ^= b
        h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def owner(key: bytes, world: int) -> int:
    return key_hash(key) % world


def owners_packed(keys: np.ndarray, key_off: np.ndarray, world: int) -> np.ndarray:
    """owner rank of every key of a packed batch (vectorised FNV-1a over ragged rows)"""
    n = len(key_off) - 1
    h = np.full(n, _FNV_OFFSET, dtype=np.uint64)
    lens = np.diff(key_off)
    with np.errstate(over="ignore"):
        for j in range(int(lens.max()) if n else 0):
            live = lens > j
            b = keys[(key_off[:-1] + j)[live]].astype(np.uint64)
            h[live] = (h[live] ^ b) * _FNV_PRIME
    return (h % np.uint64(world)).astype(np.int64)


def take_packed(data: np.ndarray, off: np.ndarray, idx: np.ndarray):
    """sub-batch of a packed batch: rows `idx` in that order"""
    lens = np.diff(off)[idx]
    new_off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    total = int(new_off[-1])
    if total == 0:
        return np.zeros(0, dtype=data.dtype), new_off
    pos = np.arange(total, dtype=np.int64) + np.repeat(off[:-1][idx] - new_off[:-1], lens)
    return data[pos], new_off


def partition(keys, key_off, vals, val_off, world: int, rank: int):
    """the records of a packed batch owned by `rank`, plus their indices in the batch"""
    own = owners_packed(keys, key_off, world)
    idx = np.nonzero(own == rank)[0]
    kd, ko = take_packed(keys, key_off, idx)
    vd, vo = take_packed(vals, val_off, idx)
    return idx, kd, ko, vd, vo
