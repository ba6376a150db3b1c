"""N>1 host logic on CPU: world-size-2 gloo run of the key partitioning / routing used by
bench.py --gpus N (per-rank stores are the CPU oracle here; the GPU store has the same API)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch

    from oracle import pyoracle as po
    from pixiu_b200 import shard, synth

    kd, ko, vd, vo = synth.gen_urls_kv(400, seed=1)          # every rank sees the same batch
    idx, skd, sko, svd, svo = shard.partition(kd, ko, vd, vo, world, rank)
    keys, vals = synth.unpack(skd, sko), synth.unpack(svd, svo)
    assert all(shard.owner(k, world) == rank for k in keys)
    store = po.OracleStore()
    for k, v in zip(keys, vals):
        store.setitem(k, v)
    # the partition is a disjoint cover of the batch
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(idx)], dtype=torch.int64))
    assert sum(int(c) for c in counts) == 400
    mark = torch.zeros(400, dtype=torch.int64)
    mark[torch.from_numpy(idx)] = 1
    dist.all_reduce(mark)
    assert bool((mark == 1).all())
    # lookups: every rank answers the queries it owns; answers are summed across ranks
    all_keys = synth.unpack(kd, ko) + [b"http://absent/%d" % i for i in range(50)]
    found = torch.zeros(len(all_keys), dtype=torch.int64)
    nbytes = torch.zeros(len(all_keys), dtype=torch.int64)
    for i, k in enumerate(all_keys):
        if shard.owner(k, world) == rank:
            d = store.getitem(k)
            if d is not None:
                found[i] = 1
                nbytes[i] = len(d)
    dist.all_reduce(found)
    dist.all_reduce(nbytes)
    assert found[:400].tolist() == [1] * 400 and found[400:].tolist() == [0] * 50
    all_vals = synth.unpack(vd, vo)
    assert nbytes[:400].tolist() == [len(po.make_doc(k, v)) for k, v in zip(all_keys[:400], all_vals)]
    # throughput aggregation the way bench.py does it: max over ranks of the elapsed time
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_partition_vectorised_matches_scalar():
    sys.path.insert(0, ROOT)
    from pixiu_b200 import shard, synth

    kd, ko, _, _ = synth.gen_urls_kv(300, seed=5)
    keys = synth.unpack(kd, ko)
    for world in (1, 2, 3, 8):
        assert shard.owners_packed(kd, ko, world).tolist() == [shard.owner(k, world) for k in keys]


def test_world_size_2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
