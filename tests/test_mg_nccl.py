"""Real two-GPU run of the sharded-window setitem: NCCL all_reduce(MAX) / all_reduce(MIN) between the
C-ABI phases (pixiu_b200/multigpu.py).  Skipped on boxes with fewer than 2 GPUs (the single-GPU
emulation in test_gpu_parity.py covers the same code with the host playing the collectives)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from golden_util import load
    from oracle import pyoracle as po
    from pixiu_b200 import ctrl, multigpu, synth

    kd, ko, vd, vo = synth.gen_html_pages(60, seed=11, max_len=30000, mean_len=15000)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    c = ctrl.PiXiuCtrl(device=rank, rotate_policy=ctrl.ROTATE_RECORDS)
    c.mg_config(rank, world)
    for a in range(0, len(keys), 17):
        rc, saved = multigpu.setitem_sharded(c, keys[a:a + 17], vals[a:a + 17])
        assert not rc.any()
    w = po.OracleWindow()
    for i, (k, v) in enumerate(zip(keys, vals)):
        assert c.encoded(0, i) == w.encode(po.make_doc(k, v)), f"rank {rank}: record {i}"
    buf, off, found = c.getitem_batch(keys)
    assert found.all()
    for i, (k, v) in enumerate(zip(keys, vals)):
        assert buf[off[i]:off[i + 1]].tobytes() == po.make_doc(k, v)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_two_gpu_sharded_window_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
