"""Real two-GPU run of the sharded-window setitem with the collectives inside the library: pixiu_mg_setitem_batch
issues ncclAllReduce(MAX) / ncclAllReduce(MIN) on the store's stream (pixiu_b200/csrc/mgcomm.cu); the ranks only
exchange the NCCL unique id through a file.  No torch.distributed anywhere.  Skipped on boxes with fewer than
2 GPUs (the single-GPU emulation in test_gpu_parity.py covers the same phases with the host playing the collectives);
`gpurun --gpus 2 -- python -m pytest tests/test_mg_nccl.py -m gpu` is the run whose log is kept under profiles/."""
import multiprocessing as mp
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import pyoracle as po
    from pixiu_b200 import ctrl, multigpu, synth

    kd, ko, vd, vo = synth.gen_html_pages(60, seed=11, max_len=30000, mean_len=15000)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    c = ctrl.PiXiuCtrl(device=rank, rotate_policy=ctrl.ROTATE_RECORDS)
    multigpu.init_comm_file(c, rank, world, os.path.join(out_dir, "nccl_id"))
    for a in range(0, len(keys), 17):
        rc, saved = c.mg_setitem_batch(keys[a:a + 17], vals[a:a + 17])
        assert not rc.any()
    # bytes identical to ONE oracle window holding every record (the shards never see each other's text)
    w = po.OracleWindow()
    for i, (k, v) in enumerate(zip(keys, vals)):
        assert c.encoded(0, i) == w.encode(po.make_doc(k, v)), f"rank {rank}: record {i}"
    buf, off, found = c.getitem_batch(keys)
    assert found.all()
    for i, (k, v) in enumerate(zip(keys, vals)):
        assert buf[off[i]:off[i + 1]].tobytes() == po.make_doc(k, v)
    # plain single-GPU updates are refused on a shard (they would desynchronise the ranks)
    with pytest.raises(ctrl.PiXiuError):
        c.setitem(b"k", b"v")
    st = c.mg_stats()
    assert st.world == world and st.batches == 4 and st.max_reduce_bytes > 0 and st.nccl_version > 0
    open(os.path.join(out_dir, f"ok{rank}"), "w").write(
        f"rank {rank}: nccl {st.nccl_version}, MAX {st.max_reduce_bytes} B in {st.max_reduce_ms:.3f} ms, "
        f"MIN {st.min_reduce_bytes} B in {st.min_reduce_ms:.3f} ms\n")
    c.free_prop()


@pytest.mark.parametrize("world", [2])
def test_two_gpu_sharded_window_nccl(tmp_path, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
    for r, p in enumerate(procs):
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0, f"rank {r} exited with {p.exitcode}"
        print(open(tmp_path / f"ok{r}").read().strip())
