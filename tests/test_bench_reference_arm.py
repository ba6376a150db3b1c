"""bench.py --impl reference (the reference's own CPU implementation, oracle/_ref) on a tiny sample: the line carries the
contract's keys, the same `config` object as the GPU arm and the steps / warm-up it was asked for.  CPU only."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpixiu_ref.so")),
                    reason="oracle/_ref is not built (needs /root/reference)")
def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-pages", "2", "--ref-procs", "2", "--no-read-side"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([x for x in out.stdout.splitlines() if x.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "setitem_raw_input_throughput" and line["unit"] == "MB/s"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"] == {"value": line["value"], "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench

    class A:
        pages, window = 10000, "reference"

    assert line["config"] == bench.workload_config(A)   # the GPU arm prints this very object
