"""Measurement aid: one getitem of the LAST key of a full C2 window (the whole chunk prefix is decoded: ~12 MB).
python profiles/prof_single_get.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402,F401

from pixiu_b200 import ctrl, synth  # noqa: E402

kd, ko, vd, vo = synth.gen_html_pages(500, seed=2)
keys = synth.unpack(kd, ko)
c = ctrl.PiXiuCtrl(rotate_policy=ctrl.ROTATE_REFERENCE)
c.setitem_batch((kd, ko), (vd, vo))
st = c.stats()
# the last record of the first (full) window
n0 = 0
while n0 < st.records and c.record_location(n0)[0] == 0:
    n0 += 1
last = n0 - 1
print("chunks", st.chunks, "records in chunk 0:", n0)
for k in (last, n0 // 2, 0):
    for rep in range(4):
        t0 = time.perf_counter()
        buf, off, found = c.getitem_batch([keys[k]])
        wall = (time.perf_counter() - t0) * 1e6
        s = c.stats()
    print(f"getitem(record {k} of chunk 0): decode kernels {s.last_getitem_gpu_ms * 1e3:.1f} us, lookup {s.last_lookup_gpu_ms * 1e3:.1f} us, "
          f"whole call {wall:.1f} us wall, {off[1]} bytes")
