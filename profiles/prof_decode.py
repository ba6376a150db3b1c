"""Measurement aid for the decode kernel: python profiles/prof_decode.py {c2|c3} N [trace]
c2: N synthetic HTML-like pages; c3: N nested 1 KB records.  Stores them, decodes every record three times (the last
call is the one ncu captures with `-k regex:k_decode_ -s 4 -c 2`), prints the kernel time and the decoder's
counters; with `trace` also the per-tile timeline (PIXIU_DEC_TRACE_FILE)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
wl, n = sys.argv[1], int(sys.argv[2])
trace = len(sys.argv) > 3 and sys.argv[3] == "trace"
if trace:
    os.environ["PIXIU_DEC_TRACE_FILE"] = "/tmp/dectrace.bin"
import torch  # noqa: E402

from pixiu_b200 import ctrl, synth  # noqa: E402

kd, ko, vd, vo = synth.gen_html_pages(n, seed=2) if wl == "c2" else synth.gen_nested(n, seed=3)
c = ctrl.PiXiuCtrl(rotate_policy=ctrl.ROTATE_REFERENCE)
c.setitem_batch((kd, ko), (vd, vo))
st = c.stats()
cap = int(st.doc_bytes + 64)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for rep in range(3):
    c.profile_enable(True)
    c.getitem_batch_dev((kd, ko), out.data_ptr(), cap)
    pr = c.profile()
    pd = pr["decode"]
    c.profile_enable(False)
    pieces, drains = c.debug_decode_counters()
    print(f"call {rep}: decode {pd['ms']:.3f} ms (literals {pr['decode_lit']['ms']:.3f} + copies {pr['decode_copy']['ms']:.3f}), "
          f"{pd['bytes'] / 1e6 / pd['ms']:.1f} GB/s, {pieces} copy pieces")
if trace:
    raw = open("/tmp/dectrace.bin", "rb").read()
    nw = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
    tr = np.frombuffer(raw[8 + 4 * nw:], dtype=np.uint64).reshape(nw, 4).astype(np.int64)
    t0 = tr[:, 0].min()
    ent, par, don, sw = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3, tr[:, 3]
    print("tiles", nw, "total us %.1f" % don.max(), "chunks", st.chunks)
    print("entry->literals done us: mean %.2f p50 %.2f p99 %.2f" % ((par - ent).mean(), np.median(par - ent), np.percentile(par - ent, 99)))
    print("literals done->complete us: mean %.2f p50 %.2f p99 %.2f" % ((don - par).mean(), np.median(don - par), np.percentile(don - par, 99)))
    print("sweeps: mean %.2f p99 %.0f; tiles with pending pieces %.3f" % (sw.mean(), np.percentile(sw, 99), (sw > 0).mean()))
    # tiles resident at a time = sum of (done - entry) / total
    print("mean tiles in flight %.0f" % ((don - ent).sum() / don.max()))
