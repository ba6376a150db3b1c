"""Measurement aid: per-tile timeline of k_decode_tiles on the C3 workload (PIXIU_DEC_TRACE_FILE makes the decode write
entry / end-of-literal-phase / done times and the sweep count of every tile).  python profiles/decode_trace.py 400000"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pixiu_b200 import ctrl, synth
n = int(sys.argv[1])
kd, ko, vd, vo = synth.gen_nested(n, seed=3)
c = ctrl.PiXiuCtrl(rotate_policy=ctrl.ROTATE_REFERENCE)
c.setitem_batch((kd, ko), (vd, vo))
st = c.stats(); cap = int(st.doc_bytes + 64)
out = torch.empty(cap, dtype=torch.uint8, device="cuda")
for rep in range(2):
    c.getitem_batch_dev((kd, ko), out.data_ptr(), cap)
os.environ["PIXIU_DEC_TRACE_FILE"] = "/tmp/dectrace.bin"
c.getitem_batch_dev((kd, ko), out.data_ptr(), cap)
print("decode ms", c.stats().last_getitem_gpu_ms)
raw = open("/tmp/dectrace.bin", "rb").read()
nw = int(np.frombuffer(raw[:8], dtype=np.uint64)[0])
rec = np.frombuffer(raw[8:8 + 4 * nw], dtype=np.uint32)
tr = np.frombuffer(raw[8 + 4 * nw:], dtype=np.uint64).reshape(nw, 4).astype(np.int64)
t0 = tr[:, 0].min()
ent, par, don, sw = (tr[:, 0] - t0) / 1e3, (tr[:, 1] - t0) / 1e3, (tr[:, 2] - t0) / 1e3, tr[:, 3] - 3
print("tiles", nw, "total us", don.max())
print("parse us: mean %.1f p50 %.1f p99 %.1f" % ((par - ent).mean(), np.median(par - ent), np.percentile(par - ent, 99)))
print("wait+copy us: mean %.1f p50 %.1f p99 %.1f" % ((don - par).mean(), np.median(don - par), np.percentile(don - par, 99)))
print("sweeps: mean %.1f p50 %.0f p99 %.0f" % (sw.mean(), np.median(sw), np.percentile(sw, 99)))
# one chunk: records of chunk 1
locs0 = c.record_location(int(rec[0]))
first = np.array([c.record_location(int(g))[0] for g in rec[:200]])
ch = 3
idx = [i for i in range(nw) if False]
# chunk of every work item through searchsorted on chunk firsts
nch = st.chunks
firsts = []
g = 0
import bisect
# derive chunk firsts by probing record_location on a coarse grid
cf = [0]
lo = 0
for cidx in range(1, nch):
    a, b = cf[-1], n - 1
    while a < b:
        m = (a + b) // 2
        if c.record_location(m)[0] >= cidx: b = m
        else: a = m + 1
    cf.append(a)
cf = np.array(cf)
chunk = np.searchsorted(cf, rec, side="right") - 1
sel = np.where(chunk == ch)[0]
o = np.argsort(rec[sel]); sel = sel[o]
r = rec[sel] - cf[ch]
print("chunk", ch, "records", len(sel))
for a in range(0, len(sel), max(len(sel) // 16, 1)):
    i = sel[a]
    print(f" rec {r[a]:6d}: entry {ent[i]:8.1f} parsed {par[i]:8.1f} done {don[i]:8.1f} sweeps {sw[i]:4d}")
d = don[sel]; e = ent[sel]
k = len(sel) // 2
print("done-time slope us/record (mid half): %.3f" % ((d[3 * k // 2] - d[k // 2]) / (r[3 * k // 2] - r[k // 2])))
print("entry-time slope us/record: %.3f" % ((e[3 * k // 2] - e[k // 2]) / (r[3 * k // 2] - r[k // 2])))
lag = d[1:] - d[:-1]
print("done(r)-done(r-1): mean %.3f p10 %.3f p50 %.3f p90 %.3f; fraction done after predecessor: %.3f" % (lag.mean(), np.percentile(lag, 10), np.median(lag), np.percentile(lag, 90), (lag > 0).mean()))
lag10 = d[10:] - d[:-10]
print("done(r)-done(r-10): p10 %.3f p50 %.3f p90 %.3f" % (np.percentile(lag10, 10), np.median(lag10), np.percentile(lag10, 90)))
print("in flight per chunk (entered, not done) at mid time:", int(((e <= np.median(d)) & (d > np.median(d))).sum()))
