"""Generates tests/golden/c2_full_enc_len.npz from the REAL reference (oracle/_ref).

Run here (where /root/reference is mounted):  python tests/golden/make_c2_full.py   (~8 min, one core)

The fixture pins the FULL BASELINE config[1] corpus (bench.py's workload: synth.gen_html_pages(10000, seed=2),
356 MB raw): per record the encoded length the unmodified reference stores, the window (chunk) serial it landed in
and its idx inside the window, i.e. the compression ratio figure README.md:53 quotes and every rotation point
(PiXiuCtrl.cpp:13-17).  A CRC32 of each record's encoded bytes is kept as well, so byte identity is checked without
shipping 290 MB.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from pixiu_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PAGES, SEED = 10000, 2


def main():
    kd, ko, vd, vo = synth.gen_html_pages(PAGES, seed=SEED)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    ref = po.Ref()
    n = len(keys)
    enc_len = np.zeros(n, dtype=np.uint16)
    chunk = np.zeros(n, dtype=np.uint16)
    idx = np.zeros(n, dtype=np.uint16)
    crc = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        rc = ref.setitem(keys[i], vals[i])
        assert rc == 0
        info = ref.last_info()
        e = ref.encoded(info["idx"])
        enc_len[i], chunk[i], idx[i] = len(e), info["chunk"], info["idx"]
        crc[i] = zlib.crc32(e)
        if i % 500 == 0:
            print(i, int(chunk[i]), flush=True)
    ref.close()
    np.savez_compressed(os.path.join(HERE, "c2_full_enc_len.npz"), pages=PAGES, seed=SEED, enc_len=enc_len, chunk=chunk,
                        idx=idx, crc32=crc, raw_bytes=int(ko[-1] + vo[-1]))
    print("stored/raw", float(enc_len.astype(np.int64).sum()) / float(ko[-1] + vo[-1]), "windows", int(chunk[-1]) + 1)


if __name__ == "__main__":
    main()
