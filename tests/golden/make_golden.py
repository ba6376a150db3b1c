"""Generates tests/golden/*.npz from the REAL reference (oracle/_ref).

Run here (where /root/reference is mounted):  python tests/golden/make_golden.py
The fixtures pin the oracle (and through it the CUDA path) to what the
unmodified reference stores for seeded inputs: per record the encoded bytes,
setitem rc, chunk serial / idx, arena pools (MemPool::nth) and blocks used in the current pool.
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from pixiu_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def fuzz_records(seed, n_docs, mode):
    rng = random.Random(seed)
    alpha = [b"ABCDE", bytes([250, 251, 252, 0, 1, 2]), bytes([65, 66, 251])][mode]
    keys, vals = [], []
    for d in range(n_docs):
        k = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 12))) + str(d).encode()
        r = rng.random()
        if r < 0.4 or not vals:
            v = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 300)))
        elif r < 0.7:
            src = rng.choice(vals)
            a = rng.randrange(len(src))
            b = rng.randint(a, min(len(src), a + 400))
            v = src[a:b] + bytes(rng.choice(alpha) for _ in range(rng.randint(0, 20))) + src[:rng.randint(0, len(src))]
        else:
            p = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 9)))
            v = (p * 400)[:rng.randint(1, 600)]
        keys.append(k)
        vals.append(v or b"x")
    return keys, vals


def dump(name, ref, keys, vals):
    ref.reset()
    r = ref.setitem_batch(keys, vals)
    assert r["chunk"][-1] == 0, "fixtures are single-window"
    encs = [ref.encoded(i) for i in range(len(keys))]
    kd, ko = synth.pack(keys)
    vd, vo = synth.pack(vals)
    ed, eo = synth.pack(encs)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), keys=kd, key_off=ko, vals=vd, val_off=vo,
                        enc=ed, enc_off=eo, rc=r["rc"], pools=r["pools"], pool_used=r["pool_used"])
    print(name, len(keys), "records", int(vo[-1]), "value bytes ->", int(eo[-1]), "encoded")


def main():
    ref = po.Ref()
    for mode in range(3):
        keys, vals = fuzz_records(100 + mode, 120, mode)
        dump("fuzz_mode%d" % mode, ref, keys, vals)
    kd, ko, vd, vo = synth.gen_urls_kv(600, seed=1)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    # duplicates exercise rc=1 (replace)
    keys += keys[:50]
    vals += vals[50:100]
    dump("c1_urls", ref, keys, vals)
    kd, ko, vd, vo = synth.gen_html_pages(6, seed=2, max_len=20000, mean_len=12000)
    dump("c2_pages", ref, synth.unpack(kd, ko), synth.unpack(vd, vo))
    kd, ko, vd, vo = synth.gen_nested(400, seed=3)
    dump("c3_nested", ref, synth.unpack(kd, ko), synth.unpack(vd, vo))
    ref.close()


if __name__ == "__main__":
    main()
