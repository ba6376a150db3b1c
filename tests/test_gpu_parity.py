"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
committed reference fixtures.  Bit-exact: encoded bytes, rc, saved bytes, decoded bytes.

Nothing here reads /root/reference; the oracle (oracle/liboracle.so) is the checker.
"""
import ctypes as C
import random

import numpy as np
import pytest

from golden_util import GOLDEN, load
from oracle import pyoracle as po
from pixiu_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctrl_mod():
    import torch  # noqa: F401  (only to fail fast with a clear message when there is no GPU)

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from pixiu_b200 import ctrl

    return ctrl


def _oracle_encode_all(keys, vals, strict):
    w = po.OracleWindow(strict251=strict)
    docs = [po.make_doc(k, v) for k, v in zip(keys, vals)]
    return docs, [w.encode(d) for d in docs]


# --------------------------------------------------------------------------- sort
@pytest.mark.parametrize("n,bits", [(1, 8), (1000, 13), (4096, 32), (100003, 47), (1 << 20, 63)])
def test_radix_sort_matches_numpy(ctrl_mod, n, bits):
    L = ctrl_mod.load_library()
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << bits, size=n, dtype=np.uint64)
    if n > 10:
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)]  # many duplicates: stability matters
    k = keys.copy()
    vout = np.zeros(n, dtype=np.uint32)
    rc = L.pixiu_debug_sort_pairs(0, k.ctypes.data_as(C.c_void_p), None, n, bits, vout.ctypes.data_as(C.c_void_p))
    assert rc == 0
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(vout, order.astype(np.uint32))


# --------------------------------------------------------------------------- encode
@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("n_batches", [1, 3])
def test_setitem_matches_reference_fixture(ctrl_mod, name, n_batches):
    g = load(name)
    keys, vals = g["keys"], g["vals"]
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS, strict251=True)
    n = len(keys)
    cuts = [round(i * n / n_batches) for i in range(n_batches + 1)]
    rcs, saveds = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        rc, saved = c.setitem_batch(keys[a:b], vals[a:b])
        rcs.append(rc)
        saveds.append(saved)
    rc = np.concatenate(rcs)
    saved = np.concatenate(saveds)
    assert rc.tolist() == g["rc"].tolist()
    for i in range(n):
        enc = c.encoded(0, i)
        assert enc == g["enc"][i], f"{name}: record {i} encoded bytes differ from the reference"
        assert saved[i] == len(po.make_doc(keys[i], vals[i])) - len(enc)
    c.free_prop()


def test_window_arrays_match_model(ctrl_mod):
    """SA / LCP / reach of a small window against the numpy model (pinpoints a failing stage)."""
    from pipeline_model import build_text, lcp_array, suffix_array

    g = load("fuzz_mode1")
    keys, vals = g["keys"][:40], g["vals"][:40]
    docs = [po.make_doc(k, v) for k, v in zip(keys, vals)]
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS, strict251=True)
    c.setitem_batch(keys, vals)
    text, is_sep, rec_start, rec_id, dist = build_text(docs)
    assert np.array_equal(c.debug_window_array("text", np.uint8), text)
    assert np.array_equal(c.debug_window_array("dist", np.uint16), dist.astype(np.uint16))
    sa = suffix_array(text, is_sep)
    assert np.array_equal(c.debug_window_array("sa", np.uint32), sa.astype(np.uint32))
    lcp = lcp_array(text, dist, sa)
    assert np.array_equal(c.debug_window_array("lcp", np.uint32), lcp.astype(np.uint32))
    c.free_prop()


def test_setitem_html_pages_vs_oracle(ctrl_mod):
    kd, ko, vd, vo = synth.gen_html_pages(40, seed=7, max_len=60000)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    docs, encs = _oracle_encode_all(keys, vals, strict=False)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS)
    rc, saved = c.setitem_batch((kd, ko), (vd, vo))
    assert not rc.any()
    for i in range(len(keys)):
        assert c.encoded(0, i) == encs[i], f"page {i}"
    assert saved.tolist() == [len(d) - len(e) for d, e in zip(docs, encs)]
    st = c.stats()
    assert st.encoded_bytes == sum(map(len, encs)) and st.raw_bytes == int(ko[-1] + vo[-1])
    # decode everything back through getitem
    buf, off, found = c.getitem_batch((kd, ko))
    assert found.all()
    for i, d in enumerate(docs):
        assert buf[off[i]:off[i + 1]].tobytes() == d, f"page {i} round trip"
    c.free_prop()


def test_setitem_nested_and_periodic_vs_oracle(ctrl_mod):
    kd, ko, vd, vo = synth.gen_nested(1500, seed=3)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    docs, encs = _oracle_encode_all(keys, vals, strict=False)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS)
    c.setitem_batch((kd, ko), (vd, vo))
    for i in range(len(keys)):
        assert c.encoded(0, i) == encs[i], f"record {i}"
    buf, off, found = c.getitem_batch((kd, ko))
    assert found.all()
    for i, d in enumerate(docs):
        assert buf[off[i]:off[i + 1]].tobytes() == d
    c.free_prop()


@pytest.mark.parametrize("piece_cap", [1, 6, 40])
def test_decode_piece_batches(ctrl_mod, piece_cap):
    """the copy kernel takes the pieces of a tile in batches (128 by default); tiny batches force the multi-batch path
    (a piece may then wait for a piece of an earlier batch of its own tile): same bytes either way"""
    # HTML-like pages (many short references), nested records (long references, self-periodic runs), escapes
    for gen in ("html", "nested", "esc"):
        if gen == "html":
            kd, ko, vd, vo = synth.gen_html_pages(24, seed=11, max_len=30000, mean_len=16000)
        elif gen == "nested":
            kd, ko, vd, vo = synth.gen_nested(1200, seed=5)
        else:
            rng = np.random.default_rng(9)
            vals = []
            for i in range(60):
                base = rng.choice(np.array([250, 251, 252, 0, 1, 2, 65, 66], dtype=np.uint8), size=3000).tobytes()
                vals.append(base if i % 3 else vals[-1][:1500] + base[:1500] if vals else base)
            keys = [b"esc%04d" % i for i in range(len(vals))]
            kd, ko = synth.pack(keys)
            vd, vo = synth.pack(vals)
        keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
        c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS)
        c.setitem_batch((kd, ko), (vd, vo))
        c.debug_set_knob("piece_cap", piece_cap)
        buf, off, found = c.getitem_batch((kd, ko))
        assert found.all()
        for i, (k, v) in enumerate(zip(keys, vals)):
            assert buf[off[i]:off[i + 1]].tobytes() == po.make_doc(k, v), f"{gen} record {i}"
        pieces, _ = c.debug_decode_counters()
        assert pieces > 0
        c.free_prop()


def test_edge_records(ctrl_mod):
    """maximum lengths (PiXiuCtrl.cpp:121-174), key-only docs, long runs of 251, long periodic values"""
    recs = [
        (b"A" * 65533, b""),                       # max key-only record: 65,535 decoded bytes
        (b"k1", b"B" * 65527),                     # max k+v
        (b"k2", bytes([251]) * 30000),             # escapes double it: 60,000 + terminators
        (b"k3", b"xyz" * 20000),
        (b"k4", bytes(range(256)) * 200),
        (bytes([251, 0, 251, 2, 251]), bytes([0, 2, 251, 251, 0])),
        (b"k5", b""),
        (b"k6", b"B" * 251 + b"!" + b"B" * 251),   # runs of exactly 251 (reference bug B1 territory)
    ]
    keys, vals = [r[0] for r in recs], [r[1] for r in recs]
    docs, encs = _oracle_encode_all(keys, vals, strict=False)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS)
    rc, saved = c.setitem_batch(keys, vals)
    for i in range(len(recs)):
        assert c.encoded(0, i) == encs[i], f"record {i}"
    buf, off, found = c.getitem_batch(keys)
    assert found.all()
    for i, d in enumerate(docs):
        assert buf[off[i]:off[i + 1]].tobytes() == d, f"record {i}"
    # oversize records are an explicit error and leave the store untouched
    with pytest.raises(ctrl_mod.PiXiuError) as e:
        c.setitem_batch([b"big"], [b"C" * 65529])
    assert e.value.code == ctrl_mod.ETOOLONG
    with pytest.raises(ctrl_mod.PiXiuError):
        c.setitem_batch([b""], [b"v"])
    assert c.stats().records == len(recs)
    c.free_prop()


@pytest.mark.gpu
@pytest.mark.parametrize("max_run,mode", [(32, 0), (32, 1), (32, 2), (200, 0), (200, 1)])
def test_runs_of_251(ctrl_mod, max_run, mode):
    """escape pairs and tile starts inside runs of 251: batches whose runs stay within 32 raw bytes take the walk-back
    path of the pair rule, longer ones the max-scan (k_doc_len decides per batch); both against the oracle, with runs
    placed around the 2,048-byte tile boundaries of the decoder"""
    rng = np.random.default_rng(251 + max_run)
    keys, vals = [], []
    lens = sorted(set(list(range(1, 12)) + [15, 16, 17, 30, 31, 32] + ([33, 61, 62, 63, 64, 65, 127, 128, 200] if max_run > 32 else [])))
    lens = [n for n in lens if n <= max_run]
    for j, n in enumerate(lens):
        for pad in (0, 1, 1000, 1019, 1020, 1021, 1022, 1023, 1024, 2045):
            filler = bytes(rng.integers(97, 123, size=pad, dtype=np.uint8))
            vals.append(filler + bytes([251]) * n + b"tail" + filler[:7] + bytes([251]) * (n // 2 + 1))
            keys.append(b"r%03d-%04d" % (n, pad) + (bytes([251]) * (n % 5)))
    # a key that ends in a run (the terminator 251 follows it directly) and repeated content (long references)
    keys.append(b"end" + bytes([251]) * min(max_run, 9))
    vals.append(vals[3] * 3)
    docs, encs = _oracle_encode_all(keys, vals, strict=False)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS)
    c.debug_set_knob("lastnon_mode", mode)           # 0: k_doc_len decides per batch, 1: always the scan, 2: always the walk
    half = len(keys) // 2
    c.setitem_batch(keys[:half], vals[:half])        # (two batches into one window)
    c.setitem_batch(keys[half:], vals[half:])
    for i in range(len(keys)):
        assert c.encoded(0, i) == encs[i], f"record {i} ({keys[i][:9]})"
    buf, off, found = c.getitem_batch(keys)
    assert found.all()
    for i, d in enumerate(docs):
        assert buf[off[i]:off[i + 1]].tobytes() == d, f"record {i}"
    c.free_prop()


def test_decode_request_split_into_passes(ctrl_mod):
    """a request whose touched chunks decode to more than the 32-bit arena of one pass is split by chunk; the knob
    shrinks the limit so that a small store needs many passes (shuffled request, device and host outputs)"""
    kd, ko, vd, vo = synth.gen_html_pages(60, seed=4, max_len=20000, mean_len=9000)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_BYTES, window_bytes=60_000)
    c.setitem_batch((kd, ko), (vd, vo))
    assert c.stats().chunks >= 8
    c.debug_set_knob("dec_arena_limit", 150000)
    order = [int(i) for i in np.random.default_rng(3).permutation(len(keys))]
    for req in (list(range(len(keys))), order, order[:17]):
        buf, off, found = c.getitem_batch([keys[i] for i in req])
        assert found.all()
        for j, i in enumerate(req):
            assert buf[off[j]:off[j + 1]].tobytes() == po.make_doc(keys[i], vals[i]), (len(req), i)
    c.free_prop()


def test_decode_zero_bytes_travel_through_the_bitmap(ctrl_mod):
    """the decoder recognises a final byte by being nonzero; zero-valued bytes are announced in a bitmap instead.
    Values that are mostly zeros: long zero runs (period-1 self references whose source is a zero), zero runs copied
    from record to record, and a swept single-byte mutation over zero-filled records (chains of zero bytes that can
    only resolve through the bitmap)"""
    rng = np.random.default_rng(21)
    keys, vals = [], []
    base = np.zeros(1500, dtype=np.uint8)
    base[rng.integers(0, 1500, size=40)] = rng.integers(1, 250, size=40).astype(np.uint8)
    cur = base.copy()
    for r in range(900):
        cur = cur.copy()
        cur[(13 * r) % 1500] = 0 if r % 3 else int(rng.integers(1, 250))
        keys.append(b"z%05d" % r)
        vals.append(cur.tobytes())
    keys += [b"allzero", b"zero-then-text", bytes([0, 0, 0]) + b"key-with-zeros"]
    vals += [bytes(40000), bytes(3000) + b"tail" * 50 + bytes(3000), bytes(5000)]
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS)
    assert not c.setitem_batch(keys, vals)[0].any()
    assert c.stats().encoded_bytes < 0.2 * c.stats().doc_bytes
    buf, off, found = c.getitem_batch(keys)
    assert found.all()
    for i, (k, v) in enumerate(zip(keys, vals)):
        assert buf[off[i]:off[i + 1]].tobytes() == po.make_doc(k, v), i
    # a partial request (the arena then holds records that were not asked for) in shuffled order
    pick = [int(x) for x in rng.permutation(len(keys))[:200]]
    buf, off, found = c.getitem_batch([keys[i] for i in pick])
    for j, i in enumerate(pick):
        assert buf[off[j]:off[j + 1]].tobytes() == po.make_doc(keys[i], vals[i]), i
    c.free_prop()


# --------------------------------------------------------------------------- decode
@pytest.mark.parametrize("name", GOLDEN)
def test_import_and_decode_chunk(ctrl_mod, name):
    g = load(name)
    docs, encs = _oracle_encode_all(g["keys"], g["vals"], strict=False)
    c = ctrl_mod.PiXiuCtrl()
    chunk = c.import_chunk(encs)
    assert chunk == 0
    buf, off = c.decode_chunk(0)
    for i, d in enumerate(docs):
        assert buf[off[i]:off[i + 1]].tobytes() == d, f"{name}: record {i}"
    # the index was rebuilt from the decoded keys: latest value wins
    latest = dict(zip(g["keys"], g["vals"]))
    ks = list(latest)[:100]
    b2, o2, found = c.getitem_batch(ks)
    assert found.all()
    for i, k in enumerate(ks):
        assert ctrl_mod.split_doc(b2[o2[i]:o2[i + 1]].tobytes()) == (k, latest[k])
    c.free_prop()


# --------------------------------------------------------------------------- CRUD (t_PiXiuCtrl, PiXiuCtrl.cpp:176-255)
def test_crud_differential(ctrl_mod):
    rng = random.Random(11)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_BYTES, window_bytes=4000)  # forces many rotations
    st = po.OracleStore()
    model = {}
    for rnd in range(30):
        ks = [bytes(rng.choice(b"ABCDE") for _ in range(rng.randint(1, 6))) for _ in range(60)]
        vs = [bytes(rng.choice(b"ABCDE") for _ in range(rng.randint(1, 50))) for _ in range(60)]
        rc, _ = c.setitem_batch(ks, vs)
        want = []
        for k, v in zip(ks, vs):
            want.append(int(k in model))
            model[k] = v
            st.setitem(k, v)
        assert rc.tolist() == want
        dk = [bytes(rng.choice(b"ABCDE") for _ in range(rng.randint(1, 4))) for _ in range(15)]
        want = []
        for k in dk:
            want.append(int(k not in model))
            model.pop(k, None)
            st.delitem(k)
        assert c.delitem_batch(dk).tolist() == want
        qk = [bytes(rng.choice(b"ABCDEF") for _ in range(rng.randint(1, 6))) for _ in range(80)]
        assert c.contains_batch(qk).tolist() == [k in model for k in qk]
        buf, off, found = c.getitem_batch(qk)
        assert found.tolist() == [k in model for k in qk]
        for i, k in enumerate(qk):
            if k in model:
                assert ctrl_mod.split_doc(buf[off[i]:off[i + 1]].tobytes()) == (k, model[k])
    assert c.stats().chunks > 3
    assert c.stats().live_records == len(model)
    for prefix in [b"", b"A", b"AB", b"E", b"ABCDEA", b"Z"]:
        want = sorted(po.make_doc(k, v) for k, v in model.items() if k.startswith(prefix))
        assert c.iter_docs(prefix) == want
        assert st.iter(prefix) == want
    # single-record API (the reference's call shapes)
    assert c.setitem(b"solo", b"value") == 0
    assert c.setitem(b"solo", b"value2") == 1
    gen = c.getitem(b"solo")
    assert gen.bytes() == po.make_doc(b"solo", b"value2")
    assert c.getitem(b"missing") is None
    assert c.delitem(b"solo") == 0 and c.delitem(b"solo") == 1
    c.free_prop()


# --------------------------------------------------------------------------- window rotation (PiXiuCtrl.cpp:13-17)
@pytest.mark.parametrize("name", GOLDEN)
def test_arena_accounting_matches_reference_fixture(ctrl_mod, name):
    """MemPool::nth / used_num predicted from the suffix array == what the reference's arena held"""
    g = load(name)
    keys, vals = g["keys"], g["vals"]
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_REFERENCE, strict251=True)
    rng = random.Random(3)
    a = 0
    while a < len(keys):
        b = min(len(keys), a + rng.randint(1, 40))
        c.setitem_batch(keys[a:b], vals[a:b])
        assert c.debug_pool_state() == (int(g["pools"][b - 1]), int(g["pool_used"][b - 1])), f"{name}: after record {b - 1}"
        a = b
    for i in range(len(keys)):
        assert c.encoded(0, i) == g["enc"][i]
    c.free_prop()


def test_reference_rotation_live(ctrl_mod, ref):
    """More than one full reference window of pages: same rotation record, same bytes on both sides."""
    kd, ko, vd, vo = synth.gen_html_pages(520, seed=2)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    ref.reset()
    r = ref.setitem_batch(keys, vals)
    chunk_of = r["chunk"]
    assert chunk_of[-1] >= 1, "the sample must make the reference rotate at least once"
    first_rot = int(np.argmax(chunk_of > 0))
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_REFERENCE, strict251=True)
    _, saved = c.setitem_batch((kd, ko), (vd, vo))
    locs = [c.record_location(i) for i in range(len(keys))]
    assert [l[0] for l in locs] == chunk_of.tolist(), f"rotation differs (reference rotates at record {first_rot})"
    assert [l[1] for l in locs] == r["idx"].tolist()
    enc_len = np.array([len(po.make_doc(k, v)) for k, v in zip(keys, vals)]) - saved
    assert enc_len.tolist() == r["enc_len"].tolist()
    # bytes of the records of the open (last) reference chunk
    last = int(chunk_of[-1])
    for i in range(len(keys)):
        if chunk_of[i] == last:
            assert c.encoded(last, int(r["idx"][i])) == ref.encoded(int(r["idx"][i]))
    c.free_prop()


def test_cpp_facade_reference_call_shapes(ctrl_mod):
    """tests/cpp/facade_test.cpp: the reference's t_PiXiuCtrl scenario against include/PiXiuCtrl.hpp"""
    import os
    import subprocess

    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "facade_test")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "facade_test ok" in r.stdout


# --------------------------------------------------------------------------- multi-GPU extended window (emulated ranks)
def _mg_roundtrip_batch(ctrl_mod, stores, keys, vals):
    """drives pixiu_mg_setitem_{begin,mid,end} on every emulated rank; the host plays the two all-reduces"""
    L = ctrl_mod.load_library()

    def d2h(p, n):
        a = np.zeros(max(n, 1), dtype=np.uint32)
        assert L.pixiu_debug_memcpy(a.ctypes.data_as(C.c_void_p), C.c_void_p(p), 4 * n, 1) == 0
        return a[:n]

    def h2d(p, a):
        assert L.pixiu_debug_memcpy(C.c_void_p(p), a.ctypes.data_as(C.c_void_p), 4 * len(a), 2) == 0

    ptrs = [s.mg_setitem_begin(keys, vals) for s in stores]
    assert len({c for _, c in ptrs}) == 1
    red = np.maximum.reduce([d2h(p, c) for p, c in ptrs])          # collective 1: MAX of M(s)
    for p, c in ptrs:
        h2d(p, red)
    ptrs = [s.mg_setitem_mid() for s in stores]
    assert len({c for _, c in ptrs}) == 1
    if ptrs[0][1]:
        red = np.minimum.reduce([d2h(p, c) for p, c in ptrs])      # collective 2: MIN of (idx << 16 | to)
        assert (red != 0xFFFFFFFF).all()
        for p, c in ptrs:
            h2d(p, red)
    return [s.mg_setitem_end() for s in stores]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["fuzz_mode1", "c1_urls", "c3_nested", "c2_pages"])
def test_sharded_window_equals_unsharded_oracle(ctrl_mod, name, world):
    """window sharded over `world` ranks + MAX/MIN reduce == one window holding everything (oracle)"""
    g = load(name)
    keys, vals = g["keys"], g["vals"]
    docs, encs = _oracle_encode_all(keys, vals, strict=False)  # default mode: every record decodes (no bug B1)
    stores = [ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS) for _ in range(world)]
    for r, s in enumerate(stores):
        s.mg_config(r, world)
    rng = random.Random(5)
    a, rcs = 0, []
    while a < len(keys):
        b = min(len(keys), a + rng.randint(1, max(2, len(keys) // 4)))
        outs = _mg_roundtrip_batch(ctrl_mod, stores, keys[a:b], vals[a:b])
        for rc, saved in outs:
            assert rc.tolist() == outs[0][0].tolist() and saved.tolist() == outs[0][1].tolist()
        rcs.append(outs[0][0])
        a = b
    assert np.concatenate(rcs).tolist() == g["rc"].tolist()
    for s in stores:
        for i in range(len(keys)):
            assert s.encoded(0, i) == encs[i], f"{name}: record {i}"
        buf, off, found = s.getitem_batch(keys[:50])
        assert found.all()
        latest = dict(zip(keys, vals))
        for i, k in enumerate(keys[:50]):
            assert ctrl_mod.split_doc(buf[off[i]:off[i + 1]].tobytes()) == (k, latest[k])
    for s in stores:
        s.free_prop()


def test_export_import_roundtrip(ctrl_mod):
    """wire format: chunks exported from one store load into another and serve the same records"""
    kd, ko, vd, vo = synth.gen_urls_kv(3000, seed=9)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    a = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_BYTES, window_bytes=120000)
    a.setitem_batch((kd, ko), (vd, vo))
    nchunks = a.stats().chunks
    assert nchunks >= 3
    b = ctrl_mod.PiXiuCtrl()
    for c in range(nchunks):
        enc, off = a.export_chunk(c)
        assert b.import_chunk((enc, off)) == c
    assert b.stats().records == a.stats().records and b.stats().encoded_bytes == a.stats().encoded_bytes
    buf, off, found = b.getitem_batch((kd, ko))
    assert found.all()
    for i in range(0, len(keys), 7):
        assert ctrl_mod.split_doc(buf[off[i]:off[i + 1]].tobytes()) == (keys[i], vals[i])
    assert not b.contains(b"http://absent")
    a.free_prop()
    b.free_prop()


def test_export_import_with_index_blob(ctrl_mod):
    """the whole store in its wire format: chunks + the serialised CritBit SoA; the copy serves lookups, getitem, iter
    and deletes without rebuilding anything (no decode on import, no per-key insert)"""
    kd, ko, vd, vo = synth.gen_urls_kv(4000, seed=12)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    a = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_BYTES, window_bytes=150000)
    a.setitem_batch((kd, ko), (vd, vo))
    a.setitem_batch(keys[:300], vals[300:600])          # replaced records: tombstones in the exported chunks
    dead = keys[1000:1100]
    assert not a.delitem_batch(dead).any()
    b = ctrl_mod.PiXiuCtrl()
    for c in range(a.stats().chunks):
        b.import_chunk_raw(a.export_chunk(c))
    assert not b.contains(keys[5])                       # nothing is indexed yet
    b.import_index(a.export_index())
    sa, sb = a.stats(), b.stats()
    assert (sb.records, sb.live_records, sb.encoded_bytes) == (sa.records, sa.live_records, sa.encoded_bytes)
    latest = dict(zip(keys, vals))
    latest.update(zip(keys[:300], vals[300:600]))
    buf, off, found = b.getitem_batch((kd, ko))
    for i, k in enumerate(keys):
        assert bool(found[i]) == (k not in dead), i
        if found[i]:
            assert ctrl_mod.split_doc(buf[off[i]:off[i + 1]].tobytes()) == (k, latest[k])
    assert [ctrl_mod.split_doc(d)[0] for d in b.iter_docs(b"http://news")] == sorted(k for k in latest if k.startswith(b"http://news") and k not in dead)
    assert b.delitem(keys[7]) == 0 and not b.contains(keys[7]) and b.setitem(b"new-key", b"v") == 0 and b.contains(b"new-key")
    with pytest.raises(ctrl_mod.PiXiuError):             # a corrupt blob is refused
        c = ctrl_mod.PiXiuCtrl()
        c.import_chunk_raw(a.export_chunk(0))
        blob = a.export_index().copy()
        blob[40:48] = 255
        c.import_index(blob)
    a.free_prop()
    b.free_prop()


def test_api_edge_cases(ctrl_mod):
    """empty store / empty batch, duplicates inside one batch (in-order semantics), key-only records,
    keys that are prefixes of each other, binary keys full of 251 / 0 / 2, deleted keys"""
    c = ctrl_mod.PiXiuCtrl()
    # empty store (CritBitTree.cpp:154-196,:271-274: NULL / false / not found)
    assert not c.contains(b"x") and c.getitem(b"x") is None and c.iter(b"") is None and c.delitem(b"x") == 1
    rc, saved = c.setitem_batch([], [])
    assert len(rc) == 0 and c.stats().records == 0
    assert c.contains_batch([]).tolist() == [] and c.delitem_batch([]).tolist() == []
    # duplicates inside one batch: later wins, rc tells which calls replaced
    rc, _ = c.setitem_batch([b"k", b"k", b"j", b"k"], [b"v1", b"v2", b"w", b"v3"])
    assert rc.tolist() == [0, 1, 0, 1]
    assert ctrl_mod.split_doc(c.getitem(b"k").bytes()) == (b"k", b"v3")
    assert c.stats().records == 4 and c.stats().live_records == 2
    # key-only records (PiXiuCtrl.cpp:41-44) and keys that are prefixes of each other
    rc, _ = c.setitem_batch([b"a", b"ab", b"abc", b"b"], [b"", b"1", b"", b"2"])
    assert rc.tolist() == [0, 0, 0, 0]
    assert c.getitem(b"a").bytes() == b"a\xfb\x00" and c.getitem(b"ab").bytes() == b"ab\xfb\x001\xfb\x02"
    # iteration order is byte order on esc(key) 251 0: "abc" < "ab" < "a" (251 sorts above 'b' and 'c')
    assert [ctrl_mod.split_doc(d)[0] for d in c.iter_docs(b"a")] == [b"abc", b"ab", b"a"]
    assert [ctrl_mod.split_doc(d)[0] for d in c.iter_docs(b"")] == [b"abc", b"ab", b"a", b"b", b"j", b"k"]
    assert c.iter_docs(b"zz") == [] and c.iter_docs(b"abcd") == []
    # binary keys/values made of the special bytes, incl. the case the reference loses (bug B5)
    bk = [bytes([251]), bytes([251, 251]), bytes([251, 0]), bytes([0]), bytes([2, 251, 0, 251]), b"a\xfb", b"a\xfb\xfb"]
    bv = [bytes([0, 2, 251]) * 5, bytes([251]) * 9, b"", bytes([251, 0, 251, 2]), bytes([2]), b"x", bytes([251])]
    rc, _ = c.setitem_batch(bk, bv)
    assert rc.tolist() == [0] * len(bk)
    buf, off, found = c.getitem_batch(bk)
    assert found.all()
    for i in range(len(bk)):
        assert ctrl_mod.split_doc(buf[off[i]:off[i + 1]].tobytes()) == (bk[i], bv[i])
        assert buf[off[i]:off[i + 1]].tobytes() == po.make_doc(bk[i], bv[i])
    assert c.contains_batch([bytes([251, 251, 251]), bytes([251, 2]), b"a"]).tolist() == [False, False, True]
    # delete, then the key is gone but the record bytes stay referencable
    assert c.delitem_batch([b"ab", b"ab", bytes([251])]).tolist() == [0, 1, 0]
    assert c.getitem(b"ab") is None and not c.contains(bytes([251]))
    rc, _ = c.setitem_batch([b"ab"], [b"1"])  # re-insert after delete: a new key again
    assert rc.tolist() == [0]
    assert c.stats().live_records == len(c.iter_docs(b""))
    # capacity hint: maps room in the compressed arena, changes nothing that can be observed
    before = [c.getitem(k).bytes() for k in (b"k", b"a", b"ab")]
    c.reserve(300 << 20)
    c.reserve(0)
    with pytest.raises(ctrl_mod.PiXiuError):
        c.reserve(-1)
    rc, _ = c.setitem_batch([b"after-reserve"], [b"v" * 100])
    assert rc.tolist() == [0] and [c.getitem(k).bytes() for k in (b"k", b"a", b"ab")] == before
    assert ctrl_mod.split_doc(c.getitem(b"after-reserve").bytes()) == (b"after-reserve", b"v" * 100)
    c.free_prop()


def test_config1_full_parity_with_live_reference(ctrl_mod, ref):
    """BASELINE config 1 in full: 10k URL keys x ~100 B values.  SET on the reference (as is, on the CPU) and
    on the GPU path: same rc, same encoded bytes for every record; GET of every key: same decoded stream
    (records this short never hit the reference decoder's bugs)."""
    kd, ko, vd, vo = synth.gen_urls_kv(10000, seed=1)
    keys, vals = synth.unpack(kd, ko), synth.unpack(vd, vo)
    ref.reset()
    r = ref.setitem_batch(keys, vals)
    assert r["chunk"][-1] == 0
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_REFERENCE, strict251=True)
    rc, saved = c.setitem_batch((kd, ko), (vd, vo))
    assert rc.tolist() == r["rc"].tolist()
    assert c.debug_pool_state() == (int(r["pools"][-1]), int(r["pool_used"][-1]))
    enc, off = c.export_chunk(0)
    assert np.diff(off).tolist() == r["enc_len"].tolist()
    for i in range(0, len(keys), 37):
        assert enc[off[i]:off[i + 1]].tobytes() == ref.encoded(i), f"record {i}"
    buf, doff, found = c.getitem_batch((kd, ko))
    assert found.all()
    for i in range(0, len(keys), 53):
        assert buf[doff[i]:doff[i + 1]].tobytes() == ref.getitem(keys[i]) == po.make_doc(keys[i], vals[i])
    absent = [b"http://nope.qq.com/%d" % i for i in range(100)]
    assert not c.contains_batch(absent).any() and not any(ref.contains(k) for k in absent)
    c.free_prop()


# --------------------------------------------------------------------------- size-independent properties at scale
def _roundtrip_all(ctrl_mod, c, kd, ko, vd, vo):
    """getitem of every key == esc(k) 251 0 esc(v) 251 2, checked vectorised for escape-free data"""
    n = len(ko) - 1
    buf, off, found = c.getitem_batch((kd, ko))
    assert found.all()
    klen, vlen = np.diff(ko), np.diff(vo)
    assert np.array_equal(np.diff(off), klen + vlen + 4)
    # key bytes, terminators and value bytes at their places
    from pixiu_b200.synth import ragged_gather

    assert np.array_equal(ragged_gather(buf, off[:-1], klen), kd)
    assert np.array_equal(ragged_gather(buf, off[:-1] + klen + 2, vlen), vd)
    t = off[:-1] + klen
    assert (buf[t] == 251).all() and (buf[t + 1] == 0).all()
    assert (buf[off[1:] - 2] == 251).all() and (buf[off[1:] - 1] == 2).all()


def test_config3_style_deep_nesting_at_scale(ctrl_mod):
    """100k x 1 KB records, each the previous one with one byte swept (nesting depth in the hundreds) plus
    self-periodic runs: every record round-trips; the compressed store is a few percent of the input"""
    kd, ko, vd, vo = synth.gen_nested(100000, seed=3)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_REFERENCE)
    rc, saved = c.setitem_batch((kd, ko), (vd, vo))
    assert not rc.any()
    st = c.stats()
    assert st.records == 100000 and st.chunks >= 2
    assert st.encoded_bytes < 0.06 * st.raw_bytes
    assert int(saved.sum()) == st.doc_bytes - st.encoded_bytes
    _roundtrip_all(ctrl_mod, c, kd, ko, vd, vo)
    c.free_prop()


def test_config4_style_lookup_at_scale(ctrl_mod):
    """300k URL keys: 90 % present / 10 % absent queries in random order, then full round trip"""
    n = 300000
    kd, ko, vd, vo = synth.gen_urls_kv(n, seed=4, val_words=20)
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_REFERENCE)
    rc, _ = c.setitem_batch((kd, ko), (vd, vo))
    assert not rc.any() and c.stats().chunks >= 3
    rng = np.random.default_rng(1)
    keys = synth.unpack(kd, ko)
    pick = rng.integers(0, n, size=90000)
    q = [keys[i] for i in pick] + [b"http://absent.qq.com/%d" % i for i in range(10000)]
    perm = rng.permutation(len(q))
    found = c.contains_batch([q[i] for i in perm])
    assert found.tolist() == (perm < 90000).tolist()
    _roundtrip_all(ctrl_mod, c, kd, ko, vd, vo)
    # delete a third, the rest stays reachable
    dk = keys[::3]
    assert not c.delitem_batch(dk).any()
    f2 = c.contains_batch(keys[:3000])
    assert f2.tolist() == [i % 3 != 0 for i in range(3000)]
    c.free_prop()


@pytest.mark.parametrize("bulk_min", [1 << 30, 1])
def test_batched_index_insert_with_conflicts(ctrl_mod, bulk_min):
    """the batched CritBit insert against a dict - GPU probes + host splices (bulk_min huge), and the bulk build of the
    whole tree on the GPU from the sorted keys (bulk_min = 1: every batch at least as large as the tree rebuilds it):
    monotone ids (most keys of a sub-batch meet on the same edges), keys that replace stored ones, duplicates inside
    one batch, deletes followed by re-inserts (slot reuse), binary keys with 251s, and prefix iteration order afterwards"""
    rng = np.random.default_rng(12)
    model = {}

    def put(keys, tag):
        vals = [b"v%s_%d" % (tag, i) for i in range(len(keys))]
        rc, _ = c.setitem_batch(keys, vals)
        want = []
        for k, v in zip(keys, vals):
            want.append(1 if k in model else 0)
            model[k] = v
        assert rc.tolist() == want, tag

    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_BYTES, window_bytes=4 << 20)
    c.debug_set_knob("bulk_min", bulk_min)
    base = [b"http://h%d.example.com/a/%08d.htm" % (i % 7, 1000 + 3 * i) for i in range(20000)]   # monotone ids
    put(base[:5], b"0")                                  # (a tiny tree first: the next batch is the larger one)
    put(base, b"a")
    mixed = ([b"http://h%d.example.com/a/%08d.htm" % (i % 7, 1000 + 3 * i + 1) for i in range(9000)]   # neighbours of stored keys
             + [base[i] for i in rng.integers(0, len(base), 4000)]                                     # replace stored keys
             + [bytes([251, 250, i % 256, 251, (i >> 8) % 256]) + b"bin" for i in range(3000)])       # 251-heavy keys
    mixed += [mixed[i] for i in rng.integers(0, len(mixed), 3000)]                                       # duplicates in the batch
    order = rng.permutation(len(mixed))
    put([mixed[i] for i in order], b"b")
    # deletes, then re-inserts and fresh keys in one batch (leaf / inner slots are reused)
    gone = [base[i] for i in range(0, len(base), 5)]
    assert not c.delitem_batch(gone).any()
    for k in gone:
        del model[k]
    put(gone[::2] + [b"http://h9.example.com/z/%07d" % i for i in range(6000)], b"c")
    big = [b"http://h%d.example.com/b/%08d.htm" % (i % 5, 7 * i) for i in range(60000)] + base[::3]   # larger than the tree again
    put(big, b"d")
    keys = list(model)
    probe = keys + gone[1::2]
    found = c.contains_batch(probe)
    assert found.tolist() == [k in model for k in probe]
    sample = [keys[i] for i in rng.integers(0, len(keys), 3000)]
    buf, off, f = c.getitem_batch(sample)
    assert f.all()
    for i, k in enumerate(sample):
        assert buf[off[i]:off[i + 1]].tobytes() == po.make_doc(k, model[k])
    got = [ctrl_mod.split_doc(d)[0] for d in c.iter_docs(b"http://h3.")]
    esc = lambda k: k.replace(b"\xfb", b"\xfb\xfb") + b"\xfb\x00"
    want = sorted((k for k in model if k.startswith(b"http://h3.")), key=esc)
    assert got == want
    c.free_prop()


# --------------------------------------------------------------------------- compaction (PiXiuCtrl::reinsert)
def test_reinsert_chunk_moves_live_records(ctrl_mod):
    """explicit pixiu_reinsert_chunk on a partially deleted, non-full chunk (the reference crashes there, bug B4):
    every live key keeps its document, the old chunk ends with no live record, moved records sit in the open chunk"""
    rng = np.random.default_rng(5)
    keys = [b"http://c%d.example.org/p/%06d" % (i % 5, i) for i in range(6000)]
    vals = [bytes(rng.integers(97, 123, size=int(rng.integers(20, 400))).astype(np.uint8)) + b"<div class=x>" * int(rng.integers(0, 6))
            for _ in keys]
    vals[7] = b""                                  # key-only record
    vals[11] = bytes([251, 0, 251, 2, 251, 251]) * 9    # escapes in the value
    keys[13] = bytes([251, 251, 0, 7]) + keys[13]
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_BYTES, window_bytes=400_000)
    assert not c.setitem_batch(keys, vals)[0].any()
    st0 = c.stats()
    assert st0.chunks >= 3
    total0, live0, dropped0 = c.chunk_info(0)
    assert live0 == total0 and not dropped0
    first_chunk = [i for i in range(len(keys)) if c.record_location(i)[0] == 0]
    gone = set(first_chunk[::3] + first_chunk[1::3])            # two thirds of chunk 0
    assert not c.delitem_batch([keys[i] for i in sorted(gone)]).any()
    assert c.chunk_info(0)[1] == total0 - len(gone)
    with pytest.raises(ctrl_mod.PiXiuError):
        c.reinsert(st0.chunks - 1)                              # never the open chunk (PiXiuCtrl.cpp:26)
    moved = c.reinsert(0)
    assert moved == total0 - len(gone)
    assert c.chunk_info(0) == (total0, 0, True)
    assert c.reinsert(0) == 0                                   # idempotent
    st1 = c.stats()
    assert st1.live_records == len(keys) - len(gone) and st1.records == st0.records + moved
    assert st1.reinserted_records == moved and st1.reclaimable_bytes > 0
    buf, off, found = c.getitem_batch(keys)
    assert found.tolist() == [i not in gone for i in range(len(keys))]
    for i in range(len(keys)):
        if i not in gone:
            assert buf[off[i]:off[i + 1]].tobytes() == po.make_doc(keys[i], vals[i]), i
    # iteration order and content unchanged
    got = [ctrl_mod.split_doc(d)[0] for d in c.iter_docs(b"http://c3.")]
    esc = lambda k: k.replace(b"\xfb", b"\xfb\xfb") + b"\xfb\x00"
    assert got == sorted((keys[i] for i in range(len(keys)) if i not in gone and keys[i].startswith(b"http://c3.")), key=esc)
    c.free_prop()


def test_auto_reinsert_matches_live_reference(ctrl_mod, ref):
    """the reference's own trigger on a FULL chunk (65,535 records; its reinsert only works there): after more than
    half of chunk 0 is deleted, the next call re-inserts the survivors.  Same return codes, same survivors, and the
    re-inserted records are byte-identical to the reference's (same order, same window)."""
    n0 = 65535 + 40
    keys = [b"k%07d" % i for i in range(n0)]
    vals = [b"v%05d-%s" % (i % 977, b"abcdefghij"[: 1 + i % 9]) for i in range(n0)]
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_REFERENCE, strict251=True, auto_reinsert=True)
    rc, _ = c.setitem_batch(keys, vals)
    ref.reset()
    r = ref.setitem_batch(keys, vals)
    assert rc.tolist() == r["rc"].tolist()
    assert c.stats().chunks == 2 and c.chunk_info(0)[0] == 65535
    kill = [keys[i] for i in range(0, 65535, 2)] + [keys[i] for i in range(1, 2000, 2)]     # 33,768 of chunk 0
    # (a delitem batch looks at the trigger before every delete, like n single calls)
    assert not c.delitem_batch(kill).any()
    assert all(ref.delitem(k) == 0 for k in kill)
    survivors = sorted(set(keys) - set(kill))
    total, live, dropped = c.chunk_info(0)
    assert dropped and live == 0, (total, live, dropped)
    extra_k, extra_v = b"k-after", b"v-after"
    assert c.setitem(extra_k, extra_v) == ref.setitem(extra_k, extra_v) == 0
    info = ref.last_info()
    st = c.stats()
    ch, idx = c.record_location(st.records - 1)
    assert idx == info["idx"], (idx, info)
    for i in range(0, idx + 1, 97):
        assert c.encoded(ch, i) == ref.encoded(i), i
    assert c.encoded(ch, idx) == ref.encoded(idx)
    found = c.contains_batch(keys)
    alive = set(survivors)
    assert found.tolist() == [k in alive for k in keys]
    sample = survivors[::211]
    buf, off, f = c.getitem_batch(sample)
    assert f.all()
    for i, k in enumerate(sample):
        assert buf[off[i]:off[i + 1]].tobytes() == ref.getitem(k) == po.make_doc(k, vals[keys.index(k)])
    c.free_prop()


def test_contains_with_device_resident_queries(ctrl_mod):
    """pixiu_contains_batch_dev (queries and found[] in HBM) answers like the host-buffer entry point, including
    absent keys, keys with escapes, tombstoned keys and an empty store"""
    import torch

    def ask(c, keys):
        kd, ko = synth.pack(keys)
        d_k = torch.from_numpy(np.concatenate([kd, np.zeros(8, dtype=np.uint8)])).cuda()
        d_o = torch.from_numpy(ko).cuda()
        d_f = torch.full((len(keys),), 7, dtype=torch.uint8, device="cuda")
        c.contains_batch_dev(d_k.data_ptr(), d_o.data_ptr(), len(keys), d_f.data_ptr())
        return d_f.cpu().numpy().astype(bool)

    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_BYTES, window_bytes=200_000)
    probe = [b"a", bytes([251, 0, 251]), b"http://x/%d" % 5]
    assert not ask(c, probe).any()                                  # empty store
    keys = [b"http://x/%d" % i for i in range(5000)] + [bytes([251, i % 256, 251, 251, 0]) + b"k%d" % i for i in range(300)]
    vals = [b"v%d" % i * (1 + i % 7) for i in range(len(keys))]
    c.setitem_batch(keys, vals)
    c.delitem_batch(keys[::5])
    q = keys[::3] + [b"http://x/%d" % i for i in range(5000, 5400)] + [bytes([251, 7, 251, 251, 0]) + b"k7x"]
    got = ask(c, q)
    want = c.contains_batch(q)
    alive = set(keys) - set(keys[::5])
    assert got.tolist() == want.tolist() == [k in alive for k in q]
    c.free_prop()


# --------------------------------------------------------------------------- full-corpus pin (BASELINE config[1])
def test_c2_full_corpus_matches_reference_fixture(ctrl_mod):
    """All 10,000 pages of the headline workload (bench.py's corpus) through setitem: every record lands in the same
    window (chunk) at the same idx as in the UNMODIFIED reference, with the same encoded length and the same bytes
    (CRC32) - i.e. the compression ratio README.md:53 quotes is identical, and so is every rotation point
    (PiXiuCtrl.cpp:13-17).  Fixture: tests/golden/c2_full_enc_len.npz (tests/golden/make_c2_full.py, ~8 min of the
    reference on one core)."""
    import os
    import zlib

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_full_enc_len.npz"))
    kd, ko, vd, vo = synth.gen_html_pages(int(g["pages"]), seed=int(g["seed"]))
    n = len(ko) - 1
    assert int(ko[-1] + vo[-1]) == int(g["raw_bytes"])
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_REFERENCE, strict251=True)
    rc, saved = c.setitem_batch((kd, ko), (vd, vo))
    assert not rc.any()
    st = c.stats()
    assert st.chunks == int(g["chunk"][-1]) + 1
    enc_len = (np.diff(ko) + np.diff(vo) + 4 - saved).astype(np.int64)   # escape-free corpus: doc = k + v + 4
    assert np.array_equal(enc_len, g["enc_len"].astype(np.int64))
    assert st.encoded_bytes == int(g["enc_len"].astype(np.int64).sum())
    r = 0
    for ch in range(st.chunks):
        enc, off = c.export_chunk(ch)
        cnt = len(off) - 1
        assert np.array_equal(g["chunk"][r:r + cnt], np.full(cnt, ch)) and np.array_equal(g["idx"][r:r + cnt], np.arange(cnt))
        b = enc.tobytes()
        crc = np.array([zlib.crc32(b[off[i]:off[i + 1]]) for i in range(cnt)], dtype=np.uint32)
        assert np.array_equal(crc, g["crc32"][r:r + cnt]), f"window {ch}: encoded bytes differ from the reference"
        r += cnt
    assert r == n
    c.free_prop()


# --------------------------------------------------------------------------- guards
def test_mg_phase_guards_and_shard_is_not_a_plain_store(ctrl_mod):
    """while a multi-GPU batch is between its phases every other entry point is refused (they would reuse its
    scratch); a shard refuses plain setitem / import / rotate / reinsert altogether"""
    c = ctrl_mod.PiXiuCtrl(rotate_policy=ctrl_mod.ROTATE_RECORDS)
    c.mg_config(0, 1)
    keys, vals = [b"alpha", b"beta"], [b"x" * 40, b"y" * 40]
    c.mg_setitem_begin(keys, vals)
    for call in (lambda: c.contains(b"alpha"), lambda: c.getitem(b"alpha"), lambda: c.delitem(b"alpha"),
                 lambda: c.setitem(b"k", b"v"), lambda: c.rotate(), lambda: c.iter_docs(b"")):
        with pytest.raises(ctrl_mod.PiXiuError) as e:
            call()
        assert e.value.code == ctrl_mod.EINVAL
    c.mg_setitem_mid()
    rc, _ = c.mg_setitem_end()
    assert not rc.any()
    assert c.contains(b"alpha") and c.delitem(b"beta") == 0          # reads and deletes are fine between batches
    for call in (lambda: c.setitem(b"k", b"v"), lambda: c.rotate(), lambda: c.reinsert(0)):
        with pytest.raises(ctrl_mod.PiXiuError):
            call()
    c.free_prop()
