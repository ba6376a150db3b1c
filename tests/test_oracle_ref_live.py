"""Live differential tests: CPU oracle vs the real reference (needs oracle/_ref)."""
import random

from oracle import pyoracle as po
from make_golden import fuzz_records


def test_encoder_fuzz_vs_reference(ref):
    for trial in range(12):
        ref.reset()
        w = po.OracleWindow(strict251=True)
        keys, vals = fuzz_records(1000 + trial, 60, trial % 3)
        for k, v in zip(keys, vals):
            ref.setitem(k, v)
            assert ref.last_encoded() == w.encode(po.make_doc(k, v))


def test_crud_vs_reference(ref):
    """mirrors t_PiXiuCtrl's CRUD differential (proj/PiXiuCtrl.cpp:176-226) at small scale"""
    ref.reset()
    st = po.OracleStore(strict251=True)
    rng = random.Random(7)
    model = {}
    for it in range(4000):
        k = bytes(rng.choice(b"ABCDE") for _ in range(rng.randint(1, 6)))
        op = rng.random()
        if op < 0.55:
            v = bytes(rng.choice(b"ABCDE") for _ in range(rng.randint(1, 50)))
            assert ref.setitem(k, v) == st.setitem(k, v) == int(k in model)
            model[k] = v
        elif op < 0.75:
            assert ref.delitem(k) == st.delitem(k) == int(k not in model)
            model.pop(k, None)
        else:
            assert ref.contains(k) == st.contains(k) == (k in model)
            got = st.getitem(k)
            if k in model:
                assert po.split_doc(got) == (k, model[k])
                assert ref.getitem(k) == got  # short records: reference decoder is sound here
            else:
                assert got is None and ref.getitem(k) is None
    for prefix in [b"", b"A", b"AB", b"E", b"ABCDEA", b"Z"]:
        want = sorted(po.make_doc(k, v) for k, v in model.items() if k.startswith(prefix))
        assert st.iter(prefix) == want
        assert ref.iter(prefix) == want
