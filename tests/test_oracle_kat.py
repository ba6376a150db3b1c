"""Pins the CPU oracle to the reference's own known-answer vectors.

Vectors are transcribed from the reference's self tests (file:line relative to
/root/reference/src) and README; when oracle/_ref is built the same vectors are
also replayed through the real reference.
"""
import numpy as np
import pytest

from oracle import pyoracle as po

PASS = -3
U = 251


def _msgs_to_arrays(msgs):
    cmd = [m[0] for m in msgs]
    pos = [m[1] for m in msgs]
    val = bytes(m[2] for m in msgs)
    return cmd, pos, val


# proj/PiXiuStr.cpp:302-311 — escape KAT
def test_escape_kat():
    src = bytes([1, U, 2, U, 4])
    assert po.escape(src, True) == bytes([1, U, U, 2, U, U, 4, U, 0])
    assert po.escape(src, False) == bytes([1, U, U, 2, U, U, 4])
    # :313-321 concat of two escaped [1,251]
    assert po.escape(src[:2], False) * 2 == bytes([1, U, U, 1, U, U])


# proj/PiXiuStr.cpp:327-339 — a 6-run stays literal, a 7-run becomes a small record
KAT_6_7 = ([(PASS, 0, 1)] * 4 + [(2, i, 3) for i in range(6)] + [(PASS, 0, 1)] + [(3, i, 4) for i in range(7)],
           bytes([1, 1, 1, 1, 3, 3, 3, 3, 3, 3, 1, U, 7, 3, 0, 7, 0]))
# :341-352 — 255-run small, 256-run big
KAT_255_256 = ([(PASS, 0, 1)] + [(2, i, 3) for i in range(255)] + [(PASS, 0, 1)] + [(3, i, 4) for i in range(256)],
               bytes([1, U, 255, 2, 0, 255, 0, 1, U, 1, 3, 0, 0, 1, 0, 0]))
# :357-376 — escape pair passed, escape pair inside a (too short) run, 11-run, big run from offset 1
KAT_PARSE = ([(PASS, 0, 1), (PASS, 0, U), (PASS, 0, U), (PASS, 0, 1), (1, 0, U), (1, 1, U), (PASS, 0, 3)]
             + [(2, i, 2) for i in range(11)] + [(PASS, 0, 3)] + [(3, i, 6) for i in range(1, 257)],
             bytes([1, U, U, 1, U, U, 3, U, 11, 2, 0, 11, 0, 3, U, 1, 3, 0, 1, 1, 1, 0]))


@pytest.mark.parametrize("kat", [KAT_6_7, KAT_255_256, KAT_PARSE], ids=["6_7", "255_256", "parse"])
def test_stream_encoder_kats(kat):
    msgs, expect = kat
    cmd, pos, val = _msgs_to_arrays(msgs)
    assert po.stream_encode(cmd, pos, val, strict251=True) == expect
    assert po.stream_encode(cmd, pos, val, strict251=False) == expect


@pytest.mark.parametrize("kat", [KAT_6_7, KAT_255_256, KAT_PARSE], ids=["6_7", "255_256", "parse"])
def test_stream_encoder_kats_on_reference(ref, kat):
    msgs, expect = kat
    assert ref.stream(msgs) == expect


# proj/PiXiuStr.cpp:378-410 — parse(1, 272) across a small and a big record
def test_range_parse_kat():
    ch = po.OracleChunk()
    ch.append(b"")                       # idx 0 (unused)
    ch.append(bytes([U, U]))             # idx 1: referenced by nothing decodable here
    ch.append(bytes([2] * 11 + [8]))     # idx 2 = i2v2
    ch.append(bytes([8] + [6] * 256 + [8]))  # idx 3 = i3v6
    cmd, pos, val = _msgs_to_arrays(KAT_PARSE[0])
    enc = po.stream_encode(cmd, pos, val)
    # the escape pair with cmd=1 is a 2-run => literal, so record 1 is never dereferenced
    idx = ch.append(enc)
    out = ch.decode(idx, 1, 272)
    assert len(out) == 271
    assert out[:6] == bytes([U, U, 1, U, U, 3])
    assert out[6:17] == bytes([2] * 11)
    assert out[17] == 3
    assert out[18:] == bytes([6] * (271 - 18))


# run of exactly 251: reference emits FB FB idx to (bug B1); default diverges (+2 B)
def test_len251_divergence():
    msgs = [(PASS, 0, 1)] + [(0, i, 7) for i in range(251)] + [(PASS, 0, 1)]
    cmd, pos, val = _msgs_to_arrays(msgs)
    strict = po.stream_encode(cmd, pos, val, strict251=True)
    fixed = po.stream_encode(cmd, pos, val, strict251=False)
    assert strict == bytes([1, U, 251, 0, 0, 251, 0, 1])
    assert fixed == bytes([1, U, 1, 0, 0, 251, 0, 0, 0, 1])


def test_len251_on_reference(ref):
    msgs = [(PASS, 0, 1)] + [(0, i, 7) for i in range(251)] + [(PASS, 0, 1)]
    assert ref.stream(msgs) == bytes([1, U, 251, 0, 0, 251, 0, 1])


# README.md:70-96 + SURVEY §8c derived vectors
README_SEQ = [(b"123", b"321"), (b"BOBO", b"https://www.zhihu.com/question/55439090"),
              (b"BOBO1", b"https://www.zhihu.com/question/22454692")]


def test_readme_example():
    w = po.OracleWindow()
    encs = [w.encode(po.make_doc(k, v)) for k, v in README_SEQ]
    assert encs[0] == bytes.fromhex("313233fb00333231fb02")
    assert len(encs[1]) == 47  # all literal: the FB 00 pair is demoted (only its 2nd byte matched)
    assert encs[2] == bytes.fromhex("424f424f31fb00fb1f010025003232343534363932fb02")
    # README: "saves 27": len("SET BOBO1::https://...") - 23 = 50 - 23
    assert len(b"SET BOBO1::https://www.zhihu.com/question/22454692") - len(encs[2]) == 27
    ch = po.OracleChunk()
    for (k, v), e in zip(README_SEQ, encs):
        i = ch.append(e)
        assert ch.decode(i) == po.make_doc(k, v)
        assert po.split_doc(ch.decode(i)) == (k, v)


def test_readme_example_on_reference(ref):
    ref.reset()
    w = po.OracleWindow(strict251=True)
    for k, v in README_SEQ:
        ref.setitem(k, v)
        assert ref.last_encoded() == w.encode(po.make_doc(k, v))


def test_self_overlap_vector():
    # value with period 15 extended to 120 B -> small record len=100 idx=0 to=109
    v = (b"abcdefghijklmnopqrst" + b"fghijklmnopqrst" * 10)[:120]
    w = po.OracleWindow()
    enc = w.encode(po.make_doc(b"kx", v))
    assert enc == bytes.fromhex("6b78fb006162636465666768696a6b6c6d6e6f7071727374fb6400006d00fb02")
    ch = po.OracleChunk()
    assert ch.decode(ch.append(enc)) == po.make_doc(b"kx", v)


def test_make_doc_limits():
    # proj/PiXiuCtrl.cpp:121-174 — max key-only record 65,533 raw bytes -> 65,535 decoded
    assert len(po.make_doc(b"A" * 65533, b"")) == 65535
    assert po.make_doc(b"A" * 65534, b"") is None
    assert len(po.make_doc(b"A" * 30000, b"B" * 35531)) == 65535
    assert po.make_doc(b"A" * 30000, b"B" * 35532) is None
    assert po.make_doc(bytes([251]) * 10, b"x") == bytes([251] * 20) + b"\xfb\x00x\xfb\x02"
