"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle.

* ``Oracle*`` classes wrap ``oracle/liboracle.so`` (the C restatement,
  ``oracle/pixiu_oracle.c``).
* ``Ref`` wraps ``oracle/_ref/libpixiu_ref.so`` — the UNMODIFIED reference
  compiled from /root/reference/src by ``oracle/Makefile`` (present only where
  it was built; it travels to the GPU box as a prebuilt file).

Nothing under ``pixiu_b200/`` imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_longlong)


def _buf(b):
    """bytes/ndarray -> (ctypes pointer, keepalive)"""
    if isinstance(b, np.ndarray):
        a = np.ascontiguousarray(b, dtype=np.uint8)
    else:
        a = np.frombuffer(bytes(b), dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)
    return a.ctypes.data_as(_u8p), a


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/libpixiu_ref.so when /root/reference is mounted)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")) or (
        os.path.exists("/root/reference/src/proj/PiXiuCtrl.cpp")
        and not os.path.exists(os.path.join(HERE, "_ref", "libpixiu_ref.so"))
    ):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.pxo_escape.argtypes = [_u8p, C.c_int, C.c_int, _u8p]
        L.pxo_make_doc.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p]
        L.pxo_stream_encode.argtypes = [C.c_int, _i32p, _i32p, _u8p, _u8p, C.c_int]
        L.pxo_window_new.restype = C.c_void_p
        L.pxo_window_free.argtypes = [C.c_void_p]
        L.pxo_window_count.argtypes = [C.c_void_p]
        L.pxo_window_encode.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p, C.c_int, _i32p, _i32p]
        L.pxo_chunk_new.restype = C.c_void_p
        L.pxo_chunk_free.argtypes = [C.c_void_p]
        L.pxo_chunk_append.argtypes = [C.c_void_p, _u8p, C.c_int]
        L.pxo_chunk_count.argtypes = [C.c_void_p]
        L.pxo_chunk_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _u8p]
        L.pxo_chunk_declen.argtypes = [C.c_void_p, C.c_int]
        L.pxo_cbt_new.restype = C.c_void_p
        L.pxo_cbt_free.argtypes = [C.c_void_p]
        for f in ("pxo_cbt_get", "pxo_cbt_del", "pxo_cbt_depth"):
            getattr(L, f).argtypes = [C.c_void_p, _u8p, C.c_int]
            getattr(L, f).restype = C.c_longlong
        L.pxo_cbt_set.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_longlong]
        L.pxo_cbt_set.restype = C.c_longlong
        L.pxo_cbt_iter.argtypes = [C.c_void_p, _u8p, C.c_int, _i64p, C.c_longlong]
        L.pxo_cbt_iter.restype = C.c_longlong
        L.pxo_cbt_size.argtypes = [C.c_void_p]
        L.pxo_cbt_size.restype = C.c_longlong
        _lib = L
    return _lib


# ----------------------------------------------------------------------------
# codec helpers
# ----------------------------------------------------------------------------
def escape(src: bytes, is_key: bool = False) -> bytes:
    p, _k = _buf(src)
    out = (C.c_uint8 * (2 * len(src) + 4))()
    n = lib().pxo_escape(p, len(src), int(is_key), out)
    return bytes(out[:n])


def unescape(esc: bytes) -> bytes:
    """inverse of escape() for a terminator-free escaped string"""
    out = bytearray()
    i = 0
    while i < len(esc):
        out.append(esc[i])
        i += 2 if esc[i] == 251 else 1
    return bytes(out)


def make_doc(k: bytes, v: bytes) -> bytes | None:
    pk, _a = _buf(k)
    pv, _b = _buf(v)
    out = (C.c_uint8 * 65536)()
    n = lib().pxo_make_doc(pk, len(k), pv, len(v), out)
    return None if n < 0 else bytes(out[:n])


def split_doc(doc: bytes) -> tuple[bytes, bytes]:
    """decoded doc `esc(k) 251 0 [esc(v) 251 2]` -> (k, v) un-escaped"""
    i = 0
    while True:
        if doc[i] == 251:
            if doc[i + 1] == 0:
                break
            i += 2
        else:
            i += 1
    k = unescape(doc[:i])
    rest = doc[i + 2:]
    if not rest:
        return k, b""
    assert rest[-2:] == b"\xfb\x02"
    return k, unescape(rest[:-2])


def stream_encode(cmd, pos, val, strict251: bool = False) -> bytes:
    n = len(val)
    c = np.ascontiguousarray(cmd, dtype=np.int32)
    p = np.ascontiguousarray(pos, dtype=np.int32)
    pv, _k = _buf(bytes(val))
    out = (C.c_uint8 * (n + 16))()
    m = lib().pxo_stream_encode(n, c.ctypes.data_as(_i32p), p.ctypes.data_as(_i32p), pv, out, int(strict251))
    return bytes(out[:m])


class OracleWindow:
    """One compression window (= chunk): records are encoded against all earlier ones."""

    def __init__(self, strict251: bool = False):
        self._w = lib().pxo_window_new()
        self.strict251 = strict251

    def __del__(self):
        if getattr(self, "_w", None):
            lib().pxo_window_free(self._w)
            self._w = None

    def __len__(self):
        return lib().pxo_window_count(self._w)

    def encode(self, doc: bytes, want_msgs: bool = False):
        p, _k = _buf(doc)
        out = (C.c_uint8 * 65536)()
        if want_msgs:
            cmd = np.zeros(len(doc) + 1, dtype=np.int32)
            pos = np.zeros(len(doc) + 1, dtype=np.int32)
            n = lib().pxo_window_encode(self._w, p, len(doc), out, int(self.strict251),
                                        cmd.ctypes.data_as(_i32p), pos.ctypes.data_as(_i32p))
            return bytes(out[:n]), cmd[:len(doc)], pos[:len(doc)]
        n = lib().pxo_window_encode(self._w, p, len(doc), out, int(self.strict251), None, None)
        return bytes(out[:n])


class OracleChunk:
    """Decoder over the encoded records of one chunk."""

    def __init__(self):
        self._c = lib().pxo_chunk_new()

    def __del__(self):
        if getattr(self, "_c", None):
            lib().pxo_chunk_free(self._c)
            self._c = None

    def append(self, enc: bytes) -> int:
        p, _k = _buf(enc)
        r = lib().pxo_chunk_append(self._c, p, len(enc))
        if r < 0:
            raise ValueError("malformed encoded record")
        return r

    def __len__(self):
        return lib().pxo_chunk_count(self._c)

    def decode(self, idx: int, frm: int = 0, to: int = 65535) -> bytes:
        out = (C.c_uint8 * 65536)()
        n = lib().pxo_chunk_decode(self._c, idx, frm, to, out)
        if n < 0:
            raise IndexError(idx)
        return bytes(out[:n])


class OracleCritBit:
    def __init__(self):
        self._t = lib().pxo_cbt_new()

    def __del__(self):
        if getattr(self, "_t", None):
            lib().pxo_cbt_free(self._t)
            self._t = None

    def __len__(self):
        return lib().pxo_cbt_size(self._t)

    def set(self, qkey: bytes, leaf_id: int) -> int:
        p, _k = _buf(qkey)
        return lib().pxo_cbt_set(self._t, p, len(qkey), leaf_id)

    def get(self, qkey: bytes) -> int:
        p, _k = _buf(qkey)
        return lib().pxo_cbt_get(self._t, p, len(qkey))

    def delete(self, qkey: bytes) -> int:
        p, _k = _buf(qkey)
        return lib().pxo_cbt_del(self._t, p, len(qkey))

    def depth(self, qkey: bytes) -> int:
        p, _k = _buf(qkey)
        return lib().pxo_cbt_depth(self._t, p, len(qkey))

    def iter(self, prefix_esc: bytes) -> list[int]:
        p, _k = _buf(prefix_esc)
        cap = max(1, len(self))
        ids = np.zeros(cap, dtype=np.int64)
        n = lib().pxo_cbt_iter(self._t, p, len(prefix_esc), ids.ctypes.data_as(_i64p), cap)
        return ids[:n].tolist()


class OracleStore:
    """PiXiuCtrl semantics (PiXiuCtrl.cpp:12-86) assembled from the oracle parts.

    Window rotation is by an explicit ``rotate()`` call or ``max_records`` — the
    reference's arena-pool trigger (PiXiuCtrl.cpp:13) is restated separately.
    """

    def __init__(self, strict251: bool = False, max_records: int = 65535):
        self.strict251 = strict251
        self.max_records = max_records
        self.cbt = OracleCritBit()
        self.chunks: list[OracleChunk] = []
        self.encoded: list[list[bytes]] = []
        self.live: dict[int, bool] = {}
        self.window = None
        self.rotate()

    def rotate(self):
        self.window = OracleWindow(self.strict251)
        self.chunks.append(OracleChunk())
        self.encoded.append([])

    def setitem(self, k: bytes, v: bytes) -> int:
        doc = make_doc(k, v)
        if doc is None:
            raise ValueError("record too long")
        if len(self.window) >= self.max_records:
            self.rotate()
        enc = self.window.encode(doc)
        c = len(self.chunks) - 1
        idx = self.chunks[c].append(enc)
        self.encoded[c].append(enc)
        leaf = (c << 16) | idx
        self.live[leaf] = True
        old = self.cbt.set(escape(k, True), leaf)
        if old >= 0:
            self.live[old] = False  # tombstone; bytes stay (PiXiuStr.cpp:178-187)
            return 1
        return 0

    def _doc(self, leaf: int) -> bytes:
        return self.chunks[leaf >> 16].decode(leaf & 0xFFFF)

    def contains(self, k: bytes) -> bool:
        return self.cbt.get(escape(k, True)) >= 0

    def getitem(self, k: bytes):
        leaf = self.cbt.get(escape(k, True))
        return None if leaf < 0 else self._doc(leaf)

    def delitem(self, k: bytes) -> int:
        leaf = self.cbt.delete(escape(k, True))
        if leaf < 0:
            return 1
        self.live[leaf] = False
        return 0

    def iter(self, prefix: bytes) -> list[bytes]:
        return [self._doc(l) for l in self.cbt.iter(escape(prefix, False))]


# ----------------------------------------------------------------------------
# the real reference (oracle/_ref), when built
# ----------------------------------------------------------------------------
REF_SO = os.path.join(HERE, "_ref", "libpixiu_ref.so")


def ref_available() -> bool:
    if not os.path.exists(REF_SO) and os.path.exists("/root/reference/src/proj/PiXiuCtrl.cpp"):
        try:
            build(force=True)
        except Exception:
            return False
    return os.path.exists(REF_SO)


def _pack(items):
    off = np.zeros(len(items) + 1, dtype=np.int64)
    np.cumsum([len(x) for x in items], out=off[1:])
    data = np.frombuffer(b"".join(items), dtype=np.uint8) if off[-1] else np.zeros(1, dtype=np.uint8)
    return data, off


class Ref:
    """The reference PiXiuCtrl (one instance per process — it is not re-entrant)."""

    _alive = False

    def __init__(self):
        assert not Ref._alive, "reference is not re-entrant: one Ref per process"
        L = C.CDLL(REF_SO)
        L.ref_setitem.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        L.ref_last_info.argtypes = [_i64p, _i32p, _i32p, _i32p, _i32p]
        L.ref_encoded.argtypes = [C.c_int, _u8p, C.c_int]
        L.ref_contains.argtypes = [_u8p, C.c_int]
        L.ref_delitem.argtypes = [_u8p, C.c_int]
        L.ref_getitem.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        L.ref_iter.argtypes = [_u8p, C.c_int, _u8p, C.c_longlong, _i64p, C.c_int]
        L.ref_setitem_batch.argtypes = [C.c_int, _u8p, _i64p, _u8p, _i64p, _i32p, _i32p, _i64p, _i32p, _i32p, _i32p]
        L.ref_setitem_batch.restype = C.c_double
        L.ref_getitem_batch.argtypes = [C.c_int, _u8p, _i64p, _i64p, _i32p]
        L.ref_getitem_batch.restype = C.c_double
        L.ref_contains_batch.argtypes = [C.c_int, _u8p, _i64p, _u8p]
        L.ref_contains_batch.restype = C.c_double
        L.ref_stream.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, C.c_int]
        L.ref_escape.argtypes = [_u8p, C.c_int, C.c_int, _u8p, C.c_int]
        self.L = L
        assert L.ref_init() == 0
        Ref._alive = True

    def close(self):
        if Ref._alive:
            self.L.ref_free()
            Ref._alive = False

    def reset(self):
        self.L.ref_free()
        assert self.L.ref_init() == 0

    def setitem(self, k: bytes, v: bytes) -> int:
        pk, _a = _buf(k)
        pv, _b = _buf(v)
        return self.L.ref_setitem(pk, len(k), pv, len(v))

    def last_info(self):
        cs = C.c_longlong()
        idx, el, pools, used = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        self.L.ref_last_info(C.byref(cs), C.byref(idx), C.byref(el), C.byref(pools), C.byref(used))
        return dict(chunk=cs.value, idx=idx.value, enc_len=el.value, pools=pools.value, pool_used=used.value)

    def encoded(self, idx: int) -> bytes:
        out = (C.c_uint8 * 65536)()
        n = self.L.ref_encoded(idx, out, 65536)
        if n < 0:
            raise IndexError(idx)
        return bytes(out[:n])

    def last_encoded(self) -> bytes:
        return self.encoded(self.last_info()["idx"])

    def contains(self, k: bytes) -> bool:
        pk, _a = _buf(k)
        return bool(self.L.ref_contains(pk, len(k)))

    def delitem(self, k: bytes) -> int:
        pk, _a = _buf(k)
        return self.L.ref_delitem(pk, len(k))

    def getitem(self, k: bytes):
        pk, _a = _buf(k)
        out = (C.c_uint8 * 70000)()
        n = self.L.ref_getitem(pk, len(k), out, 70000)
        return None if n < 0 else bytes(out[:min(n, 70000)])

    def iter(self, prefix: bytes, cap: int = 1 << 24, max_n: int = 1 << 20) -> list[bytes]:
        pp, _a = _buf(prefix)
        out = np.zeros(cap, dtype=np.uint8)
        offs = np.zeros(max_n + 1, dtype=np.int64)
        n = self.L.ref_iter(pp, len(prefix), out.ctypes.data_as(_u8p), cap, offs.ctypes.data_as(_i64p), max_n)
        assert n >= 0
        return [out[offs[i]:offs[i + 1]].tobytes() for i in range(n)]

    def setitem_batch(self, keys: list[bytes], vals: list[bytes]):
        kd, ko = _pack(keys)
        vd, vo = _pack(vals)
        n = len(keys)
        rc = np.zeros(n, dtype=np.int32)
        el = np.zeros(n, dtype=np.int32)
        ch = np.zeros(n, dtype=np.int64)
        idx = np.zeros(n, dtype=np.int32)
        pools = np.zeros(n, dtype=np.int32)
        used = np.zeros(n, dtype=np.int32)
        dt = self.L.ref_setitem_batch(n, kd.ctypes.data_as(_u8p), ko.ctypes.data_as(_i64p),
                                      vd.ctypes.data_as(_u8p), vo.ctypes.data_as(_i64p),
                                      rc.ctypes.data_as(_i32p), el.ctypes.data_as(_i32p),
                                      ch.ctypes.data_as(_i64p), idx.ctypes.data_as(_i32p),
                                      pools.ctypes.data_as(_i32p), used.ctypes.data_as(_i32p))
        return dict(seconds=dt, rc=rc, enc_len=el, chunk=ch, idx=idx, pools=pools, pool_used=used)

    def getitem_batch(self, keys: list[bytes]):
        kd, ko = _pack(keys)
        tot = C.c_longlong()
        nf = C.c_int32()
        dt = self.L.ref_getitem_batch(len(keys), kd.ctypes.data_as(_u8p), ko.ctypes.data_as(_i64p),
                                      C.byref(tot), C.byref(nf))
        return dict(seconds=dt, bytes=tot.value, found=nf.value)

    def contains_batch(self, keys: list[bytes]):
        kd, ko = _pack(keys)
        found = np.zeros(len(keys), dtype=np.uint8)
        dt = self.L.ref_contains_batch(len(keys), kd.ctypes.data_as(_u8p), ko.ctypes.data_as(_i64p),
                                       found.ctypes.data_as(_u8p))
        return dict(seconds=dt, found=found)

    def stream(self, msgs) -> bytes:
        """msgs: iterable of (cmd, pos, val); brackets with STREAM_ON/OFF"""
        out = (C.c_uint8 * 70000)()
        self.L.ref_stream(-1, 0, 0, out, 70000)
        for cmd, pos, val in msgs:
            self.L.ref_stream(cmd, pos, val, out, 70000)
        n = self.L.ref_stream(-2, 0, 0, out, 70000)
        return bytes(out[:n])

    def escape(self, src: bytes, is_key: bool) -> bytes:
        p, _a = _buf(src)
        out = (C.c_uint8 * (2 * len(src) + 8))()
        n = self.L.ref_escape(p, len(src), int(is_key), out, 2 * len(src) + 8)
        return bytes(out[:n])
