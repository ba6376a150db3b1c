"""Host-side mirror of the reference's ``PiXiuCtrl`` (proj/PiXiuCtrl.h:7-26) over the C ABI.

``PiXiuCtrl`` keeps the reference's method names and return conventions
(``setitem`` -> 0 / 1=CBT_SET_REPLACE, ``delitem`` -> 0 / 1=CBT_DEL_NOT_FOUND,
``getitem`` -> a ``PXSGen`` or None, ``iter`` -> a ``CBTGen`` or None) and adds the
batched forms the GPU path is built for.  Everything runs in
``libpixiu_b200.so`` (hand-written sm_100a CUDA); importing this module on a
machine where the library is missing raises — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PIXIU_B200_LIB") or os.path.join(HERE, "libpixiu_b200.so")   # (override: A/B builds)

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_u32p = C.POINTER(C.c_uint32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)

OK, EINVAL, ETOOLONG, ECUDA, ENOSPC, ECORRUPT, EINTERNAL, EPOISONED = 0, -1, -2, -3, -4, -5, -6, -7
NCCL_UNIQUE_ID_BYTES = 128
CBT_SET_REPLACE = 1
CBT_DEL_NOT_FOUND = 1
ROTATE_REFERENCE, ROTATE_BYTES, ROTATE_RECORDS = 0, 1, 2


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("rotate_policy", C.c_int32), ("window_bytes", C.c_int64),
                ("strict251", C.c_int32), ("auto_reinsert", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("records", C.c_int64), ("live_records", C.c_int64), ("chunks", C.c_int64),
                ("raw_bytes", C.c_int64), ("doc_bytes", C.c_int64), ("encoded_bytes", C.c_int64),
                ("window_bytes", C.c_int64), ("kernel_launches", C.c_int64),
                ("last_setitem_gpu_ms", C.c_double), ("last_getitem_gpu_ms", C.c_double),
                ("last_lookup_gpu_ms", C.c_double),
                ("reinserted_records", C.c_int64), ("reclaimable_bytes", C.c_int64),
                ("index_key_arena_bytes", C.c_int64), ("index_host_bytes", C.c_int64), ("index_device_bytes", C.c_int64),
                ("table_device_bytes", C.c_int64)]


class MgStats(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("nccl_version", C.c_int32), ("pad", C.c_int32),
                ("batches", C.c_int64), ("max_reduce_bytes", C.c_int64), ("min_reduce_bytes", C.c_int64),
                ("max_reduce_ms", C.c_double), ("min_reduce_ms", C.c_double)]


EXPORTS = [
    "pixiu_default_config", "pixiu_create", "pixiu_destroy", "pixiu_last_error", "pixiu_get_stats",
    "pixiu_setitem_batch", "pixiu_setitem_batch_dev", "pixiu_contains_batch", "pixiu_contains_batch_dev", "pixiu_delitem_batch",
    "pixiu_getitem_batch", "pixiu_getitem_batch_dev", "pixiu_iter", "pixiu_encoded_view",
    "pixiu_record_location", "pixiu_import_chunk", "pixiu_decode_chunk", "pixiu_rotate", "pixiu_reserve",
    "pixiu_profile_enable", "pixiu_profile_get", "pixiu_stream",
    "pixiu_reinsert_chunk", "pixiu_chunk_info",
    "pixiu_export_chunk", "pixiu_mg_config", "pixiu_mg_setitem_begin", "pixiu_mg_setitem_mid", "pixiu_mg_setitem_end",
    "pixiu_mg_unique_id", "pixiu_mg_comm_init", "pixiu_mg_setitem_batch", "pixiu_mg_get_stats",
    "pixiu_export_index", "pixiu_import_chunk_raw", "pixiu_import_index",
]

_lib = None


def load_library():
    """dlopen libpixiu_b200.so and declare its prototypes (fails loudly when it is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    if "PIXIU_NCCL_LIB" not in os.environ:
        # multi-GPU mode binds NCCL at run time: prefer the copy PyTorch bundles (a torchrun process has it loaded
        # anyway, and two NCCL copies in one process are best avoided); the system libnccl.so.2 is the fallback
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl") if importlib.util.find_spec("nvidia") else None
        for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["PIXIU_NCCL_LIB"] = cand
                break
    L = C.CDLL(LIB_PATH)
    L.pixiu_default_config.argtypes = [C.POINTER(Config)]
    L.pixiu_create.argtypes = [C.POINTER(Config)]
    L.pixiu_create.restype = C.c_void_p
    L.pixiu_destroy.argtypes = [C.c_void_p]
    L.pixiu_last_error.argtypes = [C.c_void_p]
    L.pixiu_last_error.restype = C.c_char_p
    L.pixiu_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.pixiu_setitem_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32p, _i32p]
    L.pixiu_setitem_batch_dev.argtypes = L.pixiu_setitem_batch.argtypes
    L.pixiu_contains_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, _u8p]
    L.pixiu_contains_batch_dev.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pixiu_debug_index_depth.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, _i32p]
    L.pixiu_delitem_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, _i32p]
    L.pixiu_getitem_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _i64p, _u8p, _i64p]
    L.pixiu_getitem_batch_dev.argtypes = L.pixiu_getitem_batch.argtypes
    L.pixiu_iter.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _i64p, C.c_int64, _i64p, _i64p]
    L.pixiu_encoded_view.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
    L.pixiu_record_location.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p]
    L.pixiu_import_chunk.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.pixiu_import_chunk.restype = C.c_int64
    L.pixiu_decode_chunk.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _i64p, _i64p]
    L.pixiu_rotate.argtypes = [C.c_void_p]
    L.pixiu_reserve.argtypes = [C.c_void_p, C.c_int64]
    L.pixiu_reinsert_chunk.argtypes = [C.c_void_p, C.c_int64]
    L.pixiu_reinsert_chunk.restype = C.c_int64
    L.pixiu_chunk_info.argtypes = [C.c_void_p, C.c_int64, _i64p, _i64p, _i32p]
    L.pixiu_export_chunk.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _i64p, _i64p, _i64p]
    L.pixiu_export_index.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, _i64p]
    L.pixiu_import_chunk_raw.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    L.pixiu_import_chunk_raw.restype = C.c_int64
    L.pixiu_import_index.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.pixiu_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.pixiu_profile_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), _i64p]
    L.pixiu_stream.argtypes = [C.c_void_p]
    L.pixiu_stream.restype = C.c_void_p
    L.pixiu_mg_config.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.pixiu_mg_setitem_begin.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_void_p), _i64p]
    L.pixiu_mg_setitem_mid.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), _i64p]
    L.pixiu_mg_setitem_end.argtypes = [C.c_void_p, _i32p, _i32p]
    L.pixiu_mg_unique_id.argtypes = [C.c_void_p, C.c_void_p]
    L.pixiu_mg_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.pixiu_mg_setitem_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _i32p, _i32p]
    L.pixiu_mg_get_stats.argtypes = [C.c_void_p, C.POINTER(MgStats)]
    # test hooks
    L.pixiu_debug_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    L.pixiu_debug_sort_pairs.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    L.pixiu_debug_window_array.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]
    L.pixiu_debug_window_array.restype = C.c_int64
    L.pixiu_debug_pool_state.argtypes = [C.c_void_p, _i32p, _i32p]
    L.pixiu_debug_set_knob.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.pixiu_debug_decode_counters.argtypes = [C.c_void_p, _i64p, _i64p]
    _lib = L
    return L


class PiXiuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pixiu error {code}: {msg}")
        self.code = code


def _pack(items):
    off = np.zeros(len(items) + 1, dtype=np.int64)
    if items:
        np.cumsum([len(x) for x in items], out=off[1:])
    data = np.frombuffer(b"".join(items), dtype=np.uint8) if off[-1] else np.zeros(1, dtype=np.uint8)
    return np.ascontiguousarray(data), off


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PXSGen:
    """Byte generator over one decoded record — the protocol of the reference's
    ``$gen(PXSGen)`` (proj/PiXiuStr.h:110-212): calling it yields ``esc(k) 251 0 esc(v) 251 2``
    one byte at a time; ``consume_repr()`` returns the visible bytes (33..126)."""

    def __init__(self, data: bytes):
        self._data = data
        self._pos = 0

    def __call__(self):
        """returns (True, byte) or (False, None) — `bool operator()(uint8_t&)`"""
        if self._pos < len(self._data):
            b = self._data[self._pos]
            self._pos += 1
            return True, b
        return False, None

    def __iter__(self):
        while self._pos < len(self._data):
            b = self._data[self._pos]
            self._pos += 1
            yield b

    def bytes(self) -> bytes:
        return self._data

    def consume_repr(self) -> str:
        out = bytes(b for b in self._data[self._pos:] if 33 <= b <= 126)
        self._pos = len(self._data)
        return out.decode("latin-1")


class CBTGen:
    """Generator of PXSGen objects in key order (``$gen(CBTGen)``, CritBitTree.h:130-157)."""

    def __init__(self, docs):
        self._docs = docs
        self._pos = 0

    def __call__(self):
        if self._pos < len(self._docs):
            g = PXSGen(self._docs[self._pos])
            self._pos += 1
            return True, g
        return False, None

    def __iter__(self):
        while self._pos < len(self._docs):
            g = PXSGen(self._docs[self._pos])
            self._pos += 1
            yield g


def unescape(esc: bytes) -> bytes:
    if 251 not in esc:
        return esc
    out = bytearray()
    i = 0
    while i < len(esc):
        out.append(esc[i])
        i += 2 if esc[i] == 251 else 1
    return bytes(out)


def split_doc(doc: bytes):
    """decoded doc -> (key, value) un-escaped (README.md:157: the caller un-escapes)"""
    i = 0
    n = len(doc)
    while i + 1 < n and not (doc[i] == 251 and doc[i + 1] == 0):
        i += 2 if doc[i] == 251 else 1
    k = unescape(doc[:i])
    rest = doc[i + 2:]
    return (k, unescape(rest[:-2])) if rest else (k, b"")


class PiXiuCtrl:
    """Drop-in for the reference's ``struct PiXiuCtrl`` — same methods, plus ``*_batch``."""

    def __init__(self, device: int = 0, rotate_policy: int = ROTATE_REFERENCE, window_bytes: int = 12_500_000,
                 strict251: bool = False, auto_reinsert: bool = False):
        self._L = load_library()
        self._cfg = Config(device, rotate_policy, window_bytes, int(strict251), int(auto_reinsert))
        self._h = None
        self.init_prop()

    # -- lifecycle (PiXiuCtrl.cpp:77-86) --
    def init_prop(self):
        if self._h:
            self.free_prop()
        self._h = self._L.pixiu_create(C.byref(self._cfg))
        if not self._h:
            raise PiXiuError(ECUDA, "pixiu_create failed (no usable CUDA device?)")

    def free_prop(self):
        if self._h:
            self._L.pixiu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free_prop()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise PiXiuError(rc, self._L.pixiu_last_error(self._h).decode("utf-8", "replace"))
        return rc

    # -- batched API --
    def setitem_batch(self, keys, vals, want_saved=True):
        """keys/vals: lists of bytes, or packed (data u8[], off i64[n+1]) tuples. -> (rc, saved)"""
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        vd, vo = vals if isinstance(vals, tuple) else _pack(vals)
        n = len(ko) - 1
        rc = np.zeros(n, dtype=np.int32)
        saved = np.zeros(n, dtype=np.int32)
        self._check(self._L.pixiu_setitem_batch(self._h, n, _ptr(kd), _ptr(ko), _ptr(vd), _ptr(vo),
                                                rc.ctypes.data_as(_i32p), saved.ctypes.data_as(_i32p)))
        return rc, saved

    def setitem_batch_dev(self, d_keys, d_koff, d_vals, d_voff, n):
        """device pointers (ints) of a packed batch already resident in HBM"""
        rc = np.zeros(n, dtype=np.int32)
        saved = np.zeros(n, dtype=np.int32)
        self._check(self._L.pixiu_setitem_batch_dev(self._h, n, d_keys, d_koff, d_vals, d_voff,
                                                    rc.ctypes.data_as(_i32p), saved.ctypes.data_as(_i32p)))
        return rc, saved

    def contains_batch(self, keys):
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        n = len(ko) - 1
        found = np.zeros(max(n, 1), dtype=np.uint8)
        self._check(self._L.pixiu_contains_batch(self._h, n, _ptr(kd), _ptr(ko), found.ctypes.data_as(_u8p)))
        return found[:n].astype(bool)

    def contains_batch_dev(self, d_keys: int, d_koff: int, n: int, d_found: int):
        """device pointers (ints): packed keys + int64 offsets in HBM, found u8[n] written on the device"""
        self._check(self._L.pixiu_contains_batch_dev(self._h, n, d_keys, d_koff, d_found))

    def debug_index_depth(self, keys):
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        n = len(ko) - 1
        out = np.zeros(max(n, 1), dtype=np.int32)
        self._check(self._L.pixiu_debug_index_depth(self._h, n, _ptr(kd), _ptr(ko), out.ctypes.data_as(_i32p)))
        return out[:n]

    def delitem_batch(self, keys):
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        n = len(ko) - 1
        rc = np.zeros(max(n, 1), dtype=np.int32)
        self._check(self._L.pixiu_delitem_batch(self._h, n, _ptr(kd), _ptr(ko), rc.ctypes.data_as(_i32p)))
        return rc[:n]

    def getitem_batch(self, keys, out=None):
        """-> (data u8[], off i64[n+1], found bool[n]); decoded docs in escaped form"""
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        n = len(ko) - 1
        off = np.zeros(n + 1, dtype=np.int64)
        found = np.zeros(max(n, 1), dtype=np.uint8)
        need = C.c_int64(0)
        cap = 0 if out is None else out.size
        buf = out if out is not None else np.zeros(1, dtype=np.uint8)
        rc = self._L.pixiu_getitem_batch(self._h, n, _ptr(kd), _ptr(ko), _ptr(buf), cap, off.ctypes.data_as(_i64p),
                                         found.ctypes.data_as(_u8p), C.byref(need))
        if rc == ENOSPC:
            buf = np.zeros(max(need.value, 1), dtype=np.uint8)
            rc = self._L.pixiu_getitem_batch(self._h, n, _ptr(kd), _ptr(ko), _ptr(buf), buf.size,
                                             off.ctypes.data_as(_i64p), found.ctypes.data_as(_u8p), C.byref(need))
        self._check(rc)
        return buf, off, found[:n].astype(bool)

    def getitem_batch_dev(self, keys, d_out, cap):
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        n = len(ko) - 1
        off = np.zeros(n + 1, dtype=np.int64)
        found = np.zeros(max(n, 1), dtype=np.uint8)
        need = C.c_int64(0)
        self._check(self._L.pixiu_getitem_batch_dev(self._h, n, _ptr(kd), _ptr(ko), d_out, cap,
                                                    off.ctypes.data_as(_i64p), found.ctypes.data_as(_u8p), C.byref(need)))
        return off, found[:n].astype(bool)

    def iter_docs(self, prefix: bytes):
        p = np.frombuffer(prefix, dtype=np.uint8) if prefix else np.zeros(1, dtype=np.uint8)
        count, need = C.c_int64(0), C.c_int64(0)
        rc = self._L.pixiu_iter(self._h, _ptr(p), len(prefix), None, 0, None, 0, C.byref(count), C.byref(need))
        if rc not in (OK, ENOSPC):
            self._check(rc)
        if count.value == 0:
            return []
        buf = np.zeros(max(need.value, 1), dtype=np.uint8)
        off = np.zeros(count.value + 1, dtype=np.int64)
        self._check(self._L.pixiu_iter(self._h, _ptr(p), len(prefix), _ptr(buf), buf.size, off.ctypes.data_as(_i64p),
                                       off.size, C.byref(count), C.byref(need)))
        b = buf.tobytes()
        return [b[off[i]:off[i + 1]] for i in range(count.value)]

    # -- the reference's single-record API (PiXiuCtrl.h:11-19) --
    def setitem(self, k: bytes, v: bytes = b"") -> int:
        rc, _ = self.setitem_batch([bytes(k)], [bytes(v)])
        return int(rc[0])

    def contains(self, k: bytes) -> bool:
        return bool(self.contains_batch([bytes(k)])[0])

    def getitem(self, k: bytes):
        buf, off, found = self.getitem_batch([bytes(k)])
        return PXSGen(buf[:off[1]].tobytes()) if found[0] else None

    def delitem(self, k: bytes) -> int:
        return int(self.delitem_batch([bytes(k)])[0])

    def iter(self, prefix: bytes = b""):
        if self.stats().live_records == 0:
            return None  # CritBitTree::iter returns NULL on an empty tree (CritBitTree.cpp:271-274)
        return CBTGen(self.iter_docs(bytes(prefix)))

    # -- inspection --
    def stats(self) -> Stats:
        s = Stats()
        self._check(self._L.pixiu_get_stats(self._h, C.byref(s)))
        return s

    def encoded(self, chunk: int, idx: int) -> bytes:
        buf = np.zeros(65536, dtype=np.uint8)
        n = self._check(self._L.pixiu_encoded_view(self._h, chunk, idx, _ptr(buf), buf.size))
        return buf[:n].tobytes()

    def record_location(self, record: int):
        c, i = C.c_int64(), C.c_int64()
        self._check(self._L.pixiu_record_location(self._h, record, C.byref(c), C.byref(i)))
        return c.value, i.value

    def reinsert(self, chunk: int) -> int:
        """PiXiuCtrl::reinsert (PiXiuCtrl.cpp:88-114): move the live records of a closed chunk into the open window
        and drop the chunk; returns the number of records moved"""
        return self._check(self._L.pixiu_reinsert_chunk(self._h, chunk))

    def chunk_info(self, chunk: int):
        """-> (records ever stored in the chunk, live records, dropped by a re-insertion)"""
        t, l, d = C.c_int64(), C.c_int64(), C.c_int32()
        self._check(self._L.pixiu_chunk_info(self._h, chunk, C.byref(t), C.byref(l), C.byref(d)))
        return t.value, l.value, bool(d.value)

    def import_chunk(self, encs) -> int:
        ed, eo = encs if isinstance(encs, tuple) else _pack(encs)
        return self._check(self._L.pixiu_import_chunk(self._h, len(eo) - 1, _ptr(ed), _ptr(eo)))

    def import_chunk_raw(self, encs) -> int:
        """a chunk's records without indexing them (the index follows with import_index)"""
        ed, eo = encs if isinstance(encs, tuple) else _pack(encs)
        return self._check(self._L.pixiu_import_chunk_raw(self._h, len(eo) - 1, _ptr(ed), _ptr(eo)))

    def export_index(self) -> np.ndarray:
        """the CritBit index in its wire format (feeds import_index of a store holding the same chunks)"""
        need = C.c_int64(0)
        rc = self._L.pixiu_export_index(self._h, None, 0, C.byref(need))
        if rc not in (OK, ENOSPC):
            self._check(rc)
        buf = np.zeros(max(need.value, 1), dtype=np.uint8)
        self._check(self._L.pixiu_export_index(self._h, _ptr(buf), buf.size, C.byref(need)))
        return buf[:need.value]

    def import_index(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._check(self._L.pixiu_import_index(self._h, _ptr(blob), blob.size))

    def export_chunk(self, chunk: int):
        """-> (enc u8[], off i64[n+1]): the chunk's wire format (feeds import_chunk of another store)"""
        count, need = C.c_int64(0), C.c_int64(0)
        rc = self._L.pixiu_export_chunk(self._h, chunk, None, 0, None, C.byref(count), C.byref(need))
        if rc not in (OK, ENOSPC):
            self._check(rc)
        buf = np.zeros(max(need.value, 1), dtype=np.uint8)
        off = np.zeros(count.value + 1, dtype=np.int64)
        self._check(self._L.pixiu_export_chunk(self._h, chunk, _ptr(buf), buf.size, off.ctypes.data_as(_i64p),
                                               C.byref(count), C.byref(need)))
        return buf[:need.value], off

    def decode_chunk(self, chunk: int):
        n_off = 65536 + 1
        off = np.zeros(n_off, dtype=np.int64)
        need = C.c_int64(0)
        rc = self._L.pixiu_decode_chunk(self._h, chunk, None, 0, off.ctypes.data_as(_i64p), C.byref(need))
        if rc not in (OK, ENOSPC):
            self._check(rc)
        buf = np.zeros(max(need.value, 1), dtype=np.uint8)
        self._check(self._L.pixiu_decode_chunk(self._h, chunk, _ptr(buf), buf.size, off.ctypes.data_as(_i64p), C.byref(need)))
        return buf, off

    # -- multi-GPU extended window: three phases around the caller's two all-reduces --
    def mg_config(self, rank: int, world: int):
        self._check(self._L.pixiu_mg_config(self._h, rank, world))

    def mg_setitem_begin(self, keys, vals):
        """-> (device pointer of uint32 M[count], count): all_reduce(MAX) it in place"""
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        vd, vo = vals if isinstance(vals, tuple) else _pack(vals)
        self._mg_n = len(ko) - 1
        p, cnt = C.c_void_p(), C.c_int64()
        self._check(self._L.pixiu_mg_setitem_begin(self._h, self._mg_n, _ptr(kd), _ptr(ko), _ptr(vd), _ptr(vo),
                                                   C.byref(p), C.byref(cnt)))
        return p.value or 0, cnt.value

    def mg_setitem_mid(self):
        """-> (device pointer of uint32 cand[count], count): all_reduce(MIN) it in place"""
        p, cnt = C.c_void_p(), C.c_int64()
        self._check(self._L.pixiu_mg_setitem_mid(self._h, C.byref(p), C.byref(cnt)))
        return p.value or 0, cnt.value

    def mg_setitem_end(self):
        rc = np.zeros(self._mg_n, dtype=np.int32)
        saved = np.zeros(self._mg_n, dtype=np.int32)
        self._check(self._L.pixiu_mg_setitem_end(self._h, rc.ctypes.data_as(_i32p), saved.ctypes.data_as(_i32p)))
        return rc, saved

    # -- the same with the collectives inside the library (NCCL on the store's stream; pixiu_b200/csrc/mgcomm.cu) --
    def mg_unique_id(self) -> bytes:
        """rank 0: a fresh ncclUniqueId (128 bytes) to hand to every rank's mg_comm_init"""
        buf = (C.c_uint8 * NCCL_UNIQUE_ID_BYTES)()
        self._check(self._L.pixiu_mg_unique_id(self._h, buf))
        return bytes(buf)

    def mg_comm_init(self, rank: int, world: int, unique_id: bytes):
        assert len(unique_id) == NCCL_UNIQUE_ID_BYTES
        buf = (C.c_uint8 * NCCL_UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._check(self._L.pixiu_mg_comm_init(self._h, rank, world, buf))

    def mg_setitem_batch(self, keys, vals):
        """one replicated batch through the sharded window (collective call: every rank, same batch)"""
        kd, ko = keys if isinstance(keys, tuple) else _pack(keys)
        vd, vo = vals if isinstance(vals, tuple) else _pack(vals)
        n = len(ko) - 1
        rc = np.zeros(n, dtype=np.int32)
        saved = np.zeros(n, dtype=np.int32)
        self._check(self._L.pixiu_mg_setitem_batch(self._h, n, _ptr(kd), _ptr(ko), _ptr(vd), _ptr(vo),
                                                   rc.ctypes.data_as(_i32p), saved.ctypes.data_as(_i32p)))
        return rc, saved

    def mg_stats(self) -> MgStats:
        st = MgStats()
        self._check(self._L.pixiu_mg_get_stats(self._h, C.byref(st)))
        return st

    def profile_enable(self, on: bool = True):
        self._check(self._L.pixiu_profile_enable(self._h, int(on)))

    def profile(self) -> dict:
        """{class: dict(ms, bytes, launches)} accumulated since profile_enable(True)"""
        out = {}
        cls = 0
        while True:
            name, ms, by, ln = C.c_char_p(), C.c_double(), C.c_double(), C.c_int64()
            rc = self._L.pixiu_profile_get(self._h, cls, C.byref(name), C.byref(ms), C.byref(by), C.byref(ln))
            if rc != 0:
                break
            out[name.value.decode()] = dict(ms=ms.value, bytes=by.value, launches=ln.value)
            cls += 1
        return out

    def stream(self) -> int:
        return int(self._L.pixiu_stream(self._h) or 0)

    def rotate(self):
        self._check(self._L.pixiu_rotate(self._h))

    def reserve(self, encoded_bytes: int):
        """capacity hint: map room for `encoded_bytes` more compressed bytes now instead of inside the batches"""
        self._check(self._L.pixiu_reserve(self._h, int(encoded_bytes)))

    def debug_set_knob(self, name: str, value: int):
        """tuning / test knob of a live store (Knobs::set in csrc/store.h)"""
        self._check(self._L.pixiu_debug_set_knob(self._h, name.encode(), int(value)))

    def debug_decode_counters(self):
        """-> (pending pieces, drain passes) of the last decode call"""
        a, b = C.c_int64(), C.c_int64()
        self._check(self._L.pixiu_debug_decode_counters(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_pool_state(self):
        nth, used = C.c_int32(), C.c_int32()
        self._L.pixiu_debug_pool_state(self._h, C.byref(nth), C.byref(used))
        return nth.value, used.value

    def debug_window_array(self, name: str, dtype):
        n = self.stats().window_bytes
        out = np.zeros(max(n, 1), dtype=dtype)
        r = self._L.pixiu_debug_window_array(self._h, name.encode(), _ptr(out), out.nbytes)
        if r < 0:
            raise PiXiuError(r, "debug_window_array")
        return out[:r]
