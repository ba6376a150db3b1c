"""The C-ABI library builds, loads and exports every symbol the headers declare (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("pixiu_b200.h", "pixiu_b200_debug.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(pixiu_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "pixiu_b200", "libpixiu_b200.so"))
    decl = _declared()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(lib, name), name


def test_python_mirror_lists_the_same_entry_points():
    from pixiu_b200 import ctrl

    decl = set(_declared())
    assert set(ctrl.EXPORTS) <= decl
    # the drop-in surface of the reference's PiXiuCtrl (proj/PiXiuCtrl.h:11-23)
    for m in ("setitem", "contains", "getitem", "iter", "delitem", "init_prop", "free_prop"):
        assert callable(getattr(ctrl.PiXiuCtrl, m))


def test_generator_protocol():
    from pixiu_b200.ctrl import CBTGen, PXSGen, split_doc

    g = PXSGen(b"ab\xfb\x00c d\xfb\x02")
    ok, b = g()
    assert ok and b == ord("a")
    assert g.consume_repr() == "bcd"   # visible bytes only (PiXiuStr.h:200-211)
    assert g() == (False, None)
    it = CBTGen([b"k\xfb\x00v\xfb\x02", b"l\xfb\x00"])
    docs = [x.bytes() for x in it]
    assert [split_doc(d) for d in docs] == [(b"k", b"v"), (b"l", b"")]
    assert split_doc(b"\xfb\xfb\xfb\x00x\xfb\xfb\xfb\x02") == (b"\xfb", b"x\xfb")


def test_fails_loudly_without_a_gpu():
    """no CPU fallback: on a machine without a CUDA device the store cannot even be created"""
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pixiu_b200 import ctrl

    with pytest.raises(ctrl.PiXiuError):
        ctrl.PiXiuCtrl()
