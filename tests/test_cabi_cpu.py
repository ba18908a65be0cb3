"""The C-ABI library loads without a GPU and exports every symbol include/agile3d_b200.h declares."""
import ctypes
import os
import re

from helpers import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "agile3d_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ag3d_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from agile3d_b200.build import build
    path = build()
    handle = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    handle.ag3d_abi_version.restype = ctypes.c_int32
    assert handle.ag3d_abi_version() == 7


def test_ctypes_signature_table_covers_the_header():
    from agile3d_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()


def test_argument_validation_without_gpu():
    """Entry points validate arguments before touching CUDA, so this runs on a CPU box."""
    from agile3d_b200 import _lib
    L = _lib.lib()
    assert L.ag3d_hash_capacity(1000) == 2048 and L.ag3d_hash_capacity(150000) == 524288
    rc = L.ag3d_spconv_fwd(None, 32, 33, None, 1, 10, None, None, 32, None, None, None, 0, None, 32, 0, 1, None, 0, None)
    assert rc == -1 and b"multiples of 32" in L.ag3d_last_error()
    rc = L.ag3d_s2c_mask_fwd(None, None, 10, None, None, None, None, None, None, 1e-5, None, None, 300, 8, 3, None,
                             None, None, None, 0, None, 0, None)
    assert rc == -1 and b"256 click queries" in L.ag3d_last_error()
    assert L.ag3d_query_blob_floats() == 14 * 128 * 128 + 2 * 128 * 1024 + 1024 + 21 * 128
