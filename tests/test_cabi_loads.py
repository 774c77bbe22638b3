"""CPU: the C-ABI library loads without a GPU and exports every symbol include/signer_gibbs.h declares."""
import ctypes as C
import os
import re

from signer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "signer_gibbs.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(signer_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported(sg):
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(_lib.EXPORTS) == names


def test_version_and_no_gpu_error_text(sg):
    L = _lib.lib()
    assert L.signer_version() == 3 and L.signer_draw_spec_version() == 7
    n = L.signer_device_count()
    if n < 0:       # CPU container: a status, not a crash
        assert n == -3 and L.signer_last_error()


def test_host_side_hooks_need_no_gpu(sg, oracle):
    L = _lib.lib()
    o = (C.c_int32 * 7)()
    for seed, sweep in [(1, 0), (2 ** 40 + 3, 9), (0, 2 ** 32 - 1)]:
        L.signer_test_perm(seed, sweep, o)
        assert list(o) == oracle.perm(seed, sweep)
        assert sorted(o) == [1, 2, 3, 4, 5, 6, 7]
    ctr = (C.c_uint32 * 4)(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)
    key = (C.c_uint32 * 2)(0xa4093822, 0x299f31d0)
    out = (C.c_uint32 * 4)()
    L.signer_test_philox(ctr, key, out)
    assert [int(x) for x in out] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
