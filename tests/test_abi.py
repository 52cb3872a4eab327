"""CPU-only: libckb200.so builds for sm_100a, loads, and exports every entry point that
include/ckb200.h declares; struct layouts of the ctypes binding match the header.  No compute
call is made here (there is no GPU in the build container and no CPU fallback in the library)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import PKG, ROOT


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, PKG)
    import build as B
    return B.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "ckb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ck_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    names = _declared()
    assert len(names) >= 30
    L = C.CDLL(built)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    L.ck_abi_version.restype = C.c_int
    assert L.ck_abi_version() == 2


def test_binding_struct_sizes(built):
    from ckb200 import lib
    # sizes computed from the header by the C compiler
    prog = r'''
    #include <stdio.h>
    #include "ckb200.h"
    int main(void) { printf("%zu %zu %zu %zu %zu %zu\n", sizeof(ck_pos), sizeof(ck_leaf), sizeof(ck_engine_cfg),
                            sizeof(ck_record), sizeof(ck_game_result), sizeof(ck_run_stats)); return 0; }
    '''
    exe = os.path.join(ROOT, "tests", "host_rules", "_build", "abi_sizes")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
    sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [lib.POS_DTYPE.itemsize, lib.LEAF_DTYPE.itemsize, C.sizeof(lib.EngineCfg),
                     lib.RECORD_DTYPE.itemsize, lib.GAME_DTYPE.itemsize, C.sizeof(lib.RunStats)]


def test_no_device_fails_loudly(built):
    """on a box without a GPU every compute entry point must error out, never fall back"""
    from ckb200 import lib
    if lib.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(lib.CkError):
        lib.movegen(np.zeros(1, dtype=lib.POS_DTYPE))
    with pytest.raises(lib.CkError):
        lib.Net(0)
    with pytest.raises(lib.CkError):
        lib.Engine(lib.make_cfg(n_slots=1, budget=1, evaluator="hash"))


def test_sass_is_sm100(built):
    out = subprocess.run(["cuobjdump", "-lelf", built], capture_output=True, text=True).stdout
    assert "sm_100a" in out
