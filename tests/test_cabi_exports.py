"""The C-ABI library loads without a GPU and exports every function include/tokengeex_b200.h declares; compute entry
points fail loudly (TGX_ERR_NO_DEVICE) instead of falling back to a CPU path.  CPU only."""
import ctypes
import os
import re

import numpy as np

from tokengeex_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "tokengeex_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tgx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = declared_functions()
    assert len(names) >= 18 and "tgx_encode_batch" in names and "tgx_model_rebuild" in names
    lib = ctypes.CDLL(os.path.join(ROOT, "tokengeex_b200", "csrc", "libtokengeex_b200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    bound = {name for name, _, _ in N.SYMBOLS}
    assert set(declared_functions()) <= bound


def test_no_cpu_fallback():
    m = N.Model([b"a", b"b", b"ab"], [-1.0, -1.0, -1.5], device=None)  # host-only model: trie queries only
    assert m.common_prefix_search(b"abx") == ([0, 2], [1, 2])
    blob, off = N.pack([b"ab"])
    L = N.lib()
    ids = np.zeros(4, np.uint32)
    id_off = np.zeros(2, np.uint64)
    rc = L.tgx_encode_batch(m._h, blob.ctypes.data_as(N.u8p), off.ctypes.data_as(N.u64p), 1, 0,
                            ids.ctypes.data_as(N.u32p), 4, id_off.ctypes.data_as(N.u64p), None, None, None)
    assert rc == N.TGX_ERR_NO_DEVICE
    assert b"no CPU compute path" in L.tgx_last_error()
    # rebuild in place works on a host-only model too
    m.rebuild([b"a", b"b", b"abx"], [-1.0, -1.0, -2.0])
    assert m.common_prefix_search(b"abx") == ([0, 2], [1, 3])
