"""Developer probe: at which step does pinned host memory lose its registration inside the Python process?"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tokengeex_b200 import _native as N, synth
from oracle import oracle as O

drv = C.CDLL("libcuda.so.1")
def attr(ptr, tag):
    v = C.c_uint(0)
    rc = drv.cuPointerGetAttribute(C.byref(v), 2, C.c_uint64(ptr))  # CU_POINTER_ATTRIBUTE_MEMORY_TYPE
    print(f"{tag}: rc={rc} memtype={v.value}", flush=True)

a = N.pinned_empty(1 << 20)
ptr = a.ctypes.data
attr(ptr, "after pinned_empty (before any model)")
blob, off = synth.corpus(2, 13, 2_000_000)
toks, sc, kp = synth.vocab(blob, off, 13, 30000, 16, 0.05)
attr(ptr, "after synth")
gm = N.Model(toks, sc, device=0)
attr(ptr, "after model create")
om = O.OracleModel(toks, sc)
attr(ptr, "after oracle model")
r = om.encode_batch(blob, off, crlf=True, threads=1)
attr(ptr, "after oracle encode 1 thread")
r = om.encode_batch(blob, off, crlf=True, threads=8)
attr(ptr, "after oracle encode 8 threads")
ids = gm.encode_batch(blob, off, crlf=True)
attr(ptr, "after gpu encode")
