import sys, time, numpy as np, ctypes as C
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tokengeex_b200 import synth, _native as N
kind = int(sys.argv[1]) if len(sys.argv) > 1 else synth.KIND_MULTILANG
vb, vo = synth.corpus(kind, 2, 96_000_000)
toks, sc, kp = synth.vocab(vb, vo, 2, 131072, 16, 0.05)
m = N.Model(toks, sc, device=None)
blob, off = synth.corpus(kind, 7, 3_000_000)
L = N.lib()
n = int(off[-1])
ids = np.zeros(64, np.uint32); lens = np.zeros(64, np.uint32); cnt = C.c_uint64(0)
maxlen = np.zeros(n, np.int32); nmatch = np.zeros(n, np.int32)
S = len(off) - 1
t = time.time()
for s in range(S):
    a, b = int(off[s]), int(off[s + 1])
    for p in range(a, b):
        e = min(b, p + 16)
        L.tgx_model_common_prefix_search(m._h, blob[p:e].ctypes.data_as(N.u8p), e - p, ids.ctypes.data_as(N.u32p), lens.ctypes.data_as(N.u32p), 64, C.byref(cnt))
        k = int(cnt.value)
        nmatch[p] = k
        maxlen[p] = lens[k - 1] if k else 0
print("walk", time.time() - t, "s; matches/pos", nmatch.mean(), "maxlen mean", maxlen.mean())
# cuts: c (absolute) is a cut inside a sample iff no token starting before c ends after c
reach = np.arange(n) + maxlen
gaps = []
for s in range(S):
    a, b = int(off[s]), int(off[s + 1])
    if b - a < 2: continue
    r = np.maximum.accumulate(reach[a:b])
    cut = r[:-1] <= np.arange(a + 1, b)   # cut at position a+1+i
    idx = np.flatnonzero(cut)
    if len(idx):
        g = np.diff(np.concatenate([[0], idx + 1, [b - a]]))
    else:
        g = np.array([b - a])
    gaps.append(g)
g = np.concatenate(gaps)
print("segments", len(g), "mean gap", g.mean(), "p50", np.percentile(g, 50), "p99", np.percentile(g, 99), "p99.9", np.percentile(g, 99.9), "max", g.max())
print("bytes in segments > 256:", g[g > 256].sum() / g.sum(), " > 1024:", g[g > 1024].sum() / g.sum(), " > 4096:", g[g > 4096].sum() / g.sum())
