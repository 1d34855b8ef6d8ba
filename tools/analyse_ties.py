import sys, time, numpy as np, ctypes as C
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from tokengeex_b200 import synth, _native as N
kind = int(sys.argv[1]) if len(sys.argv) > 1 else synth.KIND_MULTILANG
vseed = 2 if kind == synth.KIND_MULTILANG else 3
vb, vo = synth.corpus(kind, vseed, 96_000_000)
toks, sc, kp = synth.vocab(vb, vo, vseed, 131072, 16, 0.05)
print("distinct scores", len(np.unique(sc)), "of", len(sc))
m = N.Model(toks, sc, device=None)
blob, off = synth.corpus(kind, 7, 1_500_000)
L = N.lib()
n = int(off[-1])
ids = np.zeros(64, np.uint32); lens = np.zeros(64, np.uint32); cnt = C.c_uint64(0)
S = len(off) - 1
EPS = 2.0 ** -20
nseg = nfrag = ntie = 0
frag_samples = 0
seglen_frag = []
for s in range(S):
    a, b = int(off[s]), int(off[s + 1])
    nn = b - a
    M = []
    for p in range(a, b):
        e = min(b, p + 16)
        L.tgx_model_common_prefix_search(m._h, blob[p:e].ctypes.data_as(N.u8p), e - p, ids.ctypes.data_as(N.u32p), lens.ctypes.data_as(N.u32p), 64, C.byref(cnt))
        k = int(cnt.value)
        M.append([(int(lens[i]), float(sc[ids[i]])) for i in range(k)])
    # relative dp per segment, margins
    NEG = float('-inf')
    dp = [NEG] * (nn + 1); dp[0] = 0.0
    frag_cell = [False] * (nn + 1)
    tie_cell = [False] * (nn + 1)
    reach = 0
    seg_start = 0
    sample_frag = False
    for p in range(nn):
        if p >= reach and p > 0:  # cut at p: close segment [seg_start, p)
            nseg += 1
            f = any(frag_cell[seg_start + 1:p + 1]); t = any(tie_cell[seg_start + 1:p + 1])
            nfrag += f; ntie += t
            if f: seglen_frag.append(p - seg_start); sample_frag = True
            seg_start = p
            dp[p] = 0.0 if dp[p] > NEG else NEG   # rebase (relative to the cut)
        if dp[p] > NEG:
            for (l, w) in M[p]:
                cand = dp[p] + w
                q = p + l
                if dp[q] > NEG and abs(cand - dp[q]) <= EPS:
                    frag_cell[q] = True
                    if cand == dp[q]: tie_cell[q] = True
                if cand > dp[q]:
                    dp[q] = cand
        for (l, w) in M[p]:
            reach = max(reach, p + l)
        reach = max(reach, p + 1) if not M[p] else reach
    nseg += 1
    f = any(frag_cell[seg_start + 1:nn + 1]); nfrag += f; ntie += any(tie_cell[seg_start + 1:nn + 1])
    sample_frag |= f
    frag_samples += sample_frag
print("samples", S, "with fragile segment", frag_samples)
print("segments", nseg, "fragile (|diff|<=2^-20 at some cell)", nfrag, nfrag / nseg, "exact ties", ntie, ntie / nseg)
if seglen_frag: print("fragile seg len mean", np.mean(seglen_frag))
