import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
from tests.util import synth_setup
from tokengeex_b200.prune import ModelVocabularyPruner, Vocab
kind, seed, nbytes, v0, target, subiters = 2, 23, 24_000_000, 60000, 16384, 2
blob, off, toks, sc, kp = synth_setup(kind, seed, nbytes, v0, 16)
res = {}
for th in (1, 8, 16):
    om = O.OracleModel(toks, sc, kp)
    m2, iters = om.prune(blob, off, vocab_size=target, shrink=0.8, em_subiters=subiters, threads=th)
    res[th] = (set(m2.export()[0]), iters)
    print("oracle threads", th, iters, flush=True)
pr = ModelVocabularyPruner(target, shrink_factor=0.8, em_subiters=subiters, dropout=0.0)
vocab, report = pr.prune(Vocab(list(toks), np.array(sc), np.array(kp)), blob, off)
print("gpu", report.vocab_sizes)
g = set(vocab.tokens)
for th in res:
    print("gpu vs oracle", th, "sym diff", len(g ^ res[th][0]))
print("oracle1 vs oracle8", len(res[1][0] ^ res[8][0]), "oracle8 vs oracle16", len(res[8][0] ^ res[16][0]))
print("margins", report.m_margins)
print("cuts", [(bool(a[5]), float(a[6])) for a in report.audits])
