"""The EM loop of tokengeex_b200/prune.py (ModelVocabularyPruner.prune, src/prune.rs:23-57) on the CPU: the two device
passes (E-step, frequency pass) are answered by the oracle, everything else — M-step, in-place rebuilds of the host
double-array, the selection over the model's own trie, the vocabulary packed once per step — is the product's host code.
With one oracle thread (one order of the f64 sums) the schedule must land exactly where the oracle's own does.
The GPU kernels themselves are checked in tests/test_gpu_prune.py."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import synth_setup
from tokengeex_b200 import _native as N


class OracleBackedModel(N.Model):
    """Host-only product model (trie, rebuild, selection) whose device passes go to the oracle."""

    def __init__(self, tokens, scores, device=0):
        super().__init__(tokens, scores, device=None)
        self._om = O.OracleModel(list(tokens), np.asarray(scores, np.float64))
        self.rebuilds = 0

    def rebuild(self, tokens, scores, packed=None):
        assert packed is not None  # the loop hands the pack over
        super().rebuild(tokens, scores, packed=packed)
        self._om = O.OracleModel(list(tokens), np.asarray(scores, np.float64))
        self.rebuilds += 1

    def expected_counts(self, blob, off, snippet_len=81920):
        return self._om.run_e_step(blob, off, threads=1)

    def token_frequencies(self, blob, off, crlf=False):
        return self._om.token_frequencies(blob, off, threads=1), 0, -1, 0


@pytest.mark.parametrize("kind,seed,nbytes,v0,target,subiters", [(2, 31, 400_000, 3000, 1200, 2),
                                                                 (1, 32, 300_000, 2000, 1500, 1)])
def test_em_loop_on_the_host_lands_on_the_oracles_schedule(monkeypatch, kind, seed, nbytes, v0, target, subiters):
    import torch
    if torch.cuda.is_available():
        pytest.skip("the stand-in model serves the host-buffer path only")
    from tokengeex_b200 import prune as P
    blob, off, toks, sc, kp = synth_setup(kind, seed, nbytes, v0, 16)
    om = O.OracleModel(toks, sc, kp)
    want, witers = om.prune(blob, off, vocab_size=target, shrink=0.8, em_subiters=subiters, threads=1)
    wt, ws, wk = want.export()
    monkeypatch.setattr(N, "Model", OracleBackedModel)
    pruner = P.ModelVocabularyPruner(target, shrink_factor=0.8, em_subiters=subiters, dropout=0.0)
    vocab, report = pruner.prune(P.Vocab(list(toks), np.array(sc), np.array(kp)), blob, off)
    assert report.vocab_sizes == witers
    assert vocab.tokens == list(wt)
    assert np.array_equal(vocab.scores.view(np.uint64), np.asarray(ws).view(np.uint64))
    assert np.array_equal(vocab.keep, np.asarray(wk))
    assert len(report.rebuild_s) == len(witers) + 1 and len(report.audits) >= 1
