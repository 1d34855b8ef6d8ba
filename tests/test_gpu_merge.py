"""Pair-frequency pass of `tokengeex merge` (src/merge.rs:36-84) on the GPU vs the oracle: exact integer counts, same
order (count descending; ties — undefined in the reference's sort_unstable — by ids ascending in both)."""
import random

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import rand_samples, rand_vocab, synth_setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N():
    from tokengeex_b200 import _native
    return _native


def test_pair_frequencies_small(N):
    toks, sc = [b"a", b"b", b"ab", b"c"], [-1.0, -1.0, -1.5, -2.0]
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc)
    blob, off = N.pack([b"abab", b"abc", b"", b"c", b"ba"])
    ab, cnt, rc, bad, blen = gm.pair_frequencies(blob, off)
    wab, wcnt = om.pair_frequencies(blob, off)
    assert rc == 0 and ab.tolist() == wab.tolist() == [[1, 0], [2, 2], [2, 3]] and cnt.tolist() == wcnt.tolist() == [1, 1, 1]
    # nothing to count: empty batch, single-token samples
    blob, off = N.pack([b"c", b"", b"ab"])
    ab, cnt, rc, bad, blen = gm.pair_frequencies(blob, off)
    assert rc == 0 and len(cnt) == 0
    # NoPath surfaces like encode (the reference unwraps the error)
    blob, off = N.pack([b"ab", b"axb"])
    ab, cnt, rc, bad, blen = gm.pair_frequencies(blob, off)
    assert rc == N.TGX_ERR_NO_PATH and bad == 1 and blen == 3


def test_pair_frequencies_random_vs_oracle(N):
    rng = random.Random(41)
    for it in range(15):
        toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=rng.randrange(5, 120), max_len=rng.randrange(1, 9),
                                  int_scores=(it % 2 == 0))
        gm, om = N.Model(toks, scores, device=0), O.OracleModel(toks, scores)
        samples = rand_samples(rng, b"abcd", rng.randrange(1, 200), 0, 300)
        blob, off = N.pack(samples)
        ab, cnt, rc, bad, blen = gm.pair_frequencies(blob, off, cap=(3 if it % 5 == 0 else 0))  # cap 3: capacity retry
        wab, wcnt = om.pair_frequencies(blob, off, threads=4)
        assert rc == 0 and np.array_equal(ab, wab) and np.array_equal(cnt, wcnt), it


def test_pair_frequencies_synth_corpus(N):
    blob, off, toks, sc, kp = synth_setup(3, 17, 3_000_000, 20000, 16)
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc)
    ab, cnt, rc, bad, blen = gm.pair_frequencies(blob, off, crlf=False)
    wab, wcnt = om.pair_frequencies(blob, off, threads=8)
    assert rc == 0 and np.array_equal(cnt, wcnt) and np.array_equal(ab, wab)
    ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off)
    n_tok = np.diff(id_off.astype(np.int64))
    assert int(cnt.sum()) == int(np.maximum(n_tok - 1, 0).sum())  # every adjacent pair counted once
