"""Host half of the pruning loop (csrc/prune_host.cpp) against the oracle.  CPU only."""
import random

import numpy as np

from oracle import oracle as O
from tests.util import rand_samples, rand_vocab
from tokengeex_b200 import _native as N


def test_m_step_matches_oracle_bitwise():
    rng = random.Random(1)
    for it in range(30):
        toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=rng.randrange(4, 80), max_len=4)
        keep = np.array([len(t) == 1 or rng.random() < 0.1 for t in toks], np.uint8)
        ex = np.array([rng.choice([0.0, 0.1, 0.499999, 0.5, 0.5000001, 3.0, 1e6 * rng.random()]) for _ in toks])
        om = O.OracleModel(toks, scores, keep)
        want_t, want_s, want_k = om.run_m_step(ex).export()
        kept, ns = N.m_step(ex, keep)
        idx = np.flatnonzero(kept)
        assert [toks[i] for i in idx] == want_t
        assert np.array_equal(ns[idx], want_s)  # bit for bit: same digamma, same summation order
        assert np.array_equal(keep[idx], want_k)


def test_prune_select_matches_oracle():
    rng = random.Random(2)
    for it in range(12):
        toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=rng.randrange(20, 200), max_len=5,
                                  int_scores=(it % 3 == 0))
        keep = np.array([len(t) == 1 for t in toks], np.uint8)
        samples = rand_samples(rng, b"abcd", rng.randrange(50, 300), 3, 100)
        blob, off = O.pack_samples(samples)
        om = O.OracleModel(toks, scores, keep)
        fr = om.token_frequencies(blob, off)
        target = rng.randrange(8, len(toks))
        shrink = rng.choice([0.5, 0.8, 0.95])
        try:
            want, waudit = om.prune_vocab(blob, off, target, shrink)
        except RuntimeError:
            continue  # non-normal loss: the reference panics; covered below
        wt, ws, wk = want.export()
        ids, audit = N.prune_select(toks, scores, keep, fr, len(samples), target, shrink, threads=3)
        # the same through a (host-only) model's own trie: tgx_model_prune_select
        ids_m, audit_m = N.Model(toks, scores, device=None).prune_select(toks, scores, keep, fr, len(samples), target,
                                                                        shrink, threads=2)
        assert np.array_equal(ids, ids_m) and np.array_equal(audit, audit_m)
        assert [toks[i] for i in ids] == wt
        assert np.array_equal(np.asarray(scores)[ids], ws)
        assert np.array_equal(audit[:7], waudit[:7])
        # alternatives / always_keep themselves
        ak, aoff, aids = om.token_alternatives()
        assert int(audit[0]) == int((ak == 0).sum())


def test_prune_select_and_m_step_threaded_paths_match_oracle():
    """Sizes at which the host threads engage (chunked stable sorts merged pairwise, scores of the M-step split over
    threads): still the oracle's order, ties included (src/prune.rs:124-170,247-318)."""
    rng = random.Random(11)
    for it in range(3):
        toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=1500, max_len=6, int_scores=(it == 0))
        keep = np.array([len(t) == 1 for t in toks], np.uint8)
        samples = rand_samples(rng, b"abcd", 400, 20, 200)
        blob, off = O.pack_samples(samples)
        om = O.OracleModel(toks, scores, keep)
        fr = om.token_frequencies(blob, off)
        if it == 2:
            fr = np.minimum(fr, 2)  # many equal losses
        shrink = [0.6, 0.8, 0.9][it]
        want, waudit = om.prune_vocab(blob, off, 100, shrink) if it != 2 else (None, None)
        for th in (1, 4, 7):
            ids, audit = N.prune_select(toks, scores, keep, fr, len(samples), 100, shrink, threads=th)
            if th == 1:
                ids1, audit1 = ids, audit
            assert np.array_equal(ids, ids1) and np.array_equal(audit, audit1)
        if want is not None:
            wt, ws, wk = want.export()
            assert [toks[i] for i in ids1] == wt
            assert np.array_equal(np.asarray(scores)[ids1], ws)
            assert np.array_equal(audit1[:7], waudit[:7])
        ex = np.array([rng.choice([0.0, 0.1, 0.499999, 0.5, 0.5000001, 3.0, 1e6 * rng.random()]) for _ in toks])
        want_t, want_s, want_k = om.run_m_step(ex).export()
        kept, ns = N.m_step(ex, keep)
        idx = np.flatnonzero(kept)
        assert [toks[i] for i in idx] == want_t and np.array_equal(ns[idx], want_s)


def test_vocab_packed_once_and_accepted_by_rebuild_and_selection():
    """The EM loop packs a vocabulary once (prune.Vocab.packed) and hands the pack to tgx_model_rebuild and
    tgx_model_prune_select: the same results as packing inside each call."""
    from tokengeex_b200.prune import Vocab
    rng = random.Random(3)
    toks, scores = rand_vocab(rng, alphabet=b"abc", n_tok=300, max_len=5)
    keep = np.array([len(t) == 1 for t in toks], np.uint8)
    v = Vocab(list(toks), np.asarray(scores, np.float64), keep)
    pk = v.packed()
    assert pk is v.packed()
    blob, off = N.pack(toks)
    assert np.array_equal(pk[0], blob) and np.array_equal(pk[1], off)
    fr = np.array([rng.randrange(0, 50) for _ in toks], np.uint64)
    a, b = N.Model(toks, scores, device=None), N.Model(toks[:5], scores[:5], device=None)
    b.rebuild(v.tokens, v.scores, packed=pk)
    for q in (b"abcabc", b"ccc", b"a"):
        assert a.common_prefix_search(q) == b.common_prefix_search(q)
    ids_a, aud_a = a.prune_select(v.tokens, v.scores, v.keep, fr, 100, 50, 0.8, threads=2)
    ids_b, aud_b = b.prune_select(v.tokens, v.scores, v.keep, fr, 100, 50, 0.8, threads=2, packed=pk)
    assert np.array_equal(ids_a, ids_b) and np.array_equal(aud_a, aud_b)
