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


def test_selection_and_m_step_at_200k_tokens_match_oracle():
    """Row a10 at the size of the configurations (BASELINE configs[3] prunes a 250k-token vocabulary after its first
    M-step): 200 000 tokens, a 24 MB corpus for the oracle's frequency pass — the surviving tokens IN ORDER, their
    scores, the audit and the M-step's scores are the oracle's (src/prune.rs:124-170,173-319), with every host thread
    path engaged.  No GPU needed: the frequency vector is the oracle's own."""
    from tokengeex_b200 import synth
    vb, vo = synth.corpus(synth.KIND_CODE_CJK, 2, 48_000_000)
    toks, sc, kp = synth.vocab(vb, vo, 2, 200000, 16, 0.05)
    blob, off = synth.corpus(synth.KIND_CODE_CJK, 7, 24_000_000)
    keep = np.asarray(kp, np.uint8)
    om = O.OracleModel(toks, sc, keep)
    fr = om.token_frequencies(blob, off, threads=8)
    want, waudit = om.prune_vocab(blob, off, 65536, 0.8)
    wt, ws, wk = want.export()
    ids, audit = N.prune_select(toks, sc, keep, fr, len(off) - 1, 65536, 0.8, threads=8)
    assert [toks[i] for i in ids] == list(wt)
    assert np.array_equal(np.asarray(sc)[ids].view(np.uint64), np.asarray(ws).view(np.uint64))
    assert np.array_equal(audit[:7], waudit[:7])
    assert audit[4] > 50000  # tens of thousands of candidates ranked by loss
    ex = fr.astype(np.float64) * 0.37  # counts on both sides of the 0.5 threshold
    want_t, want_s, want_k = om.run_m_step(ex).export()
    kept, ns = N.m_step(ex, keep)
    idx = np.flatnonzero(kept)
    assert [toks[i] for i in idx] == list(want_t)
    assert np.array_equal(ns[idx].view(np.uint64), np.asarray(want_s).view(np.uint64))
