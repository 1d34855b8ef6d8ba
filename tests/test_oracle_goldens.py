"""Pins the CPU oracle against every golden the reference's own tests hold for
the hot path and against an independent brute-force
enumerator.  CPU only."""
import math
import random

import numpy as np
import pytest

from oracle import oracle as O
from tests import bruteforce as BF


def mk(vocab):
    return O.OracleModel([t for t, _ in vocab], [s for _, s in vocab])


# --- /root/reference/src/model.rs:209-215  test_encode -----------------------
def test_ref_model_test_encode():
    m = mk([(b"a", -3.0), (b"b", -3.0), (b"c", -3.0), (b"ab", -4.0)])
    assert m.encode(b"abc", 0.0) == [3, 2]


# --- src/model.rs:218-236  test_encode_dropout (dropout = 1.0 is deterministic)
def test_ref_model_test_encode_dropout():
    m = mk([(b"a", -3.0), (b"b", -3.0), (b"c", -3.0), (b"d", -3.0), (b"e", -3.0), (b"f", -3.0),
            (b"ab", -4.0), (b"abc", -5.0), (b"abcd", -6.0), (b"abcde", -7.0), (b"abcdef", -8.0)])
    assert m.encode(b"abcdef", 1.0) == [0, 1, 2, 3, 4, 5]
    assert m.encode(b"abcdef", 0.0) == [10]


# --- src/model.rs:243-252  test_decode_encode_invariants ----------------------
def test_ref_model_test_decode_encode_invariants():
    toks = [bytes([i]) for i in range(256)]  # src/lib.rs:206-210 new_default_vocab
    m = O.OracleModel(toks, [1.0 / 256.0] * 256)
    text = "你好，我叫罗杰斯".encode()
    ids = m.encode(text, 0.0)
    assert len(ids) == len(text)
    assert b"".join(toks[i] for i in ids) == text


# --- tie rule H2 / Q3 (SURVEY: vocab a,b,c:-3; ab,bc:-4 on "abc" -> [0,4]) ----
def test_tie_smallest_start_wins():
    m = mk([(b"a", -3.0), (b"b", -3.0), (b"c", -3.0), (b"ab", -4.0), (b"bc", -4.0)])
    assert m.encode(b"abc", 0.0) == [0, 4]


def test_empty_and_nopath():
    m = mk([(b"a", -1.0), (b"ab", -1.5)])
    assert m.encode(b"", 0.0) == []  # Q5
    with pytest.raises(O.NoPath) as ei:
        m.encode(b"abx", 0.0)
    assert (ei.value.pos, ei.value.length) == (3, 3)
    assert str(ei.value) == "no path to position 3/3"  # src/lib.rs:243-245
    # unreachable middle position is skipped: "b" alone is not a token
    assert m.encode(b"ab", 0.0) == [1]


def test_duplicate_token_last_id_wins():  # Q1
    m = mk([(b"ab", -1.0), (b"a", -5.0), (b"b", -5.0), (b"ab", -2.0)])
    assert m.encode(b"ab", 0.0) == [3]
    assert m.common_prefix_search(b"abc") == ([1, 3], [1, 2])


def test_common_prefix_search_stops_at_missing_edge():  # Q2
    m = mk([(b"a", -1.0), (b"abc", -1.0), (b"abcde", -1.0)])
    assert m.common_prefix_search(b"abcdx")[0] == [0, 1]
    assert m.common_prefix_search(b"abcde")[0] == [0, 1, 2]
    assert m.common_prefix_search(b"xbc")[0] == []


# --- src/lattice.rs:403-474 (commented-out test_lattice) ----------------------
LATTICE_VOCAB = [(b"<", -3.0), (b" value", -6.0), (b">", -3.0), (b"DC value", -8.0), (b"<DC", -4.0),
                 (b"<DC value>", -12.0)]


@pytest.mark.parametrize("literal", [True, False])
def test_ref_lattice_marginals(literal):
    m = mk(LATTICE_VOCAB)
    z, ex = m.marginal(b"<DC value>", literal=literal)
    # values in the reference's comments, src/lattice.rs:447-452 (6 dp)
    want = {b"<DC value>": 0.665241, b">": 0.334759, b"<DC": 0.244728, b" value": 0.244728,
            b"<": 0.090031, b"DC value": 0.090031}
    for (tok, _), e in zip(LATTICE_VOCAB, ex):
        assert abs(e - want[tok]) < 5e-7, (tok, e)
    # Z = e^-12 + e^-13 + e^-14 (hand-derived)
    assert abs(z - math.log(math.exp(-12) + math.exp(-13) + math.exp(-14))) < 1e-12
    assert abs(z - (-11.59239403555562)) < 1e-12


def test_q7_unreachable_positions_keep_log1():
    m = mk([(b"ab", -1.0), (b"c", -2.0), (b"bc", -3.0)])
    for literal in (True, False):
        z, ex = m.marginal(b"abc", literal=literal)
        assert abs(z - (-2.3068528194400546)) < 1e-15
        assert np.allclose(ex, [0.5, 0.5, 0.5], atol=1e-15)


def test_q21_lattice_viterbi_includes_eos():
    m = mk(LATTICE_VOCAB)
    assert m.nbest(b"<DC value>", 1) == [[5, 4294967295]]
    nb = m.nbest(b"<DC value>", 10)
    assert nb[0] == [5]
    assert sorted(map(tuple, nb)) == sorted([(5,), (4, 1, 2), (0, 3, 2)])
    assert nb == [[5], [4, 1, 2], [0, 3, 2]]  # -12, -13, -14


def test_log_sum_exp_q8():
    assert O.log_sum_exp(123.0, -7.0, True) == -7.0
    assert O.log_sum_exp(0.0, -51.0, False) == 0.0  # cutoff 50
    assert O.log_sum_exp(-51.0, 0.0, False) == 0.0
    x = O.log_sum_exp(-1.0, -2.0, False)
    assert x == -1.0 + math.log(math.exp(-1.0) + 1.0)


def test_crlf_q16():
    assert O.crlf(b"a\r\nb") == b"a\nb"
    assert O.crlf(b"\r\r\n") == b"\r\n"
    assert O.crlf(b"\r") == b"\r"
    assert O.crlf(b"\n\r") == b"\n\r"
    assert O.crlf(b"\r\n\r\n\r") == b"\n\n\r"
    assert O.crlf(b"") == b""
    rng = random.Random(0)
    for _ in range(200):
        s = bytes(rng.choice(b"\r\nab") for _ in range(rng.randrange(0, 40)))
        assert O.crlf(s) == s.replace(b"\r\n", b"\n")


# --- literal lattice == per-position form, bit for bit -----------------------
def rand_vocab(rng, alphabet=b"abc", n_tok=12, max_len=4, complete=True):
    toks = set()
    if complete:
        toks |= {bytes([c]) for c in alphabet}
    while len(toks) < n_tok:
        toks.add(bytes(rng.choice(alphabet) for _ in range(rng.randrange(1, max_len + 1))))
    toks = sorted(toks)
    rng.shuffle(toks)
    scores = [-(rng.random() * 6 + 0.5) for _ in toks]
    return toks, scores


def test_per_position_equals_literal_bitwise():
    rng = random.Random(1)
    for it in range(60):
        toks, scores = rand_vocab(rng, complete=(it % 3 != 0))
        m = O.OracleModel(toks, scores)
        text = bytes(rng.choice(b"abc") for _ in range(rng.randrange(1, 60)))
        z1, e1 = m.marginal(text, literal=True)
        z2, e2 = m.marginal(text, literal=False)
        assert (z1 == z2) or (math.isnan(z1) and math.isnan(z2))
        assert np.array_equal(e1, e2)


# --- brute force: Viterbi with first-wins ties, and marginals ------------------
def test_bruteforce_viterbi():
    rng = random.Random(2)
    for it in range(300):
        toks, scores = rand_vocab(rng, n_tok=rng.randrange(4, 14), complete=(it % 4 != 0))
        if it % 2 == 0:  # force many exact ties (H2): scores from a tiny integer set
            scores = [-float(rng.randrange(2, 5)) for _ in toks]
        m = O.OracleModel(toks, scores)
        text = bytes(rng.choice(b"abc") for _ in range(rng.randrange(0, 14)))
        want = BF.viterbi_first_wins(text, toks, scores)
        if want is None:
            with pytest.raises(O.NoPath):
                m.encode(text, 0.0)
        else:
            got = m.encode(text, 0.0)
            assert got == want
            # property (iii): no enumerated path scores strictly higher
            best = BF.path_score_left_to_right(got, scores)
            for p in BF.all_paths(text, toks):
                assert BF.path_score_left_to_right(p, scores) <= best + 1e-12


def test_bruteforce_marginals():
    rng = random.Random(3)
    for it in range(150):
        toks, scores = rand_vocab(rng, n_tok=rng.randrange(4, 12), complete=True)
        m = O.OracleModel(toks, scores)
        text = bytes(rng.choice(b"abc") for _ in range(rng.randrange(1, 13)))
        logz, ex = BF.marginals(text, toks, scores)
        z, got = m.marginal(text, literal=True)
        assert abs(z - logz) < 1e-12
        assert np.allclose(got, ex, rtol=1e-11, atol=1e-14)
        # invariant (i): sum_id expected*len == snippet length (byte-complete vocab)
        tot = sum(e * len(t) for e, t in zip(got, toks))
        assert abs(tot - len(text)) < 1e-10


def test_bruteforce_nbest2():
    rng = random.Random(4)
    for it in range(150):
        toks, scores = rand_vocab(rng, n_tok=rng.randrange(4, 12), complete=True)
        m = O.OracleModel(toks, scores)
        text = bytes(rng.choice(b"abc") for _ in range(rng.randrange(1, 9)))
        paths = BF.all_paths(text, toks)
        ranked = sorted(paths, key=lambda p: -math.fsum(scores[t] for t in p))
        nb = m.nbest(text, 2)
        assert len(nb) == min(2, len(paths))
        for k, p in enumerate(nb):
            s_got = math.fsum(scores[t] for t in p)
            s_want = math.fsum(scores[t] for t in ranked[k])
            assert abs(s_got - s_want) < 1e-12  # the k-th best score (ties: any order)


# --- prune.rs pieces ----------------------------------------------------------
def test_digamma_matches_asymptotic_reference():
    # src/prune.rs:322-335.  Cross-check against scipy's digamma (the series is
    # accurate to ~1e-12 after the x>=7 shift); exact op order is restated.
    from scipy.special import digamma as sp
    for x in [0.5, 1.0, 2.5, 6.999, 7.0, 10.0, 100.0, 111111.0, 1e9]:
        assert abs(O.digamma(x) - sp(x)) < 1e-10 * max(1.0, abs(sp(x)))


def test_m_step_q13():
    m = O.OracleModel([b"a", b"b", b"ab", b"ba"], [-1.0, -1.0, -2.0, -2.0], keep=[1, 0, 0, 0])
    ex = np.array([0.1, 0.49999, 3.0, 0.5])
    m2 = m.run_m_step(ex)
    toks, sc, kp = m2.export()
    assert toks == [b"a", b"ab", b"ba"]  # b dropped (freq<0.5, !keep); a kept by keep with freq clamped
    s = 0.5 + 3.0 + 0.5
    for t, got, f in zip(toks, sc, [0.5, 3.0, 0.5]):
        assert got == O.digamma(f) - O.digamma(s)
    assert kp.tolist() == [1, 0, 0]


def test_e_step_snippets_q10():
    # tokens never span snippet boundaries: with max_sample_length=2, "abab"
    # becomes "ab","ab" and the token "ba" can never be counted.
    m = O.OracleModel([b"a", b"b", b"ab", b"ba"], [-1.0, -1.0, -1.5, -1.5])
    blob, off = O.pack_samples([b"abab"])
    ex, rc, bad, _ = m.run_e_step(blob, off, max_sample_length=2)
    assert rc == 0 and bad == -1
    assert ex[3] == 0.0
    z, e1 = m.marginal(b"ab")
    assert np.allclose(ex, 2 * e1, rtol=0, atol=1e-15)
    # literal and per-position agree bitwise over a batch, threads or not
    rng = random.Random(5)
    samples = [bytes(rng.choice(b"ab") for _ in range(rng.randrange(1, 50))) for _ in range(40)]
    blob, off = O.pack_samples(samples)
    e_lit, _, _, _ = m.run_e_step(blob, off, threads=1, literal=True, max_sample_length=16)
    e_pp, _, _, _ = m.run_e_step(blob, off, threads=1, literal=False, max_sample_length=16)
    e_mt, _, _, _ = m.run_e_step(blob, off, threads=4, literal=False, max_sample_length=16)
    assert np.array_equal(e_lit, e_pp)
    assert np.allclose(e_pp, e_mt, rtol=1e-14)


def test_e_step_bad_z_q11():
    m = O.OracleModel([b"a"], [-1.0])
    blob, off = O.pack_samples([b"aa", b"ab", b"a"])
    ex, rc, bad, badz = m.run_e_step(blob, off)
    assert rc == 1 and bad == 1 and badz == 0.0  # z == 0.0 is not "normal"


def test_token_frequencies_and_prune_vocab_smoke():
    rng = random.Random(6)
    toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=40, max_len=4, complete=True)
    keep = [len(t) == 1 for t in toks]
    m = O.OracleModel(toks, scores, keep)
    samples = [bytes(rng.choice(b"abcd") for _ in range(rng.randrange(5, 80))) for _ in range(200)]
    blob, off = O.pack_samples(samples)
    fr = m.token_frequencies(blob, off, threads=2)
    want = np.zeros(len(toks), np.uint64)
    for s in samples:
        for i in m.encode(s):
            want[i] += 1
    assert np.array_equal(fr, want)
    assert sum(int(f) * len(t) for f, t in zip(fr, toks)) == sum(map(len, samples))  # invariant (iv)
    m2, audit = m.prune_vocab(blob, off, target=20, shrink=0.8)
    t2, s2, k2 = m2.export()
    assert set(t for t, k in zip(toks, keep) if k) <= set(t2)
    assert len(t2) <= max(int(len(toks) * 0.8), 20) or audit[4] == 0
    assert all(s2[i] >= s2[i + 1] for i in range(len(s2) - 1))


def test_full_prune_loop_runs_q20_q22():
    rng = random.Random(7)
    toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=120, max_len=5, complete=True)
    keep = [len(t) == 1 for t in toks]
    m = O.OracleModel(toks, scores, keep)
    samples = [bytes(rng.choice(b"abcd") for _ in range(rng.randrange(5, 120))) for _ in range(300)]
    blob, off = O.pack_samples(samples)
    m2, iters = m.prune(blob, off, vocab_size=30, shrink=0.8, em_subiters=2, threads=2)
    assert m2.V <= 30 or m2.V < 120
    assert iters[-1] == m2.V
    assert all(a >= b for a, b in zip(iters, iters[1:]))


def test_pair_frequencies_brute_force():
    """src/merge.rs:53-84: adjacent id pairs inside each sample, counted over a batch, frequency-descending."""
    import collections
    import random

    from tests.util import rand_samples, rand_vocab
    rng = random.Random(9)
    for it in range(10):
        toks, scores = rand_vocab(rng, alphabet=b"abc", n_tok=rng.randrange(4, 40), max_len=5)
        om = O.OracleModel(toks, scores)
        samples = rand_samples(rng, b"abc", 60, 0, 80)
        want = collections.Counter()
        for s in samples:
            ids = om.encode(s)
            for i in range(1, len(ids)):
                want[(ids[i - 1], ids[i])] += 1
        blob = np.frombuffer(b"".join(samples) or b"\0", np.uint8).copy()
        off = np.zeros(len(samples) + 1, np.uint64)
        off[1:] = np.cumsum([len(s) for s in samples])
        ab, cnt = om.pair_frequencies(blob, off, threads=3)
        got = {(int(a), int(b)): int(c) for (a, b), c in zip(ab.tolist(), cnt.tolist())}
        assert got == dict(want)
        order = sorted(want.items(), key=lambda kv: (-kv[1], kv[0]))
        assert [list(k) for k, _ in order] == ab.tolist()
