"""`dropout` in (0, 1) — Model::encode's second argument (src/model.rs:59,100).

The reference draws `rand::random::<f64>()` from an unseeded thread_rng for every multi-byte candidate of a reachable
position, so only the DISTRIBUTION of its output is defined.  The product replaces the draw
by a keyed one (include/tokengeex_b200.h, tgx_model_set_dropout).  Pinned here:
  * CPU: the oracle's keyed encode against the oracle's sequential-draw encode (the reference's own loop with a
    seeded splitmix64) — identical at the deterministic ends, same mean token count in between;
  * GPU: the CUDA path against the oracle's keyed encode, bit-exact, whatever the chunking.
"""
import random

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import rand_samples, rand_vocab, split_ids, synth_setup


# ----------------------------------------------------------------------------- CPU: the keyed draw itself
def _vocab_text(seed=3):
    rng = random.Random(seed)
    toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=60, max_len=6, complete=True)
    samples = rand_samples(rng, b"abcd", 300, 1, 400)
    return toks, scores, samples


def test_keyed_draw_deterministic_ends():
    toks, scores, samples = _vocab_text()
    om = O.OracleModel(toks, scores)
    for i, s in enumerate(samples[:80]):
        assert om.encode_keyed(s, 0.0, 123, i) == om.encode(s, 0.0)
        # dropout >= 1.0: `dropout < u` never holds for u in [0, 1)  (src/model.rs:218-236)
        assert om.encode_keyed(s, 1.0, 123, i) == om.encode(s, 1.0)
        assert all(len(toks[t]) == 1 for t in om.encode_keyed(s, 1.0, 123, i))


def test_keyed_draw_reproducible_and_keyed_by_seed_and_sample():
    toks, scores, samples = _vocab_text()
    om = O.OracleModel(toks, scores)
    s = samples[0] * 8
    a = om.encode_keyed(s, 0.4, 7, 0)
    assert a == om.encode_keyed(s, 0.4, 7, 0)
    assert a != om.encode_keyed(s, 0.4, 8, 0)      # another seed
    assert a != om.encode_keyed(s, 0.4, 7, 1)      # same text as another sample of the call
    assert b"".join(toks[t] for t in a) == s       # still a segmentation of the input


@pytest.mark.parametrize("p", [0.1, 0.5, 0.9])
def test_keyed_draw_same_distribution_as_sequential_draw(p):
    """Token counts under the keyed draw and under the reference's per-candidate sequential draw agree in the mean
    (both are i.i.d. uniform draws with the keep rule `dropout < u`)."""
    toks, scores, samples = _vocab_text(seed=11)
    om = O.OracleModel(toks, scores)
    seq = sum(len(om.encode(s, p)) for s in samples)
    keyed = sum(len(om.encode_keyed(s, p, 99, i)) for i, s in enumerate(samples))
    plain = sum(len(om.encode(s, 0.0)) for s in samples)
    nbytes = sum(len(s) for s in samples)
    assert plain < seq < nbytes and plain < keyed < nbytes  # dropout lengthens the segmentation
    assert abs(seq - keyed) / seq < 0.02, (seq, keyed)


def test_keyed_draw_is_uniform():
    """Keep rate of one two-byte token that always wins when kept = 1 - dropout."""
    om = O.OracleModel([b"a", b"b", b"ab"], [-5.0, -5.0, -1.0])
    text = b"ab" * 20000
    for p in (0.25, 0.7):
        ids = om.encode_keyed(text, p, 5, 0)
        kept = sum(1 for t in ids if t == 2)
        assert abs(kept / 20000 - (1.0 - p)) < 0.015, (p, kept)


# ----------------------------------------------------------------------------- GPU: parity against the keyed oracle
@pytest.fixture(scope="module")
def N():
    from tokengeex_b200 import _native
    return _native


@pytest.mark.gpu
@pytest.mark.parametrize("algo", [0, 1])  # pair kernel with the draw in its producers / lane-group kernels
def test_gpu_dropout_small_random_vs_keyed_oracle(N, algo):
    rng = random.Random(21 + algo)
    for it in range(12):
        toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=rng.randrange(6, 50), max_len=rng.randrange(2, 9),
                                  complete=(it % 4 != 0), int_scores=(it % 2 == 0))
        gm, om = N.Model(toks, scores, device=0), O.OracleModel(toks, scores)
        gm.set_option(3, algo)
        samples = rand_samples(rng, b"abcd", 70, 0, 90) + rand_samples(rng, b"abcd", 4, 600, 3000) + [b""]
        blob, off = N.pack(samples)
        for p, seed in ((0.3, it), (0.85, 2 ** 63 + it)):
            gm.set_dropout(p, seed)
            ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off)
            got = split_ids(ids, id_off)
            first_bad = -1
            for i, s in enumerate(samples):
                try:
                    want = om.encode_keyed(s, p, seed, i)
                    assert status[i] == 0 and got[i] == want, (it, p, i)
                except O.NoPath as e:
                    assert status[i] == N.TGX_ERR_NO_PATH and got[i] == [] and (e.pos, e.length) == (len(s), len(s))
                    first_bad = i if first_bad < 0 else first_bad
            assert bad == first_bad and (rc == 0) == (first_bad < 0)
        # switched off again: the default forward kernel, the reference's dropout 0.0 result
        gm.set_dropout(0.0, 0)
        ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off)
        for i, s in enumerate(samples):
            if status[i] == 0:
                assert split_ids(ids, id_off)[i] == om.encode(s, 0.0)


@pytest.mark.gpu
def test_gpu_dropout_synth_corpus_chunked_with_crlf(N):
    """1.2 MB of synthetic code, crlf on the device, 6000-token vocabulary: the draw is keyed by the sample's index in
    the CALL and its byte position in the PROCESSED text, so chunking the host entry point changes nothing."""
    blob, off, toks, sc, kp = synth_setup(2, 31, 1_200_000, 6000, 16)
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc)
    S = len(off) - 1
    p, seed = 0.2, 0xC0FFEE
    want = [om.encode_keyed(O.crlf(blob[int(off[i]):int(off[i + 1])].tobytes()), p, seed, i) for i in range(S)]
    gm.set_dropout(p, seed)
    for algo, chunk in ((0, 1 << 30), (0, 100_000), (1, 100_000), (0, 4096), (1, 1 << 30)):
        gm.set_option(3, algo)
        gm.set_option(7, chunk)
        ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off, crlf=True)
        assert rc == 0 and not status.any()
        assert split_ids(ids, id_off) == want, (algo, chunk)
    gm.set_option(3, 0)
    lens = np.array([len(t) for t in toks])
    assert int(lens[ids].sum()) == int(plen.sum())  # still a segmentation of every processed sample
    # more tokens than without dropout, and the frequency pass keeps encoding with dropout 0.0 (src/prune.rs:218)
    freq = gm.token_frequencies(blob, off, crlf=True)[0]
    gm.set_dropout(0.0, 0)
    ids0, id_off0, *_ = gm.encode_batch(blob, off, crlf=True)
    assert ids.size > ids0.size
    assert np.array_equal(freq, np.bincount(ids0, minlength=len(toks)).astype(freq.dtype))


@pytest.mark.gpu
def test_gpu_dropout_rejects_bad_values(N):
    gm = N.Model([b"a", b"b", b"ab"], [-1.0, -1.0, -1.5], device=0)
    for bad in (-0.1, 1.0, 1.5, float("nan")):
        with pytest.raises(N.TgxError):
            gm.set_dropout(bad, 0)


@pytest.mark.gpu
def test_tokenizer_dropout_through_the_python_surface():
    import json
    import tokengeex
    rng = random.Random(8)
    toks, scores = rand_vocab(rng, alphabet=b"abcd\r\n", n_tok=80, max_len=5, complete=True)
    vocab = [{"value": t.decode(), "score": s} for t, s in zip(toks, scores)]
    t = tokengeex.Tokenizer.from_str(json.dumps({"version": "2.0", "special_tokens": ["<|x|>"],
                                                 "processors": [{"type": "crlf"}], "vocab": vocab}))
    texts = ["".join(rng.choice("abcd\r\n") for _ in range(rng.randrange(0, 300))) + "<|x|>" +
             "".join(rng.choice("abcd") for _ in range(rng.randrange(0, 50))) for _ in range(40)]
    t.dropout_seed = 42
    a = t.encode_batch(texts, 0.5)
    assert a == t.encode_batch(texts, 0.5)                       # reproducible with a seed
    plain = t.encode_batch(texts, 0.0)
    assert a != plain and sum(map(len, a)) > sum(map(len, plain))
    for x, ids in zip(texts, a):
        assert t.decode(ids, True) == x.replace("\r\n", "\n")    # property (ii) holds under dropout
    om = O.OracleModel(toks, scores)
    # pieces of a call are numbered in order; text 0 = pieces 0 ("…" before the special) and 1 (after it)
    head = O.crlf(texts[0].split("<|x|>")[0].encode())
    assert a[0][:len(om.encode_keyed(head, 0.5, 42, 0))] == om.encode_keyed(head, 0.5, 42, 0)
    t.dropout_seed = None
    runs = {tuple(t.encode(texts[0] * 4, 0.5)) for _ in range(4)}  # unseeded like the reference: calls differ
    assert len(runs) > 1
    assert t.encode(texts[0], 0.0) == plain[0]


# ----------------------------------------------------------------------------- E-step: populate_nodes(.., dropout)
def test_estep_keyed_dropout_oracle_properties():
    """src/prune.rs:87 + src/model.rs:48-50 with the keyed draw: still a distribution over segmentations (property
    (i): sum of expected * len = bytes), fewer multi-byte counts than without dropout, same multi-byte mass as the
    sequential-draw restatement of the reference loop, reproducible, and keyed by seed and byte base."""
    toks, scores, samples = _vocab_text(seed=17)
    om = O.OracleModel(toks, scores)
    from tokengeex_b200 import _native as N
    blob, off = N.pack(samples)
    lens = np.array([len(t) for t in toks], np.float64)
    nbytes = float(off[-1])
    plain = om.run_e_step(blob, off, literal=True)[0]
    multi = lambda ex: float((ex * lens)[lens > 1].sum())
    for p in (0.1, 0.5):
        keyed, rc, bad, _ = om.run_e_step_dropout(blob, off, p, seed=4, keyed=True)
        seq = om.run_e_step_dropout(blob, off, p, seed=4, keyed=False)[0]
        assert rc == 0 and bad == -1
        assert abs(float((keyed * lens).sum()) - nbytes) < 1e-9 * nbytes
        assert multi(keyed) < multi(plain)
        assert abs(multi(keyed) - multi(seq)) / multi(seq) < 0.03, (p, multi(keyed), multi(seq))
        # the keyed draw does not depend on the threading (only the order of the f64 merge does)
        np.testing.assert_allclose(keyed, om.run_e_step_dropout(blob, off, p, seed=4, keyed=True, threads=4)[0],
                                   rtol=1e-12, atol=1e-300)
        assert not np.allclose(keyed, om.run_e_step_dropout(blob, off, p, seed=5)[0], rtol=1e-6)
        assert not np.allclose(keyed, om.run_e_step_dropout(blob, off, p, seed=4, byte_base=1000)[0], rtol=1e-6)
    # sharding: two halves with their byte bases = the whole
    S = len(off) - 1
    h = S // 2
    a = om.run_e_step_dropout(blob[:int(off[h])], off[:h + 1], 0.3, seed=9)[0]
    b = om.run_e_step_dropout(blob[int(off[h]):], (off[h:] - off[h]).astype(np.uint64), 0.3, seed=9,
                              byte_base=int(off[h]))[0]
    np.testing.assert_allclose(a + b, om.run_e_step_dropout(blob, off, 0.3, seed=9)[0], rtol=1e-12, atol=1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("g", [4, 32])
def test_gpu_estep_dropout_vs_keyed_oracle(N, g):
    rng = random.Random(40 + g)
    for it in range(5):
        toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=rng.randrange(8, 60), max_len=rng.randrange(2, 9),
                                  complete=(it != 3))
        gm, om = N.Model(toks, scores, device=0), O.OracleModel(toks, scores)
        gm.set_option(2, g)
        samples = rand_samples(rng, b"abcd", 60, 1, 200) + rand_samples(rng, b"abcd", 3, 700, 2500)
        blob, off = N.pack(samples)
        for p, seed, base, snip in ((0.3, 11, 0, 81920), (0.7, 2 ** 62 + 5, 123456, 300)):
            gm.set_option(22, base)
            gm.set_dropout(p, seed)
            ex, rc, bad, badz = gm.expected_counts(blob, off, snippet_len=snip)
            want, wrc, wbad, _ = om.run_e_step_dropout(blob, off, p, seed, keyed=True, byte_base=base,
                                                       max_sample_length=snip)
            assert (rc != 0) == (wrc != 0)
            if wrc == 0:
                np.testing.assert_allclose(ex, want, rtol=1e-9, atol=1e-300)
        gm.set_dropout(0.0, 0)
        ex, rc, *_ = gm.expected_counts(blob, off)
        if rc == 0:
            np.testing.assert_allclose(ex, om.run_e_step(blob, off)[0], rtol=1e-9, atol=1e-300)


@pytest.mark.gpu
def test_gpu_estep_dropout_synth_and_pruner(N):
    blob, off, toks, sc, kp = synth_setup(2, 13, 2_000_000, 30000, 16)
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc)
    gm.set_dropout(0.01, 77)  # the reference CLI's default dropout (src/cli.rs:687)
    ex, rc, bad, badz = gm.expected_counts(blob, off)
    gm.set_dropout(0.0, 0)
    want, wrc, *_ = om.run_e_step_dropout(blob, off, 0.01, 77, keyed=True, threads=8)
    assert rc == 0 and wrc == 0
    np.testing.assert_allclose(ex, want, rtol=1e-9, atol=1e-300)
    lens = np.array([len(t) for t in toks], np.float64)
    assert abs(float((ex * lens).sum()) - float(off[-1])) < 1e-9 * float(off[-1])
    # the EM loop with dropout: reproducible with a seed (up to the atomics' last-bit jitter in the counts)
    from tokengeex_b200.prune import ModelVocabularyPruner, Vocab
    sizes = []
    for rep in range(2):
        pruner = ModelVocabularyPruner(12000, shrink_factor=0.8, em_subiters=1, dropout=0.05, dropout_seed=5)
        v, report = pruner.prune(Vocab(list(toks), np.array(sc, np.float64), np.array(kp, np.uint8)), blob, off)
        sizes.append(len(v))
        assert len(v) <= 12000 and pruner._e_steps == len(report.e_step_s) >= 1
    assert abs(sizes[0] - sizes[1]) <= 3
