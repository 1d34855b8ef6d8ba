"""Weak secondary cross-check of the oracle: HuggingFace `tokenizers`' Unigram model is the upstream the
reference's Viterbi was "imported and modified from" (src/model.rs:1-2, src/lattice.rs:1-2).  It steps over chars and
has unk handling, so it is not an oracle — but on text whose every char is in the vocabulary its best segmentation,
including the choice among exactly tied paths, must be the reference's (smallest start wins)."""
import random

import pytest

from oracle import oracle as O

tokenizers = pytest.importorskip("tokenizers")


def _hf(toks, scores):
    from tokenizers import Tokenizer, models
    # the unk token is a char that never occurs (HF fuses consecutive unk ids)
    return Tokenizer(models.Unigram([(t.decode(), s) for t, s in zip(toks, scores)] + [("⁇", -100.0)],
                                    unk_id=len(toks), byte_fallback=False))


def _vocab(rng, chars, n_tok, max_len, int_scores):
    toks = set(chars)
    n_tok = min(n_tok, sum(len(chars) ** l for l in range(1, max_len + 1)))
    while len(toks) < n_tok:
        toks.add("".join(rng.choice(chars) for _ in range(rng.randrange(1, max_len + 1))))
    toks = sorted(toks)
    rng.shuffle(toks)
    scores = [-float(rng.randrange(2, 6)) if int_scores else -(rng.random() * 6 + 0.5) for _ in toks]
    return [t.encode() for t in toks], scores


@pytest.mark.parametrize("chars", ["abcd", "你好世界"])
def test_oracle_encode_equals_hf_unigram(chars):
    rng = random.Random(len(chars[0].encode()))
    n = 0
    for it in range(30):
        toks, scores = _vocab(rng, chars, rng.randrange(6, 60), rng.randrange(2, 6), int_scores=(it % 2 == 0))
        if max(len(t) for t in toks) > 64:
            continue
        om, hf = O.OracleModel(toks, scores), _hf(toks, scores)
        for _ in range(40):
            s = "".join(rng.choice(chars) for _ in range(rng.randrange(1, 60)))
            assert om.encode(s.encode(), 0.0) == hf.encode(s).ids, (it, s)
            n += 1
    assert n >= 1000
