"""`tokengeex.Tokenizer` end to end on the GPU: special tokens + processors + Model::encode."""
import json
import random

import pytest

from oracle import oracle as O
from tests.util import rand_vocab

pytestmark = pytest.mark.gpu


def make(toks, scores, specials, processors):
    import tokengeex
    vocab = [{"value": t.decode(), "score": s} for t, s in zip(toks, scores)]
    js = json.dumps({"version": "2.0", "special_tokens": specials, "processors": processors, "vocab": vocab})
    return tokengeex.Tokenizer.from_str(js)


def test_reference_example_flow():
    import tokengeex
    # src/model.rs:209-215 through the public Python surface
    t = make([b"a", b"b", b"c", b"ab"], [-3.0, -3.0, -3.0, -4.0], [], [])
    assert t.encode("abc", 0.0) == [3, 2]
    assert t.encode_ordinary("abc", 0.0) == [3, 2]
    assert t.decode(t.encode("abcab", 0.0), True) == "abcab"
    # src/model.rs:218-236: dropout = 1.0 keeps single-byte tokens only
    t = make([b"a", b"b", b"c", b"d", b"e", b"f", b"ab", b"abc", b"abcd", b"abcde", b"abcdef"],
             [-3.0] * 6 + [-4.0, -5.0, -6.0, -7.0, -8.0], [], [])
    assert t.encode("abcdef", 1.0) == [0, 1, 2, 3, 4, 5]
    assert t.encode("abcdef", 0.0) == [10]
    # 0 < dropout < 1: a keyed draw per multi-byte candidate (tests/test_dropout.py); always a segmentation
    assert t.decode(t.encode("abcdef", 0.5), True) == "abcdef"
    with pytest.raises(tokengeex.TokenGeeXError) as ei:
        t.encode("abx", 0.0)
    assert str(ei.value) == "no path to position 3/3"  # Display of Error::NoPath (src/lib.rs:243-245)


def test_specials_and_processors_vs_oracle():
    rng = random.Random(5)
    toks, scores = rand_vocab(rng, alphabet=b"abc\r\n", n_tok=60, max_len=5)
    specials = ["<EOS>", "ab", "<EOS_2>"]  # "ab" is also a plain substring: special wins (src/tokenizer.rs:315-347)
    t = make(toks, scores, specials, [{"type": "crlf"}])
    om = O.OracleModel(toks, scores)
    V = len(toks)
    texts = []
    for _ in range(120):
        parts = []
        for _ in range(rng.randrange(0, 6)):
            parts.append(rng.choice(specials + ["".join(rng.choice("abc\r\n") for _ in range(rng.randrange(0, 30)))]))
        texts.append("".join(parts))
    got = t.encode_batch(texts, 0.0)
    from tokengeex_b200.tokenizer import split_special_tokens
    for text, ids in zip(texts, got):
        want = []
        for sub, sp in split_special_tokens(text, specials):
            if sp:
                want.append(V + specials.index(sub))
            else:
                want.extend(om.encode(O.crlf(sub.encode()), 0.0))
        assert ids == want, text
        assert t.encode(text, 0.0) == want
        assert t.decode(ids, True) == text.replace("\r\n", "\n") or "ab" in text or "\r\n" in text
    # encode_ordinary ignores specials
    assert t.encode_ordinary_batch(["ab<EOS>"[:2]], 0.0) == [om.encode(b"ab", 0.0)]
