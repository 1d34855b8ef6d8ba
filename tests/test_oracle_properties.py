"""Property tests of the oracle with `hypothesis` (SURVEY §7 step 1, Appendix A invariants): they hold in the
reference's arithmetic for ANY vocabulary that contains every byte of the text, so they pin the restatement
independently of recorded outputs.  The same invariants are checked on the CUDA path at full size
(tests/test_gpu_full_size.py, bench.py `prune_iter.property_*`)."""
import math

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import oracle as O  # noqa: E402

ALPHABET = b"ab\r\n"


@st.composite
def vocab_and_text(draw):
    extra = draw(st.lists(st.binary(min_size=2, max_size=5).map(lambda b: bytes(ALPHABET[x % 4] for x in b)),
                          min_size=0, max_size=12, unique=True))
    toks = [bytes([c]) for c in ALPHABET] + extra
    order = draw(st.permutations(range(len(toks))))
    toks = [toks[i] for i in order]
    scores = draw(st.lists(st.one_of(st.integers(-6, -1).map(float), st.floats(-9.0, -0.25)),
                           min_size=len(toks), max_size=len(toks)))
    text = bytes(ALPHABET[x % 4] for x in draw(st.binary(min_size=0, max_size=40)))
    return toks, scores, text


@settings(max_examples=300, deadline=None)
@given(vocab_and_text())
def test_encode_is_a_segmentation_and_optimal_over_single_bytes(vt):
    toks, scores, text = vt
    om = O.OracleModel(toks, scores)
    ids = om.encode(text, 0.0)
    assert b"".join(toks[i] for i in ids) == text                      # (ii) decode(encode(x)) == x
    best = math.fsum(scores[i] for i in ids)
    single = {t: s for t, s in zip(toks, scores) if len(t) == 1}        # (duplicates: any byte path is a path)
    assert best >= math.fsum(single[bytes([c])] for c in text) - 1e-9  # (iii) no worse than the all-bytes path
    assert om.encode(O.crlf(text), 0.0) == om.encode(text.replace(b"\r\n", b"\n"), 0.0)  # Q16


@settings(max_examples=200, deadline=None)
@given(vocab_and_text())
def test_marginals_cover_every_byte_once(vt):
    toks, scores, text = vt
    if not text:
        return
    om = O.OracleModel(toks, scores)
    z, ex = om.marginal(text, literal=True)
    zp, exp_ = om.marginal(text, literal=False)
    assert z == zp and np.array_equal(ex, exp_)                        # per-position form == literal form, bitwise
    assert np.all(ex >= 0.0) and math.isfinite(z)
    lens = np.array([len(t) for t in toks], np.float64)
    assert abs(float(ex @ lens) - len(text)) <= 1e-9 * len(text)        # (i) sum expected * len == bytes
    ids = om.encode(text, 0.0)
    assert z >= math.fsum(scores[i] for i in ids) - 1e-9                # log-sum over paths >= the best path
