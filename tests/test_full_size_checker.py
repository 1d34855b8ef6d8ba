"""The numpy checker of tests/test_gpu_full_size.py, exercised on the CPU with the oracle's output: it accepts a
correct segmentation and rejects corrupted ids, offsets and lengths."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_full_size import check_segmentation
from tests.util import synth_setup


def test_checker_accepts_oracle_output_and_rejects_corruption():
    from tokengeex_b200 import _native as N
    blob, off, toks, sc, kp = synth_setup(2, 31, 1_200_000, 6000, 16)
    om = O.OracleModel(toks, sc)
    ids, id_off, status, plen, bad = om.encode_batch(blob, off, crlf=True, threads=4)
    samples = [O.crlf(blob[int(off[i]):int(off[i + 1])].tobytes()) for i in range(len(off) - 1)]
    pblob, poff = N.pack(samples)
    assert int(poff[-1]) < int(off[-1])  # the corpus has crlf line ends
    check_segmentation(blob, off, toks, ids, id_off, plen, pblob, poff)
    lens = np.array([len(t) for t in toks])
    k = int(np.flatnonzero(lens[ids] > 2)[1000])
    other_len = ids.copy()
    other_len[k] = int(np.flatnonzero(lens == lens[ids[k]] - 1)[0])
    with pytest.raises(AssertionError):
        check_segmentation(blob, off, toks, other_len, id_off, plen, pblob, poff)
    same_len = ids.copy()
    tk = toks[ids[k]]
    same_len[k] = next(i for i, t in enumerate(toks) if len(t) == len(tk) and t[0] != tk[0])
    with pytest.raises(AssertionError):
        check_segmentation(blob, off, toks, same_len, id_off, plen, pblob, poff)
    bad_plen = plen.copy()
    bad_plen[3] += 1
    with pytest.raises(AssertionError):
        check_segmentation(blob, off, toks, ids, id_off, bad_plen, pblob, poff)
