"""tests/golden/hotpath_small.json (written by tools/make_golden.py) against the oracle (CPU) and the CUDA path (GPU)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

FX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_small.json")


@pytest.fixture(scope="module")
def fx():
    with open(FX) as f:
        d = json.load(f)
    d["toks"] = [bytes.fromhex(v["hex"]) for v in d["vocab"]]
    d["sc"] = np.array([float.fromhex(v["score"]) for v in d["vocab"]])
    d["kp"] = np.array([v["keep"] for v in d["vocab"]], np.uint8)
    d["samples"] = [bytes.fromhex(s) for s in d["samples_hex"]]
    d["proc"] = [bytes.fromhex(s) for s in d["prune_samples_hex"]]
    d["ex"] = np.array([float.fromhex(x) for x in d["expected_counts"]])
    d["ex64"] = np.array([float.fromhex(x) for x in d["expected_counts_snippet64"]])
    return d


def test_oracle_reproduces_fixture(fx):
    om = O.OracleModel(fx["toks"], fx["sc"], fx["kp"])
    blob, off = O.pack_samples(fx["samples"])
    ids, id_off, status, plen, bad = om.encode_batch(blob, off, crlf=True, threads=2)
    assert ids.tolist() == fx["encode_crlf"]["ids"] and id_off.tolist() == fx["encode_crlf"]["id_off"]
    assert status.tolist() == fx["encode_crlf"]["status"] and plen.tolist() == fx["encode_crlf"]["proc_len"]
    ids, id_off, status, _, _ = om.encode_batch(blob, off, crlf=False, threads=1)
    assert ids.tolist() == fx["encode_raw"]["ids"] and id_off.tolist() == fx["encode_raw"]["id_off"]
    pblob, poff = O.pack_samples(fx["proc"])
    ex, rc, _, _ = om.run_e_step(pblob, poff, threads=1, literal=True)
    assert rc == 0 and np.array_equal(ex, fx["ex"])  # single thread: bit for bit
    ex, rc, _, _ = om.run_e_step(pblob, poff, threads=1, literal=True, max_sample_length=64)
    assert rc == 0 and np.array_equal(ex, fx["ex64"])
    assert om.token_frequencies(pblob, poff, threads=3).tolist() == fx["token_frequencies"]
    mt, ms, mk = om.run_m_step(fx["ex"]).export()
    assert [t.hex() for t in mt] == fx["m_step"]["tokens_hex"]
    assert [float(x).hex() for x in ms] == fx["m_step"]["scores"]
    pv, _ = om.prune_vocab(pblob, poff, fx["prune_vocab"]["target"], fx["prune_vocab"]["shrink"], threads=2)
    pt, ps, pk = pv.export()
    assert [t.hex() for t in pt] == fx["prune_vocab"]["tokens_hex"]


def test_fixture_invariants(fx):
    """Size-independent properties the domain offers (SURVEY appendix A invariants ii and iv)."""
    lens = np.array([len(t) for t in fx["toks"]])
    ids = np.array(fx["encode_crlf"]["ids"])
    off = fx["encode_crlf"]["id_off"]
    for i, s in enumerate(fx["samples"]):
        if fx["encode_crlf"]["status"][i] == 0:
            dec = b"".join(fx["toks"][t] for t in ids[off[i]:off[i + 1]])
            assert dec == s.replace(b"\r\n", b"\n")
    assert int((np.array(fx["token_frequencies"]) * lens).sum()) == sum(len(s) for s in fx["proc"])
    assert abs(float((fx["ex"] * lens).sum()) - sum(len(s) for s in fx["proc"])) < 1e-6


@pytest.mark.gpu
def test_cuda_reproduces_fixture(fx):
    from tokengeex_b200 import _native as N
    gm = N.Model(fx["toks"], fx["sc"], device=0)
    blob, off = N.pack(fx["samples"])
    for algo in (0, 2, 3, 1):  # match rows / pair-CTA kernel / lane teams / lane groups
        gm.set_option(3, algo)
        ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off, crlf=True)
        assert ids.tolist() == fx["encode_crlf"]["ids"] and id_off.tolist() == fx["encode_crlf"]["id_off"]
        assert status.tolist() == fx["encode_crlf"]["status"] and plen.tolist() == fx["encode_crlf"]["proc_len"]
        ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off, crlf=False)
        assert ids.tolist() == fx["encode_raw"]["ids"] and id_off.tolist() == fx["encode_raw"]["id_off"]
    gm.set_option(3, 0)
    pblob, poff = N.pack(fx["proc"])
    ex, rc, _, _ = gm.expected_counts(pblob, poff)
    from tests.util import counts_rel_err
    assert rc == 0 and counts_rel_err(ex, fx["ex"]) < 1e-9  # north_star tolerance
    ex, rc, _, _ = gm.expected_counts(pblob, poff, snippet_len=64)
    assert rc == 0 and counts_rel_err(ex, fx["ex64"]) < 1e-9
    fr, rc, _, _ = gm.token_frequencies(pblob, poff)
    assert rc == 0 and fr.tolist() == fx["token_frequencies"]
    kept, ns = N.m_step(fx["ex"], fx["kp"])
    idx = np.flatnonzero(kept)
    assert [fx["toks"][i].hex() for i in idx] == fx["m_step"]["tokens_hex"]
    assert [float(x).hex() for x in ns[idx]] == fx["m_step"]["scores"]
