"""BASELINE.json configurations at their full sizes.

configs[0] (64 MB code corpus, 32k vocabulary made with the generate defaults: max token length 24, crlf) is small
enough for the oracle to encode all of it: bit-exact ids.  configs[1] (1 GB multi-language corpus, 131k vocabulary)
is checked through size-independent properties ((ii) decode(encode(x)) = crlf(x), (iv) sum of
freq * len = bytes), agreement of independent device paths (emit vs histogram, chunked vs unchunked host entry
point), and the oracle on a seeded sample of the corpus that includes its longest sample.
"""
import random

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N():
    from tokengeex_b200 import _native
    return _native


def test_config0_64mb_32k_vocab_bit_exact(N):
    from tokengeex_b200 import synth
    blob, off = synth.corpus(synth.KIND_CODE, 1, 64_000_000)
    toks, sc, kp = synth.vocab(blob, off, 1, 32768, 24, 0.05)  # src/cli.rs:675 default --max-token-length 24
    assert max(map(len, toks)) > 16  # the lane-group forward kernels, whatever the options
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc)
    wids, wid_off, wstatus, wplen, wbad = om.encode_batch(blob, off, crlf=True, threads=16)
    ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off, crlf=True)
    assert rc == 0 and not wstatus.any()
    assert np.array_equal(id_off, wid_off) and np.array_equal(ids, wids) and np.array_equal(plen, wplen)
    fr, rc, *_ = gm.token_frequencies(blob, off, crlf=True)
    assert rc == 0 and np.array_equal(fr, om.token_frequencies(*_crlf_corpus(blob, off), threads=16))


def check_segmentation(blob, off, toks, ids, id_off, plen, pblob, poff):
    """Size-independent checks that `ids` segment crlf(sample) for every sample (property (ii)), in numpy only."""
    NB = int(off[-1])
    lens = np.array([len(t) for t in toks], np.uint32)
    # crlf: every "\r\n" inside a sample loses one byte (Q16) — counted independently on the host
    cr = np.flatnonzero(blob[:max(NB - 1, 0)] == 13)
    n_crlf = int((blob[cr + 1] == 10).sum())
    ends = off[1:-1].astype(np.int64)
    ends = ends[(ends > 0) & (ends < NB)]
    n_crlf -= int(((blob[ends - 1] == 13) & (blob[ends] == 10)).sum())  # pairs that straddle two samples
    assert int(plen.sum()) == NB - n_crlf == int(poff[-1])
    assert np.array_equal(np.diff(poff.astype(np.int64)), plen.astype(np.int64))
    # token lengths add up to the processed length of every sample ...
    io = id_off.astype(np.int64)
    csum = np.zeros(ids.size + 1, np.uint32)  # < 2^32 bytes per call
    np.cumsum(lens[ids], dtype=np.uint32, out=csum[1:])
    assert np.array_equal((csum[io[1:]] - csum[io[:-1]]).astype(np.int64), plen.astype(np.int64))
    # ... and token k of a sample starts where the processed text has the token's first (and last) byte
    first = np.array([t[0] for t in toks], np.uint8)
    last = np.array([t[-1] for t in toks], np.uint8)
    shift = (poff[:-1].astype(np.int64) - csum[io[:-1]].astype(np.int64))  # per sample: text offset - token offset
    starts = csum[:-1].astype(np.int64)
    starts += np.repeat(shift, np.diff(io))
    assert np.array_equal(pblob[starts], first[ids])
    starts += lens[ids]
    starts -= 1
    assert np.array_equal(pblob[starts], last[ids])


def _crlf_corpus(blob, off):
    samples = [O.crlf(blob[int(off[i]):int(off[i + 1])].tobytes()) for i in range(len(off) - 1)]
    from tokengeex_b200 import _native as N
    return N.pack(samples)


def test_config1_1gb_131k_vocab_properties(N):
    import bench
    from tokengeex_b200 import synth
    toks, sc, kp = bench.build_vocab(synth)
    blob, off, _ = bench.workload(synth, 1, 0, 1_000_000_000)
    S, NB = len(off) - 1, int(off[-1])
    assert NB >= 999_000_000 and len(toks) == 131072
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc)
    lens = np.array([len(t) for t in toks], np.int64)
    ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off, crlf=True)
    assert rc == 0 and bad == -1 and not status.any()
    pblob, poff = gm.crlf_batch(blob, off)
    check_segmentation(blob, off, toks, ids, id_off, plen, pblob, poff)
    # (iv) + two independent device paths: histogram mode of emit == bincount of the emitted ids
    fr, rc, *_ = gm.token_frequencies(blob, off, crlf=True)
    assert rc == 0 and np.array_equal(fr.astype(np.int64), np.bincount(ids, minlength=len(toks)))
    assert int((fr.astype(np.int64) * lens).sum()) == int(plen.sum())
    # chunking of the host entry point does not change a single id
    gm.set_option(7, 96 << 20)
    ids2, id_off2, *_ = gm.encode_batch(blob, off, crlf=True)
    assert np.array_equal(id_off2, id_off) and np.array_equal(ids2, ids)
    # the oracle on a seeded sample of the corpus + its longest samples: bit-exact
    rng = random.Random(1)
    order = np.argsort(np.diff(off.astype(np.int64)))
    pick = sorted(set(rng.sample(range(S), 400)) | set(order[-3:].tolist()) | set(order[:3].tolist()))
    for i in pick:
        text = O.crlf(blob[int(off[i]):int(off[i + 1])].tobytes())
        assert ids[int(id_off[i]):int(id_off[i + 1])].tolist() == om.encode(text, 0.0), i


def test_config2_1gb_code_chinese_131k_vocab(N):
    """configs[2]: one GPU's 1 GB shard of the code+Chinese mix (long multi-byte tokens, max token length 16) with the
    vocabulary bench.py builds for N > 1: segmentation properties over the whole shard, every forward pass agrees,
    the oracle on a seeded sample + the longest samples."""
    import bench
    from tokengeex_b200 import synth
    toks, sc, kp = bench.build_vocab(synth, 2)
    blob, off, _ = bench.workload(synth, 2, 1, 1_000_000_000)  # rank 1's shard of the N = 2 run
    S, NB = len(off) - 1, int(off[-1])
    assert NB >= 999_000_000 and len(toks) == 131072 and max(map(len, toks)) <= 16
    assert any(len(t) >= 9 and t[0] >= 0xE4 for t in toks)  # multi-scalar CJK tokens are in the vocabulary
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc)
    lens = np.array([len(t) for t in toks], np.int64)
    ids, id_off, status, plen, rc, bad = gm.encode_batch(blob, off, crlf=True)
    assert rc == 0 and bad == -1 and not status.any()
    pblob, poff = gm.crlf_batch(blob, off)
    check_segmentation(blob, off, toks, ids, id_off, plen, pblob, poff)
    fr, rc, *_ = gm.token_frequencies(blob, off, crlf=True)
    assert rc == 0 and np.array_equal(fr.astype(np.int64), np.bincount(ids, minlength=len(toks)))
    assert int((fr.astype(np.int64) * lens).sum()) == int(plen.sum())
    for algo in (2, 0):  # pair-CTA kernel / match stream: not one id differs
        gm.set_option(3, algo)
        ids2, id_off2, *_ = gm.encode_batch(blob, off, crlf=True)
        assert np.array_equal(id_off2, id_off) and np.array_equal(ids2, ids), algo
    rng = random.Random(2)
    order = np.argsort(np.diff(off.astype(np.int64)))
    pick = sorted(set(rng.sample(range(S), 400)) | set(order[-3:].tolist()) | set(order[:3].tolist()))
    for i in pick:
        text = O.crlf(blob[int(off[i]):int(off[i + 1])].tobytes())
        assert ids[int(id_off[i]):int(id_off[i + 1])].tolist() == om.encode(text, 0.0), i


def test_config3_500k_vocab_e_step_frequencies_selection(N):
    """configs[3] at its vocabulary size (500 000 tokens, code+Chinese): E-step against the oracle on a sample of the
    corpus (1e-9 relative), bit-identical counts between two runs and between two chunkings, frequency pass exact,
    M-step + prune_vocab selection (incl. the n-best alternatives over the model's own trie) equal to the oracle's."""
    import bench
    from tokengeex_b200 import prune as P
    from tokengeex_b200 import synth
    from tests.util import counts_rel_err
    vb, vo = synth.corpus(synth.KIND_CODE_CJK, 4, bench.PRUNE_VOCAB_SAMPLE_BYTES)
    toks, sc, kp = synth.vocab(vb, vo, 4, bench.PRUNE_VOCAB, 16, 0.05)
    del vb, vo
    assert len(toks) == 500_000
    blob, off = synth.corpus(synth.KIND_CODE_CJK, 4, 48_000_000)
    gm, om = N.Model(toks, sc, device=0), O.OracleModel(toks, sc, kp)
    want = om.run_e_step(blob, off, threads=16)[0]
    ex, rc, _, _ = gm.expected_counts(blob, off)
    assert rc == 0 and counts_rel_err(ex, want) < 1e-9
    ex2 = gm.expected_counts(blob, off)[0]
    assert np.array_equal(ex.view(np.uint64), ex2.view(np.uint64))  # integer accumulation: no run-to-run jitter
    k = len(off) // 2  # two calls over the two halves = one call (what sharding over GPUs relies on)
    import torch
    d_text = torch.from_numpy(blob).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_off2 = torch.from_numpy((off[k:] - off[k]).view(np.int64)).cuda()
    limbs = torch.zeros(5 * len(toks), dtype=torch.int64, device="cuda")
    gm.expected_counts_fixed_dev(d_text.data_ptr(), d_off.data_ptr(), k, int(off[k]), limbs.data_ptr())
    d_text2 = d_text[int(off[k]):].clone()
    gm.expected_counts_fixed_dev(d_text2.data_ptr(), d_off2.data_ptr(), len(off) - 1 - k, int(off[-1] - off[k]),
                                 limbs.data_ptr())
    d_ex = torch.empty(len(toks), dtype=torch.float64, device="cuda")
    gm.counts_from_limbs_dev(limbs.data_ptr(), len(toks), d_ex.data_ptr())
    assert np.array_equal(d_ex.cpu().numpy().view(np.uint64), ex.view(np.uint64))
    # frequency pass at 500k tokens: exact
    fr, rc, *_ = gm.token_frequencies(blob, off)
    wfr = om.token_frequencies(blob, off, threads=16)
    assert rc == 0 and np.array_equal(fr, wfr)
    # M-step, rebuild, frequency pass and selection at ~250k tokens
    kept, ns = N.m_step(want, kp)
    idx = np.flatnonzero(kept)
    v2 = P.Vocab([toks[i] for i in idx], ns[idx].copy(), np.asarray(kp, np.uint8)[idx].copy())
    om2 = om.run_m_step(want)
    t2, s2, k2 = om2.export()
    assert t2 == v2.tokens and np.array_equal(np.asarray(s2).view(np.uint64), v2.scores.view(np.uint64))
    gm.rebuild(v2.tokens, v2.scores)
    fr2 = gm.token_frequencies(blob, off)[0]
    assert np.array_equal(fr2, om2.token_frequencies(blob, off, threads=16))
    ids, audit = gm.prune_select(v2.tokens, v2.scores, v2.keep, fr2, len(off) - 1, 65536, 0.8, threads=16)
    wv, waudit = om2.prune_vocab(blob, off, 65536, 0.8, threads=16)
    assert [v2.tokens[i] for i in ids] == list(wv.export()[0])
