#!/usr/bin/env python
"""Writes tests/golden/hotpath_small.json: a self-contained fixture (vocabulary, samples and the
oracle's outputs for every hot-path function) used by the CPU and GPU test-suites.

The reference is a Rust crate and this image has no cargo/rustc, so the fixture cannot be
produced by the reference itself; it is produced by the CPU oracle (oracle/tgx_oracle.cpp, a
line-by-line restatement pinned against the reference's own unit-test goldens in
tests/test_oracle_goldens.py).  Floats are stored as C99 hex strings (bit-exact).

    python tools/make_golden.py            # rewrites the fixture
"""
import json
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tokengeex_b200 import synth  # noqa: E402


def main():
    blob, off = synth.corpus(synth.KIND_CODE_CJK, 77, 400_000)
    toks, sc, kp = synth.vocab(blob, off, 77, 600, 16, 0.05)
    rng = random.Random(77)
    raw = blob.tobytes()
    samples = []
    for i in range(len(off) - 1):
        s = raw[int(off[i]):int(off[i + 1])]
        samples.append(s[:rng.randrange(1, 700)])
    samples = samples[:48]
    samples += [b"", b"a\r\nb\r\r\n", "你好，我叫罗杰斯".encode(), b"    " * 40, b"\r\n" * 9, bytes(range(32, 127))]
    sblob, soff = O.pack_samples(samples)
    om = O.OracleModel(toks, sc, kp)
    ids, id_off, status, plen, bad = om.encode_batch(sblob, soff, crlf=True, threads=1)
    ids_raw, id_off_raw, status_raw, _, _ = om.encode_batch(sblob, soff, crlf=False, threads=1)
    # prune path operates on pre-processed samples (src/cli.rs:279-285): crlf first, drop empties
    proc = [O.crlf(s) for s in samples]
    proc = [s for s in proc if s]
    pblob, poff = O.pack_samples(proc)
    ex, rc, _, _ = om.run_e_step(pblob, poff, threads=1, literal=True)
    assert rc == 0
    ex_short, rc2, _, _ = om.run_e_step(pblob, poff, threads=1, literal=True, max_sample_length=64)
    assert rc2 == 0
    fr = om.token_frequencies(pblob, poff, threads=1)
    m2 = om.run_m_step(ex)
    mt, ms, mk = m2.export()
    pv, audit = om.prune_vocab(pblob, poff, 400, 0.8, threads=1)
    pt, ps, pk = pv.export()
    fx = {
        "about": "oracle-generated fixture (see tools/make_golden.py); the reference (Rust) cannot run in this image",
        "vocab": [{"hex": t.hex(), "score": float(s).hex(), "keep": int(k)} for t, s, k in zip(toks, sc, kp)],
        "samples_hex": [s.hex() for s in samples],
        "encode_crlf": {"ids": ids.tolist(), "id_off": id_off.tolist(), "status": status.tolist(),
                        "proc_len": plen.tolist()},
        "encode_raw": {"ids": ids_raw.tolist(), "id_off": id_off_raw.tolist(), "status": status_raw.tolist()},
        "prune_samples_hex": [s.hex() for s in proc],
        "expected_counts": [float(x).hex() for x in ex],
        "expected_counts_snippet64": [float(x).hex() for x in ex_short],
        "token_frequencies": fr.tolist(),
        "m_step": {"tokens_hex": [t.hex() for t in mt], "scores": [float(x).hex() for x in ms], "keep": [int(k) for k in mk]},
        "prune_vocab": {"target": 400, "shrink": 0.8, "tokens_hex": [t.hex() for t in pt],
                        "scores": [float(x).hex() for x in ps]},
    }
    out = os.path.join(ROOT, "tests", "golden", "hotpath_small.json")
    with open(out, "w") as f:
        json.dump(fx, f, separators=(",", ":"))
    print(out, os.path.getsize(out), "bytes;", len(toks), "tokens,", len(samples), "samples,", ids.size, "ids")


if __name__ == "__main__":
    main()
