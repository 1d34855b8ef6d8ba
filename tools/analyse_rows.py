#!/usr/bin/env python
"""Match rows on the bench corpus / vocabulary (CPU only; DESIGN.md §4 "match rows").

For every start position: the deepest vocabulary token that prefixes the text there (its "row" lists every shorter
token on the same trie path), matches per position, and how many positions the rows of the K smallest ids cover
(vocabularies are score-sorted, so the first bytes of the row table are the hot ones a kernel would stage in shared
memory)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=600_000)
    ap.add_argument("--vocab", type=int, default=131072)
    ap.add_argument("--kind", type=int, default=1)
    ap.add_argument("--seed", type=int, default=2)
    args = ap.parse_args()
    from tokengeex_b200 import synth
    vb, vo = synth.corpus(args.kind, args.seed, 96_000_000)
    toks, sc, kp = synth.vocab(vb, vo, args.seed, args.vocab, 16, 0.05)
    tid = {}
    prefixes = set()
    for i, t in enumerate(toks):
        if t:
            tid[t] = i
        for l in range(1, len(t) + 1):
            prefixes.add(t[:l])
    # row of token t = every prefix of t that is a token
    row_units = {}
    for t, i in tid.items():
        k = sum(1 for l in range(1, len(t) + 1) if t[:l] in tid)
        row_units[i] = 1 + k
    blob, off = synth.corpus(args.kind, args.seed, 3_000_000)
    text = blob[:args.bytes].tobytes()
    n = len(text) - 16
    deepest = np.full(n, -1, np.int64)
    nmatch = np.zeros(n, np.int32)
    probes = np.zeros(n, np.int32)
    for p in range(n):
        d = 0
        while d < 16 and text[p:p + d + 1] in prefixes:
            d += 1
            i = tid.get(text[p:p + d])
            if i is not None:
                deepest[p] = i
                nmatch[p] += 1
        probes[p] = min(d + 1, 16)
    print(f"V={len(toks)} positions={n}: {probes.mean():.2f} probes, {nmatch.mean():.2f} matches per start; "
          f"no match at {int((deepest < 0).sum())} positions")
    print("matches/position histogram:", np.bincount(nmatch, minlength=17).tolist())
    ru = np.array([row_units.get(i, 0) for i in range(len(toks))])
    units16 = (ru + 1) // 2 * 2  # rows padded to 16 bytes
    print(f"row table: {int((ru > 0).sum())} rows, {units16.sum() * 8 / 1e6:.2f} MB (16-byte aligned rows)")
    cum = np.cumsum(units16) * 8
    d = deepest[deepest >= 0]
    for kb in (16, 32, 64, 96, 128, 160, 192):
        lim = int(np.searchsorted(cum, kb * 1024))
        print(f"  first {kb:3d} KB of rows = ids < {lim:6d}: covers {100.0 * (d < lim).mean():.1f}% of positions")
    # ideal: rows sorted by how often they are the deepest row on THIS text
    cnt = np.bincount(d, minlength=len(toks))
    order = np.argsort(-cnt)
    cumo = np.cumsum(units16[order]) * 8
    cc = np.cumsum(cnt[order]) / max(1, d.size)
    for kb in (16, 32, 64, 96, 128, 160, 192):
        lim = int(np.searchsorted(cumo, kb * 1024))
        print(f"  oracle order, {kb:3d} KB: covers {100.0 * cc[min(lim, len(cc) - 1)]:.1f}% of positions")


if __name__ == "__main__":
    main()
