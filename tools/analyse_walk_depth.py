#!/usr/bin/env python
"""Trie-walk depth on the bench corpus / vocabulary (CPU only): probes per start position and the deepest walk among R
consecutive starts — the producer side of a pair-kernel round waits for the deepest of its 64."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=400_000)
    ap.add_argument("--vocab", type=int, default=131072)
    args = ap.parse_args()
    from tokengeex_b200 import synth
    vb, vo = synth.corpus(synth.KIND_MULTILANG, 2, 96_000_000)
    toks, sc, kp = synth.vocab(vb, vo, 2, args.vocab, 16, 0.05)
    prefixes = set()
    for t in toks:
        for l in range(1, len(t) + 1):
            prefixes.add(t[:l])
    blob, off = synth.corpus(1, 2, 3_000_000)
    text = blob[:args.bytes].tobytes()
    n = len(text)
    depth = np.zeros(n, np.int32)
    for p in range(n - 16):
        d = 0
        while d < 16 and text[p:p + d + 1] in prefixes:
            d += 1
        depth[p] = min(d + 1, 16)  # probes incl. the failing one
    print(f"V={len(toks)} positions={n}: {depth.mean():.2f} probes per start; histogram by probes (0..16):")
    print("  ", np.bincount(depth, minlength=17).tolist())
    for R in (32, 64, 128, 576):
        m = depth[:n // R * R].reshape(-1, R).max(1)
        print(f"deepest of {R:3d} consecutive starts: mean {m.mean():.2f}  p90 {np.percentile(m, 90):.0f}")


if __name__ == "__main__":
    main()
