#!/usr/bin/env python
"""Developer probe: cost of dropout > 0 (lane-group kernels with the keyed draw) against the default kernels, on the
bench's 1 GB corpus and 131k vocabulary: encode (host entry point, pinned buffers) and one E-step (host buffers)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import bench
    from tokengeex_b200 import _native as N, synth
    toks, sc, kp = bench.build_vocab(synth)
    m = N.Model(toks, sc, device=0)
    h_text = N.pinned_empty(1_000_000_000)
    blob, off, _ = bench.workload(synth, 1, 0, 1_000_000_000, out=h_text)
    NB = int(off[-1])
    h_ids = N.pinned_empty(4 * (NB + 16)).view(np.uint32)
    for p in (0.0, 0.1):
        m.set_dropout(p, 1)
        best, ntok = 1e9, 0
        for i in range(3):
            t = time.perf_counter()
            ids, *_ = m.encode_batch(blob, off, crlf=True, ids_out=h_ids)
            dt = time.perf_counter() - t
            ntok = ids.size
            if i:
                best = min(best, dt)
        print(f"encode dropout={p}: {best * 1e3:.1f} ms end to end ({NB / best / 1e9:.2f} GB/s), forward {m.stat(1):.1f} ms "
              f"of the last chunk, {ntok} tokens", flush=True)
    for p in (0.0, 0.01):
        m.set_dropout(p, 1)
        best = 1e9
        for i in range(3):
            t = time.perf_counter()
            ex, rc, *_ = m.expected_counts(blob, off)
            dt = time.perf_counter() - t
            if i:
                best = min(best, dt)
        print(f"E-step dropout={p}: {best * 1e3:.1f} ms with host buffers, device {m.stat(4):.1f} ms, rc={rc}", flush=True)


if __name__ == "__main__":
    main()
