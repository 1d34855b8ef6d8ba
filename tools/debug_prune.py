#!/usr/bin/env python
"""Developer tool: steps the GPU prune loop and the oracle prune loop side by side and reports the first divergence."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
from tests.util import synth_setup
from tokengeex_b200 import _native as N


def main():
    kind, seed, nbytes, v0, target, subiters = 2, 21, 1_500_000, 4000, 2000, 2
    if len(sys.argv) > 1:  # kind seed bytes vocab target subiters
        kind, seed, nbytes, v0, target, subiters = (int(x) for x in sys.argv[1:7])
    blob, off, toks, sc, kp = synth_setup(kind, seed, nbytes, v0, 16)
    toks = list(toks); sc = np.array(sc); kp = np.array(kp)
    om = O.OracleModel(toks, sc, kp)
    step = 0
    while len(toks) > target:
        for sub in range(subiters):
            gm = N.Model(toks, sc, device=0)
            ex, rc, bad, bz = gm.expected_counts(blob, off)
            wex, wrc, _, _ = om.run_e_step(blob, off, threads=8)
            wex1, _, _, _ = om.run_e_step(blob, off, threads=1)
            nz = wex > 0
            rel = np.abs(ex[nz] - wex[nz]) / wex[nz]
            rel1 = np.abs(wex1[nz] - wex[nz]) / wex[nz]
            flip = np.flatnonzero((ex >= 0.5) != (wex >= 0.5))
            flip1 = np.flatnonzero((wex1 >= 0.5) != (wex >= 0.5))
            print(f"step {step} sub {sub}: V={len(toks)} max rel gpu-vs-oracle8 {rel.max():.3e}  oracle1-vs-oracle8 {rel1.max():.3e} "
                  f"flips gpu {flip.tolist()} oracle1 {flip1.tolist()}  zero-mismatch {int(((ex>0)!=(wex>0)).sum())}", flush=True)
            for i in flip[:10]:
                print("   flip", i, toks[i], repr(ex[i]), repr(wex[i]), repr(wex1[i]), "keep", kp[i])
            near = np.flatnonzero(np.abs(wex - 0.5) < 1e-9)
            print("   tokens within 1e-9 of 0.5:", len(near), [(int(i), repr(wex[i]), repr(ex[i])) for i in near[:6]])
            kept, ns = N.m_step(ex, kp)
            idx = np.flatnonzero(kept)
            om2 = om.run_m_step(wex)
            wt, ws, wk = om2.export()
            gt = [toks[i] for i in idx]
            print(f"   M-step: gpu {len(gt)} oracle {len(wt)} same_tokens {gt == wt}", flush=True)
            # continue on the ORACLE's trajectory so that later steps are compared on equal inputs
            toks, sc, kp, om = list(wt), np.array(ws), np.array(wk), om2
            gm.close()
        gm = N.Model(toks, sc, device=0)
        fr, rc, bad, bl = gm.token_frequencies(blob, off)
        wfr = om.token_frequencies(blob, off, threads=8)
        print(f"step {step} freq equal {np.array_equal(fr, wfr)}", flush=True)
        ids, audit = N.prune_select(toks, sc, kp, fr, len(off) - 1, target, 0.8)
        om2, waudit = om.prune_vocab(blob, off, target, 0.8, threads=8)
        wt, ws, wk = om2.export()
        gt = [toks[i] for i in ids]
        print(f"   prune_select: gpu {len(gt)} oracle {len(wt)} same {gt == wt} audit {audit.tolist()} oracle audit {waudit.tolist()}", flush=True)
        toks, sc, kp, om = list(wt), np.array(ws), np.array(wk), om2
        gm.close()
        step += 1


if __name__ == "__main__":
    main()
