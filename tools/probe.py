#!/usr/bin/env python
"""Developer probe: times encode / frequency pass / E-step variants on one GPU (not the bench contract)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bytes", type=int, default=100_000_000)
    ap.add_argument("--vocab", type=int, default=131072)
    ap.add_argument("--kind", type=int, default=1)
    ap.add_argument("--what", default="encode,estep")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--single", type=int, default=0, help="encode only the N longest samples")
    ap.add_argument("--maxlen", type=int, default=0, help="drop the samples of at least this many bytes")
    ap.add_argument("--minlen", type=int, default=0, help="drop the samples shorter than this")
    ap.add_argument("--estep-cfgs", default="", help="comma list of G:threshold pairs for the E-step")
    ap.add_argument("--snippet", type=int, default=81920, help="E-step snippet length (throughput experiments)")
    ap.add_argument("--algos", default="", help="comma list of forward algos to time (default all)")
    ap.add_argument("--opts", default=None, help="encode: ONE configuration instead of the sweep: algo:key=value,key=value")
    args = ap.parse_args()
    import torch
    from tokengeex_b200 import _native as N, synth
    vb, vo = synth.corpus(synth.KIND_MULTILANG, 2, 96_000_000)
    toks, sc, kp = synth.vocab(vb, vo, 2, args.vocab, 16, 0.05)
    m = N.Model(toks, sc, device=0)
    blob, off = synth.corpus(args.kind, 2, args.bytes)
    if args.single:
        lens = np.diff(off.astype(np.int64))
        idx = np.argsort(-lens)[:args.single]
        parts = [blob[int(off[i]):int(off[i + 1])].tobytes() for i in idx]
        blob, off = N.pack(parts)
    if args.maxlen or args.minlen:
        lens = np.diff(off.astype(np.int64))
        keep = np.flatnonzero((lens >= args.minlen) & ((lens < args.maxlen) if args.maxlen else True))
        parts = [blob[int(off[i]):int(off[i + 1])].tobytes() for i in keep]
        blob, off = N.pack(parts)
    S, NB = len(off) - 1, int(off[-1])
    d_text = torch.from_numpy(blob).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_ids = torch.empty(NB + 4, dtype=torch.int32, device="cuda")
    d_id_off = torch.empty(S + 1, dtype=torch.int64, device="cuda")
    d_ex = torch.zeros(len(toks), dtype=torch.float64, device="cuda")
    d_fr = torch.zeros(len(toks), dtype=torch.int64, device="cuda")
    print(f"V={len(toks)} S={S} N={NB} slots={m.info().trie_slots}", flush=True)
    what = args.what.split(",")
    if "encode" in what:
        K = 1024
        cfgs = [(2, {14: 0}), (0, {}), (3, {32: 131072}), (3, {32: 65536}), (3, {32: 32768}), (3, {32: 200000}), (3, {32: 1 << 30})]
        if args.algos:
            cfgs = [c for c in cfgs if str(c[0]) in args.algos.split(",")]
        if args.opts is not None:  # "algo:key=value,key=value;algo:..."
            cfgs = []
            for one in args.opts.split(";"):
                a, _, kv = one.partition(":")
                cfgs.append((int(a), {int(x.split("=")[0]): int(x.split("=")[1]) for x in kv.split(",") if x}))
        defaults = {6: 0, 13: 2, 14: 0, 32: 65536, 33: 8, 35: 160 << 10, 37: 1, 38: 20, 39: 1, 43: 4}
        for algo, opts in cfgs:
            m.set_option(3, algo)
            for k, v in {**defaults, **opts}.items():  # (options are sticky: every configuration starts from the defaults)
                m.set_option(k, v)
            best = 1e9
            for _ in range(args.reps + 1):
                tot, rc, bad = m.encode_batch_dev(d_text.data_ptr(), d_off.data_ptr(), S, NB, True, d_ids.data_ptr(),
                                                  NB + 4, d_id_off.data_ptr())
                best = min(best, m.stat(4))
            chk = int(d_ids[:tot].to(torch.int64).sum()) if tot else 0
            print(f"encode algo={algo} opts={opts}: {best:.2f} ms  {NB / best / 1e6:.2f} GB/s  forward {m.stat(1):.2f} ms "
                  f"match {m.stat(7):.2f} ms all-forward {m.stat(8):.2f} ms side {m.stat(9):.2f} ms backtrack {m.stat(5):.2f} ms emit {m.stat(6):.2f} ms tokens={tot} idsum={chk} rc={rc}", flush=True)
        m.set_option(3, 0)
    if "freq" in what:
      for eh in (0, 1):
        m.set_option(16, eh)
        for _ in range(args.reps):
            d_fr.zero_()
            rc, bad, bl = m.token_frequencies_dev(d_text.data_ptr(), d_off.data_ptr(), S, NB, False, d_fr.data_ptr())
            print(f"freq emit_hash={eh}: {m.stat(4):.2f} ms fwd {m.stat(1):.2f} back {m.stat(5):.2f} emit {m.stat(6):.2f}  {NB / m.stat(4) / 1e6:.2f} GB/s sum={int(d_fr.sum())}", flush=True)
    if "estep" in what:
        cfgs = [(4, 16384), (4, 32768), (4, 65536), (8, 32768), (8, 65536), (2, 16384), (2, 32768), (4, 8192)]
        if args.single == 0 and args.estep_cfgs:
            cfgs = [tuple(int(x) for x in c.split(":")) for c in args.estep_cfgs.split(",")]
        for cfg in cfgs:
            g, thr = cfg[0], cfg[1]
            lane = cfg[2] if len(cfg) > 2 else 0
            m.set_option(18, cfg[3] if len(cfg) > 3 else 0)
            m.set_option(19, cfg[4] if len(cfg) > 4 else 1)
            m.set_option(20, cfg[5] if len(cfg) > 5 else 64)
            m.set_option(21, cfg[6] if len(cfg) > 6 else 4096)
            if os.environ.get("TGX_CUT"):
                c1, c2 = (int(x) for x in os.environ["TGX_CUT"].split(","))
                m.set_option(30, c1)
                m.set_option(31, c2)
            m.set_option(2, g)
            m.set_option(5, thr)
            m.set_option(17, lane)
            for _ in range(args.reps):
                d_ex.zero_()
                rc, bad, bz = m.expected_counts_dev(d_text.data_ptr(), d_off.data_ptr(), S, NB, d_ex.data_ptr(), args.snippet)
            print(f"estep G={g} thr={thr} lane={lane} blocks/SM={cfg[3] if len(cfg) > 3 else 0} split={cfg[4] if len(cfg) > 4 else 1} hot={cfg[5:] if len(cfg) > 5 else ''}: total {m.stat(4):.2f} ms fwd {m.stat(2):.2f} bwd {m.stat(3):.2f}  "
                  f"{NB / m.stat(4) / 1e6:.3f} GB/s sum={float(d_ex.sum()):.3f}", flush=True)


if __name__ == "__main__":
    main()
