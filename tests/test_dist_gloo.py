"""N > 1 host logic on CPU: two gloo ranks shard a corpus by sample, each computes its shard's count vectors (here
with the oracle standing in for the GPU kernels, which need a device), the vectors are summed through
tokengeex_b200.dist.Collective, and every rank must end up with the whole-corpus result and the same M-step."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
from tests.util import rand_samples, rand_vocab
from tokengeex_b200 import _native as N
from tokengeex_b200.dist import Collective, shard_ranges, take_shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup():
    rng = random.Random(5)
    toks, scores = rand_vocab(rng, alphabet=b"abcd", n_tok=120, max_len=6, complete=True)
    keep = np.array([len(t) == 1 for t in toks], np.uint8)
    samples = rand_samples(rng, b"abcd", 90, 1, 400)
    return toks, scores, keep, samples


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        toks, scores, keep, samples = _setup()
        blob, off = O.pack_samples(samples)
        sub, soff, first = take_shard(blob, off, rank, world)
        om = O.OracleModel(toks, scores, keep)
        coll = Collective()
        assert (coll.rank, coll.world_size, coll.device) == (rank, world, "cpu")
        ex, rc, _, _ = om.run_e_step(sub, soff, threads=1)
        ex = coll(ex)
        fr = coll(om.token_frequencies(sub, soff, threads=1))
        n_samples = coll.sum_int(len(soff) - 1)
        kept, ns = N.m_step(ex, keep)
        # E-step dropout (src/prune.rs:87): one seed for the job (rank 0's, as cli.prune_cmd shares it) and the keyed
        # draw's byte base = the shard's offset in the corpus, so the sharded sum is the unsharded lattice's
        box = [1234567 + 1000 * rank]
        dist.broadcast_object_list(box, src=0)
        base = int(off[shard_ranges(off, world)[rank][0]])
        exd = coll(om.run_e_step_dropout(sub, soff, 0.2, box[0], keyed=True, byte_base=base)[0])
        q.put((rank, ex, fr, n_samples, kept, ns, exd))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_balance_by_bytes():
    off = np.array([0, 10, 20, 1000, 1010, 1020, 2000, 2010], np.uint64)
    r = shard_ranges(off, 2)
    assert r[0][0] == 0 and r[-1][1] == 7 and r[0][1] == r[1][0]
    b = [int(off[hi]) - int(off[lo]) for lo, hi in r]
    assert abs(b[0] - b[1]) <= 1000
    for w in (1, 3, 8, 16):  # every sample exactly once, also with more ranks than samples
        rr = shard_ranges(off, w)
        assert rr[0][0] == 0 and rr[-1][1] == 7 and all(rr[i][1] == rr[i + 1][0] for i in range(w - 1))
    assert shard_ranges(np.zeros(1, np.uint64), 4) == [(0, 0)] * 4


@pytest.mark.timeout(180)
def test_two_ranks_allreduce_counts():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=150) for _ in range(world)), key=lambda x: x[0])
    for p in ps:
        p.join(30)
        assert p.exitcode == 0
    toks, scores, keep, samples = _setup()
    blob, off = O.pack_samples(samples)
    om = O.OracleModel(toks, scores, keep)
    want_ex, _, _, _ = om.run_e_step(blob, off, threads=1)
    want_fr = om.token_frequencies(blob, off, threads=1)
    want_exd = om.run_e_step_dropout(blob, off, 0.2, 1234567, keyed=True)[0]
    assert not np.allclose(want_exd, want_ex, rtol=1e-6)
    for rank, ex, fr, n_samples, kept, ns, exd in res:
        assert np.allclose(exd, want_exd, rtol=1e-12, atol=0)
        assert np.allclose(ex, want_ex, rtol=1e-12, atol=0)
        assert np.array_equal(fr, want_fr) and n_samples == len(samples)
    # identical inputs on every rank -> identical host steps ("replicas only")
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][4], res[1][4])
    assert np.array_equal(res[0][5], res[1][5])
