"""EM vocabulary pruning driven by the GPU hot path.

Mirrors ``ModelVocabularyPruner`` (/root/reference/src/prune.rs:6-57): the E-step
(run_e_step, :64-120) and the frequency pass of prune_vocab (:205-246) run on the GPU
through the C ABI; run_m_step (:124-170) and the rest of prune_vocab (:173-319) run on the
host in C++ (csrc/prune_host.cpp).  After every step the model is rebuilt from the new
vocabulary exactly as ``*model = Model::from(vocab)`` does (:48, :53) — here that means a
new double-array trie uploaded to the device.

With more than one rank (torch.distributed), every rank holds a shard of the samples; the
expected-count vector and the frequency vector are all-reduced (one collective per E-step
/ frequency pass) so that every rank runs the identical host steps on identical inputs.

When torch sees a CUDA device the corpus shard is uploaded ONCE and stays resident in HBM for
the whole schedule (every E-step / frequency pass reads it through the `_dev` entry points);
the count vectors are accumulated on the device and, under NCCL, all-reduced there before the
single D2H copy of V values per step.
"""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _native as N

log = logging.getLogger("tokengeex_b200.prune")


@dataclass
class Vocab:
    """Vec<ScoredToken> (/root/reference/src/lib.rs:27-31)."""
    tokens: List[bytes]
    scores: np.ndarray  # f64[V]
    keep: np.ndarray    # u8[V]
    _packed: Optional[tuple] = field(default=None, repr=False, compare=False)

    def __len__(self):
        return len(self.tokens)

    def packed(self):
        """(blob u8[], offsets u64[V + 1]) of the tokens as the C ABI takes them, made once per vocabulary (a Vocab is
        not modified after it is built: every step of the loop returns a new one)."""
        if self._packed is None:
            self._packed = N.pack(self.tokens)
        return self._packed


@dataclass
class PruneReport:
    vocab_sizes: List[int] = field(default_factory=list)   # after every M-step / prune step
    e_step_s: List[float] = field(default_factory=list)
    m_step_s: List[float] = field(default_factory=list)
    freq_s: List[float] = field(default_factory=list)
    select_s: List[float] = field(default_factory=list)
    rebuild_s: List[float] = field(default_factory=list)
    allreduce_s: List[float] = field(default_factory=list)  # the collective of every E-step / frequency pass
    audits: List[np.ndarray] = field(default_factory=list)
    # margin audit: how close every discrete decision of the schedule came to going the other way
    m_margins: List[dict] = field(default_factory=list)    # per M-step: distance of the counts from the 0.5 threshold


class ModelVocabularyPruner:
    """new(vocab_size, shrink_factor, em_subiters, dropout) — src/prune.rs:13-21.

    dropout is the E-step's `populate_nodes(.., self.dropout)` (src/prune.rs:87): 0.0 = the default kernels (every
    BASELINE configuration); in (0, 1) every multi-byte match is dropped with that probability by a keyed draw
    (tgx_model_set_dropout: key = seed of this E-step, byte offset in the whole corpus, token length — so the result
    does not depend on how the corpus is sharded), on the lane-group E-step kernels.  The reference's draw is an
    unseeded thread_rng: `dropout_seed=None` takes a fresh seed per prune(), an int makes the run reproducible.
    The frequency pass encodes with dropout 0.0 like the reference (src/prune.rs:218).
    """

    def __init__(self, vocab_size: int, shrink_factor: float = 0.8, em_subiters: int = 1, dropout: float = 0.0,
                 device: int = 0, allreduce: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                 n_samples_global: Optional[int] = None, dropout_seed: Optional[int] = None, byte_base: int = 0):
        if not (0.0 <= dropout < 1.0):
            raise ValueError("dropout must be in [0, 1)")
        self.dropout = float(dropout)
        self.dropout_seed = dropout_seed
        self.byte_base = int(byte_base)  # offset of this rank's shard in the whole corpus (keyed dropout draw)
        self._e_steps = 0
        self.vocab_size = vocab_size
        self.shrink_factor = shrink_factor
        self.em_subiters = em_subiters
        self.device = device
        self.allreduce = allreduce
        self.n_samples_global = n_samples_global

    # -- device-resident corpus ------------------------------------------------------------------
    def _upload(self, blob: np.ndarray, off: np.ndarray):
        """Corpus shard -> HBM, once per prune() (None when torch has no CUDA device: host-buffer API)."""
        try:
            import torch
            if not torch.cuda.is_available():
                return None
        except ImportError:
            return None
        dev = torch.device("cuda", self.device)
        d_text = torch.from_numpy(np.ascontiguousarray(blob)).to(dev)
        d_off = torch.from_numpy(off.astype(np.uint64).view(np.int64)).to(dev)
        return {"torch": torch, "dev": dev, "text": d_text, "off": d_off, "S": len(off) - 1, "N": int(off[-1])}

    def _reduce_dev(self, t, torch=None):
        """Integer tensor summed across ranks — on the device when the collective runs there (NCCL), else via the
        host.  Returns a tensor on t's device; the time of the collective goes to self.last_allreduce_s."""
        coll = self.allreduce
        self.last_allreduce_s = 0.0
        if coll is None:
            return t
        if torch is not None:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if hasattr(coll, "allreduce_tensor") and str(getattr(coll, "device", "cpu")).startswith("cuda"):
            t = coll.allreduce_tensor(t)
            if torch is not None:
                torch.cuda.synchronize()
        else:
            dev = t.device
            t = (torch or __import__("torch")).from_numpy(coll(t.cpu().numpy().view(np.uint64)).view(np.int64)).to(dev)
        self.last_allreduce_s = time.perf_counter() - t0
        return t

    # -- steps --------------------------------------------------------------------------------
    def run_e_step(self, model: N.Model, blob: np.ndarray, off: np.ndarray, dev=None) -> np.ndarray:
        if self.dropout > 0.0:  # a new set of draws for every E-step, like the reference's running generator
            if self.dropout_seed is None:
                import secrets
                self.dropout_seed = secrets.randbits(63)
            model.set_option(22, self.byte_base)
            model.set_dropout(self.dropout, self.dropout_seed + self._e_steps)
            self._e_steps += 1
            try:
                return self._run_e_step(model, blob, off, dev)
            finally:
                model.set_dropout(0.0, 0)
        return self._run_e_step(model, blob, off, dev)

    def _run_e_step(self, model: N.Model, blob: np.ndarray, off: np.ndarray, dev=None) -> np.ndarray:
        if dev is not None:
            # Counts as exact integer limbs (tgx_expected_counts_fixed_dev): the all-reduce is an integer sum, so the
            # result is bit-identical for every number of ranks and every sharding of the corpus.
            torch = dev["torch"]
            V = max(model.V, 1)
            d_limbs = torch.zeros(5 * V, dtype=torch.int64, device=dev["dev"])
            # one call takes fewer than 2^32 bytes: larger shards go in parts of whole samples, summed into the same limbs
            lo = 0
            while lo < dev["S"]:
                hi = int(np.searchsorted(off, int(off[lo]) + (1 << 32) - (1 << 20), side="right")) - 1
                hi = min(max(hi, lo + 1), dev["S"])
                b0, b1 = int(off[lo]), int(off[hi])
                d_off = dev["off"] if (lo == 0 and hi == dev["S"]) else (dev["off"][lo:hi + 1] - b0).contiguous()
                rc, bad, badz = model.expected_counts_fixed_dev(dev["text"].data_ptr() + b0, d_off.data_ptr(), hi - lo,
                                                                b1 - b0, d_limbs.data_ptr())
                if rc == N.TGX_ERR_BAD_Z:
                    raise FloatingPointError(f"normalization constant is f64::NaN (z={badz}, sample={lo + bad})")
                lo = hi
            d_limbs = self._reduce_dev(d_limbs, torch)
            d_ex = torch.empty(V, dtype=torch.float64, device=dev["dev"])
            model.counts_from_limbs_dev(d_limbs.data_ptr(), model.V, d_ex.data_ptr())
            return d_ex.cpu().numpy()[:model.V]
        ex, rc, bad, badz = model.expected_counts(blob, off)
        if rc == N.TGX_ERR_BAD_Z:  # the reference panics (src/prune.rs:90-96)
            raise FloatingPointError(f"normalization constant is f64::NaN (z={badz}, sample={bad})")
        if self.allreduce is not None:
            ex = self.allreduce(ex)
        return ex

    def run_m_step(self, vocab: Vocab, expected: np.ndarray, report: Optional[PruneReport] = None) -> Vocab:
        if report is not None:  # `freq < 0.5 && !keep` drops a token (src/prune.rs:132)
            ex = np.asarray(expected, np.float64)
            free = np.asarray(vocab.keep) == 0
            d = np.abs(ex[free] - 0.5) if free.any() else np.array([np.inf])
            report.m_margins.append({"min_abs_distance_to_0.5": float(d.min()),
                                     "tokens_within_1e-9": int((d < 1e-9).sum()),
                                     "tokens_within_1e-6": int((d < 1e-6).sum()),
                                     "dropped": int(((ex < 0.5) & free).sum())})
        kept, ns = N.m_step(expected, vocab.keep)
        idx = np.flatnonzero(kept)
        return Vocab([vocab.tokens[i] for i in idx], ns[idx].copy(), vocab.keep[idx].copy())

    def prune_vocab(self, model: N.Model, vocab: Vocab, blob: np.ndarray, off: np.ndarray, report: PruneReport,
                    dev=None) -> Vocab:
        t = time.perf_counter()
        if dev is not None:
            torch = dev["torch"]
            d_fr = torch.zeros(max(model.V, 1), dtype=torch.int64, device=dev["dev"])
            rc, bad, blen = model.token_frequencies_dev(dev["text"].data_ptr(), dev["off"].data_ptr(), dev["S"],
                                                        dev["N"], False, d_fr.data_ptr())
            if rc == N.TGX_ERR_NO_PATH:
                raise RuntimeError(f"no path to position {blen}/{blen}")
            fr = self._reduce_dev(d_fr, torch).cpu().numpy()[:model.V].view(np.uint64)
        else:
            fr, rc, bad, blen = model.token_frequencies(blob, off)
            if rc == N.TGX_ERR_NO_PATH:
                raise RuntimeError(f"no path to position {blen}/{blen}")  # Error::NoPath, src/prune.rs:218-221
            if self.allreduce is not None:
                fr = self.allreduce(fr)
        self.last_freq = fr
        report.freq_s.append(time.perf_counter() - t)
        report.allreduce_s.append(getattr(self, "last_allreduce_s", 0.0))
        t = time.perf_counter()
        n_samples = self.n_samples_global if self.n_samples_global is not None else len(off) - 1
        # the model was rebuilt from `vocab` just before (src/prune.rs:48): its trie serves the n-best alternatives.
        # Every rank holds the same inputs, so with a collective ONE rank runs the selection on all the cores of the box
        # and broadcasts the surviving ids (eight replicas sharing the cores took 0.105 s against 0.057 s alone).
        import os
        coll = self.allreduce
        cores = len(os.sched_getaffinity(0))
        if coll is not None and hasattr(coll, "broadcast_u32") and coll.world_size > 1:
            ids, audit = None, np.zeros(8)
            if coll.rank == 0:
                ids, audit = model.prune_select(vocab.tokens, vocab.scores, vocab.keep, fr, n_samples, self.vocab_size,
                                                self.shrink_factor, threads=max(2, cores - coll.world_size + 1),
                                                packed=vocab.packed())
            ids = coll.broadcast_u32(ids, 0)
            audit = coll.allreduce(np.ascontiguousarray(audit, np.float64))  # (zeros on the other ranks)
        else:
            ids, audit = model.prune_select(vocab.tokens, vocab.scores, vocab.keep, fr, n_samples, self.vocab_size,
                                            self.shrink_factor, threads=max(2, cores), packed=vocab.packed())
        report.select_s.append(time.perf_counter() - t)
        report.audits.append(audit)
        return Vocab([vocab.tokens[i] for i in ids], vocab.scores[ids].copy(), vocab.keep[ids].copy())

    # -- src/prune.rs:23-57 ------------------------------------------------------------------------
    def prune(self, vocab: Vocab, blob: np.ndarray, off: np.ndarray) -> (Vocab, PruneReport):
        report = PruneReport()
        dev = self._upload(blob, off)
        t = time.perf_counter()
        model = N.Model(vocab.tokens, vocab.scores, device=self.device)
        report.rebuild_s.append(time.perf_counter() - t)
        while len(vocab) > self.vocab_size:
            for subiter in range(self.em_subiters):
                log.info("EM subiter %d/%d", subiter + 1, self.em_subiters)
                t = time.perf_counter()
                expected = self.run_e_step(model, blob, off, dev)
                report.e_step_s.append(time.perf_counter() - t)
                report.allreduce_s.append(getattr(self, "last_allreduce_s", 0.0))
                log.info("E-step completed subiter=%d vocab_size=%d", subiter, len(vocab))
                t = time.perf_counter()
                new_vocab = self.run_m_step(vocab, expected, report)
                report.m_step_s.append(time.perf_counter() - t)
                log.info("M-step completed subiter=%d vocab_size=%d alternative_vocab_size=%d", subiter, len(vocab),
                         len(new_vocab))
                vocab = new_vocab
                t = time.perf_counter()
                model.rebuild(vocab.tokens, vocab.scores, packed=vocab.packed())  # *model = Model::from(vocab)  (src/prune.rs:48)
                report.rebuild_s.append(time.perf_counter() - t)
                report.vocab_sizes.append(len(vocab))
            before = len(vocab)
            vocab = self.prune_vocab(model, vocab, blob, off, report, dev)
            log.info("Pruning vocabulary from=%d to=%d", before, len(vocab))
            t = time.perf_counter()
            model.rebuild(vocab.tokens, vocab.scores, packed=vocab.packed())  # (src/prune.rs:53)
            report.rebuild_s.append(time.perf_counter() - t)
            report.vocab_sizes.append(len(vocab))
        model.close()
        return vocab, report
