"""Deterministic synthetic corpora / initial vocabularies (ctypes over libtgx_synth.so).

Workloads for tests and bench.py — see tokengeex_b200/csrc/synth.cpp and SURVEY.md §8d.
Host-only tooling; not part of the encode / prune hot path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "csrc", "libtgx_synth.so")

KIND_CODE, KIND_MULTILANG, KIND_CODE_CJK = 0, 1, 2

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError(f"{_LIB} not built; run `make -C tokengeex_b200/csrc` or __graft_entry__.build()")
        L = C.CDLL(_LIB)
        L.tgx_synth_sample_lengths.restype = C.c_uint64
        L.tgx_synth_sample_lengths.argtypes = [C.c_uint64, C.c_uint64, _u64p, C.c_uint64]
        L.tgx_synth_corpus.argtypes = [C.c_int, C.c_uint64, _u64p, C.c_uint64, _u8p, C.c_int]
        L.tgx_synth_corpus_range.argtypes = [C.c_int, C.c_uint64, _u64p, C.c_uint64, C.c_uint64, _u8p, C.c_int]
        L.tgx_synth_allow_exact.restype = C.c_int
        L.tgx_synth_allow_exact.argtypes = [_u8p, C.c_uint64]
        L.tgx_synth_vocab.restype = C.c_int64
        L.tgx_synth_vocab.argtypes = [_u8p, _u64p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_double,
                                      C.c_int, _u8p, C.c_uint64, _u64p, _f64p, _u8p]
        _lib = L
    return _lib


def n_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def corpus(kind: int, seed: int, total_bytes: int, threads: int = 0, out: np.ndarray | None = None
           ) -> Tuple[np.ndarray, np.ndarray]:
    """→ (blob u8[total_bytes], offsets u64[S+1]).  `out` may be a preallocated (e.g. pinned) buffer."""
    L = lib()
    cap = total_bytes // 16 + 2
    lens = np.zeros(cap, np.uint64)
    S = int(L.tgx_synth_sample_lengths(seed, total_bytes, lens.ctypes.data_as(_u64p), cap))
    off = np.zeros(S + 1, np.uint64)
    np.cumsum(lens[:S], out=off[1:])
    n = int(off[-1])
    blob = np.empty(n, np.uint8) if out is None else out[:n]
    L.tgx_synth_corpus(kind, seed, off.ctypes.data_as(_u64p), S, blob.ctypes.data_as(_u8p), threads or n_threads())
    return blob, off


def corpus_offsets(seed: int, total_bytes: int) -> np.ndarray:
    """Offsets u64[S+1] of the corpus `corpus(kind, seed, total_bytes)` would generate (no text)."""
    L = lib()
    cap = total_bytes // 16 + 2
    lens = np.zeros(cap, np.uint64)
    S = int(L.tgx_synth_sample_lengths(seed, total_bytes, lens.ctypes.data_as(_u64p), cap))
    off = np.zeros(S + 1, np.uint64)
    np.cumsum(lens[:S], out=off[1:])
    return off


def corpus_shard(kind: int, seed: int, total_bytes: int, rank: int, world_size: int, threads: int = 0,
                 out: np.ndarray | None = None) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """Rank `rank`'s byte-balanced shard (dist.shard_ranges) of ONE corpus, generated without the rest of it.
    → (blob, offsets rebased to 0, index of the shard's first sample, samples in the whole corpus)."""
    from .dist import shard_ranges
    L = lib()
    off = corpus_offsets(seed, total_bytes)
    lo, hi = shard_ranges(off, world_size)[rank]
    n = int(off[hi] - off[lo])
    blob = np.empty(max(n, 1), np.uint8) if out is None else out[:max(n, 1)]
    L.tgx_synth_corpus_range(kind, seed, off.ctypes.data_as(_u64p), lo, hi - lo, blob.ctypes.data_as(_u8p),
                             threads or n_threads())
    return blob[:n] if n else blob[:1], (off[lo:hi + 1] - off[lo]).astype(np.uint64), lo, len(off) - 1


def vocab(blob: np.ndarray, off: np.ndarray, seed: int, vocab_size: int, max_token_length: int = 24,
          insert_probability: float = 0.01, threads: int = 0):
    """generate.rs restated → (tokens: List[bytes], scores f64[V], keep u8[V])."""
    L = lib()
    cap = vocab_size * (max_token_length + 1) + 1024
    tb = np.zeros(cap, np.uint8)
    to = np.zeros(vocab_size + 1, np.uint64)
    sc = np.zeros(vocab_size, np.float64)
    kp = np.zeros(vocab_size, np.uint8)
    V = int(L.tgx_synth_vocab(blob.ctypes.data_as(_u8p), off.ctypes.data_as(_u64p), len(off) - 1, seed, vocab_size,
                              max_token_length, insert_probability, threads or n_threads(),
                              tb.ctypes.data_as(_u8p), cap, to.ctypes.data_as(_u64p), sc.ctypes.data_as(_f64p),
                              kp.ctypes.data_as(_u8p)))
    if V < 0:
        raise RuntimeError("token blob capacity too small")
    raw = tb.tobytes()
    toks: List[bytes] = [raw[int(to[i]):int(to[i + 1])] for i in range(V)]
    return toks, sc[:V].copy(), kp[:V].copy()


def allow_exact(s: bytes) -> bool:
    a = np.frombuffer(s, np.uint8) if s else np.zeros(1, np.uint8)
    return bool(lib().tgx_synth_allow_exact(a.ctypes.data_as(_u8p), len(s)))
