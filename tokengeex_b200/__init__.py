"""tokengeex_b200 — B200-native (sm_100a) implementation of TokenGeeX's data-parallel hot path:
UnigramLM Viterbi encode and the forward-backward / frequency passes of EM vocabulary pruning.

    _native     ctypes binding of the C ABI (include/tokengeex_b200.h) — CUDA kernels live behind it
    tokenizer   `tokengeex.Tokenizer` mirror (JSON v2.0, special tokens, encode*/decode*)
    prune       ModelVocabularyPruner mirror (E-step and frequency pass on the GPU)
    dist        one-process-per-GPU sharding over torch.distributed
    synth       deterministic synthetic corpora / vocabularies for tests and bench.py
"""
from .tokenizer import TokenGeeXError, Tokenizer  # noqa: F401

__all__ = ["Tokenizer", "TokenGeeXError"]
