"""Independent brute-force segmentation enumerator (pure Python, tiny inputs).

Derives Viterbi and forward-backward results from first principles — by
enumerating every segmentation of the input — without sharing code with the
oracle.  Used to pin the oracle where the reference has no golden.
"""
import math
from typing import Dict, List, Sequence, Tuple


def all_paths(text: bytes, tokens: Sequence[bytes]) -> List[List[int]]:
    # duplicate byte strings: the LAST id wins (src/trie.rs:19, src/model.rs:20-23)
    last: Dict[bytes, int] = {}
    for i, t in enumerate(tokens):
        if len(t) > 0:
            last[t] = i
    n = len(text)
    out: List[List[int]] = []

    def rec(pos: int, acc: List[int]):
        if pos == n:
            out.append(list(acc))
            return
        # trie semantics: a token is reachable only if every proper prefix is a
        # trie node, which holds automatically (prefixes of tokens are nodes).
        for l in range(1, n - pos + 1):
            tid = last.get(text[pos:pos + l])
            if tid is not None:
                acc.append(tid)
                rec(pos + l, acc)
                acc.pop()

    rec(0, [])
    return out


def path_score_left_to_right(path: List[int], scores: Sequence[float]) -> float:
    s = 0.0
    for tid in path:
        s = s + scores[tid]  # same left-to-right f64 chain as src/model.rs:98
    return s


def viterbi_first_wins(text: bytes, tokens: Sequence[bytes], scores: Sequence[float]):
    """Best path under the reference's tie rule: for every end position the
    candidate with the smallest start wins ties (src/model.rs:100-101).  Done as
    a DP over enumerated prefixes to stay independent of the oracle's code."""
    n = len(text)
    last: Dict[bytes, int] = {}
    for i, t in enumerate(tokens):
        if len(t) > 0:
            last[t] = i
    best: List[Tuple[float, List[int]]] = [None] * (n + 1)  # type: ignore
    best[0] = (0.0, [])
    for e in range(1, n + 1):
        for s in range(0, e):  # ascending start; replace only on strictly greater
            if best[s] is None:
                continue
            tid = last.get(text[s:e])
            if tid is None:
                continue
            sc = best[s][0] + scores[tid]
            if best[e] is None or sc > best[e][0]:
                best[e] = (sc, best[s][1] + [tid])
    return None if best[n] is None else best[n][1]


def marginals(text: bytes, tokens: Sequence[bytes], scores: Sequence[float]):
    """Sum over all complete paths: Z and per-token expected counts."""
    paths = all_paths(text, tokens)
    if not paths:
        return None, None
    logs = [math.fsum(scores[t] for t in p) for p in paths]
    m = max(logs)
    Z = math.fsum(math.exp(l - m) for l in logs)
    logz = m + math.log(Z)
    exp = [0.0] * len(tokens)
    for p, l in zip(paths, logs):
        w = math.exp(l - logz)
        for t in p:
            exp[t] += w
    return logz, exp
