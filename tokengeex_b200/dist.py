"""One process per GPU over torch.distributed (NCCL on the GPUs, gloo in the CPU tests).

The hot path shards by SAMPLE (rayon's only parallel axis in the reference:
/root/reference/src/tokenizer.rs:107-110, src/prune.rs:72,211-212):

  * encode / encode_batch: contiguous sample ranges balanced by bytes, no data-path collective;
  * E-step: every rank accumulates expected[V] over its shard, then ONE all-reduce (f64 sum) replaces
    the RwLock merge of src/prune.rs:104-112;
  * frequency pass of prune_vocab: same with an exact integer sum (src/prune.rs:231-236).

After the collective every rank holds identical vectors and runs the identical host M-step /
prune_vocab, so no further communication is needed ("replicas only").
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_ranges(off: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous sample ranges [lo, hi) per rank, balanced by bytes (not by count)."""
    S = len(off) - 1
    total = int(off[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = (total * r) // world_size
        i = int(np.searchsorted(off, target, side="left"))
        cuts.append(min(max(i, cuts[-1]), S))
    cuts.append(S)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def take_shard(blob: np.ndarray, off: np.ndarray, rank: int, world_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """→ (blob view, offsets rebased to 0, index of the shard's first sample)."""
    lo, hi = shard_ranges(off, world_size)[rank]
    b0, b1 = int(off[lo]), int(off[hi])
    sub = blob[b0:b1] if b1 > b0 else np.zeros(1, np.uint8)
    return sub, (off[lo:hi + 1] - off[lo]).astype(np.uint64), lo


class Collective:
    """Sum all-reduce of the count vectors through the default process group.

    Host vectors travel through a tensor on `device` (cuda:N under NCCL, cpu under gloo); device
    tensors (the pruner's device-resident path) are reduced in place.  Integer counts are reduced as
    int64 (exact)."""

    def __init__(self, device: Optional[str] = None, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.torch, self.dist, self.group = torch, dist, group
        backend = dist.get_backend(group)
        self.device = device or ("cuda" if backend == "nccl" else "cpu")
        self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)

    def allreduce_tensor(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce(self, v: np.ndarray) -> np.ndarray:
        torch = self.torch
        if v.dtype == np.uint64:
            t = torch.from_numpy(v.view(np.int64).copy()).to(self.device)
            return self.allreduce_tensor(t).cpu().numpy().view(np.uint64)
        t = torch.from_numpy(np.ascontiguousarray(v)).to(self.device)
        return self.allreduce_tensor(t).cpu().numpy()

    def broadcast_u32(self, v, src: int = 0) -> np.ndarray:
        """Rank `src`'s uint32 vector on every rank (the others pass None): length first, then the values."""
        torch = self.torch
        n = torch.tensor([len(v) if self.rank == src else 0], dtype=torch.int64, device=self.device)
        self.dist.broadcast(n, src, group=self.group)
        if self.rank == src:
            t = torch.from_numpy(np.ascontiguousarray(v, np.uint32).view(np.int32).copy()).to(self.device)
        else:
            t = torch.empty(int(n[0]), dtype=torch.int32, device=self.device)
        self.dist.broadcast(t, src, group=self.group)
        return t.cpu().numpy().view(np.uint32)

    def sum_int(self, x: int) -> int:
        return int(self.allreduce(np.array([x], np.uint64))[0])

    def __call__(self, v: np.ndarray) -> np.ndarray:
        return self.allreduce(v)


def bind_to_gpu_numa_node(device: int) -> Optional[int]:
    """Pin the calling process (and the threads it starts later) to the CPUs of the NUMA node the GPU hangs off, so
    that pinned host buffers are allocated next to the GPU's PCIe root (first touch) and the H2D / D2H copies of
    several ranks do not cross the socket interconnect.  Returns the node, or None when it cannot be determined
    (single-node machine, sysfs not available): then nothing is changed."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bus = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{getattr(p, 'pci_device_id', 0):02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
