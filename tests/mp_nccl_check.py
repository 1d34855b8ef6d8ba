"""Worker of tests/test_gpu_multi.py: two (or more) NCCL ranks, the REAL kernels on a shard of ONE corpus each, one
integer all-reduce per pass — compared, on rank 0, with the same corpus on one GPU (src/prune.rs:104-112,231-236: the
RwLock merges the all-reduce replaces).  Launched with torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from tokengeex_b200 import _native as N, prune as P, synth
    from tokengeex_b200.dist import Collective
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = int(os.environ.get("TGX_MP_BYTES", "24000000"))
    vb, vo = synth.corpus(synth.KIND_CODE_CJK, 4, 12_000_000)
    toks, sc, kp = synth.vocab(vb, vo, 4, 30000, 16, 0.05)
    vocab = P.Vocab(list(toks), np.asarray(sc, np.float64), np.asarray(kp, np.uint8))
    blob, off, first, n_samples = synth.corpus_shard(synth.KIND_CODE_CJK, 4, total, rank, world)
    coll = Collective(device=f"cuda:{local}")
    for dropout in (0.0, 0.05):
        pr = P.ModelVocabularyPruner(20000, 0.8, 1, dropout, device=local, allreduce=coll, n_samples_global=n_samples,
                                     dropout_seed=77, byte_base=int(synth.corpus_offsets(4, total)[first]))
        model = N.Model(vocab.tokens, vocab.scores, device=local)
        d = pr._upload(blob, off)
        ex = pr.run_e_step(model, blob, off, d)
        rep = P.PruneReport()
        pruned = pr.prune_vocab(model, vocab, blob, off, rep, d)
        fr = pr.last_freq
        model.close()
        if rank == 0:  # the whole corpus on this GPU alone
            wb, wo = synth.corpus(synth.KIND_CODE_CJK, 4, total)
            pr1 = P.ModelVocabularyPruner(20000, 0.8, 1, dropout, device=local, n_samples_global=n_samples,
                                          dropout_seed=77)
            m1 = N.Model(vocab.tokens, vocab.scores, device=local)
            d1 = pr1._upload(wb, wo)
            ex1 = pr1.run_e_step(m1, wb, wo, d1)
            rep1 = P.PruneReport()
            pruned1 = pr1.prune_vocab(m1, vocab, wb, wo, rep1, d1)
            m1.close()
            assert np.array_equal(ex.view(np.uint64), ex1.view(np.uint64)), f"expected counts differ (dropout {dropout})"
            assert np.array_equal(fr, pr1.last_freq), "frequencies differ"
            assert pruned.tokens == pruned1.tokens, "pruned vocabularies differ"
            print(f"OK dropout={dropout} world={world} V={len(vocab)} pruned={len(pruned)} sum={float(ex.sum()):.6f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
