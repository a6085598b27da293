"""Multi-GPU parity worker (one process per GPU): ShardedBPETokenizer against the CPU oracle.  Run by
tests/test_multi_gpu.py (torch.multiprocessing spawn) or directly:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mg_worker.py"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _merges(t):
    return [[a.index, b.index, c.weight] for a, b, c in t.merge_tokens]


def run_checks(rank: int, world: int, zipf_bytes: int = 1_000_000, zipf_merges: int = 500, fuzz_cases: int = 10, weak_bytes: int = 0,
               weak_merges: int = 300) -> None:
    import torch

    from bpe_tokenizer_b200.sharded import ShardedBPETokenizer
    from oracle import LiteralTokenizer
    from oracle.int_oracle import IntOracleTokenizer

    dev = torch.cuda.current_device()
    # ---- fuzz: small alphabets, many ties, runs, empty documents; every rank sees every document ----
    for case in range(fuzz_cases):
        rng = random.Random(9000 + case)
        alphabet = "ab" if case % 3 == 0 else "abcde "
        docs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, rng.choice([20, 100, 400])))) for _ in range(rng.randint(1, 9))]
        opts = {"min_weight": rng.choice([None, 2, 4]), "max_length": rng.choice([None, 5, 9]), "max_iterations": rng.choice([None, None, 7])}
        lit, gpu = LiteralTokenizer(), ShardedBPETokenizer(dev)
        for d in docs:
            lit.addToCorpus(d)
            gpu.addToCorpus(d)
        if case == 0 and world > 1:
            # a shard alone cannot answer for the corpus: the single-step calls are refused, in the class and in the C ABI
            import ctypes as C

            from bpe_tokenizer_b200 import BpeError, _abi

            try:
                gpu.findNextMerge()
                raise AssertionError("findNextMerge on a sharded corpus must raise")
            except BpeError:
                pass
            m, found = _abi.bpe_merge(), C.c_int()
            assert gpu._lib.bpe_find_next_merge(gpu._h, 2, 0, C.byref(m), C.byref(found)) == _abi.BPE_E_INVALID
        lit.mergeUntil(opts)
        gpu.mergeUntil(opts)
        assert _merges(gpu) == _merges(lit), (case, rank, _merges(gpu)[:8], _merges(lit)[:8])
        assert gpu.toJSON() == lit.toJSON(), (case, rank)
        ids, off = gpu.corpusIdsAllRanks()
        codes = (ids.astype(np.uint32) + 1).tolist()
        got = ["".join(map(chr, codes[off[d]:off[d + 1]])) for d in range(len(off) - 1)]
        assert got == lit.corpus_in_code, (case, rank)
        for d in docs[:3]:
            try:
                want = lit.encodeToVector(d)
            except ValueError:
                continue
            assert gpu.encodeToVector(d) == want
        gpu.close()
    # ---- Zipf corpus: bulk local shards, merge log against the compiled oracle ----
    if zipf_bytes:
        from bpe_tokenizer_b200.sharded import shard_bounds
        from bpe_tokenizer_b200.synth import first_appearance_ids, synth_corpus

        text, off = synth_corpus(zipf_bytes, seed=43)
        ids, alphabet = first_appearance_ids(text)
        gpu, orc = ShardedBPETokenizer(dev), IntOracleTokenizer()
        for t in (gpu, orc):
            t.addToCorpus("".join(chr(c) for c in alphabet))
            if hasattr(t, "_pending"):
                t._pending = []
            else:
                t._o.clear_corpus()
            for tk in t.token_table:
                tk.weight = 0
                tk.original_weight = 0
        b = shard_bounds(np.diff(off), world)
        lo, hi = b[rank], b[rank + 1]
        gpu.addDocuments(ids[off[lo]:off[hi]], off[lo:hi + 1] - off[lo], local_shard=True)
        orc.add_ids(ids, off)
        n1 = gpu.mergeUntil({"max_iterations": zipf_merges})
        n2 = orc.mergeUntil({"max_iterations": zipf_merges})
        assert n1 == n2 == zipf_merges, (n1, n2)
        assert _merges(gpu) == _merges(orc), rank
        assert gpu.toJSON() == orc.toJSON(), rank
        got_ids, _ = gpu.corpusIdsAllRanks()
        want = np.concatenate([orc._o.document(d) for d in range(orc._o.num_documents())])
        assert np.array_equal(got_ids, want), rank
        s = gpu.stats()
        if rank == 0:
            print("mg parity ok: world %d, %d merges on %d B, ties %d, merge_until %.1f ms" % (world, n1, zipf_bytes, s["tie_breaks"], s["ms_last_merge_until"]), flush=True)
        gpu.close()
    # ---- weak construction (bench.py --workload cfg3w): every rank GENERATES its own shard (seed 43 + 1000 * rank), the corpus
    # is the concatenation in rank order; token indices follow first appearance in that concatenation ----
    if weak_bytes:
        from bpe_tokenizer_b200.synth import first_appearance_ids, synth_corpus

        shards = [synth_corpus(weak_bytes, seed=43 + 1000 * r) for r in range(world)]
        text = np.concatenate([t for t, _ in shards])
        base = np.cumsum([0] + [t.size for t, _ in shards])
        off = np.concatenate([shards[0][1]] + [o[1:] + base[r] for r, (_, o) in enumerate(shards) if r > 0])
        ids, alphabet = first_appearance_ids(text)
        gpu, orc = ShardedBPETokenizer(dev), IntOracleTokenizer()
        for t in (gpu, orc):
            t.addToCorpus("".join(chr(c) for c in alphabet))
            if hasattr(t, "_pending"):
                t._pending = []
            else:
                t._o.clear_corpus()
            for tk in t.token_table:
                tk.weight = 0
                tk.original_weight = 0
        my_off = shards[rank][1]
        gpu.addDocuments(ids[base[rank]:base[rank + 1]], my_off, local_shard=True)
        orc.add_ids(ids, off)
        n1 = gpu.mergeUntil({"max_iterations": weak_merges})
        n2 = orc.mergeUntil({"max_iterations": weak_merges})
        assert n1 == n2 == weak_merges, (n1, n2)
        assert _merges(gpu) == _merges(orc), rank
        assert gpu.toJSON() == orc.toJSON(), rank
        if rank == 0:
            print("mg weak parity ok: world %d, %d merges on %d x %d B" % (world, n1, world, weak_bytes), flush=True)
        gpu.close()


def _spawn_entry(rank: int, world: int, port: int, kwargs: dict) -> None:
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        run_checks(rank, world, **kwargs)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    run_checks(rank, world, zipf_bytes=int(os.environ.get("MG_ZIPF", "1000000")), zipf_merges=int(os.environ.get("MG_MERGES", "500")),
               fuzz_cases=int(os.environ.get("MG_FUZZ", "10")))
    dist.destroy_process_group()
