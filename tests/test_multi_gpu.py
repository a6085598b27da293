"""GPU, >= 2 devices: exact training on a corpus sharded by document (bpe_tokenizer_b200/sharded.py, csrc/mg_kernels.cuh)
against the CPU oracle.  Skipped on single-GPU boxes."""
import pytest

pytestmark = pytest.mark.gpu


def test_sharded_training_matches_oracle_world2():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import mg_worker

    mp.spawn(mg_worker._spawn_entry, args=(2, 29533, {"zipf_bytes": 400_000, "zipf_merges": 300, "fuzz_cases": 6, "weak_bytes": 250_000, "weak_merges": 250}), nprocs=2, join=True)


def test_two_engines_on_two_devices_in_one_process():
    """bpe_create(device) allows engines on different GPUs inside one process: per-device state (the dynamic shared memory
    limit of the encode kernels, the encode scratch) must belong to the engine, not to the process or the thread."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from bpe_tokenizer_b200 import BPETokenizer
    from oracle import LiteralTokenizer

    docs = ["the quick brown fox jumps over the lazy dog " * 3, "pack my box with five dozen liquor jugs", "aaaa bbbb aaaa", ""]
    lit = LiteralTokenizer()
    gpus = [BPETokenizer(0), BPETokenizer(1)]
    for t in [lit] + gpus:
        for d in docs:
            t.addToCorpus(d)
        t.mergeUntil({"max_iterations": 40})
    for g in gpus:
        assert g.toJSON() == lit.toJSON()
    for rep in range(3):  # alternate between the devices: scratch and attributes of one must not leak into the other
        for g in gpus:
            for d in docs[:3]:
                assert g.encodeToVector(d) == lit.encodeToVector(d)
                assert g.decodeVector(lit.encodeToVector(d)) == d
    for g in gpus:
        g.close()
