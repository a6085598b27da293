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

    mp.spawn(mg_worker._spawn_entry, args=(2, 29533, {"zipf_bytes": 400_000, "zipf_merges": 300, "fuzz_cases": 6}), nprocs=2, join=True)
