"""Sharded training only (no encode leg): python -m torch.distributed.run --nproc-per-node N tools/mg_bench.py"""
import os, sys, time, ctypes as C, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
sys.argv = ["bench.py"]
import bench
from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200._abi import MERGE_DTYPE, p32, p64
from bpe_tokenizer_b200.sharded import exchange_pair_counts, shard_bounds
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
size = int(os.environ.get("SIZE", "1000000000")); merges = int(os.environ.get("MERGES", "32000")); reps = int(os.environ.get("REPS", "2"))
lib = _abi.load_library()
text, off = bench.synth(lib, size, 43)
lut, alphabet = bench.alphabet_lut(text)
b = shard_bounds(np.diff(off), world); lo, hi = b[rank], b[rank + 1]
ids = torch.from_numpy(lut[text[off[lo]:off[hi]]]).cuda(); offs = np.ascontiguousarray(off[lo:hi + 1] - off[lo]); del text
h = C.c_void_p(); assert lib.bpe_create(local, C.byref(h)) == 0
if world > 1:
    handle = C.create_string_buffer(64); assert lib.bpe_mg_init(h, rank, world, handle) == 0
    hs = [b""] * world; dist.all_gather_object(hs, bytes(handle.raw)); assert lib.bpe_mg_connect(h, b"".join(hs)) == 0; dist.barrier()
len16 = np.ones(len(alphabet), dtype=np.int32); log = np.zeros(merges, dtype=MERGE_DTYPE); nd = C.c_int64()
for rep in range(reps):
    lib.bpe_clear_corpus(h); lib.bpe_set_tokens(h, p32(len16), len(len16)); lib.bpe_load_merges(h, None, 0)
    torch.cuda.synchronize(); t0 = time.time()
    assert lib.bpe_add_documents_dev(h, C.c_void_p(ids.data_ptr()), p64(offs), len(offs) - 1) == 0
    exchange_pair_counts(lib, h, rank, world, torch.device("cuda", local))
    torch.cuda.synchronize(); t1 = time.time()
    rc = lib.bpe_merge_until(h, 2, 0, merges, log.ctypes.data_as(C.c_void_p), merges, C.byref(nd))
    t2 = time.time()
    if rank == 0:
        print("rep", rep, "rc", rc, "world", world, "ingest+K1+exchange %.1f ms" % ((t1 - t0) * 1e3), "merge_until %.1f ms" % ((t2 - t1) * 1e3),
              "merges", nd.value, "sha", hashlib.sha1(log[:nd.value].tobytes()).hexdigest()[:12], flush=True)
lib.bpe_destroy(h)
if world > 1: dist.destroy_process_group()
