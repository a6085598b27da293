"""Train cfg3 (1 GB Zipf corpus, 32k merges) on the GPU and save the merge list (prototyping input for encode work)."""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = ["bench.py"]
import bench
from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200._abi import MERGE_DTYPE, p32, p64
import torch
size = int(os.environ.get("SIZE", "1000000000")); merges = int(os.environ.get("MERGES", "32000"))
lib = _abi.load_library()
text, off = bench.synth(lib, size, 43)
lut, alphabet = bench.alphabet_lut(text)
ids = torch.from_numpy(lut[text]).cuda(); del text
h = C.c_void_p(); assert lib.bpe_create(0, C.byref(h)) == 0
len16 = np.ones(len(alphabet), dtype=np.int32)
log = np.zeros(merges, dtype=MERGE_DTYPE); nd = C.c_int64()
lib.bpe_set_tokens(h, p32(len16), len(len16))
assert lib.bpe_add_documents_dev(h, C.c_void_p(ids.data_ptr()), p64(off), len(off) - 1) == 0
t1 = time.time()
rc = lib.bpe_merge_until(h, 2, 0, merges, log.ctypes.data_as(C.c_void_p), merges, C.byref(nd))
print("rc", rc, "merges", nd.value, "wall", time.time() - t1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "merges_%d_%d.npy" % (size, merges)), log[: nd.value])
np.save(os.path.join(ROOT, "gpurun_out", "alphabet_%d.npy" % size), np.array(alphabet, dtype=np.int32))
