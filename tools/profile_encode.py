"""Short encode-only run for ncu: train a table on 4 MB (1500 merges), encode 20 MB of unseen text."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
sys.argv = ["bench.py"]
import bench
from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200._abi import MERGE_DTYPE, bpe_stats, p32, p64
lib = _abi.load_library()
train = int(os.environ.get("TRAIN", 4_000_000)); merges = int(os.environ.get("MERGES", 1500)); enc = int(os.environ.get("ENC", 20_000_000))
text, off = bench.synth(lib, train, 43)
lut, alphabet = bench.alphabet_lut(text)
ids = torch.from_numpy(lut[text]).cuda()
h = C.c_void_p(); assert lib.bpe_create(0, C.byref(h)) == 0
len16 = np.ones(len(alphabet), dtype=np.int32)
lib.bpe_set_tokens(h, p32(len16), len(len16))
assert lib.bpe_add_documents_dev(h, C.c_void_p(ids.data_ptr()), p64(off), len(off) - 1) == 0
log = np.zeros(merges, dtype=MERGE_DTYPE); nd = C.c_int64()
assert lib.bpe_merge_until(h, 2, 0, merges, log.ctypes.data_as(C.c_void_p), merges, C.byref(nd)) == 0
text2, off2 = bench.synth(lib, enc, 44)
ids2 = torch.from_numpy(lut[text2]).cuda(); off2d = torch.from_numpy(off2).cuda()
out = torch.empty(ids2.numel(), dtype=torch.int32, device="cuda"); ooff = torch.empty(len(off2), dtype=torch.int64, device="cuda")
n_out = C.c_int64()
for rep in range(3):
    assert lib.bpe_encode_batch_dev(h, C.c_void_p(ids2.data_ptr()), C.c_void_p(off2d.data_ptr()), len(off2) - 1, ids2.numel(), int(np.diff(off2).max()),
                                    None, 0, C.c_void_p(out.data_ptr()), C.c_void_p(ooff.data_ptr()), None, C.byref(n_out)) == 0
    s = bpe_stats(); lib.bpe_get_stats(h, C.byref(s))
    print("encode %d chars -> %d tokens in %.2f ms = %.2f GB/s" % (ids2.numel(), n_out.value, s.ms_encode, ids2.numel() / s.ms_encode / 1e6))
