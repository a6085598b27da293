"""Encode-only run with a saved merge table (tools/data/merges_cfg3_abc.npy = the 32 000 merges of cfg3): time
bpe_encode_batch_dev on ENC bytes of seed-44 Zipf text.  (Parity of the same path against the oracle: tests/test_gpu_parity.py.)"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
sys.argv = ["bench.py"]
import bench
from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200._abi import bpe_stats, p32, p64
lib = _abi.load_library()
enc = int(os.environ.get("ENC", 100_000_000)); reps = int(os.environ.get("REPS", 3))
abc = np.load(os.path.join(ROOT, "tools", "data", "merges_cfg3_abc.npy"))
nm = int(os.environ.get("MERGES", len(abc))); abc = np.ascontiguousarray(abc[:nm])
alphabet = np.load(os.path.join(ROOT, "tools", "data", "alphabet_cfg3.npy"))
lut = np.full(256, -1, dtype=np.int32); lut[alphabet] = np.arange(len(alphabet), dtype=np.int32)
h = C.c_void_p(); assert lib.bpe_create(0, C.byref(h)) == 0
n_tok = len(alphabet) + nm
len16 = np.ones(n_tok, dtype=np.int32)
assert lib.bpe_set_tokens(h, p32(len16), n_tok) == 0
assert lib.bpe_load_merges(h, p32(abc.reshape(-1)), nm) == 0, lib.bpe_last_error(h)
text2, off2 = bench.synth(lib, enc, 44)
ids2h = lut[text2]
ids2 = torch.from_numpy(ids2h).cuda(); off2d = torch.from_numpy(off2).cuda()
out = torch.empty(ids2.numel(), dtype=torch.int32, device="cuda"); ooff = torch.empty(len(off2), dtype=torch.int64, device="cuda")
n_out = C.c_int64()
maxdoc = int(np.diff(off2).max())
for rep in range(reps):
    rc = lib.bpe_encode_batch_dev(h, C.c_void_p(ids2.data_ptr()), C.c_void_p(off2d.data_ptr()), len(off2) - 1, ids2.numel(), maxdoc,
                                  None, 0, C.c_void_p(out.data_ptr()), C.c_void_p(ooff.data_ptr()), None, C.byref(n_out))
    assert rc == 0, lib.bpe_last_error(h)
    s = bpe_stats(); lib.bpe_get_stats(h, C.byref(s))
    print("encode %d chars (%d docs, max %d) -> %d tokens in %.2f ms = %.2f GB/s" % (ids2.numel(), len(off2) - 1, maxdoc, n_out.value, s.ms_encode, ids2.numel() / s.ms_encode / 1e6), flush=True)

# ---- host-buffer calls (three-stream pipeline): E2E_CHUNKS="16M 64M ..." times bpe_encode_batch / bpe_encode_text_batch from pinned
# buffers for each chunk size (BPE_ENC_CHUNK is read when an engine first encodes from host buffers, so one engine per setting)
if os.environ.get("E2E_CHUNKS"):
    ids_host = torch.from_numpy(ids2h).pin_memory()
    text_host = torch.from_numpy(np.ascontiguousarray(text2)).pin_memory()
    out_host = torch.empty(ids2h.size, dtype=torch.int32).pin_memory()
    ooff_host = np.zeros(len(off2), dtype=np.int64)
    for spec in os.environ["E2E_CHUNKS"].split():
        os.environ["BPE_ENC_CHUNK"] = str(int(float(spec[:-1]) * (1 << 20)) if spec.endswith("M") else int(spec))
        g = C.c_void_p(); assert lib.bpe_create(0, C.byref(g)) == 0
        assert lib.bpe_set_tokens(g, p32(len16), n_tok) == 0 and lib.bpe_load_merges(g, p32(abc.reshape(-1)), nm) == 0
        cps = np.asarray(alphabet, dtype=np.int32); idx = np.arange(len(alphabet), dtype=np.int32)
        assert lib.bpe_set_chars(g, p32(cps), p32(idx), len(cps)) == 0
        for name in ("ids", "text"):
            best = 1e9
            for rep in range(reps + 1):
                t0 = time.perf_counter()
                if name == "ids":
                    rc = lib.bpe_encode_batch(g, C.cast(ids_host.data_ptr(), _abi.i32p), p64(off2), len(off2) - 1, None, 0,
                                              C.cast(out_host.data_ptr(), _abi.i32p), out_host.numel(), p64(ooff_host), None, C.byref(n_out))
                else:
                    rc = lib.bpe_encode_text_batch(g, C.cast(text_host.data_ptr(), _abi.u8p), p64(off2), len(off2) - 1, None, 0,
                                                   C.cast(out_host.data_ptr(), _abi.i32p), out_host.numel(), p64(ooff_host), None, C.byref(n_out), None, None)
                dt = time.perf_counter() - t0
                assert rc == 0, lib.bpe_last_error(g)
                if rep: best = min(best, dt)
            s = bpe_stats(); lib.bpe_get_stats(g, C.byref(s))
            print("e2e chunk %s from %s: %.2f ms = %.2f GB/s (encode kernels %.2f ms, %d tokens)" % (spec, name, best * 1e3, ids2h.size / best / 1e9, s.ms_encode, n_out.value), flush=True)
        lib.bpe_destroy(g)
