import hashlib, os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = ["bench.py"]
import bench
from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200._abi import MERGE_DTYPE, bpe_stats, p32, p64
import torch
size = int(os.environ.get("SIZE", "1000000000")); merges = int(os.environ.get("MERGES", "32000"))
lib = _abi.load_library()
text, off = bench.synth(lib, size, 43)
lut, alphabet = bench.alphabet_lut(text)
ids = torch.from_numpy(lut[text]).cuda(); del text
h = C.c_void_p(); assert lib.bpe_create(0, C.byref(h)) == 0
len16 = np.ones(len(alphabet), dtype=np.int32)
log = np.zeros(merges, dtype=MERGE_DTYPE); nd = C.c_int64()
sweep = os.environ.get("ENVSWEEP", "").split()  # e.g. "BPE_LOOP_PREFETCH=0 BPE_LOOP_PREFETCH=1": rep i runs with entry i % len
for rep in range(int(os.environ.get("REPS", "3"))):
    if sweep:
        for kv in sweep[rep % len(sweep)].split(","):  # "A=1,B=2": several knobs per rep
            k, v = kv.split("=")
            os.environ[k] = v
        print("rep", rep, "with", sweep[rep % len(sweep)], file=sys.stderr)
    t0 = time.time()
    lib.bpe_clear_corpus(h); lib.bpe_set_tokens(h, p32(len16), len(len16)); lib.bpe_load_merges(h, None, 0)
    assert lib.bpe_add_documents_dev(h, C.c_void_p(ids.data_ptr()), p64(off), len(off) - 1) == 0
    t1 = time.time()
    rc = lib.bpe_merge_until(h, 2, 0, merges, log.ctypes.data_as(C.c_void_p), merges, C.byref(nd))
    t2 = time.time()
    s = bpe_stats(); lib.bpe_get_stats(h, C.byref(s))
    print("rep", rep, "rc", rc, "ingest %.1f ms" % ((t1 - t0) * 1e3), "merge_until wall %.1f ms" % ((t2 - t1) * 1e3), "k1 total ms %.1f" % s.ms_index_build,
          "phases", [round(x, 1) for x in s.ms_loop_phase], "launches", s.kernel_launches, "rebuilds", s.hot_rebuilds,
          "sha1", hashlib.sha1(log[: nd.value].tobytes()).hexdigest()[:12], file=sys.stderr)
