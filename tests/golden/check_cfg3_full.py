"""Checks the GPU's FULL merge log of BASELINE config 3 (1 GB seeded Zipf corpus, 32 000 merges: tools/data/merges_cfg3_abc.npy +
merges_cfg3_weights.npy, SHA-1 add92aa1...) against the incremental CPU oracle (oracle/fast_oracle.cpp, itself pinned to the literal
restatement of core.ts by tests/test_oracle_golden.py) and writes tests/golden/cfg3_full_check.json.  ~40 GB of RAM, one core.
Run from the repo root:  python tests/golden/check_cfg3_full.py [bytes=1000000000] [merges=32000]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bpe_tokenizer_b200 import _abi, synth  # noqa: E402
from bpe_tokenizer_b200._abi import MERGE_DTYPE  # noqa: E402
from oracle.fast_oracle import FastOracle  # noqa: E402


def main(train_bytes=1_000_000_000, merges=32000):
    sys.argv = ["bench.py"]
    import bench

    lib = _abi.load_library()
    text, off = bench.synth(lib, train_bytes, synth.TRAIN_SEED)
    lut, alphabet = bench.alphabet_lut(text)
    ids = lut[text]
    n_bytes, n_docs = int(text.size), len(off) - 1
    del text
    o = FastOracle()
    o.set_len16(np.ones(len(alphabet), dtype=np.int32))
    t0 = time.time()
    o.add_documents(ids, off)
    del ids
    la, lb, lw = o.merge_until(2, 0, merges, len(alphabet), merges)
    seconds = time.time() - t0
    log = np.zeros(len(la), dtype=MERGE_DTYPE)
    log["a"], log["b"], log["weight"] = la, lb, lw
    log["c"] = len(alphabet) + np.arange(len(la))
    sha = hashlib.sha1(log.tobytes()).hexdigest()
    out = {"workload": "%d B Zipf-word corpus (seed %d, %d docs), mergeUntil to %d merges" % (n_bytes, synth.TRAIN_SEED, n_docs, merges),
           "oracle": "oracle/fast_oracle.cpp (incremental CPU oracle)", "merges": int(len(la)), "sha1": sha, "oracle_seconds": round(seconds, 1),
           "tokens_left": int(o.L.fast_total_tokens(o.h))}
    if train_bytes == 1_000_000_000:
        abc = np.load(os.path.join(ROOT, "tools", "data", "merges_cfg3_abc.npy"))[: len(la)]
        w = np.load(os.path.join(ROOT, "tools", "data", "merges_cfg3_weights.npy"))[: len(la)]
        same = np.array_equal(abc[:, 0], la) and np.array_equal(abc[:, 1], lb) and np.array_equal(w.astype(np.int64), lw)
        out["equals_gpu_merge_log"] = bool(same)
        if not same:
            bad = np.flatnonzero((abc[:, 0] != la) | (abc[:, 1] != lb) | (w.astype(np.int64) != lw))
            out["first_difference"] = int(bad[0])
        with open(os.path.join(ROOT, "tests", "golden", "cfg3_full_check.json"), "w") as f:
            json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:3]))
