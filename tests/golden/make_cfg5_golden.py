"""Generates tests/golden/cfg5_merge_log.json: the merge log of BASELINE config 5 (256 MB seeded Zipf corpus, max_length 8, 8 192
merges) by the incremental CPU oracle (oracle/fast_oracle.cpp, pinned to the literal restatement).  tools/run_cfg5.py compares
the GPU's uninterrupted run with it (same hash as its `sha()`).  Run from the repo root:  python tests/golden/make_cfg5_golden.py"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bpe_tokenizer_b200 import _abi, synth  # noqa: E402
from oracle.fast_oracle import FastOracle  # noqa: E402

TRAIN_BYTES, MERGES, MAX_LENGTH = 256_000_000, 8192, 8


def main():
    sys.argv = ["bench.py"]
    import bench

    lib = _abi.load_library()
    text, off = bench.synth(lib, TRAIN_BYTES, synth.TRAIN_SEED)
    lut, alphabet = bench.alphabet_lut(text)
    ids = lut[text]
    n_bytes, n_docs = int(text.size), len(off) - 1
    del text
    o = FastOracle()
    o.set_len16(np.ones(len(alphabet), dtype=np.int32))
    t0 = time.time()
    o.add_documents(ids, off)
    del ids
    la, lb, lw = o.merge_until(2, MAX_LENGTH, MERGES, len(alphabet), MERGES)
    rows = [[int(a), int(b), int(w)] for a, b, w in zip(la, lb, lw)]
    out = {"workload": "cfg5: %d B Zipf-word corpus (seed %d, %d docs), mergeUntil({max_length: %d}) to %d merges" % (n_bytes, synth.TRAIN_SEED, n_docs, MAX_LENGTH, MERGES),
           "merges": len(rows), "sha16_of_repr": hashlib.sha1(repr(rows).encode()).hexdigest()[:16], "first": rows[:8], "last": rows[-4:],
           "tokens_left": int(o.L.fast_total_tokens(o.h)), "oracle_seconds": round(time.time() - t0, 1)}
    with open(os.path.join(ROOT, "tests", "golden", "cfg5_merge_log.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
