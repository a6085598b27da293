"""Generates tests/golden/cfg2_merge_log.json: the merge log of BASELINE config 2 (10 MB seeded Zipf corpus, 4 000 merges) as the CPU
restatement of core.ts computes it (oracle/int_oracle.cpp, ~5 minutes on one core), in the byte layout bench.py hashes
(`bpe_merge` records: a, b, c, reserved, weight).  The GPU test test_cfg2_full_size_merge_log_matches_the_oracle compares the
product's log with it.  Run from the repo root:  python tests/golden/make_cfg2_golden.py"""
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
from oracle.int_oracle import IntOracle  # noqa: E402

TRAIN_BYTES, MERGES = 10_000_000, 4000


def main():
    sys.argv = ["bench.py"]
    import bench

    lib = _abi.load_library()
    text, off = bench.synth(lib, TRAIN_BYTES, synth.TRAIN_SEED)
    lut, alphabet = bench.alphabet_lut(text)
    ids = lut[text]
    o = IntOracle()
    o.set_len16(np.ones(len(alphabet) + MERGES + 1, dtype=np.int32))
    o.add_documents(ids, off)
    t0 = time.time()
    la, lb, lw = o.merge_until(2, 0, MERGES, len(alphabet), MERGES)
    log = np.zeros(len(la), dtype=MERGE_DTYPE)
    log["a"], log["b"], log["weight"] = la, lb, lw
    log["c"] = len(alphabet) + np.arange(len(la))
    out = {
        "workload": "cfg2: %d B Zipf-word corpus (seed %d, %d docs), mergeUntil to %d merges" % (text.size, synth.TRAIN_SEED, len(off) - 1, MERGES),
        "alphabet": len(alphabet), "merges": int(len(la)), "sha1": hashlib.sha1(log.tobytes()).hexdigest(),
        "first": [[int(a), int(b), int(w)] for a, b, w in zip(la[:8], lb[:8], lw[:8])],
        "last": [[int(a), int(b), int(w)] for a, b, w in zip(la[-4:], lb[-4:], lw[-4:])],
        "weights_sha1": hashlib.sha1(np.asarray(lw, dtype=np.int64).tobytes()).hexdigest(),
        "oracle_seconds": round(time.time() - t0, 1),
    }
    with open(os.path.join(ROOT, "tests", "golden", "cfg2_merge_log.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
