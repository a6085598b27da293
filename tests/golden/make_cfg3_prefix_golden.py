"""Generates tests/golden/cfg3_prefix.json: the FIRST merges of BASELINE config 3 (the full 1 GB seeded Zipf corpus) as the CPU
restatement of core.ts computes them (oracle/int_oracle.cpp; ~8 s per merge on one core, 5 GB of RAM) -- SURVEY.md section 8(d),
parity at scale (1): "literal-oracle prefix".  tests/test_oracle_golden.py compares the committed merge table of the GPU run
(tools/data/merges_cfg3_abc.npy, written by tools/dump_merges.py) with it.  Run from the repo root:
    python tests/golden/make_cfg3_prefix_golden.py [merges=16]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from bpe_tokenizer_b200 import _abi, synth  # noqa: E402
from oracle.int_oracle import IntOracle  # noqa: E402

TRAIN_BYTES = 1_000_000_000


def main(merges=16):
    sys.argv = ["bench.py"]
    import bench

    lib = _abi.load_library()
    text, off = bench.synth(lib, TRAIN_BYTES, synth.TRAIN_SEED)
    lut, alphabet = bench.alphabet_lut(text)
    ids = lut[text]
    n_bytes, n_docs = int(text.size), len(off) - 1
    del text
    o = IntOracle()
    o.set_len16(np.ones(len(alphabet) + merges + 1, dtype=np.int32))
    o.add_documents(ids, off)
    del ids
    t0 = time.time()
    la, lb, lw = o.merge_until(2, 0, merges, len(alphabet), merges)
    out = {
        "workload": "cfg3: %d B Zipf-word corpus (seed %d, %d docs), first %d merges" % (n_bytes, synth.TRAIN_SEED, n_docs, merges),
        "alphabet": [int(x) for x in alphabet],
        "merges": [[int(a), int(b), len(alphabet) + i, int(w)] for i, (a, b, w) in enumerate(zip(la, lb, lw))],
        "oracle_seconds": round(time.time() - t0, 1),
    }
    with open(os.path.join(ROOT, "tests", "golden", "cfg3_prefix.json"), "w") as f:
        json.dump(out, f)
    print(out)


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:2]))
