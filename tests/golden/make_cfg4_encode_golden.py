"""Generates tests/golden/cfg4_encode.json: BASELINE config 4 at FULL size -- the 1 GB seed-44 text encoded with the 32 000 merges of
cfg3 by the LITERAL CPU restatement of encodeToCode (core.ts:404-406: every merge in order, replaceAll over the document; one call per
document), spread over the host cores (about 3.5 core-hours).  Recorded: the number of tokens, a SHA-1 over the per-document token
counts and a SHA-1 over the token stream (per block of 65 536 documents, then over the block digests), which bench.py prints for the
GPU's output of the same text as `encode.output_sha1`.  Run from the repo root:  python tests/golden/make_cfg4_encode_golden.py [workers=7]"""
import hashlib
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
BLOCK = 65536
ENCODE_BYTES = 1_000_000_000
_state = {}


def stream_sha1(values: np.ndarray, offsets: np.ndarray) -> str:
    """SHA-1 over the per-block SHA-1 digests of the int32 token stream (blocks of BLOCK documents) -- shared with bench.py"""
    h = hashlib.sha1()
    n_docs = len(offsets) - 1
    for d0 in range(0, n_docs, BLOCK):
        d1 = min(n_docs, d0 + BLOCK)
        h.update(hashlib.sha1(np.ascontiguousarray(values[offsets[d0]:offsets[d1]], dtype=np.int32).tobytes()).digest())
    return h.hexdigest()


def _init():
    from bpe_tokenizer_b200 import _abi, synth
    from oracle.int_oracle import IntOracle

    sys.argv = ["bench.py"]
    import bench

    lib = _abi.load_library()
    text, off = bench.synth(lib, ENCODE_BYTES, synth.ENCODE_SEED)
    alphabet = np.load(os.path.join(ROOT, "tools", "data", "alphabet_cfg3.npy"))
    lut = np.full(256, -1, dtype=np.int32)
    lut[alphabet] = np.arange(len(alphabet), dtype=np.int32)
    o = IntOracle()
    o.load_merges(np.load(os.path.join(ROOT, "tools", "data", "merges_cfg3_abc.npy")))
    _state.update(text=text, off=off, lut=lut, o=o)


def _block(b):
    text, off, lut, o = _state["text"], _state["off"], _state["lut"], _state["o"]
    n_docs = len(off) - 1
    d0, d1 = b * BLOCK, min(n_docs, (b + 1) * BLOCK)
    outs = [o.encode(lut[text[off[d]:off[d + 1]]]) for d in range(d0, d1)]
    lens = np.array([x.size for x in outs], dtype=np.int32)
    stream = np.concatenate(outs) if outs else np.zeros(0, np.int32)
    return b, hashlib.sha1(stream.astype(np.int32).tobytes()).digest(), lens


def main(workers=7):
    from bpe_tokenizer_b200 import _abi, synth

    sys.argv = ["bench.py"]
    import bench

    lib = _abi.load_library()
    text, off = bench.synth(lib, ENCODE_BYTES, synth.ENCODE_SEED)
    n_bytes, n_docs = int(text.size), len(off) - 1
    del text
    n_blocks = (n_docs + BLOCK - 1) // BLOCK
    t0 = time.time()
    digests, lens = [None] * n_blocks, [None] * n_blocks
    with Pool(workers, initializer=_init) as pool:
        for b, dg, ln in pool.imap_unordered(_block, range(n_blocks)):
            digests[b], lens[b] = dg, ln
            done = sum(x is not None for x in digests)
            if done % 8 == 0:
                print("blocks %d/%d after %.0f s" % (done, n_blocks, time.time() - t0), flush=True)
    h = hashlib.sha1()
    for dg in digests:
        h.update(dg)
    lens = np.concatenate(lens)
    out = {"workload": "cfg4: %d B Zipf-word text (seed %d, %d docs) encoded with the 32000 merges of cfg3" % (n_bytes, synth.ENCODE_SEED, n_docs),
           "oracle": "oracle/int_oracle.cpp orc_encode (literal encodeToCode), one call per document", "tokens_out": int(lens.astype(np.int64).sum()),
           "doc_lengths_sha1": hashlib.sha1(lens.astype(np.int32).tobytes()).hexdigest(), "output_sha1": h.hexdigest(), "block_docs": BLOCK,
           "oracle_core_seconds": round((time.time() - t0) * workers, 0)}
    with open(os.path.join(ROOT, "tests", "golden", "cfg4_encode.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:2]))
