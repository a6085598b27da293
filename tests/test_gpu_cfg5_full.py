"""GPU, full size: BASELINE.json config 5 -- max_length=8 constrained merging with resume on the 256 MB corpus.

  A: uninterrupted mergeUntil({max_length: 8, max_iterations: 8192})
  B: 4096 merges -> toJSON -> new engine fromJSON -> restoreToCorpus (every document) -> 4096 more   (core.ts:213-216)
  C: addToCorpus + restoreMerge(compactMerge(...)) for the first 4096 -> 4096 more                     (core.ts:477-494)

A must equal the incremental CPU oracle's merge log (tests/golden/cfg5_merge_log.json, generator make_cfg5_golden.py;
the oracle is pinned to the literal restatement by tests/test_oracle_golden.py); B and C must equal A: merges, weights,
token table, merge codes and the final corpus.  The result lines go to gpurun_out/cfg5_full.log (kept under profiles/)."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "cfg5_merge_log.json")
SIZE, HALF = 256_000_000, 4096


def _rows(t):
    return [[a.index, b.index, c.original_weight] for a, b, c in t.merge_tokens]


def _sha(t):
    return hashlib.sha1(repr(_rows(t)).encode()).hexdigest()[:16]


def test_cfg5_full_size_resume_routes_equal_the_cpu_oracle():
    from bpe_tokenizer_b200 import BPETokenizer, synth

    golden = json.load(open(GOLDEN))
    text, off = synth.native_corpus(SIZE, seed=synth.TRAIN_SEED)
    text = np.asarray(text, dtype=np.uint8)
    off = np.asarray(off, dtype=np.int64)
    # first-appearance alphabet (core.ts:186-199)
    _, first_pos = np.unique(text, return_index=True)
    alphabet = text[np.sort(first_pos)]
    lut = np.full(256, -1, dtype=np.int32)
    lut[alphabet] = np.arange(alphabet.size, dtype=np.int32)
    ids = lut[text]
    assert "%d B" % text.size in golden["workload"] and "%d docs" % (off.size - 1) in golden["workload"]
    del text
    lines = []

    def say(msg):
        lines.append(msg)
        print(msg, flush=True)

    def fresh():
        t = BPETokenizer()
        t.addToCorpus("".join(chr(int(c)) for c in alphabet))
        t.corpus_in_code = []
        for tk in t.token_table:
            tk.weight = 0
            tk.original_weight = 0
        return t

    opts = {"max_length": 8}
    t0 = time.time()
    A = fresh()
    A.addDocuments(ids, off)
    n = A.mergeUntil(dict(opts, max_iterations=2 * HALF))
    say("A uninterrupted: %d merges, %.1f s, sha %s (CPU oracle: %s)" % (n, time.time() - t0, _sha(A), golden["sha16_of_repr"]))
    assert n == golden["merges"] == 2 * HALF
    assert _rows(A)[:8] == golden["first"] and _rows(A)[-4:] == golden["last"]
    assert _sha(A) == golden["sha16_of_repr"]
    ja = A.toJSON()
    ia, oa = A.corpusIds()
    assert int(ia.size) == golden["tokens_left"]
    A.close()

    t0 = time.time()
    B1 = fresh()
    B1.addDocuments(ids, off)
    assert B1.mergeUntil(dict(opts, max_iterations=HALF)) == HALF
    snap = json.loads(json.dumps(B1.toJSON()))
    log = [[a.code, b.code, c.original_weight] for a, b, c in B1.merge_tokens]
    assert _rows(B1) == _rows_prefix(ja, HALF)
    B1.close()
    B = BPETokenizer()
    B.fromJSON(snap)
    B.restoreDocuments(ids, off)
    assert B.mergeUntil(dict(opts, max_iterations=HALF)) == HALF
    say("B toJSON -> fromJSON -> restoreToCorpus -> continue: %.1f s, sha %s" % (time.time() - t0, _sha(B)))
    assert _sha(B) == golden["sha16_of_repr"]
    jb = B.toJSON()
    ib, ob = B.corpusIds()
    B.close()
    assert [r[0] for r in jb["token_table"]] == [r[0] for r in ja["token_table"]]
    assert jb["merge_codes"] == ja["merge_codes"]
    assert jb == ja  # weights included: the snapshot carried them, the second half replayed the same merges
    assert np.array_equal(ia, ib) and np.array_equal(oa, ob)

    t0 = time.time()
    C = fresh()
    C.addDocuments(ids, off)
    C.restoreMerges(log)
    assert C.mergeUntil(dict(opts, max_iterations=HALF)) == HALF
    say("C addToCorpus + restoreMerges(log) -> continue: %.1f s, sha %s" % (time.time() - t0, _sha(C)))
    jc = C.toJSON()
    ic, oc = C.corpusIds()
    C.close()
    assert _sha(C) == golden["sha16_of_repr"]
    assert jc == ja  # the full snapshot: characters counted by addToCorpus, weights replayed by restoreMerge
    assert np.array_equal(ia, ic) and np.array_equal(oa, oc)
    say("cfg5 full size: A == CPU oracle, B == A, C == A (merges, weights, tables, corpus of %d tokens)" % ia.size)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "cfg5_full.log"), "w") as f:
        f.write("\n".join(lines) + "\n")


def _rows_prefix(snapshot, n):
    """[a.index, b.index, weight] of the first n merges of a toJSON snapshot (core.ts:112-127)."""
    table = snapshot["token_table"]
    return [[ord(a) - 1, ord(b) - 1, table[ord(c) - 1][2]] for a, b, c in snapshot["merge_codes"][:n]]
