"""BASELINE.json config 5 at full size: max_length=8 constrained merging with resume on a 256 MB corpus.
  A: uninterrupted mergeUntil({max_length:8, max_iterations:2*HALF})
  B: HALF merges -> toJSON -> new engine fromJSON -> restoreDocuments (every document) -> HALF more   (core.ts:213-216 route)
  C: addDocuments + restoreMerges(merge log of the first HALF) -> HALF more                              (core.ts:477-494 route)
All three must produce the same merge sequence and snapshot."""
import os, sys, time, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sys.argv = ["bench.py"]
import bench
from bpe_tokenizer_b200 import BPETokenizer, _abi

size = int(os.environ.get("SIZE", 256_000_000)); half = int(os.environ.get("HALF", 4096))
lib = _abi.load_library()
text, off = bench.synth(lib, size, 43)
lut, alphabet = bench.alphabet_lut(text)
ids = lut[text]; del text

def fresh():
    t = BPETokenizer()
    t.addToCorpus("".join(chr(c) for c in alphabet)); t._pending = []
    for tk in t.token_table: tk.weight = 0; tk.original_weight = 0
    return t

def sha(t):
    return hashlib.sha1(repr([[a.index, b.index, c.original_weight] for a, b, c in t.merge_tokens]).encode()).hexdigest()[:16]

opts = {"max_length": 8}
t0 = time.time(); A = fresh(); A.addDocuments(ids, off); n = A.mergeUntil(dict(opts, max_iterations=2 * half)); tA = time.time() - t0
print("A uninterrupted: %d merges, %.1f s, sha %s" % (n, tA, sha(A)), flush=True)
golden_path = os.path.join(ROOT, "tests", "golden", "cfg5_merge_log.json")  # the incremental CPU oracle's log of this exact run
if size == 256_000_000 and half == 4096 and os.path.exists(golden_path):
    import json
    golden = json.load(open(golden_path))
    print("A equals the CPU oracle's merge log (%s): %s" % (golden["sha16_of_repr"], golden["sha16_of_repr"] == sha(A) and golden["merges"] == n), flush=True)
t0 = time.time(); B1 = fresh(); B1.addDocuments(ids, off); B1.mergeUntil(dict(opts, max_iterations=half)); snap = B1.toJSON()
log = [[a.code, b.code, c.original_weight] for a, b, c in B1.merge_tokens]; B1.close()
B = BPETokenizer(); B.fromJSON(snap); B.restoreDocuments(ids, off); B.mergeUntil(dict(opts, max_iterations=half)); tB = time.time() - t0
print("B toJSON -> fromJSON -> restoreToCorpus -> continue: %.1f s, sha %s" % (tB, sha(B)), flush=True)
t0 = time.time(); Cc = fresh(); Cc.addDocuments(ids, off); Cc.restoreMerges(log); Cc.mergeUntil(dict(opts, max_iterations=half)); tC = time.time() - t0
print("C addToCorpus + restoreMerges(log) -> continue: %.1f s, sha %s" % (tC, sha(Cc)), flush=True)
ja, jb, jc = A.toJSON(), B.toJSON(), Cc.toJSON()
same_tables = [r[0] for r in ja["token_table"]] == [r[0] for r in jb["token_table"]] == [r[0] for r in jc["token_table"]]
print("merge sequences equal:", sha(A) == sha(B) == sha(Cc), "| token chars equal:", same_tables, "| merge_codes equal:", ja["merge_codes"] == jb["merge_codes"] == jc["merge_codes"],
      "| full snapshot A == C:", ja == jc)
ia, _ = A.corpusIds(); ib, _ = B.corpusIds(); ic, _ = Cc.corpusIds()
print("corpora equal:", bool(np.array_equal(ia, ib) and np.array_equal(ia, ic)), "tokens", ia.size)
