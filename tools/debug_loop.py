"""GPU debug helper: compare persistent / host-driven merge loops against the golden zipf_6k case and time cfg2."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from bpe_tokenizer_b200 import BPETokenizer
from bpe_tokenizer_b200.synth import synth_corpus, first_appearance_ids

cases = json.load(open(os.path.join(ROOT, "tests/golden/literal_cases.json")))
case = [c for c in cases if c["name"] == "zipf_6k"][0]
for mode in (4, 0, 4 | 2, 2):
    t = BPETokenizer()
    t._lib.bpe_set_profiling(t._h, mode)
    for d in case["docs"]:
        t.addToCorpus(d)
    t.mergeUntil(case["options"])
    got = [[a.index, b.index, c.weight] for a, b, c in t.merge_tokens]
    bad = [i for i, (x, y) in enumerate(zip(got, case["merges"])) if x != y]
    print("mode", mode, "merges", len(got), "first mismatch", bad[:1], got[bad[0]] if bad else None, case["merges"][bad[0]] if bad else None, t.stats()["kernel_launches"])

text, off = synth_corpus(10_000_000)
ids, alphabet = first_appearance_ids(text)
for mode in (4, 0, 0):
    t = BPETokenizer()
    t._lib.bpe_set_profiling(t._h, mode)
    t.addToCorpus("".join(chr(c) for c in alphabet)); t._pending = []
    t.addDocuments(ids, off)
    t.findNextMerge()
    t0 = time.time(); n = t.mergeUntil({"max_iterations": 4000}); dt = time.time() - t0
    s = t.stats()
    import ctypes as C
    print("mode", mode, "cfg2 merges", n, "wall %.1f ms" % (dt * 1e3), "dev %.1f ms" % s["ms_last_merge_until"], "launches", s["kernel_launches"], "hot_rebuilds", s["hot_rebuilds"], "ties", s["tie_breaks"], "k1 ms %.2f" % s["ms_index_build"], "pairs", s["distinct_pairs"], "phases ms", [round(x, 1) for x in s["ms_loop_phase"]])
