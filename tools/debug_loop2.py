import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bpe_tokenizer_b200 import BPETokenizer
cases = json.load(open(os.path.join(ROOT, "tests/golden/literal_cases.json")))
case = [c for c in cases if c["name"] == "zipf_6k"][0]
for rep in range(3):
    t = BPETokenizer()
    for d in case["docs"]:
        t.addToCorpus(d)
    n = t.mergeUntil({"max_iterations": 12})
    got = [[a.index, b.index, c.weight] for a, b, c in t.merge_tokens]
    print("got ", got)
print("want", case["merges"][:12])
t = BPETokenizer()
for d in case["docs"]:
    t.addToCorpus(d)
for k in range(6):
    n = t.mergeUntil({"max_iterations": 1})
    print(k, [[a.index, b.index, c.weight] for a, b, c in t.merge_tokens][-1], t.stats()["kernel_launches"])
