"""How predictable is the next winner?  (planning aid for the merge loop, not a test and not on the product path)

k_merge_loop is bound by a chain of dependent memory round trips per merge, so the only way to go much below its ~26 us
per merge is to start the site pass of merge t+1 before merge t has finished -- which needs the winner of t+1 early.
This script runs plain BPE training (numpy, counts by sort; the (x,x) run-parity rule is honoured in the replacement,
approximated in the counts) on a small seeded Zipf corpus and reports, per merge t:

  * whether winner(t+1) is the runner-up of decision t (what prefetch_runner_up speculates on),
  * whether it is among the top 3 / top 4 of decision t,
  * whether it shares a token with merge t (then its count may have moved, and its sites may touch those of merge t).

usage: python tests/proto/proto_speculation_stats.py [bytes=2000000] [merges=2000]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bpe_tokenizer_b200 import synth  # noqa: E402


def main(n_bytes=2_000_000, merges=2000):
    text, off = synth.synth_corpus(n_bytes, seed=43)
    ids, alphabet = synth.first_appearance_ids(text)
    tok = ids.astype(np.int64)
    start = np.zeros(tok.size, dtype=bool)
    start[off[:-1]] = True  # first token of a document: no pair reaches across it
    n_tok = len(alphabet)
    prev_top, prev_merge = None, None
    hit1 = hit3 = hit4 = shares = total = 0
    for t in range(merges):
        valid = ~start[1:]
        key = (tok[:-1] << 20) | tok[1:]
        k, c = np.unique(key[valid], return_counts=True)
        order = np.lexsort((k, -c))[:4]  # by count, ties by key (the reference's tie-break differs; irrelevant for the statistics)
        top = [int(k[i]) for i in order]
        if c[order[0]] < 2:
            break
        w = top[0]
        a, b = w >> 20, w & 0xFFFFF
        if prev_top is not None:
            total += 1
            hit1 += w == prev_top[1] if len(prev_top) > 1 else 0
            hit3 += w in prev_top[1:3]
            hit4 += w in prev_top[1:4]
            pa, pb, pc = prev_merge
            shares += len({a, b} & {pa, pb, pc}) > 0
        # left-to-right, non-overlapping replacement (runs of a == b: every other pair)
        m = np.flatnonzero((tok[:-1] == a) & (tok[1:] == b) & valid)
        if a == b and m.size:
            keep, last = [], -2
            for i in m.tolist():
                if i != last + 1:
                    keep.append(i)
                    last = i
                else:
                    last = -2
            m = np.array(keep, dtype=np.int64)
        tok[m] = n_tok
        alive = np.ones(tok.size, dtype=bool)
        alive[m + 1] = False
        tok, start = tok[alive], start[alive]
        prev_top, prev_merge = top, (a, b, n_tok)
        n_tok += 1
    print("merges analysed: %d (corpus %d bytes)" % (total, n_bytes))
    print("winner(t+1) == runner-up(t):        %.1f %%" % (100.0 * hit1 / total))
    print("winner(t+1) in places 2-3 of t:     %.1f %%" % (100.0 * hit3 / total))
    print("winner(t+1) in places 2-4 of t:     %.1f %%" % (100.0 * hit4 / total))
    print("winner(t+1) shares a token with t:  %.1f %%" % (100.0 * shares / total))


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:3]))
