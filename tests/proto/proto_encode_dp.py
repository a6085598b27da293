"""PROTOTYPE (test infrastructure): exact encodeToCode (core.ts:392-409) by a forward dynamic programme over positions.

The reference applies every merge in training order with replaceAll over the whole string.  For a vocabulary built by that
very process the result can be computed left to right: last[i] = the token the encoding of the prefix text[:i] ends with is
the LONGEST token t ending at i for which (last[i - len(t)], t) is a pair the sequential process leaves standing
("compatible"); the encoding of the whole text is the chain from the end.  Compatibility of two tokens is decided on their
merge trees alone (no text): walk down the right spine of t1 and the left spine of t2 and make sure no rule across the
boundary would have fired before the rules that built the two tokens.

This file checks the formulation against the literal oracle (oracle/ref_literal.py) on random small-alphabet corpora --
runs, chains, ties -- and on Zipf text; the CUDA kernel (csrc/encode_dp.cuh) is the same algorithm.
    python tests/proto/proto_encode_dp.py [cases]
"""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


class DpEncoder:
    def __init__(self, n_chars, merges):
        """merges: [(a, b, c)] in training order; c must be n_chars + rank (token indices grow with the rank)."""
        self.n_chars = n_chars
        self.ok = True
        self.split = {}
        self.pair = {}
        self.text = {i: (i,) for i in range(n_chars)}
        for r, (a, b, c) in enumerate(merges):
            if c != n_chars + r or a >= c or b >= c:
                self.ok = False  # (addToCorpus after merges: characters younger than merged tokens) -- not this path
                return
            self.split[c] = (a, b)
            if (a, b) not in self.pair:
                self.pair[(a, b)] = c  # (a pair can only be learnt once: after its merge it no longer occurs)
            self.text[c] = self.text[a] + self.text[b]
        self.tok_of = {}
        for t, s in self.text.items():
            if s in self.tok_of:
                self.ok = False  # two tokens with the same characters: the string does not identify the token
                return
            self.tok_of[s] = t
        self.maxlen = max(len(s) for s in self.text.values())

    def compatible(self, t1, t2):
        """Would the sequential process, run over text(t1) + text(t2), end with exactly [t1, t2]?"""
        n = self.n_chars
        limit = 1 << 30
        while True:
            c = self.pair.get((t1, t2))
            if c is not None and c < limit:
                return False
            if t1 > t2:
                limit = t1
                if t1 < n:
                    return True  # (both are characters and no rule joins them)
                t1 = self.split[t1][1]
                # the right part may itself be what the boundary rule needs: loop
            else:
                limit = t2 + 1
                if t2 < n:
                    return True
                t2 = self.split[t2][0]

    def encode(self, doc):
        n = len(doc)
        if n == 0:
            return []
        last = [None] * (n + 1)
        for i in range(1, n + 1):
            for L in range(min(i, self.maxlen), 0, -1):
                t = self.tok_of.get(tuple(doc[i - L:i]))
                if t is None:
                    continue
                if i - L == 0 or self.compatible(last[i - L], t):
                    last[i] = t
                    break
            assert last[i] is not None, (i, doc)
        out = []
        i = n
        while i > 0:
            t = last[i]
            out.append(t)
            i -= len(self.text[t])
        return out[::-1]


def literal_encode(n_chars, merges, doc):
    """core.ts:404-406 on token indices."""
    cur = list(doc)
    for a, b, c in merges:
        out, i = [], 0
        while i < len(cur):
            if i + 1 < len(cur) and cur[i] == a and cur[i + 1] == b:
                out.append(c)
                i += 2
            else:
                out.append(cur[i])
                i += 1
        cur = out
    return cur


def train(docs, n_chars, max_merges, rng):
    """plain BPE with the reference's counting rule, enough to produce realistic merge tables (ties broken at random: any
    table a run of the reference could have produced is fair game for the encoder)"""
    corpus = [list(d) for d in docs]
    merges = []
    for r in range(max_merges):
        cnt = {}
        for d in corpus:
            i, run = 0, 0
            while i + 1 < len(d):
                a, b = d[i], d[i + 1]
                if a == b:
                    if run % 2 == 0:
                        cnt[(a, b)] = cnt.get((a, b), 0) + 1
                    run += 1
                else:
                    cnt[(a, b)] = cnt.get((a, b), 0) + 1
                    run = 0
                i += 1
        if not cnt:
            break
        best = max(cnt.values())
        if best < 2:
            break
        a, b = rng.choice([p for p, v in cnt.items() if v == best])
        c = n_chars + r
        merges.append((a, b, c))
        corpus = [literal_encode(n_chars, [(a, b, c)], d) for d in corpus]
    return merges


def main(cases=300):
    rng = random.Random(7)
    bad = 0
    for case in range(cases):
        k = rng.choice([1, 2, 2, 3, 5])
        docs = [[rng.randrange(k) for _ in range(rng.randint(0, rng.choice([8, 30, 120])))] for _ in range(rng.randint(1, 6))]
        merges = train(docs, k, rng.choice([3, 10, 40]), rng)
        enc = DpEncoder(k, merges)
        assert enc.ok
        tests = docs + [[rng.randrange(k) for _ in range(rng.randint(0, 60))] for _ in range(6)]
        for d in tests:
            want = literal_encode(k, merges, d)
            got = enc.encode(d)
            if got != want:
                bad += 1
                if bad <= 5:
                    print("MISMATCH", case, "merges", merges, "doc", d, "want", want, "got", got)
    print("cases", cases, "mismatches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 300) else 0)
