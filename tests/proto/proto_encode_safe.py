"""CPU prototype of the multi-merge-per-round encode rule (to be ported to k_encode), fuzzed against the
literal oracle.  A pair (x,y) of rank r may be merged NOW iff x cannot be consumed from the left and y cannot be
consumed from the right strictly before time r (then the sequential process merges exactly this pair at time r)."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import LiteralTokenizer

INF = 1 << 30

def encode_safe(ids, merges, K=4, stats=None):
    rank = {}
    for r, (a, b, c) in enumerate(merges):
        rank.setdefault((a, b), (r, c))
    min_as_right, min_as_left = {}, {}
    for (a, b), (r, c) in rank.items():
        min_as_left[a] = min(min_as_left.get(a, INF), r)
        min_as_right[b] = min(min_as_right.get(b, INF), r)
    n = len(ids)
    tok = list(ids); nxt = [i + 1 if i + 1 < n else -1 for i in range(n)]; prv = [i - 1 for i in range(n)]
    alive = [True] * n
    def rk_of(i):
        j = nxt[i]
        if j < 0: return INF, -1
        return rank.get((tok[i], tok[j]), (INF, -1))
    rk = [rk_of(i) for i in range(n)]
    def stable_left(j, r):
        for _ in range(K + 1):
            if prv[j] < 0 or min_as_right.get(tok[j], INF) >= r: return True
            p = prv[j]
            if rk[p][0] < r: return False
            j = p
        return False
    def stable_right(j, r):
        for _ in range(K + 1):
            if nxt[j] < 0 or min_as_left.get(tok[j], INF) >= r: return True
            if rk[j][0] < r: return False
            j = nxt[j]
        return False
    rounds = 0
    while True:
        live = [i for i in range(n) if alive[i]]
        gmin = min((rk[i][0] for i in live), default=INF)
        if gmin >= INF: break
        rounds += 1
        sel = []
        for i in live:
            r, c = rk[i]
            if r >= INF: continue
            j = nxt[i]
            if tok[i] != tok[j]:
                if r == gmin or (stable_left(i, r) and stable_right(j, r)): sel.append(i)
            else:
                s = i; off = 0
                while prv[s] >= 0 and tok[prv[s]] == tok[i]: s = prv[s]; off += 1
                e = j
                while nxt[e] >= 0 and tok[nxt[e]] == tok[i]: e = nxt[e]
                # a run must not be eaten at either end, and must not GROW at its left end (parity!) before time r
                if off % 2 == 0 and (r == gmin or (stable_left(s, r) and (prv[s] < 0 or stable_left(prv[s], r)) and stable_right(e, r))): sel.append(i)
        assert sel
        new = []
        for i in sel:
            j = nxt[i]; jn = nxt[j]
            tok[i] = rk[i][1]; nxt[i] = jn
            if jn >= 0: prv[jn] = i
            alive[j] = False; rk[j] = (INF, -1); new.append(i)
        for i in new:
            rk[i] = rk_of(i)
            if prv[i] >= 0: rk[prv[i]] = rk_of(prv[i])
    if stats is not None: stats.append(rounds)
    return [tok[i] for i in range(n) if alive[i]]

def fuzz(n_cases, seed):
    rng = random.Random(seed); bad = 0; rounds = []
    for case in range(n_cases):
        alphabet = "abcd"[: rng.randint(1, 4)]
        docs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 40))) for _ in range(rng.randint(1, 4))]
        t = LiteralTokenizer()
        for d in docs: t.addToCorpus(d)
        t.mergeUntil({"max_length": rng.choice([0, 0, 4, 8]), "min_weight": rng.choice([0, 2, 3])})
        merges = [(a.index, b.index, c.index) for a, b, c in t.merge_tokens]
        # also scramble: random extra texts over the same alphabet
        known = [ch for ch in alphabet if ch in t.char_to_token]
        for _ in range(6):
            text = "".join(rng.choice(known) for _ in range(rng.randint(0, 60))) if known else ""
            want = [ord(ch) - 1 for ch in t.encodeToCode(text)]
            ids = [t.char_to_token[ch].index for ch in text]
            for K in (0, 1, 4):
                got = encode_safe(ids, merges, K, rounds if K == 4 else None)
                if got != want:
                    bad += 1
                    if bad < 5: print("MISMATCH K", K, repr(text), merges, got, want)
    return bad, rounds

if __name__ == "__main__":
    total = 0
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
        bad, rounds = fuzz(400, seed)
        total += bad
        print("seed", seed, "bad", bad, "mean rounds", sum(rounds) / max(1, len(rounds)))
    print("TOTAL BAD", total)
