"""CPU prototype of the lane-sequential encode kernel (k_encode_lanes): a batch of whole documents is laid out over
NL "lanes", every lane owns a contiguous segment and walks it sequentially; lanes only exchange clamp composites
(prefix/suffix scans) and a few boundary flags per round.  Mirrors the CUDA control flow one to one and is fuzzed
against the literal oracle (sequential replaceAll per merge, core.ts:404-406).

Every pair record j = (tok_j, rk_j, rs_j, ls_j) comes from ONE table probe of (tok_j, tok_j+1):
  rk = rank of the rule (tok_j, tok_j+1) or NONE
  rs = lowest rank of any rule (y, tok_j+1) with tok_j on the RIGHT spine of y (y = tok_j included): nothing that
       can ever be built ending in tok_j takes tok_j+1 from the left before time rs
  ls = lowest rank of any rule (tok_j, z) with tok_j+1 on the LEFT spine of z
Stability values (all computed on the state at the start of the round):
  T[j]  = INF if j ends a document else min(rk[j], SL[j]),  SL[j+1] = max(rs[j], T[j])            (left to right)
  SR[j] = INF if j ends a document else max(ls[j], min(rk[j], SR[j+1]))                            (right to left)
A pair j of rank r is merged this round iff r <= SL[j] and r <= SR[j+1]; pairs (x,x) additionally follow the
replaceAll parity inside their run and need r <= T[s-1] at the run start s (the run must not grow at its left end).
"""
import math, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

INF = 0xFFFF
NONE = 0xFFFF
BOUNDARY = 0xFFFE
DIRTY = 0xFFFD


class Tables:
    def __init__(self, merges):
        self.rank = {}
        self.rule = []
        for r, (a, b, c) in enumerate(merges):
            self.rule.append(c)
            self.rank.setdefault((a, b), r)
        defn = {}
        for (a, b, c) in merges:
            defn.setdefault(c, (a, b))
        self.RS, self.LS = {}, {}
        for (a, b), r in self.rank.items():
            u = a
            while True:
                if self.RS.get((u, b), INF) > r:
                    self.RS[(u, b)] = r
                if u not in defn:
                    break
                u = defn[u][1]
            w = b
            while True:
                if self.LS.get((a, w), INF) > r:
                    self.LS[(a, w)] = r
                if w not in defn:
                    break
                w = defn[w][0]

    def lookup(self, a, b):
        k = (a, b)
        return self.rank.get(k, NONE), self.RS.get(k, INF), self.LS.get(k, INF)


def clamp_apply(f, x):
    return max(f[0], min(f[1], x))


def encode_batch(docs, T, NL=32, stats=None):
    """docs: list of non-empty id lists.  Returns list of encoded id lists."""
    recs = []  # record = [tok, rk, rs, ls]
    for d in docs:
        for i, t in enumerate(d):
            recs.append([t, BOUNDARY if i == len(d) - 1 else DIRTY, INF, INF])
    n = len(recs)
    if n == 0:
        return []
    L = (n + NL - 1) // NL
    lanes = [recs[l * L:(l + 1) * L] for l in range(NL)]
    rounds = visits = probes = 0

    def nxt_nonempty(l):
        for m in range(l + 1, NL):
            if lanes[m]:
                return m
        return -1

    def prv_nonempty(l):
        for m in range(l - 1, -1, -1):
            if lanes[m]:
                return m
        return -1

    def f_of(rec):  # left-to-right step of record j: SL[j] -> SL[j+1]
        return (INF, INF) if rec[1] == BOUNDARY else (rec[2], max(rec[1], rec[2]))

    def g_of(rec):  # right-to-left step: SR[j+1] -> SR[j]
        return (INF, INF) if rec[1] == BOUNDARY else (rec[3], max(rec[1], rec[3]))

    while True:
        # ---- probe loop + composites (F over all records but the last, G over all) ----
        Fb = [None] * NL
        G = [None] * NL
        anyvalid = False
        for l in range(NL):
            m = nxt_nonempty(l)
            nf = lanes[m][0][0] if m >= 0 else None
            f = (0, INF)
            g = (0, INF)
            cnt = len(lanes[l])
            for j, rec in enumerate(lanes[l]):
                visits += 1
                if rec[1] == DIRTY:
                    b = lanes[l][j + 1][0] if j + 1 < cnt else nf
                    assert b is not None
                    rec[1], rec[2], rec[3] = T.lookup(rec[0], b)
                    probes += 1
                if rec[1] < DIRTY:
                    anyvalid = True
                if j + 1 < cnt:
                    fj = f_of(rec)
                    f = (clamp_apply(fj, f[0]), clamp_apply(fj, f[1]))  # f_j o F
                gj = g_of(rec)
                g = (clamp_apply(g, gj[0]), clamp_apply(g, gj[1]))      # G o g_j
            Fb[l], G[l] = f, g
        if not anyvalid:
            break
        rounds += 1
        # neighbour records (pre-round state)
        prev_last = [None] * NL
        next_first = [None] * NL
        next_lane = [-1] * NL
        for l in range(NL):
            p = prv_nonempty(l)
            if p >= 0:
                prev_last[l] = list(lanes[p][-1])
            m = nxt_nonempty(l)
            next_lane[l] = m
            if m >= 0:
                next_first[l] = list(lanes[m][0])
        # shifted composite: (my records but the last) o (last record of the previous non-empty lane)
        x_in = [INF] * NL   # SL of the previous non-empty lane's last token
        acc = INF
        for l in range(NL):
            x_in[l] = acc
            if lanes[l]:
                if prev_last[l] is not None:
                    acc = clamp_apply(f_of(prev_last[l]), acc)
                acc = clamp_apply(Fb[l], acc)
        y_in = [INF] * NL   # SR of the next non-empty lane's first token
        acc = INF
        for l in range(NL - 1, -1, -1):
            y_in[l] = acc
            acc = clamp_apply(G[l], acc)
        # ---- pass 2: SR values ----
        S = [None] * NL
        for l in range(NL):
            y = y_in[l]
            s = [0] * len(lanes[l])
            for j in range(len(lanes[l]) - 1, -1, -1):
                visits += 1
                y = clamp_apply(g_of(lanes[l][j]), y)
                s[j] = y
            S[l] = s
        # ---- pass 3: decide + compact ----
        new_lanes = [None] * NL
        took_straddle = [False] * NL
        first_is_new = [False] * NL
        for l in range(NL):
            P = lanes[l]
            cnt = len(P)
            out = []
            if prev_last[l] is None:
                sl, g, prk = INF, INF, BOUNDARY
            else:
                pr = prev_last[l]
                prk = pr[1]
                g = INF if prk == BOUNDARY else min(prk, x_in[l])        # T of the previous token
                sl = INF if prk == BOUNDARY else max(pr[2], g)           # SL of my first token
            par = 0
            run_ok = False
            consumed = False
            for j in range(cnt):
                visits += 1
                t, r, rs, ls = P[j]
                last = j == cnt - 1
                nx = next_first[l] if last else P[j + 1]
                srn = y_in[l] if last else S[l][j + 1]
                take = False
                if r < DIRTY:
                    if r == prk:
                        par ^= 1
                        if j == 0:
                            run_ok = False  # run carried in from the previous lane: parity unknown here
                        take = par == 0 and run_ok and r <= srn
                    else:
                        par = 0
                        isxx = t == nx[0]
                        run_ok = (r <= g) if isxx else True
                        take = r <= (g if isxx else sl) and r <= srn
                    if consumed:
                        take = False
                if not consumed:
                    if take:
                        if out:
                            if out[-1][1] != BOUNDARY:
                                out[-1][1] = DIRTY
                        else:
                            first_is_new[l] = True
                        out.append([T.rule[r], BOUNDARY if nx[1] == BOUNDARY else DIRTY, INF, INF])
                        if last:
                            took_straddle[l] = True
                    else:
                        out.append([t, r, rs, ls])
                consumed = take
                g = INF if r == BOUNDARY else min(r, sl)
                sl = INF if r == BOUNDARY else max(rs, g)
                prk = r
            new_lanes[l] = out
        # ---- post: consumed first tokens, dirty marks across lanes ----
        for l in range(NL):
            if took_straddle[l]:
                m = next_lane[l]
                assert m >= 0
                assert not first_is_new[m], "conflict: straddle and (first,second) both taken"
                new_lanes[m].pop(0)
        lanes = new_lanes
        for l in range(NL):
            if not lanes[l]:
                continue
            m = nxt_nonempty(l)
            if m >= 0 and first_is_new[m] and lanes[l][-1][1] != BOUNDARY:
                lanes[l][-1][1] = DIRTY
    if stats is not None:
        stats.append((n, rounds, visits, probes))
    out, cur = [], []
    for l in range(NL):
        for rec in lanes[l]:
            cur.append(rec[0])
            if rec[1] == BOUNDARY:
                out.append(cur)
                cur = []
    assert not cur
    return out


def fuzz(n_cases, seed):
    from oracle import LiteralTokenizer
    rng = random.Random(seed)
    bad = 0
    for case in range(n_cases):
        alphabet = "abcd"[: rng.randint(1, 4)]
        docs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 40))) for _ in range(rng.randint(1, 4))]
        t = LiteralTokenizer()
        for d in docs:
            t.addToCorpus(d)
        t.mergeUntil({"max_length": rng.choice([0, 0, 4, 8]), "min_weight": rng.choice([0, 2, 3])})
        merges = [(a.index, b.index, c.index) for a, b, c in t.merge_tokens]
        T = Tables(merges)
        known = [ch for ch in alphabet if ch in t.char_to_token]
        if not known:
            continue
        for _ in range(4):
            texts = ["".join(rng.choice(known) for _ in range(rng.randint(1, 50))) for _ in range(rng.randint(1, 5))]
            want = [[ord(ch) - 1 for ch in t.encodeToCode(x)] for x in texts]
            ids = [[t.char_to_token[ch].index for ch in x] for x in texts]
            for NL in (1, 2, 3, 5, 32):
                got = encode_batch(ids, T, NL)
                if got != want:
                    bad += 1
                    if bad < 5:
                        print("MISMATCH NL", NL, texts, merges, got, want)
    return bad


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "stats":
        import numpy as np
        from bpe_tokenizer_b200.synth import synth_corpus
        log = np.load(os.path.join(ROOT, "tools", "data", "merges_cfg3_abc.npy"))
        alphabet = np.load(os.path.join(ROOT, "tools", "data", "alphabet_cfg3.npy"))
        lut = np.full(256, -1, dtype=np.int64)
        lut[alphabet] = np.arange(len(alphabet))
        nm = int(sys.argv[2]) if len(sys.argv) > 2 else len(log)
        merges = [tuple(int(x) for x in m) for m in log[:nm]]
        T = Tables(merges)
        text, off = synth_corpus(120_000, seed=44)
        docs = [lut[text[off[d]:off[d + 1]]].tolist() for d in range(len(off) - 1)]
        for cap, NL in ((1024, 32), (1536, 32), (600, 32)):
            stats = []
            i = 0
            tot_in = tot_out = 0
            while i < len(docs):
                batch, s = [], 0
                while i < len(docs) and s + len(docs[i]) <= cap:
                    batch.append(docs[i]); s += len(docs[i]); i += 1
                out = encode_batch(batch, T, NL, stats)
                tot_in += s; tot_out += sum(len(o) for o in out)
            n = sum(s[0] for s in stats)
            print("cap", cap, "batches", len(stats), "chars", n, "tokens", tot_out,
                  "rounds mean %.1f max %d" % (sum(s[1] for s in stats) / len(stats), max(s[1] for s in stats)),
                  "visits/char %.2f" % (sum(s[2] for s in stats) / n), "probes/char %.2f" % (sum(s[3] for s in stats) / n))
    else:
        total = 0
        for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
            b = fuzz(300, seed)
            total += b
            print("seed", seed, "bad", b)
        print("TOTAL BAD", total)
