"""CPU prototype of the cascading lane-sequential encode walk (see proto_encode_lanes.py for the base rule).
Per round: pass 2 (right to left) computes SR on the pre-round state, pass 3 walks left to right keeping ONE current
left token `cur` that greedily absorbs the following pre-round tokens while the merge is provably the one the
sequential process performs (new tokens included: right-cascade), exactly like replaceAll consumes a run."""
import os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.proto.proto_encode_lanes import Tables, clamp_apply, INF, NONE, BOUNDARY, DIRTY

NOBLOCK = 0x1FFFF


def encode_batch(docs, T, NL=32, stats=None, cascade=True):
    toks, rks = [], []
    for d in docs:
        for i, t in enumerate(d):
            toks.append(t)
            rks.append(BOUNDARY if i == len(d) - 1 else DIRTY)
    n = len(toks)
    if n == 0:
        return []
    L = (n + NL - 1) // NL
    # record = [tok, rk, mL, mR]
    lanes = [[[t, r, T.minL.get(t, INF), T.minR.get(t, INF)] for t, r in zip(toks[l * L:(l + 1) * L], rks[l * L:(l + 1) * L])] for l in range(NL)]
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

    # initial ranks
    for l in range(NL):
        m = nxt_nonempty(l)
        for j, rec in enumerate(lanes[l]):
            if rec[1] == DIRTY:
                b = lanes[l][j + 1][0] if j + 1 < len(lanes[l]) else lanes[m][0][0]
                rec[1] = T.lookup(rec[0], b); probes += 1
    while True:
        anyvalid = False
        F, G = [None] * NL, [None] * NL
        for l in range(NL):
            f, g = (0, INF), (0, INF)
            for rec in lanes[l]:
                visits += 1
                t, r, ml, mr = rec
                assert r != DIRTY
                if r < DIRTY:
                    anyvalid = True
                if r == BOUNDARY:
                    fj = gj = (INF, INF)
                else:
                    fj = (min(r, mr), r)
                    gj = (ml, max(r, ml))
                f = (clamp_apply(fj, f[0]), clamp_apply(fj, f[1]))
                g = (clamp_apply(g, gj[0]), clamp_apply(g, gj[1]))
            F[l], G[l] = f, g
        if not anyvalid:
            break
        rounds += 1
        x_in, y_in = [INF] * NL, [INF] * NL
        acc = INF
        for l in range(NL):
            x_in[l] = acc; acc = clamp_apply(F[l], acc)
        acc = INF
        for l in range(NL - 1, -1, -1):
            y_in[l] = acc; acc = clamp_apply(G[l], acc)
        prev_rk = [BOUNDARY] * NL
        next_first = [None] * NL
        next_lane = [-1] * NL
        for l in range(NL):
            p = prv_nonempty(l)
            if p >= 0:
                prev_rk[l] = lanes[p][-1][1]
            m = nxt_nonempty(l)
            next_lane[l] = m
            if m >= 0:
                next_first[l] = list(lanes[m][0])
        S = []
        for l in range(NL):
            y = y_in[l]
            s = [0] * len(lanes[l])
            for j in range(len(lanes[l]) - 1, -1, -1):
                visits += 1
                t, r, ml, mr = lanes[l][j]
                y = INF if r == BOUNDARY else max(ml, min(r, y))
                s[j] = y
            S.append(s)
        new_lanes = [None] * NL
        took_straddle = [False] * NL
        first_changed = [False] * NL   # first emitted record is a new token
        last_new = [False] * NL
        for l in range(NL):
            P = lanes[l]
            cnt = len(P)
            E = []
            if cnt == 0:
                new_lanes[l] = E
                continue
            g = x_in[l]                 # T' value left of cur
            blocked = prev_rk[l] if prev_rk[l] < DIRTY else NOBLOCK
            cur = list(P[0]); cur_new = False
            cur_sl = max(cur[3], g)
            rank_left_known = True      # rank(prevE, cur) is what prevE's record holds
            prevE = None                # pending record (emitted when cur is final)
            prevE_sl = INF
            prevE_new = False
            visits += 1
            for idx in range(1, cnt + 1):
                visits += 1
                if idx < cnt:
                    nx = P[idx]; nx_sr = S[l][idx]
                else:
                    nx = next_first[l]; nx_sr = y_in[l]
                # ---- rank of (cur, nx) ----
                if cur[1] == BOUNDARY or nx is None:
                    rc = BOUNDARY
                elif cur_new:
                    rc = T.lookup(cur[0], nx[0]); probes += 1
                else:
                    rc = cur[1]
                take = False
                if rc < DIRTY:
                    if rc == blocked:
                        take = False
                    else:
                        isxx = cur[0] == nx[0]
                        if cur_new and not rank_left_known and (isxx or rc > cur[3]):
                            # lazily refine the left stability of a new token
                            if prevE is None:
                                pass  # left neighbour lives in the previous lane: keep the conservative bound
                            else:
                                rl = BOUNDARY if prevE[1] == BOUNDARY else T.lookup(prevE[0], cur[0]); probes += 1
                                prevE[1] = rl
                                rank_left_known = True
                                g = INF if rl == BOUNDARY else min(rl, prevE_sl)
                                cur_sl = max(cur[3], g)
                        if cur_new and not rank_left_known:
                            lim = 0 if isxx else cur[3]
                        else:
                            lim = g if isxx else cur_sl
                        take = rc <= lim and rc <= nx_sr
                        blocked = rc if (isxx and not take) else NOBLOCK
                else:
                    blocked = NOBLOCK
                if take and (cascade or not cur_new):
                    d = T.rule[rc]
                    cur = [d, BOUNDARY if nx[1] == BOUNDARY else DIRTY, T.minL.get(d, INF), T.minR.get(d, INF)]
                    cur_new = True
                    rank_left_known = False
                    cur_sl = cur[3]
                    if idx == cnt:
                        took_straddle[l] = True
                    continue
                if take:
                    blocked = NOBLOCK
                # ---- cur is final: fix prevE's rank, emit prevE, shift ----
                if prevE is not None:
                    if not rank_left_known:
                        prevE[1] = BOUNDARY if prevE[1] == BOUNDARY else T.lookup(prevE[0], cur[0]); probes += 1
                        g = INF if prevE[1] == BOUNDARY else min(prevE[1], prevE_sl)
                        cur_sl = max(cur[3], g)
                    E.append(prevE)
                else:
                    first_changed[l] = cur_new
                if cur[1] != BOUNDARY:
                    cur[1] = rc if idx < cnt or not cur_new else DIRTY
                if idx == cnt:
                    # last record: its right neighbour lives in the next lane
                    last_new[l] = cur_new
                    E.append(cur)
                    break
                prevE, prevE_sl, prevE_new = cur, cur_sl, cur_new
                g = INF if cur[1] == BOUNDARY else min(rc if rc != DIRTY else NONE, cur_sl)
                cur = list(nx); cur_new = False
                cur_sl = max(cur[3], g)
                rank_left_known = True
            else:
                pass
            if took_straddle[l]:
                # cur absorbed the next lane's first token and is the last record
                if prevE is not None:
                    if not rank_left_known:
                        prevE[1] = BOUNDARY if prevE[1] == BOUNDARY else T.lookup(prevE[0], cur[0]); probes += 1
                    E.append(prevE)
                else:
                    first_changed[l] = True
                last_new[l] = True
                E.append(cur)
            new_lanes[l] = E
        for l in range(NL):
            if took_straddle[l]:
                m = next_lane[l]
                assert not first_changed[m], "conflict"
                new_lanes[m].pop(0)
                first_changed[m] = True   # its first surviving record has a new left neighbour
        lanes = new_lanes
        # last records: right neighbour changed?
        for l in range(NL):
            if not lanes[l]:
                continue
            rec = lanes[l][-1]
            if rec[1] == BOUNDARY:
                continue
            m = nxt_nonempty(l)
            assert m >= 0
            if last_new[l] or first_changed[m] or rec[1] == DIRTY or took_straddle[l]:
                rec[1] = T.lookup(rec[0], lanes[m][0][0]); probes += 1
    if stats is not None:
        stats.append((n, rounds, visits, probes))
    out, cur = [], []
    for l in range(NL):
        for rec in lanes[l]:
            cur.append(rec[0])
            if rec[1] == BOUNDARY:
                out.append(cur); cur = []
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
                try:
                    got = encode_batch(ids, T, NL)
                except AssertionError as ex:
                    got = "ASSERT %s" % ex
                if got != want:
                    bad += 1
                    if bad < 4:
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
        merges = [tuple(int(x) for x in m) for m in log]
        T = Tables(merges)
        text, off = synth_corpus(120_000, seed=44)
        docs = [lut[text[off[d]:off[d + 1]]].tolist() for d in range(len(off) - 1)]
        from tests.proto import proto_encode_lanes as base
        for casc in (True, False):
            stats = []
            i = 0
            ok = True
            while i < len(docs):
                batch, s = [], 0
                while i < len(docs) and s + len(docs[i]) <= 1024:
                    batch.append(docs[i]); s += len(docs[i]); i += 1
                out = encode_batch(batch, T, 32, stats, cascade=casc)
                ok &= out == base.encode_batch(batch, T, 32)
            n = sum(s[0] for s in stats)
            print("cascade", casc, "same as base:", ok, "rounds mean %.1f max %d" % (sum(s[1] for s in stats) / len(stats), max(s[1] for s in stats)),
                  "visits/char %.2f" % (sum(s[2] for s in stats) / n), "probes/char %.2f" % (sum(s[3] for s in stats) / n))
    else:
        total = 0
        for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
            b = fuzz(300, seed)
            total += b
            print("seed", seed, "bad", b)
        print("TOTAL BAD", total)
