"""CPU: the pair table the encode kernel works from -- rank of every rule, and for every pair on a spine of a rule the lowest rank
that can consume its right / left half (csrc/encode_lanes.cuh, built on the host by bpe_b200.cu build_lane_entries) -- against the
independent Python construction of the prototype the algorithm was proven with (tests/proto/proto_encode_lanes.py Tables, fuzzed
against the sequential replaceAll of core.ts:404-406).  No device: bpe_debug_lane_table runs the product's host code only."""
import ctypes as C
import os
import random
import sys

import numpy as np
import pytest

from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200._abi import p32

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "proto"))
from proto_encode_lanes import INF, NONE, Tables  # noqa: E402


def product_table(merges, n_tokens):
    lib = _abi.load_library()
    abc = np.ascontiguousarray(np.array(merges, dtype=np.int32).reshape(-1))
    cap = 8 * len(merges) + 16
    cols = [np.zeros(cap, dtype=np.int32) for _ in range(6)]
    n = C.c_int64()
    rc = lib.bpe_debug_lane_table(p32(abc) if len(merges) else None, len(merges), n_tokens, *[p32(x) for x in cols], cap, C.byref(n))
    assert rc == _abi.BPE_OK, rc
    a, b, rk, c, rs, ls = [x[: n.value].tolist() for x in cols]
    return {(a[i], b[i]): (rk[i], c[i], rs[i], ls[i]) for i in range(n.value)}


def check(merges, n_tokens):
    got = product_table(merges, n_tokens)
    T = Tables(merges)
    keys = set(T.rank) | set(T.RS) | set(T.LS)
    assert set(got) == keys
    for k in keys:
        rk, rs, ls = T.lookup(*k)
        want = (-1 if rk == NONE else rk, -1 if rk == NONE else T.rule[rk], -1 if rs == INF else rs, -1 if ls == INF else ls)
        assert got[k] == want, (k, got[k], want)


def random_merges(rng, n_chars, n_merges, dup=0.1):
    """merge lists shaped like training output: merge r creates token n_chars + r from two EXISTING tokens; a pair may be
    listed twice (a later duplicate rule can never fire: the first one removed every occurrence)"""
    merges, pairs = [], []
    for r in range(n_merges):
        n = n_chars + r
        if pairs and rng.random() < dup:
            a, b = rng.choice(pairs)
        else:
            a, b = rng.randrange(n), rng.randrange(n)
        pairs.append((a, b))
        merges.append((a, b, n))
    return merges


@pytest.mark.parametrize("seed", range(40))
def test_lane_table_matches_prototype_on_random_merge_lists(seed):
    rng = random.Random(500 + seed)
    n_chars = rng.choice([1, 2, 3, 5, 29])
    merges = random_merges(rng, n_chars, rng.choice([0, 1, 2, 10, 60, 400]), dup=rng.choice([0.0, 0.1, 0.4]))
    check(merges, n_chars + len(merges))


def test_lane_table_matches_prototype_on_the_cfg3_table():
    abc = np.load(os.path.join(ROOT, "tools", "data", "merges_cfg3_abc.npy"))
    merges = [tuple(int(x) for x in row) for row in abc]
    check(merges, 29 + len(merges))


def test_lane_table_rejects_malformed_merge_lists():
    lib = _abi.load_library()
    n = C.c_int64()
    dummy = np.zeros(64, dtype=np.int32)
    cols = [p32(dummy)] * 6
    bad_ref = np.array([0, 9, 2], dtype=np.int32)  # token 9 does not exist
    assert lib.bpe_debug_lane_table(p32(bad_ref), 1, 3, *cols, 64, C.byref(n)) == _abi.BPE_E_INVALID
    two_defs = np.array([0, 1, 2, 1, 0, 2], dtype=np.int32)  # token 2 produced by two different merges
    assert lib.bpe_debug_lane_table(p32(two_defs), 2, 3, *cols, 64, C.byref(n)) == _abi.BPE_E_INVALID
    small = np.array([0, 1, 2], dtype=np.int32)
    assert lib.bpe_debug_lane_table(p32(small), 1, 3, *cols, 0, C.byref(n)) == _abi.BPE_E_CAPACITY and n.value >= 1
