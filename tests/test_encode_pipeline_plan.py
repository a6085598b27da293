"""CPU: how the host-buffer encode calls (bpe_encode_batch / bpe_encode_text_batch) cut a batch into chunks of whole documents.
The staging buffers of the three-stream pipeline are sized from a BOUND on the number of chunks before the first copy starts
(bpe_b200.cu pipe_plan); these properties guard that bound -- a chunk list that outgrew it would write past the buffers.
bpe_debug_plan_chunks runs the planner of the product (same functions) without touching a device."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200._abi import p64


def plan(lengths, chunk_units, base=0):
    lib = _abi.load_library()
    off = np.zeros(len(lengths) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lengths)
    off += base
    first = np.full(len(lengths) + 1, -1, dtype=np.int64)
    n, bounds = C.c_int64(), np.zeros(5, dtype=np.int64)
    rc = lib.bpe_debug_plan_chunks(p64(off), len(lengths), chunk_units, p64(first), first.size, C.byref(n), p64(bounds))
    assert rc == _abi.BPE_OK
    return off, first[: n.value], bounds


def check(lengths, chunk_units, base=0):
    off, first, (max_chunks, unit_cap, units_used, off_cap, off_used) = plan(lengths, chunk_units, base)
    n_docs, total = len(lengths), int(off[-1] - off[0])
    # the chunks partition the documents in order, none is empty
    assert first[0] == 0 and np.all(np.diff(first) > 0) and first[-1] < n_docs
    # the bound the buffers were sized with holds, and so do the buffers
    assert len(first) <= max_chunks
    assert total <= units_used <= unit_cap
    assert off_used == n_docs + len(first) and off_used <= off_cap
    # no chunk but a single over-long document exceeds the full-size target
    ends = np.append(first[1:], n_docs)
    units = off[ends] - off[first]
    many = (ends - first) > 1
    assert np.all(units[many] <= max(chunk_units, 1))
    return first, units


@settings(max_examples=300, deadline=None)
@given(st.lists(st.one_of(st.just(0), st.integers(0, 40), st.integers(0, 3000), st.integers(0, 200000)), min_size=1, max_size=400),
       st.one_of(st.integers(1, 64), st.integers(1, 5000), st.integers(1, 1 << 22)), st.integers(0, 1 << 40))
def test_chunk_plan_properties(lengths, chunk_units, base):
    check(lengths, chunk_units, base)


@pytest.mark.parametrize("chunk_units", [1, 7, 97, 1000, 1 << 20, 1 << 62])
def test_chunk_plan_edge_cases(chunk_units):
    check([0], chunk_units)
    check([0] * 50, chunk_units)
    check([5], chunk_units)
    check([1] * 1000, chunk_units)
    check([0, 0, 10**6, 0, 0], chunk_units)
    check([3000] * 300 + [0, 0], chunk_units, base=12345)


def test_chunk_sizes_ramp_up_and_down():
    """1/8, 1/4, 1/2 of the full size first (the first copy in is exposed), halving again towards the end (so are the last encode and
    the last copy out); in between, full-size chunks."""
    first, units = check([250] * 40000, 1_000_000)  # 10 M units, 1 M per full-size chunk
    assert units[0] <= 125_000 < units[1] <= 250_000 < units[2] <= 500_000 < units[3] <= 1_000_000
    assert units.max() <= 1_000_000 and (units > 900_000).sum() >= 5
    assert units[-1] <= 250_000 and units[-1] <= units[-2] <= units[-3]


def test_chunk_plan_rejects_decreasing_offsets():
    lib = _abi.load_library()
    off = np.array([0, 10, 5, 20], dtype=np.int64)
    n, bounds = C.c_int64(), np.zeros(5, dtype=np.int64)
    assert lib.bpe_debug_plan_chunks(p64(off), 3, 100, None, 0, C.byref(n), p64(bounds)) == _abi.BPE_E_INVALID
