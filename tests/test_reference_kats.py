"""The reference's own known-answer tests for this path, restated against the CPU oracles and the product's host layer.

* dune/grid/test/yasp/test-yaspgrid-entityshifttable.cc:84-179 — the literal golden bitsets of Yasp::entityShift /
  Yasp::entityMove (dim 1..4) and the binomial table (via subEnt = C(d,c)*2^c);
* dune/grid/test/yasp/test-yaspgrid-partitioner.cc:15-92 — DefaultPartitioning equals the legacy optimiser for overlap 0
  (sizes <= 4^d, every P), results positive otherwise.
The product is checked too: dgb_partition through the C ABI, and the sub-entity tables through dgb_mapper_subindex
(GPU, tests/test_gpu_parity.py)."""
import itertools
import math

import pytest

from tests.common import oracle_kinds
from oracle import oracle as orc

SHIFT = {
    1: [0b1, 0b0, 0b0],
    2: [0b11, 0b10, 0b10, 0b01, 0b01, 0b00, 0b00, 0b00, 0b00],
    3: [0b111, 0b110, 0b110, 0b101, 0b101, 0b011, 0b011, 0b100, 0b100, 0b100, 0b100, 0b010, 0b010, 0b001, 0b001, 0b010, 0b010,
        0b001, 0b001, 0b000, 0b000, 0b000, 0b000, 0b000, 0b000, 0b000, 0b000],
    4: [0b1111, 0b1110, 0b1110, 0b1101, 0b1101, 0b1011, 0b1011, 0b0111, 0b0111, 0b1100, 0b1100, 0b1100, 0b1100, 0b1010, 0b1010,
        0b1001, 0b1001, 0b1010, 0b1010, 0b1001, 0b1001, 0b0110, 0b0110, 0b0101, 0b0101, 0b0011, 0b0011, 0b0110, 0b0110, 0b0101,
        0b0101, 0b0011, 0b0011, 0b1000, 0b1000, 0b1000, 0b1000, 0b1000, 0b1000, 0b1000, 0b1000, 0b0100, 0b0100, 0b0100, 0b0100,
        0b0010, 0b0010, 0b0001, 0b0001, 0b0010, 0b0010, 0b0001, 0b0001, 0b0100, 0b0100, 0b0100, 0b0100, 0b0010, 0b0010, 0b0001,
        0b0001, 0b0010, 0b0010, 0b0001, 0b0001] + [0] * 16,
}
MOVE = {
    1: [0b0, 0b0, 0b1],
    2: [0b00, 0b00, 0b01, 0b00, 0b10, 0b00, 0b01, 0b10, 0b11],
    3: [0b000, 0b000, 0b001, 0b000, 0b010, 0b000, 0b100, 0b000, 0b001, 0b010, 0b011, 0b000, 0b001, 0b000, 0b010, 0b100, 0b101,
        0b100, 0b110, 0b000, 0b001, 0b010, 0b011, 0b100, 0b101, 0b110, 0b111],
    4: [0b0000, 0b0000, 0b0001, 0b0000, 0b0010, 0b0000, 0b0100, 0b0000, 0b1000, 0b0000, 0b0001, 0b0010, 0b0011, 0b0000, 0b0001,
        0b0000, 0b0010, 0b0100, 0b0101, 0b0100, 0b0110, 0b0000, 0b0001, 0b0000, 0b0010, 0b0000, 0b0100, 0b1000, 0b1001, 0b1000,
        0b1010, 0b1000, 0b1100, 0b0000, 0b0001, 0b0010, 0b0011, 0b0100, 0b0101, 0b0110, 0b0111, 0b0000, 0b0001, 0b0010, 0b0011,
        0b0000, 0b0001, 0b0000, 0b0010, 0b0100, 0b0101, 0b0100, 0b0110, 0b1000, 0b1001, 0b1010, 0b1011, 0b1000, 0b1001, 0b1000,
        0b1010, 0b1100, 0b1101, 0b1100, 0b1110] + list(range(16)),
}


def kinds():
    k = oracle_kinds()
    assert k, "no oracle built"
    return k


@pytest.mark.parametrize("kind", ["ref", "port"])
def test_entity_shift_and_move_tables(kind):
    if kind not in kinds():
        pytest.skip(f"oracle '{kind}' is not built here")
    o = orc.load(kind)
    for dim in (1, 2, 3, 4):
        got_shift, got_move = [], []
        for codim in range(dim + 1):
            n = o.sub_entities(dim, dim, codim)
            assert n == math.comb(dim, codim) << codim                 # binomial table of the same test file (:84-100)
            for j in range(n):
                got_shift.append(o.entity_shift(dim, j, codim))
                got_move.append(o.entity_move(dim, j, codim))
        assert got_shift == SHIFT[dim], dim
        assert got_move == MOVE[dim], dim


def legacy_partition(size, P):
    """old_optimize_dims of test-yaspgrid-partitioner.cc:15-56"""
    d = len(size)
    best, opt = None, 1e100
    trial = [0] * d

    def rec(i, procs):
        nonlocal best, opt
        if i > 0:
            for k in range(1, procs + 1):
                if procs % k == 0:
                    trial[i] = k
                    rec(i - 1, procs // k)
        else:
            trial[0] = procs
            m = -1.0
            for k in range(d):
                mm = size[k] / trial[k]
                if math.fmod(size[k], trial[k]) > 0.0001:
                    mm *= 3
                m = max(m, mm)
            if m < opt:
                opt, best = m, list(trial)

    rec(d - 1, P)
    return tuple(best)


def partitioners():
    import dune_grid_b200 as dg
    import ctypes as C

    def product(size, P, o):
        d = len(size)
        N = (C.c_int32 * 3)(*(list(size) + [1] * (3 - d)))
        fixed = (C.c_int32 * 3)(1, 1, 1)
        out = (C.c_int32 * 3)(0, 0, 0)
        rc = dg.capi.lib.dgb_partition(d, N, P, o, 0, fixed, out)
        if rc != 0:
            raise RuntimeError(dg.capi.lib.dgb_last_error().decode())
        return tuple(out)[:d]

    res = [("product", product)]
    for k in kinds():
        ol = orc.load(k)
        res.append((k, lambda size, P, o, ol=ol: ol.partition(list(size), P, o)))
    return res


@pytest.mark.parametrize("d", [2, 3])
def test_default_partitioning_matches_legacy(d):
    for name, part in partitioners():
        for size in itertools.product(range(1, 5), repeat=d):
            maxP = math.prod(size)
            maxO = max(size) // 2
            for o in range(maxO + 1):
                for P in range(1, maxP + 1):
                    try:
                        dims = part(size, P, o)
                    except Exception:
                        continue
                    assert all(x > 0 for x in dims), (name, size, P, o)
                    assert math.prod(dims) == P
                    if o == 0:
                        assert dims == legacy_partition(size, P), (name, size, P, dims)


def test_partitioners_agree_on_failures_and_results():
    """the three implementations accept / reject the same (size, P, overlap) and return the same dims"""
    ps = partitioners()
    for size in itertools.product((1, 2, 3, 4, 6, 9), repeat=3):
        for P in (1, 2, 3, 4, 6, 8, 12):
            for o in (0, 1, 2):
                res = []
                for name, part in ps:
                    try:
                        res.append(part(size, P, o))
                    except Exception:
                        res.append(None)
                assert all(r == res[0] for r in res), (size, P, o, res)
