"""The oracle's ring buffer (oracle/shim/ring: a restatement of the subset of mrc-ide/ring
that the reference links to; the real source is not under /root/reference) against the
behaviour mrc-ide/ring documents for its C API (ring/inst/include/ring/ring.h) and the
behaviour the reference's call sites rely on:

  * capacity `size` elements; `used` counts committed elements; the head is always writable
  * head_advance on a full buffer: OVERFLOW_OVERWRITE drops the oldest element (the tail
    moves), OVERFLOW_GROW enlarges the buffer and keeps every element
  * tail_offset(k): the k-th oldest element, head_offset(k): the k-th newest (0 = the one
    most recently committed); NULL once k reaches `used`
  * search_bisect(i, pred): for a predicate that holds on a prefix of the elements in tail ->
    head order, the LAST element it holds on, whatever the starting guess i; NULL if it
    fails on the tail (src/dopri.c:756-767: "did not find time in history")
  * read(n): the n oldest elements in order, non-destructively; NULL if there are fewer
  * mirror: a copy with the same contents and positions (src/dopri.c:122-124)

TEST INFRASTRUCTURE for test infrastructure: a mistake in the shim would be reproduced by both
the reference checker and the device code (ADVICE, round 1), so its contract is pinned here."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "oracle" / "shim" / "ring" / "ring.c"
OVERWRITE, GROW = 0, 1
PRED = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_void_p)


@pytest.fixture(scope="module")
def ring(tmp_path_factory):
    so = tmp_path_factory.mktemp("ring") / "libring_shim.so"
    subprocess.run(["gcc", "-O1", "-shared", "-fPIC", "-I", str(SRC.parent.parent), str(SRC),
                    "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.ring_buffer_create.restype = C.c_void_p
    lib.ring_buffer_create.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
    lib.ring_buffer_destroy.argtypes = [C.c_void_p]
    for name in ("ring_buffer_size", "ring_buffer_used"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.c_void_p, C.c_bool]
    lib.ring_buffer_is_empty.restype = C.c_bool
    lib.ring_buffer_is_empty.argtypes = [C.c_void_p]
    for name in ("ring_buffer_tail_offset", "ring_buffer_head_offset"):
        getattr(lib, name).restype = C.c_void_p
        getattr(lib, name).argtypes = [C.c_void_p, C.c_size_t]
    lib.ring_buffer_head_advance.restype = C.c_void_p
    lib.ring_buffer_head_advance.argtypes = [C.c_void_p]
    lib.ring_buffer_search_bisect.restype = C.c_void_p
    lib.ring_buffer_search_bisect.argtypes = [C.c_void_p, C.c_size_t, PRED, C.c_void_p]
    lib.ring_buffer_mirror.argtypes = [C.c_void_p, C.c_void_p]
    lib.ring_buffer_read.restype = C.c_void_p
    lib.ring_buffer_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    return lib


class Buf:
    """a ring of doubles (stride 8); push() writes the head element and commits it"""

    def __init__(self, lib, size, policy):
        self.lib, self.b = lib, lib.ring_buffer_create(size, 8, policy)
        self.head = C.cast(C.cast(self.b, C.POINTER(C.c_void_p))[5], C.c_void_p).value  # ->head

    def push(self, x):
        C.cast(self.head, C.POINTER(C.c_double))[0] = x
        self.head = self.lib.ring_buffer_head_advance(self.b)

    def used(self):
        return self.lib.ring_buffer_used(self.b, False)

    def at(self, p):
        return None if not p else C.cast(p, C.POINTER(C.c_double))[0]

    def tail(self, k):
        return self.at(self.lib.ring_buffer_tail_offset(self.b, k))

    def newest(self, k):
        return self.at(self.lib.ring_buffer_head_offset(self.b, k))

    def last_le(self, x, guess=0):
        pred = PRED(lambda p, _: C.cast(p, C.POINTER(C.c_double))[0] <= x)
        return self.at(self.lib.ring_buffer_search_bisect(self.b, guess, pred, None))

    def contents(self):
        return [self.tail(k) for k in range(self.used())]


def test_fill_and_offsets(ring):
    b = Buf(ring, 4, OVERWRITE)
    assert ring.ring_buffer_is_empty(b.b) and b.used() == 0
    assert ring.ring_buffer_size(b.b, False) == 4 and ring.ring_buffer_size(b.b, True) == 32
    assert b.tail(0) is None and b.newest(0) is None
    for k, x in enumerate([1.0, 2.0, 3.0]):
        b.push(x)
        assert b.used() == k + 1 and not ring.ring_buffer_is_empty(b.b)
        assert b.newest(0) == x and b.tail(0) == 1.0
    assert b.contents() == [1.0, 2.0, 3.0]
    assert [b.newest(k) for k in range(3)] == [3.0, 2.0, 1.0]
    assert b.tail(3) is None and b.newest(3) is None


def test_overwrite_drops_the_oldest(ring):
    b = Buf(ring, 4, OVERWRITE)
    for x in range(1, 11):
        b.push(float(x))
        want = [float(v) for v in range(max(1, x - 3), x + 1)]
        assert b.contents() == want and b.used() == len(want)
        assert b.newest(0) == float(x)
    assert ring.ring_buffer_size(b.b, False) == 4


def test_grow_keeps_everything(ring):
    b = Buf(ring, 3, GROW)
    for x in range(1, 40):
        b.push(float(x))
        assert b.contents() == [float(v) for v in range(1, x + 1)]
    assert ring.ring_buffer_size(b.b, False) >= 39


@pytest.mark.parametrize("policy,n_push", [(OVERWRITE, 5), (OVERWRITE, 23), (GROW, 23)])
def test_search_bisect_finds_the_last_true(ring, policy, n_push):
    b = Buf(ring, 8, policy)
    for x in range(n_push):
        b.push(10.0 * x)
    have = b.contents()
    assert have == sorted(have)
    for guess in range(b.used()):
        assert b.last_le(have[0] - 1.0, guess) is None  # fails on the tail: not found
        assert b.last_le(have[-1] + 5.0, guess) == have[-1]  # holds everywhere: the newest
        for v in have:
            assert b.last_le(v, guess) == v
            assert b.last_le(v + 4.0, guess) == v
    empty = Buf(ring, 8, policy)
    assert empty.last_le(1.0) is None


def test_read_and_mirror(ring):
    b = Buf(ring, 5, OVERWRITE)
    for x in range(8):
        b.push(float(x))
    out = np.full(5, np.nan)
    assert ring.ring_buffer_read(b.b, out.ctypes.data, 6) is None  # more than there is
    assert ring.ring_buffer_read(b.b, out.ctypes.data, 3) is not None
    np.testing.assert_array_equal(out[:3], [3.0, 4.0, 5.0])
    assert b.used() == 5  # non-destructive
    c = Buf(ring, 5, OVERWRITE)
    ring.ring_buffer_mirror(b.b, c.b)
    assert c.contents() == b.contents() == [3.0, 4.0, 5.0, 6.0, 7.0]
    assert c.newest(0) == 7.0
