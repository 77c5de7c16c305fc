"""dde_pow / dde_exp (include/dde_b200/dde_math.h), host build, against this
platform's libm -- the pow the reference's step controller links to
(src/dopri.c:500,517,521,569).  Bit-exact on every input; the same source runs
on the device."""
import ctypes as C
import ctypes.util
import math
import struct

import numpy as np
import pytest

from dde_b200 import _lib

libm = C.CDLL(ctypes.util.find_library("m"))
libm.pow.restype = C.c_double
libm.pow.argtypes = [C.c_double, C.c_double]
libm.exp.restype = C.c_double
libm.exp.argtypes = [C.c_double]


def bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def same(a, b):
    return (math.isnan(a) and math.isnan(b)) or bits(a) == bits(b)


def check_pow(xs, ys):
    lib = _lib.load()
    bad = []
    for x, y in zip(xs, ys):
        a, b = lib.dde_pow_host(x, y), libm.pow(x, y)
        if not same(a, b):
            bad.append((x, y, a, b))
    assert not bad, "%d mismatches, first: pow(%r, %r) = %r, libm %r" % ((len(bad),) + bad[0])


def test_pow_controller_domain():
    # err^0.125, err^0.17 (src/dopri.c:147,153), fac_old^0.04, (0.01/der12)^(1/order)
    rng = np.random.default_rng(7)
    n = 60000
    err = np.exp(rng.uniform(-40, 12, n))
    for e in (0.125, 0.2 - 0.04 * 0.75, 0.04, 0.2, 1.0 / 5, 1.0 / 8):
        check_pow(err.tolist(), [e] * n)


def test_pow2_shares_the_logarithm_and_nothing_else():
    """dde_pow2(x, y1, y2): the dopri5 controller's err^0.17 and err^0.04 from one log(err);
    each result the bits of libm's pow, special bases included"""
    lib = _lib.load()
    rng = np.random.default_rng(17)
    xs = np.concatenate([np.exp(rng.uniform(-60, 20, 40000)), rng.uniform(0.9, 1.1, 5000),
                         [0.0, -0.0, 1.0, 1e-4, np.inf, np.nan, 5e-324, 2.2e-308, -1.0, 1e308]])
    out = (C.c_double * 2)()
    for y1, y2 in ((0.2 - 0.04 * 0.75, 0.04), (0.125, 0.0), (3.0, -2.5)):
        bad = []
        for x in xs.tolist():
            lib.dde_pow2_host(x, y1, y2, out)
            if not (same(out[0], libm.pow(x, y1)) and same(out[1], libm.pow(x, y2))):
                bad.append((x, out[0], out[1]))
        assert not bad, "%d mismatches, first %r" % (len(bad), bad[0])


def test_pow_model_domain():
    # Mackey-Glass: y^n with y in (0, 2.5), n in [8, 12]
    rng = np.random.default_rng(8)
    n = 200000
    check_pow(rng.uniform(1e-3, 2.5, n).tolist(), rng.uniform(8.0, 12.0, n).tolist())


def test_pow_wide_domain():
    rng = np.random.default_rng(9)
    n = 100000
    x = np.exp(rng.uniform(-700, 700, n))
    y = rng.uniform(-40, 40, n)
    check_pow(x.tolist(), y.tolist())
    # negative bases with integer and non-integer exponents
    check_pow((-x[:5000]).tolist(), np.round(y[:5000]).tolist())
    check_pow((-x[:2000]).tolist(), y[:2000].tolist())


def test_pow_special_values():
    inf, nan = math.inf, math.nan
    sub = 5e-324
    vals = [0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 3.0, -3.0, inf, -inf, nan, sub, -sub,
            2.2250738585072014e-308, 1.7976931348623157e308, 1e-300, 1e300, 1.0 + 2 ** -52,
            1.0 - 2 ** -53, 2.0 ** -70, -(2.0 ** -70), 2.0 ** 64, 1023.0, 1024.0, -1074.0, -1075.0,
            0.125, 1e-4, 709.78, -745.13]
    xs = [x for x in vals for _ in vals]
    ys = [y for _ in vals for y in vals]
    check_pow(xs, ys)


def test_exp():
    lib = _lib.load()
    rng = np.random.default_rng(10)
    xs = np.concatenate([rng.uniform(-750, 750, 100000), rng.uniform(-2, 2, 100000),
                         np.exp(rng.uniform(-60, 0, 20000)),
                         [0.0, -0.0, math.inf, -math.inf, 709.782712893384, 709.7827128933841,
                          -745.1332191019411, -745.1332191019412, -708.3964185322641, 1e-320]])
    bad = [(x, lib.dde_exp_host(x), libm.exp(x)) for x in xs.tolist()
           if not same(lib.dde_exp_host(x), libm.exp(x))]
    assert not bad, "%d mismatches, first %r" % (len(bad), bad[0])
    assert math.isnan(lib.dde_exp_host(math.nan))
