"""GPU parity of the block-per-trajectory kernel (dopri_coop_kernel.cuh: large
models, cooperative model ABI) against the CPU checkers: bit-exact results,
identical statistics and return codes -- the weighted norms are summed in the
reference's order, so there is no tolerance here either."""
import numpy as np
import pytest

import dde_b200
from dde_b200 import workloads
from oracle import pyoracle
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


def check(backend, model, y, times, parms, tcrit=None, **kw):
    got = dde_b200.dopri_batch(y, times, model, parms, tcrit=tcrit, **kw)
    want = pyoracle.dopri_batch(backend, model, y, times, parms, tcrit=tcrit, **kw)
    np.testing.assert_array_equal(got.code, want.code)
    np.testing.assert_array_equal(got.statistics, want.statistics)
    assert_bits_equal(got.t_last, want.t_last, "t_last")
    assert_bits_equal(got.step_size, want.step_size, "step_size")
    assert_bits_equal(got.y, want.y, "y")
    return got


@pytest.mark.parametrize("method", ["dopri853", "dopri5"])
def test_c5_seir_age(backend, method):
    # BASELINE.json configs[4], reduced: n = 512, two lags through ylag_vec_int
    model, y0, parms, times, kw = workloads.seir_age(12)
    kw["method"] = method
    got = check(backend, model, y0, times, parms, **kw)
    assert (got.code == 1).all()
    # the population is conserved by the model (births = deaths)
    np.testing.assert_allclose(got.y[:, -1].sum(axis=1), 1e7, rtol=1e-9)


def test_seir_age_history_underflow(backend):
    # the ring is too short for the 21-day lag: "did not find time in history"
    model, y0, parms, times, kw = workloads.seir_age(3)
    kw["n_history"] = 3
    got = check(backend, model, y0, times, parms, **kw)
    assert (got.code == -6).all()


def test_seir_age_history_export(backend):
    model, y0, parms, times, kw = workloads.seir_age(1, t_end=40.0, n_times=5)
    res = dde_b200.dopri(y0[0], times, model, parms[0], **kw)
    want = pyoracle.dopri_history(backend, model, y0[0], times, parms[0], **kw)
    assert_bits_equal(res.history, want.history, "history")
    assert_bits_equal(np.asarray(res)[:, 1:], want.y, "y")


@pytest.mark.parametrize("method", ["dopri5", "dopri853"])
@pytest.mark.parametrize("n", [1, 33, 70, 128])
def test_warp_per_trajectory_models(backend, method, n):
    # one warp per trajectory, runtime n incl. ragged widths
    rng = np.random.default_rng(n)
    y0 = rng.uniform(0.5, 2.0, size=(9, n))
    r = rng.uniform(0.1, 1.0, size=(9, n))
    times = np.linspace(0, 3, 13)
    check(backend, "exponential_wide", y0, times, r, method=method)
    check(backend, "exponential_wide", y0, times, r, method=method, tcrit=[0.7, 2.2], stiff_check=1)
    check(backend, "delayed_growth_wide", y0, times, r, method=method, n_history=200)
    # backward in time (dopri5; the dop853 quirk of SURVEY.md A.1#1 gives -4)
    got = check(backend, "exponential_wide", y0, -times, r, method=method)
    assert (got.code == (1 if method == "dopri5" else -4)).all()


def test_coop_controls(backend):
    rng = np.random.default_rng(5)
    y0 = rng.uniform(0.5, 2.0, size=(4, 40))
    r = rng.uniform(0.1, 1.0, size=(4, 40))
    times = np.linspace(0, 3, 7)
    got = check(backend, "exponential_wide", y0, times, r, step_max_n=3)
    assert (got.code == -3).all()
    check(backend, "exponential_wide", y0, times, r, step_size_initial=1e-3)
    check(backend, "exponential_wide", y0, times, r, step_size_max=0.05, return_initial=False)
    got = check(backend, "exponential_wide", y0, times, r, step_size_min=0.5)
    assert (got.code == -4).all()
    check(backend, "exponential_wide", y0, times, r, step_size_min=0.5, step_size_min_allow=True)
    y_bad = y0.copy()
    y_bad[2, 7] = np.inf
    got = check(backend, "exponential_wide", y_bad, times, r)
    assert got.code.tolist() == [1, 1, -8, 1]
    got = check(backend, "exponential_wide", y0[:1], [0.0, 10.0], 1e5 * np.ones((1, 40)),
                stiff_check=1)
    assert got.code[0] == -7
    with pytest.raises(dde_b200.DdeError):  # wider than the registered block
        dde_b200.dopri_batch(np.ones((1, 200)), times, "exponential_wide", np.ones((1, 200)))
