"""The device code, compiled by g++ and run one lane at a time (tests/emu), against
the oracle: control flow and arithmetic of the stepping kernels checked bit for bit
where there is no GPU.  Nothing warp-wide is exercised here (a warp of one lane);
the GPU parity tests are the gate for that."""
import numpy as np
import pytest

from conftest import assert_bits_equal
from dde_b200 import workloads
from oracle import pyoracle

pyemu = pytest.importorskip("emu.pyemu")

needs_oracle = pytest.mark.skipif(not pyoracle.available("oracle"),
                                  reason="oracle/liboracle.so not built (make -C oracle oracle)")


def compare(got, want):
    np.testing.assert_array_equal(got.code, want.code)
    np.testing.assert_array_equal(got.statistics, want.statistics)
    assert_bits_equal(got.t_last, want.t_last, "t_last")
    assert_bits_equal(got.step_size, want.step_size, "step_size")
    assert_bits_equal(got.y, want.y, "y")


@needs_oracle
@pytest.mark.parametrize("mode", [0, 2, 4])
def test_mackey_glass(mode):
    model, y0, parms, times, kw = workloads.mackey_glass(192)
    got = pyemu.dopri_batch(model, y0, times, parms, mode=mode, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, **kw)
    assert (want.code != 1).any(), "the sample should include failing trajectories"
    compare(got, want)


@needs_oracle
@pytest.mark.parametrize("n_history", [3, 8])
def test_mackey_glass_small_ring(n_history):
    """ring overwrite / underflow: ERR_YLAG_FAIL where the reference reports it"""
    model, y0, parms, times, kw = workloads.mackey_glass(64)
    kw = dict(kw, n_history=n_history)
    got = pyemu.dopri_batch(model, y0, times, parms, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, **kw)
    compare(got, want)


@needs_oracle
def test_mackey_glass_dopri5_and_tcrit():
    model, y0, parms, times, kw = workloads.mackey_glass(48)
    kw = dict(kw, method="dopri5")
    tcrit = np.array([17.0, 34.0, 51.0])
    got = pyemu.dopri_batch(model, y0, times, parms, tcrit=tcrit, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, tcrit=tcrit, **kw)
    compare(got, want)


@needs_oracle
@pytest.mark.parametrize("mode", [0, 1])
def test_lorenz_sweep(mode):
    model, y0, parms, times, kw = workloads.lorenz_sweep(32, n_times=101, t_end=2.0)
    got = pyemu.dopri_batch(model, y0, times, parms, n_out=2, mode=mode, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, n_out=2, **kw)
    compare(got, want)
    assert_bits_equal(got.output, want.output, "output")


@needs_oracle
@pytest.mark.parametrize("n_times,return_initial", [(12, False), (12, True), (14, True), (102, True),
                                                     (2, False), (2, True)])
def test_lorenz_rows_by_sector(n_times, return_initial):
    """the unrolled kernel writes its rows as aligned 32-byte sectors of the result array
    (dopri_kernels.cuh, STAGE_OUT): blocks of rows that start or end inside a sector"""
    model, y0, parms, times, kw = workloads.lorenz_sweep(37, n_times=n_times, t_end=2.0)
    kw = dict(kw, return_initial=return_initial)
    got = pyemu.dopri_batch(model, y0, times, parms, mode=1, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, **kw)
    compare(got, want)


@needs_oracle
@pytest.mark.parametrize("kwargs", [
    dict(step_size_max=0.5, step_max_n=400),
    dict(step_size_min=0.05, step_size_min_allow=True),
    dict(step_size_min=0.05),
    dict(step_size_initial=0.01),
    dict(rtol=1e-9, atol=1e-9, n_history=200),
    dict(rtol=1e-3, atol=1e-3),
    dict(return_initial=False),
    dict(n_history=3),
    dict(n_history=8),
    dict(n_history=1),
])
def test_mackey_glass_dde1_kernel_options(kwargs):
    """dopri_dde1_kernel (dop853, declared lag) under the controls that steer its branches,
    and with rings small enough to wrap below the lag (ERR_YLAG_FAIL where the reference
    reports it, the redundant evaluation's look-up included)"""
    model, y0, parms, times, kw = workloads.mackey_glass(40)
    kw = dict(kw, **kwargs)
    got = pyemu.dopri_batch(model, y0, times, parms, mode=4, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, **kw)
    compare(got, want)


@needs_oracle
def test_mackey_glass_dde1_kernel_dense_times():
    """many output rows per step, and a start time that is not zero"""
    model, y0, parms, times, kw = workloads.mackey_glass(24)
    times = np.linspace(3.0, 203.0, 801)
    got = pyemu.dopri_batch(model, y0, times, parms, mode=4, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, **kw)
    compare(got, want)


@needs_oracle
@pytest.mark.parametrize("kwargs", [
    dict(tcrit=np.array([17.0, 34.0, 51.0, 400.0])),
    dict(stiff_check=1),
    dict(backward=True),
])
def test_mackey_glass_runs_the_dde1_kernel_does_not_take(kwargs):
    """critical times, the stiffness test and backward time (where the reference's dop853
    error norm goes negative) stay on the general kernel"""
    model, y0, parms, times, kw = workloads.mackey_glass(16)
    kwargs = dict(kwargs)
    tcrit = kwargs.pop("tcrit", None)
    if kwargs.pop("backward", False):
        times = -times
    kw = dict(kw, **kwargs)
    got = pyemu.dopri_batch(model, y0, times, parms, tcrit=tcrit, **kw)
    want = pyoracle.dopri_batch("oracle", model, y0, times, parms, tcrit=tcrit, **kw)
    compare(got, want)
    with pytest.raises(RuntimeError, match="does not take"):
        pyemu.dopri_batch(model, y0, times, parms, tcrit=tcrit, mode=4, **kw)
