"""BASELINE.json's configurations at FULL size on one GPU, checked through
properties that do not need the CPU checker to integrate the whole batch:

  * slice invariance: a trajectory's bits do not depend on the batch it is in
    (the same trajectories re-run as a small batch; a sharded run is the
    concatenation of its shards for the same reason);
  * the reference itself on a strided sample across the full index range;
  * identities of the statistics the solver returns (evaluation counts follow
    from step counts, src/dopri_853.c / src/dopri_5.c);
  * dense-output consistency: the exported history, interpolated at the output
    times, reproduces the output rows;
  * an invariant of the model (the delayed SIR map conserves S + I + R).
"""
import numpy as np
import pytest

import dde_b200
from dde_b200 import workloads
from oracle import pyoracle
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    return "ref" if pyoracle.available("ref") else "oracle"


def _sample(n_total, k, seed):
    rng = np.random.default_rng(seed)
    idx = np.unique(np.concatenate([[0, n_total - 1], rng.integers(0, n_total, k)]))
    return idx


def test_c2_mackey_glass_full_batch(backend):
    n_traj = 1 << 20
    model, y0, parms, times, kw = workloads.mackey_glass(n_traj)
    b = dde_b200.DopriBatch(model, 1, n_traj, n_parms=3, **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    full = b.download()
    st, code = full.statistics.astype(np.int64), full.code
    # codes: complete, or one of the two failures the reference produces on this workload
    assert set(np.unique(code)) <= {1, -5, -6}
    ok = code == 1
    assert ok.mean() > 0.97
    # dop853: 11 evaluations per attempt + 5 per accepted step + k1 and the step-size probe
    np.testing.assert_array_equal(st[ok, 0], 2 + 11 * st[ok, 1] + 5 * st[ok, 2])
    # every attempt is accepted or rejected; rejections before the first accept do not count
    assert (st[ok, 1] >= st[ok, 2] + st[ok, 3]).all()
    assert (st[ok, 1] - st[ok, 2] - st[ok, 3] <= 30).all()
    assert (full.t_last[ok] == times[-1]).all()
    assert np.isfinite(full.y[ok]).all()
    # failed trajectories stop early and leave the unreached rows zero (src/r_dopri.c:169)
    bad = np.flatnonzero(~ok)
    assert (full.t_last[bad] < times[-1]).all()
    assert (full.y[bad, -1, 0] == 0.0).all()

    # slice invariance: 2048 scattered trajectories re-run as a batch of their own
    idx = _sample(n_traj, 2048, 1)
    small = dde_b200.dopri_batch(y0[idx], times, model, parms[idx], **kw)
    np.testing.assert_array_equal(small.code, code[idx])
    np.testing.assert_array_equal(small.statistics, full.statistics[idx])
    assert_bits_equal(small.y, full.y[idx], "y of the re-run sample")
    assert_bits_equal(small.step_size, full.step_size[idx], "step_size")

    # the reference on a strided sample over the whole index range, failures included
    idx = np.unique(np.concatenate([_sample(n_traj, 160, 2), bad[:32]]))
    want = pyoracle.dopri_batch(backend, model, y0[idx], times, parms[idx], **kw)
    np.testing.assert_array_equal(code[idx], want.code)
    np.testing.assert_array_equal(full.statistics[idx], want.statistics)
    assert_bits_equal(full.y[idx], want.y, "y vs the reference")
    assert_bits_equal(full.t_last[idx], want.t_last, "t_last")

    # dense-output consistency: the last 64 steps of a few completed trajectories,
    # interpolated at the output times they cover, give the output rows again
    for j in np.flatnonzero(ok)[[0, 12345, -1]]:
        h = b.history(int(j))
        assert h.shape[0] == min(64, st[j, 2])
        t_old, hh = h[:, -2], h[:, -1]
        np.testing.assert_array_equal(t_old[1:], t_old[:-1] + hh[:-1])  # entries abut, bit for bit
        covered = times[(times >= t_old[0]) & (times <= t_old[-1] + hh[-1])]
        if covered.size:
            hist = h.view(dde_b200.api.DopriHistory)
            hist.n = 1
            yi = dde_b200.dopri_interpolate(hist, covered)
            rows = full.y[j, np.isin(times, covered), 0]
            assert_bits_equal(np.asarray(yi).ravel(), rows, "dense output vs output rows")
    b.close()


def test_c3_sweeps_full_batch(backend):
    """2 x 2,097,152 trajectories x 1000 output rows (50 GB of rows per model stay on the
    device; sampled ranges come back through dde_dopri_batch_download_range)."""
    for make in (workloads.lorenz_sweep, workloads.rossler_sweep):
        n_traj = 1 << 21
        model, y0, parms, times, kw = make(n_traj)
        b = dde_b200.DopriBatch(model, 3, n_traj, n_parms=3, **kw)
        b.upload(y0, parms)
        b.set_times(times)
        b.run()
        meta = b.download(want_y=False)
        st = meta.statistics.astype(np.int64)
        assert (meta.code == 1).all()
        np.testing.assert_array_equal(st[:, 0], 2 + 6 * st[:, 1])  # dopri5: 6 per attempt
        assert (meta.t_last == times[-1]).all()
        for first in (0, 777_777, n_traj - 8):
            got = b.download_range(first, 8)
            want = pyoracle.dopri_batch(backend, model, y0[first:first + 8], times,
                                        parms[first:first + 8], **kw)
            np.testing.assert_array_equal(got.statistics, want.statistics)
            np.testing.assert_array_equal(got.statistics, meta.statistics[first:first + 8])
            assert_bits_equal(got.y, want.y, "%s rows %d.." % (model, first))
        b.close()


def test_c4_delayed_sir_full_batch(backend):
    n_rep = 1 << 20
    model, y0, parms, steps, kw = workloads.sir_delay(n_rep)  # 10 000 steps each
    y = dde_b200.difeq_replicate(n_rep, y0, steps, model, parms, n_history=kw["n_history"],
                                 return_initial=False, return_step=False)
    assert (y.code == 1).all()
    final = np.asarray(y)[0]  # [3, n_rep]
    assert np.isfinite(final).all()
    # S + I + R is conserved by the map up to rounding, whatever the parameters
    np.testing.assert_allclose(final.sum(axis=0), 1.0, rtol=0, atol=1e-12)
    assert (final >= -1e-12).all()
    idx = _sample(n_rep, 48, 3)
    want = pyoracle.difeq_batch(backend, model, y0[idx], steps, parms[idx],
                                n_history=kw["n_history"], return_initial=False)
    assert_bits_equal(final[:, idx].T, want.y[:, 0, :], "final state vs the reference")


def test_c5_seir_age_batch(backend):
    """n = 512 with two lags, block per trajectory, at the size one GPU of eight integrates of
    the configuration's 65 536 (8192 trajectories, 34 GB of history rings; the sharded whole:
    tools/run_c5_sharded.py): many waves of resident blocks (444), slice invariance and the
    reference on a few of them."""
    n_traj = 8192
    model, y0, parms, times, kw = workloads.seir_age(n_traj)
    full = dde_b200.dopri_batch(y0, times, model, parms, **kw)
    assert (full.code == 1).all()
    st = full.statistics.astype(np.int64)
    np.testing.assert_array_equal(st[:, 0], 2 + 11 * st[:, 1] + 5 * st[:, 2])
    assert np.isfinite(full.y).all() and (full.y[:, -1, :] > -1e-6).all()
    idx = _sample(n_traj, 4, 4)
    small = dde_b200.dopri_batch(y0[idx], times, model, parms[idx], **kw)
    assert_bits_equal(small.y, full.y[idx], "slice invariance")
    want = pyoracle.dopri_batch(backend, model, y0[idx[:3]], times, parms[idx[:3]], **kw)
    np.testing.assert_array_equal(full.statistics[idx[:3]], want.statistics)
    assert_bits_equal(full.y[idx[:3]], want.y, "y vs the reference")
