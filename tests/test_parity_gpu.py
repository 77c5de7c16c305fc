"""GPU parity: the CUDA path through the C-ABI against the CPU checkers
(oracle/_ref = the unmodified reference; oracle/liboracle.so = the C
restatement) on identical inputs.  Bar: bit-exact outputs, identical
per-trajectory statistics and return codes."""
import numpy as np
import pytest

import dde_b200
from dde_b200 import workloads
from oracle import pyoracle
from conftest import assert_bits_equal

pytestmark = pytest.mark.gpu


def check_dopri(backend, model, y, times, parms, n_out=0, tcrit=None, **kw):
    got = dde_b200.dopri_batch(y, times, model, parms, n_out=n_out, tcrit=tcrit, **kw)
    want = pyoracle.dopri_batch(backend, model, y, times, parms, n_out=n_out, tcrit=tcrit, **kw)
    np.testing.assert_array_equal(got.code, want.code)
    np.testing.assert_array_equal(got.statistics, want.statistics)
    assert_bits_equal(got.t_last, want.t_last, "t_last")
    assert_bits_equal(got.step_size, want.step_size, "step_size")
    assert_bits_equal(got.y, want.y, "y")
    if n_out:
        assert_bits_equal(got.output, want.output, "output")
    return got, want


def test_c1_lorenz_single(backend):
    # BASELINE.json configs[0]; SURVEY.md A.4: 4565 steps / 4365 acc / 200 rej / 27392 evals
    got, _ = check_dopri(backend, "lorenz", [[10.0, 1.0, 1.0]], np.linspace(0, 100, 1001),
                         [10.0, 28.0, 8.0 / 3.0])
    assert got.statistics[0].tolist() == [27392, 4565, 4365, 200]
    assert got.code[0] == 1


@pytest.mark.parametrize("method", ["dopri5", "dopri853"])
def test_lorenz_output_function(backend, method):
    check_dopri(backend, "lorenz", [[10.0, 1.0, 1.0], [1.0, 1.0, 1.0]], np.linspace(0, 5, 51),
                [10.0, 28.0, 8.0 / 3.0], n_out=2, method=method)


@pytest.mark.parametrize("return_initial", [True, False])
def test_c3_sweeps(backend, return_initial):
    for gen in (workloads.lorenz_sweep, workloads.rossler_sweep):
        model, y0, parms, times, kw = gen(1024, n_times=101)
        kw["return_initial"] = return_initial
        check_dopri(backend, model, y0, times, parms, **kw)


def test_c2_mackey_glass(backend):
    model, y0, parms, times, kw = workloads.mackey_glass(4096)
    got, want = check_dopri(backend, model, y0, times, parms, **kw)
    # the failing trajectories are part of parity
    assert (got.code != 1).sum() > 0


def test_mackey_glass_known_answer(backend):
    # 173 steps / 114 acc / 59 rej / 2475 evals, y(500) = 1.059911058006217
    got, _ = check_dopri(backend, "mackey_glass", [[1.2]], np.linspace(0, 500, 11),
                         [0.2, 0.1, 10.0], method="dopri853", n_history=1000)
    assert got.statistics[0].tolist() == [2475, 173, 114, 59]
    assert got.y[0, -1, 0] == 1.059911058006217


@pytest.mark.parametrize("model", ["seir", "seir_int"])
@pytest.mark.parametrize("method,steps", [("dopri5", [500, 83, 74, 9]), ("dopri853", None)])
def test_seir(backend, model, method, steps):
    # SURVEY.md A.4 KAT; seir == seir_int is pinned by tests/testthat/test-dde.R:272-283
    y0 = [[1e7 - 1, 0.0, 1.0, 0.0]]
    got, _ = check_dopri(backend, model, y0, np.linspace(0, 30, 301), None, n_out=1, method=method,
                         rtol=1e-7, atol=1e-7, n_history=1000)
    if steps:
        assert got.statistics[0].tolist() == steps


def test_seir_history_underflow(backend):
    # tests/testthat/test-dde.R:266-270: n_history = 2 -> "did not find time in history"
    got, _ = check_dopri(backend, "seir", [[1e7 - 1, 0.0, 1.0, 0.0]], np.linspace(0, 30, 301), None,
                         rtol=1e-7, atol=1e-7, n_history=2)
    assert got.code[0] == -6
    assert "did not find time in history" in got.message(0)


def test_ylag_without_history(backend):
    # tests/testthat/test-dde.R:285-295
    got, _ = check_dopri(backend, "seir", [[1e7 - 1, 0.0, 1.0, 0.0]], np.linspace(0, 30, 31), None)
    assert got.code[0] == -6
    assert got.message(0) == "Integration failure: can't use ylag in model with no history"


@pytest.mark.parametrize("method", ["dopri5", "dopri853"])
def test_growth_models_runtime_n(backend, method):
    rng = np.random.default_rng(1)
    for n in (1, 3, 5):
        y0 = rng.uniform(0.5, 2.0, size=(7, n))
        r = rng.uniform(0.1, 1.0, size=(7, n))
        check_dopri(backend, "exponential", y0, np.linspace(0, 3, 13), r, method=method)
        check_dopri(backend, "linear", y0, np.linspace(0, 3, 13), r, method=method)
        check_dopri(backend, "delayed_growth", y0, np.linspace(0, 3, 13), r, method=method,
                    n_history=200)
        # the same through ylag_vec (size_t indices), mirrored components
        check_dopri(backend, "delayed_growth_vec", y0, np.linspace(0, 3, 13), r, method=method,
                    n_history=200)
        # ... and through ylag_all by a thread-per-trajectory model (src/dopri.c:790-800;
        # lag_ctx.cuh ylag_all: pre-history, staged window and ring paths)
        check_dopri(backend, "delayed_growth_all", y0, np.linspace(0, 3, 13), r, method=method,
                    n_history=200)
        check_dopri(backend, "delayed_growth_all", y0, np.linspace(0, 6, 7), r, method=method,
                    n_history=4)  # a ring the lag outruns: ERR_YLAG_FAIL where the reference has it


def test_a_wrong_lag_declaration_is_reported_not_integrated():
    """the kernel for one delay equation with a declared lag (dopri_dde1_kernel.cuh) answers the
    declared look-up itself; a model that then asks for something else ends with
    DDE_ERR_LAG_UNDECLARED, on the general kernel (which does not use the declaration) it runs"""
    model, y0, parms, times, kw = workloads.mackey_glass(300)
    got = dde_b200.dopri_batch(y0, times, "mg_wrong_lag", parms, **kw)
    # (where the first step is longer than the declared lag, the declared look-up itself
    # already fails: ERR_YLAG_FAIL comes first, as it would in the reference)
    assert np.isin(got.code, (-11, -6)).all() and (got.code == -11).sum() > 250
    assert (got.statistics[:, 1] == 1).all() and (got.statistics[:, 2] == 0).all()
    assert "did not declare" in dde_b200.strerror(-11)
    ok = dde_b200.dopri_batch(y0, times, "mg_wrong_lag", parms, **dict(kw, stiff_check=1000000))
    right = dde_b200.dopri_batch(y0, times, model, parms, **dict(kw, stiff_check=1000000))
    np.testing.assert_array_equal(ok.code, right.code)
    assert_bits_equal(ok.y, right.y, "y")


def test_reference_example_model_as_plugin(backend):
    """/root/reference/inst/examples/seir.c itself, compiled as a model library with only the
    DDE_MODEL_FN decoration injected (tests/models/build_ref_example.sh), loaded at run time:
    the same bits as the reference solver running the reference's build of that file."""
    from conftest import ROOT
    plug = ROOT / "tests" / "models" / "_ref" / "libdde_ref_example.so"
    if not plug.exists():
        pytest.skip("tests/models/_ref/libdde_ref_example.so not built (needs the reference tree)")
    dde_b200.load_model_library(plug)
    assert "seir_reference_example" in dde_b200.models()
    y0 = np.array([[1e7 - 1, 0.0, 1.0, 0.0], [9e6, 5e5, 2e5, 3e5]])
    times = np.linspace(0, 365, 74)
    for method in ("dopri5", "dopri853"):
        kw = dict(n_out=1, method=method, n_history=1000)
        got = dde_b200.dopri_batch(y0, times, "seir_reference_example", None, **kw)
        want = pyoracle.dopri_batch(backend, "seir", y0, times, None, **kw)
        np.testing.assert_array_equal(got.statistics, want.statistics)
        np.testing.assert_array_equal(got.code, want.code)
        assert_bits_equal(got.y, want.y, "y")
        assert_bits_equal(got.output, want.output, "output")


def test_negative_time(backend):
    # tests/testthat/test-ode.R:511-591 (dopri5); the dop853 backward quirk is
    # SURVEY.md A.1#1 and must be reproduced too
    y0 = [[1.0, 2.0]]
    r = [[0.5, 0.25]]
    check_dopri(backend, "exponential", y0, -np.linspace(0, 2, 9), r, method="dopri5")
    got, _ = check_dopri(backend, "exponential", y0, -np.linspace(0, 2, 9), r, method="dopri853")
    assert got.code[0] == -4


def test_critical_times(backend):
    check_dopri(backend, "lorenz", [[10.0, 1.0, 1.0]], np.linspace(0, 2, 21), [10.0, 28.0, 8.0 / 3.0],
                tcrit=[0.55, 1.0, 1.25, 7.0])
    check_dopri(backend, "lorenz", [[10.0, 1.0, 1.0]], np.linspace(0, 2, 21), [10.0, 28.0, 8.0 / 3.0],
                tcrit=[-1.0, 0.0, 0.3], method="dopri853")


def test_controls(backend):
    lor = ("lorenz", [[10.0, 1.0, 1.0]], np.linspace(0, 3, 31), [10.0, 28.0, 8.0 / 3.0])
    # n_eval drops by exactly 1 with step_size_initial (tests/testthat/test-ode.R:398-404)
    # when the given initial step is the one dopri_h_init would have chosen: its one
    # extra evaluation (src/dopri.c:555) is all that goes
    # (tests/testthat/test-ode.R:384-404: the first step read off the history)
    tt = np.linspace(0, 1, 200)
    lor1 = ("lorenz", [[10.0, 1.0, 1.0]], tt, [10.0, 28.0, 8.0 / 3.0])
    a, _ = check_dopri(backend, *lor1, n_history=500)
    hb = dde_b200.DopriBatch("lorenz", 3, 1, n_parms=3, n_history=500)
    hb.upload(np.array([[10.0, 1.0, 1.0]]), np.array([[10.0, 28.0, 8.0 / 3.0]]))
    hb.set_times(tt)
    hb.run()
    h0 = float(hb.history(0)[0, -1])
    hb.close()
    b, _ = check_dopri(backend, *lor1, n_history=500, step_size_initial=h0)
    assert b.statistics[0, 0] == a.statistics[0, 0] - 1
    np.testing.assert_array_equal(b.statistics[0, 1:], a.statistics[0, 1:])
    assert_bits_equal(a.y, b.y, "y")
    check_dopri(backend, *lor, step_size_initial=1e-3)
    got, _ = check_dopri(backend, *lor, step_max_n=10)
    assert got.code[0] == -3
    check_dopri(backend, *lor, step_size_max=0.01)
    got, _ = check_dopri(backend, *lor, step_size_min=0.05)
    assert got.code[0] == -4
    check_dopri(backend, *lor, step_size_min=0.05, step_size_min_allow=True)
    check_dopri(backend, *lor, rtol=1e-9, atol=1e-11, method="dopri853")
    for method in ("dopri5", "dopri853"):
        check_dopri(backend, *lor, stiff_check=1, method=method)
        # a stiff problem trips the detector: y' = -r y with huge r
        got, _ = check_dopri(backend, "exponential", [[1.0]], [0.0, 10.0], [[1e5]], stiff_check=1,
                             method=method)
        assert got.code[0] == -7


def test_time_validation(backend):
    got, _ = check_dopri(backend, "lorenz", [[10.0, 1.0, 1.0]], [0.0, 1.0, 0.0], [10.0, 28.0, 2.0])
    assert got.code[0] == -1
    got, _ = check_dopri(backend, "lorenz", [[10.0, 1.0, 1.0]], [0.0, 2.0, 1.0, 3.0], [10.0, 28.0, 2.0])
    assert got.code[0] == -2
    # duplicated times are allowed (tests/testthat/test-ode.R:23-28)
    check_dopri(backend, "lorenz", [[10.0, 1.0, 1.0]], [0.0, 0.5, 0.5, 1.0, 1.0], [10.0, 28.0, 2.0])


def test_nonfinite_initial_derivative(backend):
    got, _ = check_dopri(backend, "exponential", [[np.inf, 1.0]], [0.0, 1.0], [[1.0, 1.0]])
    assert got.code[0] == -8


@pytest.mark.parametrize("method", ["dopri5", "dopri853"])
def test_history_and_interpolate(backend, method):
    # tests/testthat/test-ode.R:50-59: history has one entry per accepted step;
    # dense output from the history equals direct output
    y0 = [10.0, 1.0, 1.0]
    p = [10.0, 28.0, 8.0 / 3.0]
    times = np.linspace(0, 2, 41)
    res = dde_b200.dopri(y0, times, "lorenz", p, method=method, n_history=1000,
                         return_statistics=True)
    want = pyoracle.dopri_history(backend, "lorenz", y0, times, p, method=method, n_history=1000)
    assert res.history.shape == want.history.shape
    assert res.history.shape[0] == res.statistics["n_accept"]
    assert_bits_equal(res.history, want.history, "history")
    assert_bits_equal(np.asarray(res)[:, 1:], want.y, "y")
    dense = dde_b200.dopri_interpolate(res, times)
    # R's dopri_interpolate uses the entry with t_entry <= t (findInterval),
    # the solver the step that crossed t: equal to 1e-10 (test-ode.R:56-59)
    np.testing.assert_allclose(dense, want.y, rtol=1e-10, atol=1e-12)
    with pytest.raises(dde_b200.DdeError, match="outside of range"):
        dde_b200.dopri_interpolate(res, [2.5])


@pytest.mark.parametrize("method", ["dopri5", "dopri853"])
def test_interpolate_matches_the_restated_r_function(method):
    """dopri_interpolate (R/dopri.R:577-642) restated in numpy (oracle/pyinterp.py), against
    both interpolation entry points, bit for bit: the host-history one
    (dde_dopri_interpolate_batch) and the one that reads the rings of a live batch handle in
    device memory (dde_dopri_batch_interpolate), rings that have wrapped included."""
    from oracle import pyinterp
    rng = np.random.default_rng(17)
    # a three-equation system, histories that fit the ring
    n_traj = 9
    model, y0, parms, times, kw = workloads.lorenz_sweep(n_traj, n_times=11, t_end=2.0)
    kw = dict(kw, method=method, n_history=2000)
    b = dde_b200.DopriBatch(model, 3, n_traj, n_parms=3, **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    tq = np.concatenate([[0.0, 2.0], rng.uniform(0.0, 2.0, 300)])
    got = b.interpolate(tq)                       # [n_traj, n, n_t]
    for j in range(n_traj):
        h = b.history(j)
        want = pyinterp.dopri_interpolate(h, 3, tq)   # [n_t, n]
        assert_bits_equal(got[j].T, want, "from the rings, trajectory %d" % j)
        hist = h.view(dde_b200.DopriHistory)
        hist.n = 3
        assert_bits_equal(dde_b200.dopri_interpolate(hist, tq), want, "from a host history")
    with pytest.raises(dde_b200.DdeError, match="outside of range of known history"):
        b.interpolate([2.5])
    b.close()
    # the Mackey-Glass batch with rings that wrap: only the surviving entries answer
    model, y0, parms, times, kw = workloads.mackey_glass(96)
    kw = dict(kw, method=method, n_history=16)
    probe = dde_b200.dopri_batch(y0, times, model, parms, **kw)
    keep = np.flatnonzero(probe.code == 1)       # trajectories that reach t = 500
    assert 10 < keep.shape[0] < 96 or keep.shape[0] == 96
    y0, parms, n_traj = y0[keep], parms[keep], keep.shape[0]
    b = dde_b200.DopriBatch(model, 1, n_traj, n_parms=3, **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    hs = [b.history(j) for j in range(n_traj)]
    assert all(h.shape[0] == 16 for h in hs)
    lo = max(h[0, -2] for h in hs)                # inside every surviving history
    tq = np.concatenate([[lo, 500.0], rng.uniform(lo, 500.0, 200)])
    got = b.interpolate(tq)
    for j in range(n_traj):
        assert_bits_equal(got[j].T, pyinterp.dopri_interpolate(hs[j], 1, tq), "wrapped ring")
    # before the oldest surviving entry of some trajectory: refused, as R does
    with pytest.raises(dde_b200.DdeError, match="outside of range of known history"):
        b.interpolate([min(h[0, -2] for h in hs) - 1.0])
    b.close()


def test_ring_overwrite_history(backend):
    # overwrite policy: the ring keeps the newest n_history entries, oldest first
    y0 = [10.0, 1.0, 1.0]
    p = [10.0, 28.0, 8.0 / 3.0]
    times = np.linspace(0, 2, 5)
    res = dde_b200.dopri(y0, times, "lorenz", p, n_history=7)
    want = pyoracle.dopri_history(backend, "lorenz", y0, times, p, n_history=7)
    assert res.history.shape[0] == 7
    assert_bits_equal(res.history, want.history, "history")


# ---- difeq -------------------------------------------------------------------


def check_difeq(backend, model, y, steps, parms, n_out=0, n_history=0, return_initial=True):
    y = np.atleast_2d(np.asarray(y, dtype=np.float64))
    st = np.asarray(steps, dtype=np.uint64)
    b = dde_b200.DifeqBatch(model, y.shape[1], y.shape[0], n_parms=0 if parms is None else
                            np.atleast_2d(parms).shape[1], n_out=n_out, n_history=n_history)
    pa = None if parms is None else np.ascontiguousarray(np.atleast_2d(parms), dtype=np.float64)
    shared = pa is not None and pa.shape[0] == 1 and y.shape[0] > 1
    b.upload(np.ascontiguousarray(y), pa, shared_parms=shared)
    b.set_steps(st, return_initial)
    b.run()
    y_out, out, code, y_final = b.download()
    b.close()
    want = pyoracle.difeq_batch(backend, model, y, st, None if pa is None else
                                (pa[0] if shared else pa), n_out=n_out, n_history=n_history,
                                return_initial=return_initial)
    np.testing.assert_array_equal(code, want.code)
    assert_bits_equal(y_out, want.y, "y")
    ok = want.code == 1
    assert_bits_equal(y_final[ok], want.y_final[ok], "y_final")
    if n_out:
        assert_bits_equal(out, want.output, "output")
    return y_out, out, code


def test_difeq_logistic(backend):
    # tests/testthat/test-difeq.R:222-233: r = 1.5, y0 = 0.1, 20 steps
    y, _, _ = check_difeq(backend, "logistic", [[0.1]], np.arange(21), [[1.5]])
    ref = [0.1]
    for _ in range(20):
        ref.append(1.5 * ref[-1] * (1 - ref[-1]))
    np.testing.assert_array_equal(y[0, :, 0], np.array(ref))
    rng = np.random.default_rng(3)
    check_difeq(backend, "logistic", rng.uniform(0.1, 0.9, (33, 4)), [0, 5, 6, 50],
                rng.uniform(1.0, 3.9, (33, 4)), return_initial=False)


def test_difeq_additive_outputs(backend):
    # tests/testthat/test-difeq.R:13-50: y + p, with outputs paired to their row
    for hist in (0, 8):
        for ri in (True, False):
            check_difeq(backend, "additive", [[1.0, 2.0, 3.0]], np.arange(11), [[1.0, 2.0, 3.0]],
                        n_out=2, n_history=hist, return_initial=ri)
            check_difeq(backend, "additive", [[1.0, 2.0, 3.0]], [2, 5, 6, 11], [[1.0, 2.0, 3.0]],
                        n_out=2, n_history=hist, return_initial=ri)


def test_difeq_fibonacci(backend):
    # tests/testthat/test-difeq-delay.R:3-24: 1 2 3 5 8 ... 144 with n_history = 2
    y, _, _ = check_difeq(backend, "fibonacci", [[1.0] * 5], np.arange(11), None, n_history=2)
    assert y[0, :, 0].tolist() == [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144]
    y, out, _ = check_difeq(backend, "fib_out", [[1.0]], np.arange(11), None, n_out=1, n_history=2)
    assert y[0, :, 0].tolist() == [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144]
    assert out[0, :, 0].tolist() == [1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89]


@pytest.mark.parametrize("model", ["fib_all", "fib_vec"])
@pytest.mark.parametrize("n_history", [2, 3, 40, 400])
def test_difeq_yprev_all_and_vec(backend, model, n_history):
    """yprev_all and yprev_vec (size_t indices): on-chip history, and (400 entries) the ring
    in HBM"""
    rng = np.random.default_rng(5)
    for n in (1, 4, 7):
        y0 = np.floor(rng.uniform(1, 9, (6, n)))
        y, _, code = check_difeq(backend, model, y0, np.arange(0, 40), None, n_history=n_history)
        assert (code == 1).all()


def test_difeq_history_miss(backend):
    # the lag reaches further back than the ring holds: src/difeq.c:267-270
    model, y0, parms, steps, kw = workloads.sir_delay(4, n_steps=100)
    _, _, code = check_difeq(backend, model, y0, steps, parms, n_history=8,
                             return_initial=kw["return_initial"])
    assert (code == -9).all()


def test_c4_delayed_sir(backend):
    model, y0, parms, steps, kw = workloads.sir_delay(2048, n_steps=2000)
    check_difeq(backend, model, y0, steps, parms, **kw)
    # every replicate equals an independent difeq() run (R/difeq_replicate.R:1-4)
    single = dde_b200.difeq(y0[5], [0, 2000], model, parms[5], n_history=16)
    many = dde_b200.difeq_replicate(2048, y0, [0, 2000], model, parms, n_history=16)
    assert_bits_equal(np.asarray(single), np.asarray(many)[:, :, 5], "replicate slice")


# ---- multi-GPU sharding ---------------------------------------------------------


def test_sharded_equals_single_device(backend):
    # dde_dopri_batch_sharded / dde_difeq_batch_sharded over every visible GPU
    # (one here, eight on a full box): same results as one device, row for row
    model, y0, parms, times, kw = workloads.mackey_glass(1000)
    one = dde_b200.dopri_batch(y0, times, model, parms, **kw)
    for devices in ([0], "all"):
        many = dde_b200.dopri_batch(y0, times, model, parms, devices=devices, **kw)
        np.testing.assert_array_equal(one.statistics, many.statistics)
        np.testing.assert_array_equal(one.code, many.code)
        assert_bits_equal(one.y, many.y, "y")
    model, y0, parms, steps, kw = workloads.sir_delay(500, n_steps=300)
    a = dde_b200.difeq_replicate(500, y0, steps, model, parms, **kw)
    b = dde_b200.difeq_replicate(500, y0, steps, model, parms, devices="all", **kw)
    assert_bits_equal(np.asarray(a), np.asarray(b), "difeq")
    with pytest.raises(dde_b200.DdeError, match="not visible"):
        dde_b200.dopri_batch(y0[:, :1], times, "mackey_glass", parms[:, :1].repeat(3, 1),
                             devices=[99], method="dopri853", n_history=8)


def test_shards_handle_and_sharded_events(backend):
    """the resident form of the sharded batch (dde_dopri_shards_*: one handle per device,
    allocated once, run twice) and events through the sharded entry points"""
    model, y0, parms, times, kw = workloads.mackey_glass(3000)
    one = dde_b200.dopri_batch(y0, times, model, parms, **kw)
    nt = times.shape[0]
    for devices in ([0], None, [0, 0, 0]):  # the same device three times: three shards
        h = dde_b200.DopriShards(model, 1, 3000, devices=devices, n_parms=3, **kw)
        h.set_times(times)
        y = np.empty((3000, nt, 1))
        stats = np.empty((3000, 4), dtype=np.int32)
        code = np.empty((3000,), dtype=np.int32)
        t_last = np.empty((3000,))
        step = np.empty((3000,))
        for rep in range(2):
            h.upload(y0, parms)
            y.fill(np.nan)
            h.run_download_into(y, None, stats, code, t_last, step)
            np.testing.assert_array_equal(stats, one.statistics)
            np.testing.assert_array_equal(code, one.code)
            assert_bits_equal(y, one.y, "y")
            assert_bits_equal(t_last, one.t_last, "t_last")
            assert_bits_equal(step, one.step_size, "step_size")
        assert h.last_ms() > 0
        if devices == [0, 0, 0]:
            assert h.n_shards == 3
        h.close()
    with pytest.raises(dde_b200.DdeError, match="not visible"):
        dde_b200.DopriShards(model, 1, 10, devices=[99], n_parms=3, **kw)
    # events (src/dopri.c:476-491) on a sharded batch
    rng = np.random.default_rng(5)
    y0 = rng.uniform(0.5, 2.0, size=(7, 3))
    r = rng.uniform(0.1, 1.0, size=(7, 3))
    times = np.linspace(0, 3, 13)
    for method in ("dopri5", "dopri853"):
        want = pyoracle.dopri_batch(backend, "exponential", y0, times, r, tcrit=[0.5, 1.0, 2.0],
                                    events=[1, -1, 3], method=method)
        for devices in ("all", [0, 0]):
            got = dde_b200.dopri_batch(y0, times, "exponential", r, tcrit=[0.5, 1.0, 2.0],
                                       events=[1, -1, 3], method=method, devices=devices)
            np.testing.assert_array_equal(got.statistics, want.statistics)
            assert_bits_equal(got.y, want.y, "y")
        h = dde_b200.DopriShards("exponential", 3, 7, devices=[0, 0], n_parms=3, method=method)
        h.upload(y0, r)
        h.set_times(times, [0.5, 1.0, 2.0])
        h.set_events([1, -1, 3])
        y = np.empty((7, 13, 3))
        stats = np.empty((7, 4), dtype=np.int32)
        h.run_download_into(y, None, stats, None, None, None)
        np.testing.assert_array_equal(stats, want.statistics)
        assert_bits_equal(y, want.y, "y")
        h.close()


# ---- the staged lag window under stress -------------------------------------------


@pytest.mark.parametrize("n_history", [3, 5, 8, 16, 33])
def test_mackey_glass_small_rings(backend, n_history):
    # the ring recycles entries the window still holds; look-ups that reach
    # behind the tail must fail exactly where the reference's do
    model, y0, parms, times, kw = workloads.mackey_glass(768, first=2048)
    kw["n_history"] = n_history
    got, _ = check_dopri(backend, model, y0, times, parms, **kw)
    if n_history <= 8:
        assert (got.code == -6).sum() > 0


@pytest.mark.parametrize("method,rtol", [("dopri5", 1e-6), ("dopri853", 1e-9), ("dopri853", 1e-3),
                                         ("dopri5", 1e-4)])
def test_mackey_glass_methods_and_tolerances(backend, method, rtol):
    model, y0, parms, times, kw = workloads.mackey_glass(512, first=9000)
    kw.update(method=method, rtol=rtol, atol=rtol, n_history=200)
    check_dopri(backend, model, y0, times, parms, **kw)


def test_mackey_glass_dense_output_and_tcrit(backend):
    # many output rows per step and critical times at the lag's discontinuities
    model, y0, parms, _, kw = workloads.mackey_glass(256, first=123)
    times = np.linspace(0, 120, 481)
    check_dopri(backend, model, y0, times, parms, tcrit=[17.0, 34.0, 51.0, 68.0], **kw)


def test_backward_delay_queries_ahead_of_history(backend):
    # integrating backwards, t - 1 lies AHEAD of the solver: every query either
    # extrapolates the newest entry or, on the first step, fails
    rng = np.random.default_rng(11)
    y0 = rng.uniform(0.5, 2.0, size=(5, 2))
    r = rng.uniform(0.1, 1.0, size=(5, 2))
    for method in ("dopri5", "dopri853"):
        check_dopri(backend, "delayed_growth", y0, -np.linspace(0, 3, 13), r, method=method,
                    n_history=50)


def test_seir_many_parameter_free_replicates_and_long_lag(backend):
    # n = 4, two components lagged: window holds times only (coefficients from
    # the ring), long horizon so the window slides through hundreds of entries
    y0 = np.tile([1e7 - 1, 0.0, 1.0, 0.0], (6, 1))
    y0[:, 2] = np.arange(1, 7)
    y0[:, 0] = 1e7 - y0[:, 2]
    for method in ("dopri5", "dopri853"):
        check_dopri(backend, "seir", y0, np.linspace(0, 400, 41), None, n_out=1, method=method,
                    n_history=40)


# ---- restart: dopri_continue / dopri_data_copy -------------------------------------


def check_continue(backend, model, y, times1, times2, parms, y_new=None, n_out=0, tcrit2=None,
                   **kw):
    y = np.atleast_2d(np.asarray(y, dtype=np.float64))
    pa = None if parms is None else np.ascontiguousarray(np.atleast_2d(parms), dtype=np.float64)
    b = dde_b200.DopriBatch(model, y.shape[1], y.shape[0],
                            n_parms=0 if pa is None else pa.shape[1], n_out=n_out, **kw)
    shared = pa is not None and pa.shape[0] == 1 and y.shape[0] > 1
    b.upload(np.ascontiguousarray(y), pa, shared_parms=shared)
    b.set_times(times1)
    b.run()
    b.continue_(times2, y=y_new, tcrit=tcrit2)
    got = b.download()
    want = pyoracle.dopri_continue(backend, model, y, times1, times2,
                                   None if pa is None else (pa[0] if shared else pa),
                                   y_new=y_new, n_out=n_out, tcrit2=tcrit2, **kw)
    np.testing.assert_array_equal(got.code, want.code)
    np.testing.assert_array_equal(got.statistics, want.statistics)
    assert_bits_equal(got.t_last, want.t_last, "t_last")
    assert_bits_equal(got.step_size, want.step_size, "step_size")
    assert_bits_equal(got.y, want.y, "y")
    if n_out:
        assert_bits_equal(got.output, want.output, "output")
    return b, got


def test_continue_ode(backend):
    lor_p = [[10.0, 28.0, 8.0 / 3.0]]
    y0 = [[10.0, 1.0, 1.0], [1.0, 1.0, 1.0], [-3.0, 2.0, 20.0]]
    for method in ("dopri5", "dopri853"):
        b, got = check_continue(backend, "lorenz", y0, np.linspace(0, 2, 11), np.linspace(2, 5, 16),
                                lor_p, n_out=2, method=method)
        b.close()
        # a new state at the restart (r_dopri_continue's y argument)
        b, _ = check_continue(backend, "lorenz", y0, np.linspace(0, 2, 11), np.linspace(2, 3, 6),
                              lor_p, y_new=np.array(y0) * 0.5, method=method, tcrit2=[2.5])
        b.close()
    # restart == uninterrupted run to solver tolerance (tests/testthat/test-ode.R restart
    # tests; the split only moves one step boundary)
    b, got = check_continue(backend, "exponential", [[1.0, 2.0]], np.linspace(0, 1, 5),
                            np.linspace(1, 2, 5), [[0.5, 0.25]], rtol=1e-9, atol=1e-9)
    b.close()
    full = dde_b200.dopri_batch([[1.0, 2.0]], np.linspace(0, 2, 9), "exponential", [[0.5, 0.25]],
                                rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(got.y[0, -1], full.y[0, -1], rtol=1e-8)


def test_continue_dde_keeps_history(backend):
    # the ring, its oldest surviving time (t0) and the step size carry over
    # (tests/testthat/test-dde.R:301-350); failed trajectories cannot continue
    model, y0, parms, _, kw = workloads.mackey_glass(512, first=4096)
    b, got = check_continue(backend, model, y0, np.linspace(0, 200, 5), np.linspace(200, 500, 7),
                            parms, **kw)
    assert (got.code == -10).sum() > 0 and (got.code == 1).sum() > 400
    assert "Incorrect initial time" in got.message(int(np.argmax(got.code == -10)))
    # a second continuation from the same handle
    b.continue_(np.linspace(500, 600, 3))
    third = b.download()
    assert (third.code[got.code != 1] == -10).all()
    b.close()
    # the wrong start time is every trajectory's error, the wrong direction the batch's
    b = dde_b200.DopriBatch(model, 1, 4, n_parms=3, **kw)
    b.upload(y0[:4], parms[:4])
    b.set_times(np.linspace(0, 50, 3))
    b.run()
    b.continue_([60.0, 70.0])
    assert (b.download().code == -10).all()
    with pytest.raises(dde_b200.DdeError, match="Incorrect sign"):
        b.continue_([50.0, 40.0])
    b.close()
    for m in ("seir", "seir_int"):
        bb, _ = check_continue(backend, m, [[1e7 - 1, 0.0, 1.0, 0.0]], np.linspace(0, 20, 21),
                               np.linspace(20, 60, 41), None, n_out=1, method="dopri853",
                               rtol=1e-7, atol=1e-7, n_history=1000)
        bb.close()


def test_continue_large_model(backend):
    model, y0, parms, _, kw = workloads.seir_age(3)
    b, _ = check_continue(backend, model, y0, np.linspace(0, 60, 4), np.linspace(60, 120, 4),
                          parms, **kw)
    b.close()


def test_copy_is_independent(backend):
    model, y0, parms, _, kw = workloads.mackey_glass(64, first=77)
    a = dde_b200.DopriBatch(model, 1, 64, n_parms=3, **kw)
    a.upload(y0, parms)
    a.set_times(np.linspace(0, 100, 3))
    a.run()
    c = a.copy()
    first = c.download()  # the copy carries the results of the run it was taken after
    assert_bits_equal(first.y, a.download().y, "copied results")
    c.continue_(np.linspace(100, 300, 5))
    rc = c.download()
    a.continue_(np.linspace(100, 300, 5))  # the original was not disturbed by the copy's run
    ra = a.download()
    np.testing.assert_array_equal(ra.statistics, rc.statistics)
    assert_bits_equal(ra.y, rc.y, "y")
    want = pyoracle.dopri_continue(backend, model, y0, np.linspace(0, 100, 3),
                                   np.linspace(100, 300, 5), parms, **kw)
    assert_bits_equal(ra.y, want.y, "y vs checker")
    a.close()
    c.close()


def check_difeq_continue(backend, model, y, steps1, steps2, parms, y_new=None, n_out=0,
                         n_history=0, return_initial=True):
    y = np.atleast_2d(np.asarray(y, dtype=np.float64))
    pa = None if parms is None else np.ascontiguousarray(np.atleast_2d(parms), dtype=np.float64)
    b = dde_b200.DifeqBatch(model, y.shape[1], y.shape[0], n_parms=0 if pa is None else pa.shape[1],
                            n_out=n_out, n_history=n_history, restartable=True)
    b.upload(np.ascontiguousarray(y), pa)
    b.set_steps(steps1, return_initial)
    b.run()
    b.continue_(steps2, y=y_new, return_initial=return_initial)
    y_out, out, code, y_final = b.download()
    want = pyoracle.difeq_continue(backend, model, y, steps1, steps2, pa, y_new=y_new, n_out=n_out,
                                   n_history=n_history, return_initial=return_initial)
    np.testing.assert_array_equal(code, want.code)
    assert_bits_equal(y_out, want.y, "y")
    ok = want.code == 1
    assert_bits_equal(y_final[ok], want.y_final[ok], "y_final")
    if n_out:
        assert_bits_equal(out, want.output, "output")
    return b, y_out, code


def test_difeq_continue(backend):
    # tests/testthat/test-difeq-delay.R:182-222: a restarted run continues the history
    model, y0, parms, _, kw = workloads.sir_delay(300, n_steps=10)
    for hist in (16, 1000, 0):
        if hist == 0:  # the delayed model needs history; use the plain map instead
            b, _, _ = check_difeq_continue(backend, "logistic", [[0.1, 0.2]], [0, 3, 7], [7, 9, 20],
                                           [[1.5, 2.5]])
            b.close()
            continue
        b, y2, code = check_difeq_continue(backend, model, y0, [0, 100, 150], [150, 200, 300], parms,
                                           n_history=hist, return_initial=False)
        assert (code == 1).all()
        # restart == uninterrupted run, exactly: no step boundary moves in a map
        full = dde_b200.difeq_replicate(300, y0, [0, 300], model, parms, n_history=hist,
                                        return_step=False)
        assert_bits_equal(np.asarray(full)[-1].T, y2[:, -1], "restart vs uninterrupted")
        # and once more from the same handle, with a new state
        b.continue_([300, 310], y=y2[:, -1] * 0.5, return_initial=False)
        assert (b.download()[2] == 1).all()
        with pytest.raises(dde_b200.DdeError, match="Incorrect initial step"):
            b.continue_([5, 10])
        c = b.copy()
        c.continue_([310, 400])
        b.continue_([310, 400])
        assert_bits_equal(c.download()[0], b.download()[0], "copy")
        b.close()
        c.close()
    # the reference's tiny-ring quirk: step0 follows the oldest surviving entry,
    # so the first look-back after a restart is the restart state itself
    b, y2, _ = check_difeq_continue(backend, "fib_out", [[1.0]], np.arange(6), np.arange(5, 11),
                                    None, n_out=1, n_history=2)
    assert y2[0, :, 0].tolist() == [13, 26, 39, 65, 104, 169]
    b.close()
    # a run that was not restartable dropped its on-chip history
    nb = dde_b200.DifeqBatch("fib_out", 1, 1, n_out=1, n_history=2)
    nb.upload(np.ones((1, 1)), None)
    nb.set_steps(np.arange(6))
    nb.run()
    with pytest.raises(dde_b200.DdeError, match="not restartable"):
        nb.continue_(np.arange(5, 11))
    nb.close()
    # replicates that failed cannot continue
    model, y0, parms, _, _ = workloads.sir_delay(4, n_steps=10)
    b, _, code = check_difeq_continue(backend, model, y0, [0, 100], [100, 120], parms, n_history=8)
    assert (code == -10).all()
    b.close()


# ---- events at critical times ---------------------------------------------------------


@pytest.mark.parametrize("method", ["dopri5", "dopri853"])
def test_events(backend, method):
    # tests/testthat/test-events.R with the reference's growth.c fixtures: events
    # that change the state (double / halve) or the parameters, mixed with plain
    # critical times; event order = identity, double_variables, halve_variables,
    # double_parameters, halve_parameters
    rng = np.random.default_rng(21)
    for n in (1, 3):
        y0 = rng.uniform(0.5, 2.0, size=(6, n))
        r = rng.uniform(0.1, 1.0, size=(6, n))
        times = np.linspace(0, 3, 13)
        for model in ("exponential", "linear"):
            for tcrit, ev in (([0.5, 1.0, 2.0], [1, -1, 3]), ([0.25, 1.1, 1.1, 2.9], [2, 4, 0, 1]),
                              ([1.0], [3]), ([-1.0, 0.0, 1.5], [1, 1, 2])):
                got = dde_b200.dopri_batch(y0, times, model, r, tcrit=tcrit, events=ev,
                                           method=method)
                want = pyoracle.dopri_batch(backend, model, y0, times, r, tcrit=tcrit, events=ev,
                                            method=method)
                np.testing.assert_array_equal(got.statistics, want.statistics)
                np.testing.assert_array_equal(got.code, want.code)
                assert_bits_equal(got.y, want.y, "y")
    # analytic: doubling the state at t = 1 doubles the solution afterwards
    # (tests/testthat/test-events.R:3-59)
    got = dde_b200.dopri_batch([[1.0]], [0.0, 1.0, 2.0], "exponential", [[0.5]], tcrit=[1.0],
                               events=[1], method=method, rtol=1e-9, atol=1e-9)
    want = pyoracle.dopri_batch(backend, "exponential", [[1.0]], [0.0, 1.0, 2.0], [[0.5]],
                                tcrit=[1.0], events=[1], method=method, rtol=1e-9, atol=1e-9)
    assert_bits_equal(got.y, want.y, "y")
    np.testing.assert_array_equal(got.code, want.code)
    if method == "dopri5":
        # (the reference does not re-evaluate k1 after an event: dop853 fails on
        # this input there too, and so does the device)
        np.testing.assert_allclose(got.y[0, :, 0], [1.0, np.exp(-0.5), 2 * np.exp(-1.0)],
                                   rtol=1e-6)
    # no such event / events without critical times
    with pytest.raises(dde_b200.DdeError, match="event functions"):
        dde_b200.dopri_batch([[1.0]], [0.0, 1.0], "exponential", [[0.5]], tcrit=[0.5], events=[7])
    with pytest.raises(dde_b200.DdeError, match="event functions"):
        dde_b200.dopri_batch([[10.0, 1.0, 1.0]], [0.0, 1.0], "lorenz", [10.0, 28.0, 2.0],
                             tcrit=[0.5], events=[0])
    with pytest.raises(dde_b200.DdeError, match="one entry per critical time"):
        dde_b200.dopri_batch([[1.0]], [0.0, 1.0], "exponential", [[0.5]], tcrit=[0.5],
                             events=[0, 1])


def test_dopri_continue_r_style(backend):
    # dopri(..., restartable = TRUE) then dopri_continue(): R/dopri.R:469-511
    p = [10.0, 28.0, 8.0 / 3.0]
    first = dde_b200.dopri([10.0, 1.0, 1.0], np.linspace(0, 2, 11), "lorenz", p, restartable=True)
    assert first.ptr is not None
    second = dde_b200.dopri_continue(first, np.linspace(2, 4, 11), copy=True,
                                     return_statistics=True)
    again = dde_b200.dopri_continue(first, np.linspace(2, 4, 11), copy=True)
    assert_bits_equal(np.asarray(second), np.asarray(again), "copy leaves the original alone")
    want = pyoracle.dopri_continue(backend, "lorenz", [10.0, 1.0, 1.0], np.linspace(0, 2, 11),
                                   np.linspace(2, 4, 11), p)
    assert_bits_equal(np.asarray(second)[:, 1:], want.y[0], "y")
    assert second.statistics["n_step"] == want.statistics[0, 1]
    with pytest.raises(dde_b200.DdeError, match="Incorrect initial time"):
        dde_b200.dopri_continue(first, np.linspace(3, 4, 3))
    with pytest.raises(dde_b200.DdeError, match="restartable"):
        dde_b200.dopri_continue(second, np.linspace(4, 5, 3))
    third = dde_b200.dopri_continue(first, np.linspace(2, 4, 11))  # advances `first` itself
    assert_bits_equal(np.asarray(third), np.asarray(second), "in place")
    fourth = dde_b200.dopri_continue(first, np.linspace(4, 5, 6), restartable=True)
    assert fourth.ptr is not None and np.asarray(fourth)[0, 0] == 4.0


@pytest.mark.parametrize("n_traj", [100, 32768, 100000])
def test_streamed_download_equals_run_then_download(backend, n_traj):
    """dde_dopri_batch_run_download copies results block by block while the kernel runs:
    same bits as run + download (several blocks, a ragged last one, failed trajectories);
    a second streamed run on the same handle starts from clean flags."""
    model, y0, parms, times, kw = workloads.mackey_glass(n_traj, first=1000)
    b = dde_b200.DopriBatch(model, 1, n_traj, n_parms=3, **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    want = b.download()
    for _ in range(2):
        y = np.full((n_traj, b.nt, 1), -1.0)
        stats = np.full((n_traj, 4), -1, dtype=np.int32)
        code = np.full((n_traj,), -99, dtype=np.int32)
        t_last = np.full((n_traj,), -1.0)
        step_size = np.full((n_traj,), -1.0)
        b.run_download_into(y, None, stats, code, t_last, step_size)
        np.testing.assert_array_equal(code, want.code)
        np.testing.assert_array_equal(stats, want.statistics)
        assert_bits_equal(y, want.y, "y")
        assert_bits_equal(t_last, want.t_last, "t_last")
        assert_bits_equal(step_size, want.step_size, "step_size")
    b.close()
    if n_traj == 100:  # against the reference itself, and the one-shot call uses the same path
        ref = pyoracle.dopri_batch(backend, model, y0, times, parms, **kw)
        np.testing.assert_array_equal(code, ref.code)
        assert_bits_equal(y, ref.y, "y vs reference")
        got = dde_b200.dopri_batch(y0, times, model, parms, **kw)
        assert_bits_equal(got.y, ref.y, "one-shot y vs reference")


@pytest.mark.parametrize("model_fn,n_times", [("lorenz_sweep", 41), ("rossler_sweep", 1001),
                                              ("lorenz_sweep", 6001)])
def test_staged_rows_of_the_unrolled_kernel(backend, model_fn, n_times):
    """Small models without history run the unrolled kernel, whose output rows are held
    back in shared memory until they fill an aligned sector of the result array
    (dopri_kernels.cuh, MODE 1): same bits as the reference for sparse grids, dense grids,
    and grids so dense that one step crosses many rows; the streamed download (several
    completion blocks, a ragged last one) sees every row before its block is flagged."""
    n_traj = 70001 if n_times == 41 else 300
    model, y0, parms, times, kw = getattr(workloads, model_fn)(n_traj, first=17, n_times=n_times)
    got = dde_b200.dopri_batch(y0, times, model, parms, **kw)  # the streamed path
    b = dde_b200.DopriBatch(model, 3, n_traj, n_parms=3, **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    held = b.download()
    b.close()
    assert_bits_equal(got.y, held.y, "streamed vs run + download")
    np.testing.assert_array_equal(got.statistics, held.statistics)
    k = 64  # the reference on both ends of the batch
    for sl in (slice(0, k), slice(n_traj - k, n_traj)):
        want = pyoracle.dopri_batch(backend, model, y0[sl], times, parms[sl], **kw)
        np.testing.assert_array_equal(got.code[sl], want.code)
        np.testing.assert_array_equal(got.statistics[sl], want.statistics)
        assert_bits_equal(got.y[sl], want.y, "y vs reference")


@pytest.mark.parametrize("n_times,return_initial", [(12, False), (12, True), (14, True),
                                                     (1002, True), (2, False)])
def test_rows_by_sector_blocks_that_straddle_sectors(backend, n_times, return_initial):
    """The unrolled kernel writes rows as aligned 32-byte sectors of the result array; a
    trajectory's block of rows that starts or ends inside a sector (n * n_rows not a multiple
    of four, or the initial row written by the start-up kernel) shares that sector with its
    neighbour's block, written by another thread: every double, once, by its owner."""
    n_traj = 5000
    model, y0, parms, times, kw = workloads.lorenz_sweep(n_traj, first=3, n_times=n_times,
                                                         t_end=3.0)
    kw = dict(kw, return_initial=return_initial)
    got = dde_b200.dopri_batch(y0, times, model, parms, **kw)
    assert not np.isnan(got.y[got.code == 1]).any()
    k = 96
    for sl in (slice(0, k), slice(n_traj - k, n_traj)):
        want = pyoracle.dopri_batch(backend, model, y0[sl], times, parms[sl], **kw)
        np.testing.assert_array_equal(got.code[sl], want.code)
        np.testing.assert_array_equal(got.statistics[sl], want.statistics)
        assert_bits_equal(got.y[sl], want.y, "y vs reference")
    # the whole batch against a run without its first trajectory: the same blocks of rows
    # at other offsets in the result array
    half = dde_b200.dopri_batch(y0[1:], times, model, parms[1:], **kw)
    assert_bits_equal(got.y[1:], half.y, "the same trajectories at shifted offsets")


@pytest.mark.parametrize("workload", ["mackey_glass", "lorenz_sweep"])
def test_hand_out_order_does_not_change_results(backend, workload):
    """dde_dopri_batch_set_order / _order_by_last_run: the order the device's threads take
    trajectories in changes when a batch ends, never what it returns (both stepping kernels
    of the thread-per-trajectory models; the streamed download under an arbitrary order)."""
    n_traj = 40000
    if workload == "mackey_glass":
        model, y0, parms, times, kw = workloads.mackey_glass(n_traj, first=11)
    else:
        model, y0, parms, times, kw = workloads.lorenz_sweep(n_traj, first=11, n_times=23, t_end=3.0)
    b = dde_b200.DopriBatch(model, y0.shape[1], n_traj, n_parms=parms.shape[1], **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    want = b.download()
    rng = np.random.default_rng(5)

    def check(tag):
        b.run()
        got = b.download()
        np.testing.assert_array_equal(got.code, want.code)
        np.testing.assert_array_equal(got.statistics, want.statistics)
        assert_bits_equal(got.y, want.y, tag)
        y = np.empty_like(want.y)
        code = np.empty((n_traj,), dtype=np.int32)
        b.run_download_into(y, None, None, code, None, None)  # completion flags, any order
        np.testing.assert_array_equal(code, want.code)
        assert_bits_equal(y, want.y, tag + " (streamed)")

    b.set_order(rng.permutation(n_traj))
    check("a random order")
    b.order_by_last_run()
    check("longest first")
    b.set_order(None)
    check("index order again")
    with pytest.raises(dde_b200.DdeError, match="permutation"):
        b.set_order(np.zeros(n_traj, dtype=np.uint32))
    b.close()
    ref = pyoracle.dopri_batch(backend, model, y0[:64], times, parms[:64], **kw)
    assert_bits_equal(want.y[:64], ref.y, "y vs reference")


def test_shards_hand_out_order(backend):
    """dde_dopri_shards_order_by_last_run: every shard takes its order from its own last
    run; the results do not change"""
    n_traj = 30000
    model, y0, parms, times, kw = workloads.mackey_glass(n_traj, first=5)
    sh = dde_b200.DopriShards(model, 1, n_traj, devices=None, n_parms=3, **kw)
    sh.set_times(times)
    nt = times.shape[0]
    outs = []
    for mode in (None, True, False):
        if mode is not None:
            sh.order_by_last_run(mode)
        y = np.empty((n_traj, nt, 1))
        stats = np.empty((n_traj, 4), dtype=np.int32)
        code = np.empty((n_traj,), dtype=np.int32)
        sh.upload(y0, parms)
        sh.run_download_into(y, None, stats, code, None, None)
        outs.append((y, stats, code))
    sh.close()
    for y, stats, code in outs[1:]:
        np.testing.assert_array_equal(code, outs[0][2])
        np.testing.assert_array_equal(stats, outs[0][1])
        assert_bits_equal(y, outs[0][0], "y")
    want = pyoracle.dopri_batch(backend, model, y0[:64], times, parms[:64], **kw)
    assert_bits_equal(outs[0][0][:64], want.y, "y vs reference")


@pytest.mark.parametrize("n", [1, 2, 4])
@pytest.mark.parametrize("n_times,return_initial", [(8, False), (8, True), (11, True), (402, False)])
def test_rows_by_sector_other_widths(backend, n, n_times, return_initial):
    """rows of one, two and four doubles through the unrolled kernel's sector buffer (four
    rows, two rows, one row per 32-byte sector), blocks that straddle sectors included: the
    exponential fixture registered with n fixed at compile time, against the reference's
    run of the same right-hand side"""
    n_traj = 3001
    rng = np.random.default_rng(10 * n + n_times)
    y0 = rng.uniform(0.5, 2.0, size=(n_traj, n))
    r = rng.uniform(0.1, 3.0, size=(n_traj, n))
    times = np.linspace(0.0, 4.0, n_times)
    kw = dict(method="dopri5", rtol=1e-7, atol=1e-7, return_initial=return_initial)
    got = dde_b200.dopri_batch(y0, times, "exponential%d" % n, r, **kw)
    assert (got.code == 1).all()
    for sl in (slice(0, 80), slice(n_traj - 80, n_traj)):
        want = pyoracle.dopri_batch(backend, "exponential", y0[sl], times, r[sl], **kw)
        np.testing.assert_array_equal(got.statistics[sl], want.statistics)
        assert_bits_equal(got.y[sl], want.y, "y vs reference")
    again = dde_b200.dopri_batch(y0[1:], times, "exponential%d" % n, r[1:], **kw)
    assert_bits_equal(got.y[1:], again.y, "the same trajectories at shifted offsets")


def test_streamed_download_large_model(backend):
    """the block-per-trajectory kernel raises the same completion flags"""
    model, y0, parms, times, kw = workloads.seir_age(8)
    b = dde_b200.DopriBatch(model, y0.shape[1], 8, n_parms=parms.shape[1], **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    want = b.download()
    y = np.empty_like(want.y)
    code = np.empty((8,), dtype=np.int32)
    b.run_download_into(y, None, None, code, None, None)
    np.testing.assert_array_equal(code, want.code)
    assert_bits_equal(y, want.y, "y")
    b.close()


def test_grow_history_dopri(backend):
    """grow_history (src/dopri.c:92-95): a ring that grows instead of overwriting.  On
    the device the run is repeated with larger rings until none has wrapped; results,
    statistics and the full history are those of the reference's growing ring."""
    model, y0, parms, times, kw = workloads.mackey_glass(48, first=300)
    kw = dict(kw, n_history=10)
    want = pyoracle.dopri_batch(backend, model, y0, times, parms, grow_history=True, **kw)
    got = dde_b200.dopri_batch(y0, times, model, parms, grow_history=True, **kw)
    np.testing.assert_array_equal(got.code, want.code)
    np.testing.assert_array_equal(got.statistics, want.statistics)
    assert_bits_equal(got.y, want.y, "y")
    # without growing, ten entries lose the lagged history: the results differ
    lost = dde_b200.dopri_batch(y0, times, model, parms, **kw)
    assert (lost.code != 1).any()
    b = dde_b200.DopriBatch(model, 1, 48, n_parms=3, grow_history=True, **kw)
    b.upload(y0, parms)
    b.set_times(times)
    b.run()
    res = b.download()
    np.testing.assert_array_equal(res.statistics, want.statistics)
    for j in (0, 17, 47):
        ref = pyoracle.dopri_history(backend, model, y0[j:j + 1], times, parms[j:j + 1],
                                     grow_history=True, max_entries=2000, **kw)
        h = b.history(j)
        assert h.shape[0] == ref.n_entries == want.statistics[j, 2]
        assert_bits_equal(h, ref.history, "history %d" % j)
    assert b.n_history >= want.statistics[:, 2].max()
    b.close()
    # restart with a growing ring (dopri_continue on a grow_history object): the rings are
    # outgrown during the continuation, which is then repeated from a copy of its start
    t1, t2 = np.linspace(0, 100, 3), np.linspace(100, 500, 5)
    b = dde_b200.DopriBatch(model, 1, 48, n_parms=3, grow_history=True, **kw)
    b.upload(y0, parms)
    b.set_times(t1)
    b.run()
    cap1 = int(b._lib.dde_dopri_batch_n_history(b._h))
    b.continue_(t2)
    got2 = b.download()
    assert int(b._lib.dde_dopri_batch_n_history(b._h)) > cap1
    want2 = pyoracle.dopri_continue(backend, model, y0, t1, t2, parms, grow_history=True, **kw)
    np.testing.assert_array_equal(got2.code, want2.code)
    np.testing.assert_array_equal(got2.statistics, want2.statistics)
    assert_bits_equal(got2.y, want2.y, "continued y")
    # ... and equals the continuation of a run whose ring was large from the start
    big = dde_b200.DopriBatch(model, 1, 48, n_parms=3, **dict(kw, n_history=2000))
    big.upload(y0, parms)
    big.set_times(t1)
    big.run()
    big.continue_(t2)
    assert_bits_equal(got2.y, big.download().y, "grown vs large")
    for j in (3, 40):
        assert_bits_equal(b.history(j), big.history(j), "history %d" % j)
    big.close()
    b.close()
    # the single-trajectory interface, history attribute included
    one = dde_b200.dopri(y0[5], times, model, parms[5], grow_history=True, return_history=True,
                         return_statistics=True, **kw)
    ref = pyoracle.dopri_history(backend, model, y0[5:6], times, parms[5:6], grow_history=True,
                                 max_entries=2000, **kw)
    assert_bits_equal(np.asarray(one.history), ref.history, "dopri(...).history")


@pytest.mark.parametrize("grow", [False, True])
def test_difeq_history(backend, grow):
    """the `history` attribute of a difference equation (src/r_difeq.c:178-186) and
    grow_history for difference equations"""
    steps = np.arange(0, 41)
    got = dde_b200.difeq([0.1], steps, "logistic", [1.5], n_history=5, grow_history=grow)
    ref = pyoracle.difeq_history(backend, "logistic", [[0.1]], steps, [1.5], n_history=5,
                                 grow_history=grow, max_entries=64)
    assert got.history.shape[0] == ref.n_entries == (41 if grow else 5)
    assert_bits_equal(got.history, ref.history, "history")
    # a delay model with gaps between the stored steps, many replicates in one batch
    model, y0, parms, st, kw = workloads.sir_delay(8, n_steps=300)
    b = dde_b200.DifeqBatch(model, 3, 8, n_parms=2, restartable=True,
                            n_history=301 if grow else 16)
    b.upload(y0, parms)
    b.set_steps(st, return_initial=False)
    b.run()
    y, _, code, _ = b.download()
    assert (code == 1).all()
    for j in (0, 7):
        ref = pyoracle.difeq_history(backend, model, y0[j:j + 1], st, parms[j:j + 1], n_history=16,
                                     grow_history=grow, return_initial=False, max_entries=512)
        assert_bits_equal(y[j], ref.y, "y")
        assert_bits_equal(b.history(j), ref.history, "history %d" % j)
    b.close()
    # with outputs: the out columns are what the reference's lagging out_next pointer leaves
    # in them (src/difeq.c:203-213,249-255) -- every step asked for, a few, with and without
    # the initial row, rings that hold the run and rings that wrap.  (The plain-C
    # restatement keeps NA there, oracle/dde_oracle.c: against it, steps and states only.)
    def same_history(got, ref, n, what):
        if backend == "ref":
            assert_bits_equal(got, ref, what)
        else:
            assert_bits_equal(got[:, :1 + n], ref[:, :1 + n], what)

    for steps in (np.arange(0, 13), np.array([0, 3, 4, 9, 12]), np.array([2, 3, 11, 30]),
                  np.array([0, 40])):
        for return_initial in (True, False):
            for n_history in (64, 4, 2):
                b = dde_b200.DifeqBatch("fib_out", 1, 3, n_out=1, n_history=n_history,
                                        restartable=True)
                b.upload(np.array([[1.0], [2.0], [0.5]]), None)
                b.set_steps(steps, return_initial=return_initial)
                b.run()
                for j, y0 in enumerate((1.0, 2.0, 0.5)):
                    ref = pyoracle.difeq_history(backend, "fib_out", [[y0]], steps, None, n_out=1,
                                                 n_history=n_history, max_entries=64,
                                                 return_initial=return_initial)
                    same_history(b.history(j), ref.history, 1,
                                 "history %s %s %d" % (steps.tolist(), return_initial, n_history))
                b.close()
    # the additive model: two outputs per step, several replicates
    rng = np.random.default_rng(3)
    y0 = rng.integers(1, 9, size=(4, 2)).astype(float)
    p = rng.integers(1, 5, size=(4, 2)).astype(float)
    steps = np.array([0, 1, 5, 6, 20])
    b = dde_b200.DifeqBatch("additive", 2, 4, n_parms=2, n_out=3, n_history=8, restartable=True)
    b.upload(y0, p)
    b.set_steps(steps, return_initial=True)
    b.run()
    for j in range(4):
        ref = pyoracle.difeq_history(backend, "additive", y0[j:j + 1], steps, p[j:j + 1], n_out=3,
                                     n_history=8, max_entries=64)
        same_history(b.history(j), ref.history, 2, "additive history %d" % j)
    b.close()
