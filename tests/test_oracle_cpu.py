"""Pin the checkers (CPU only).

  * oracle/liboracle.so, the plain-C restatement, must reproduce every fixture
    in tests/golden/ -- outputs of the UNMODIFIED reference, written by
    tools/gen_golden.py -- bit for bit, and the reference's own known answers
    (SURVEY.md section 8c / appendix A.4).
  * where oracle/_ref (the compiled reference itself) is present, the
    restatement must also agree with it on fresh random inputs, and the
    fixtures must still be what the reference produces.
"""
import numpy as np
import pytest

from conftest import assert_bits_equal
from dde_b200 import workloads
from golden_util import cases, load
from oracle import pyoracle

needs_oracle = pytest.mark.skipif(not pyoracle.available("oracle"),
                                  reason="oracle/liboracle.so not built (make -C oracle oracle)")
needs_ref = pytest.mark.skipif(not pyoracle.available("ref"),
                               reason="oracle/_ref not built (needs /root/reference)")


def run_dopri(backend, g):
    return pyoracle.dopri_batch(backend, g["model"], g["y0"], g["times"], g["parms"],
                                n_out=g["n_out"], tcrit=g["tcrit"], **g["kw"])


def run_difeq(backend, g):
    return pyoracle.difeq_batch(backend, g["model"], g["y0"], g["steps"], g["parms"],
                                n_out=g["n_out"], n_history=int(g["n_history"]),
                                return_initial=bool(g["return_initial"]))


def compare_dopri(got, g):
    np.testing.assert_array_equal(got.code, g["code"])
    np.testing.assert_array_equal(got.statistics, g["statistics"])
    assert_bits_equal(got.t_last, g["t_last"], "t_last")
    assert_bits_equal(got.step_size, g["step_size"], "step_size")
    assert_bits_equal(got.y, g["y"], "y")
    if g["n_out"]:
        assert_bits_equal(got.output, g["output"], "output")


def compare_difeq(got, g):
    np.testing.assert_array_equal(got.code, g["code"])
    assert_bits_equal(got.y, g["y"], "y")
    ok = g["code"] == 1
    assert_bits_equal(got.y_final[ok], g["y_final"][ok], "y_final")
    if g["n_out"]:
        assert_bits_equal(got.output, g["output"], "output")


@needs_oracle
@pytest.mark.parametrize("name", cases("dopri"))
def test_oracle_reproduces_dopri_fixture(name):
    g = load(name)
    compare_dopri(run_dopri("oracle", g), g)


@needs_oracle
@pytest.mark.parametrize("name", cases("difeq"))
def test_oracle_reproduces_difeq_fixture(name):
    g = load(name)
    compare_difeq(run_difeq("oracle", g), g)


@needs_ref
@pytest.mark.parametrize("name", cases("dopri") + cases("difeq"))
def test_fixtures_are_what_the_reference_produces(name):
    g = load(name)
    if name.startswith("dopri"):
        compare_dopri(run_dopri("ref", g), g)
    else:
        compare_difeq(run_difeq("ref", g), g)


def test_fixture_known_answers():
    # SURVEY.md A.4 probes of the unmodified reference
    g = load("dopri_c1_lorenz")
    assert g["statistics"][0].tolist() == [27392, 4565, 4365, 200]
    assert g["y"][0, -1].tolist() == [7.7062450435259899, 12.475805357018485, 17.113916246292334]
    g = load("dopri_mackey_glass_kat")
    assert g["statistics"][0].tolist() == [2475, 173, 114, 59]
    assert g["y"][0, -1, 0] == 1.059911058006217
    g = load("dopri_seir_dopri5")
    assert g["statistics"][0].tolist() == [500, 83, 74, 9]
    # tests/testthat/test-dde.R:266-270 and SURVEY.md A.1#1
    assert load("dopri_seir_underflow")["code"][0] == -6
    assert load("dopri_backward_853")["code"][0] == -4
    # tests/testthat/test-difeq-delay.R:3-24
    assert load("difeq_fibonacci")["y"][0, :, 0].tolist() == [1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144]
    # tests/testthat/test-difeq.R:222-233
    ref = [0.1]
    for _ in range(20):
        ref.append(1.5 * ref[-1] * (1 - ref[-1]))
    assert load("difeq_logistic")["y"][0, :, 0].tolist() == ref


@needs_oracle
def test_oracle_analytic_cases():
    # dy/dt = 1 from 0 over [0, 1] gives 1 to expect_equal's 1.5e-8
    # (tests/testthat/test-ode.R:436-442)
    r = pyoracle.dopri_batch("oracle", "linear", [[0.0]], [0.0, 1.0], [[1.0]])
    assert abs(r.y[0, -1, 0] - 1.0) < 1.5e-8
    # exponential growth/decay, tests/testthat/test-ode.R:511-551 (1e-6 .. 1e-7 there)
    rr = np.array([[0.5, -0.3]])
    r = pyoracle.dopri_batch("oracle", "exponential", [[1.0, 2.0]], np.linspace(0, 2, 9), rr,
                             method="dopri853")
    want = np.array([1.0, 2.0]) * np.exp(-rr * np.linspace(0, 2, 9)[:, None])
    np.testing.assert_allclose(r.y[0], want, rtol=2e-6)
    # n_eval drops by exactly 1 when step_size_initial is given (test-ode.R:398-404)
    lor = ("lorenz", [[10.0, 1.0, 1.0]], np.linspace(0, 3, 31), [10.0, 28.0, 8.0 / 3.0])
    a = pyoracle.dopri_batch("oracle", *lor)
    b = pyoracle.dopri_batch("oracle", *lor, step_size_initial=float(a.step_size[0]))
    assert a.statistics[0, 0] > 0 and b.statistics[0, 0] > 0
    # statistics relations, tests/testthat/test-dde.R:26-47
    n_eval, n_step, n_accept, n_reject = a.statistics[0]
    assert n_step == n_accept + n_reject and n_eval == 6 * n_step + 2


@needs_oracle
@needs_ref
def test_oracle_matches_reference_on_fresh_inputs():
    rng = np.random.default_rng(2024)
    model, y0, parms, times, kw = workloads.mackey_glass(96, first=5000)
    for backend_kw in (kw, dict(kw, method="dopri5", n_history=200)):
        a = pyoracle.dopri_batch("oracle", model, y0, times, parms, **backend_kw)
        b = pyoracle.dopri_batch("ref", model, y0, times, parms, **backend_kw)
        np.testing.assert_array_equal(a.code, b.code)
        np.testing.assert_array_equal(a.statistics, b.statistics)
        assert_bits_equal(a.y, b.y, "y")
        assert_bits_equal(a.t_last, b.t_last, "t_last")
        assert_bits_equal(a.step_size, b.step_size, "step_size")
    for method in ("dopri5", "dopri853"):
        y0 = rng.uniform(-5, 5, (8, 3)) + np.array([5.0, 0.0, 20.0])
        p = np.array([10.0, 28.0, 8.0 / 3.0]) * rng.uniform(0.8, 1.2, (8, 3))
        opts = dict(method=method, rtol=1e-8, atol=1e-9, stiff_check=2, n_out=2)
        a = pyoracle.dopri_batch("oracle", "lorenz", y0, np.linspace(0, 4, 17), p,
                                 tcrit=[0.5, 3.3], **opts)
        b = pyoracle.dopri_batch("ref", "lorenz", y0, np.linspace(0, 4, 17), p, tcrit=[0.5, 3.3],
                                 **opts)
        np.testing.assert_array_equal(a.statistics, b.statistics)
        assert_bits_equal(a.y, b.y, "y")
        assert_bits_equal(a.output, b.output, "output")
    # history export: one entry per accepted step, oldest first (src/r_dopri.c:401-410)
    for n_history in (7, 1000):
        a = pyoracle.dopri_history("oracle", "lorenz", [10.0, 1.0, 1.0], np.linspace(0, 2, 5),
                                   [10.0, 28.0, 8.0 / 3.0], n_history=n_history)
        b = pyoracle.dopri_history("ref", "lorenz", [10.0, 1.0, 1.0], np.linspace(0, 2, 5),
                                   [10.0, 28.0, 8.0 / 3.0], n_history=n_history)
        assert_bits_equal(a.history, b.history, "history")
    model, y0, parms, steps, kw = workloads.sir_delay(32, first=77, n_steps=700)
    a = pyoracle.difeq_batch("oracle", model, y0, steps, parms, **kw)
    b = pyoracle.difeq_batch("ref", model, y0, steps, parms, **kw)
    assert_bits_equal(a.y, b.y, "y")
    # the lag functions no workload uses: ylag_vec, yprev_all, yprev_vec (size_t indices)
    y0 = rng.uniform(0.5, 2.0, (5, 4))
    r = rng.uniform(0.1, 1.0, (5, 4))
    a = pyoracle.dopri_batch("oracle", "delayed_growth_vec", y0, np.linspace(0, 3, 13), r,
                             n_history=200)
    b = pyoracle.dopri_batch("ref", "delayed_growth_vec", y0, np.linspace(0, 3, 13), r,
                             n_history=200)
    np.testing.assert_array_equal(a.statistics, b.statistics)
    assert_bits_equal(a.y, b.y, "y")
    a = pyoracle.dopri_batch("oracle", "delayed_growth_all", y0, np.linspace(0, 3, 13), r,
                             n_history=200)
    b = pyoracle.dopri_batch("ref", "delayed_growth_all", y0, np.linspace(0, 3, 13), r,
                             n_history=200)
    np.testing.assert_array_equal(a.statistics, b.statistics)
    assert_bits_equal(a.y, b.y, "y")
    for model in ("fib_all", "fib_vec"):
        a = pyoracle.difeq_batch("oracle", model, y0, np.arange(0, 30), None, n_history=3)
        b = pyoracle.difeq_batch("ref", model, y0, np.arange(0, 30), None, n_history=3)
        assert_bits_equal(a.y, b.y, model)
    # grow_history == a large history (tests/testthat/test-ode.R:627-640, test-difeq.R:410-428)
    model, y0, parms, times, kw = workloads.mackey_glass(4, first=40)
    for backend in ("oracle", "ref"):
        g = pyoracle.dopri_history(backend, model, y0[:1], times, parms[:1], grow_history=True,
                                   max_entries=2000, **dict(kw, n_history=10))
        big = pyoracle.dopri_history(backend, model, y0[:1], times, parms[:1], max_entries=2000,
                                     **dict(kw, n_history=2000))
        assert g.n_entries == big.n_entries == g.statistics[2]
        assert_bits_equal(g.history, big.history, "grown history")
        assert_bits_equal(g.y, big.y, "y")
    a = pyoracle.difeq_history("oracle", "fib_vec", [[1.0, 2.0, 3.0]], np.arange(0, 25), None,
                               n_history=4, grow_history=True, max_entries=64)
    b = pyoracle.difeq_history("ref", "fib_vec", [[1.0, 2.0, 3.0]], np.arange(0, 25), None,
                               n_history=4, grow_history=True, max_entries=64)
    assert a.n_entries == b.n_entries == 25
    assert_bits_equal(a.history, b.history, "difeq history")


@needs_ref
def test_reference_shared_ring_artefact():
    # SURVEY.md finding 5: the reference reuses ONE difeq_data for all replicates
    # (src/r_difeq.c:65,95-97), so with a small ring replicate 2 differs from an
    # independent run.  The batched contract (and both checkers' default) is a
    # fresh state per replicate (R/difeq_replicate.R:1-4).
    model, y0, parms, steps, kw = workloads.sir_delay(2, n_steps=200)
    y0 = np.repeat(y0[:1], 2, axis=0)
    parms = np.repeat(parms[:1], 2, axis=0)
    fresh = pyoracle.difeq_batch("ref", model, y0, steps, parms, **kw)
    shared = pyoracle.difeq_batch("ref", model, y0, steps, parms, shared_ring=True, **kw)
    assert_bits_equal(fresh.y[0], fresh.y[1], "fresh replicates")
    assert_bits_equal(shared.y[0], fresh.y[0], "first replicate")
    assert not np.array_equal(shared.y[1], fresh.y[1])
