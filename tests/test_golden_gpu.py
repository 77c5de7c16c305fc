"""GPU parity against the committed reference fixtures (tests/golden/, made by
tools/gen_golden.py from the unmodified reference): the CUDA path through the
C-ABI must reproduce results, statistics and return codes bit for bit.  Needs
no checker library on the GPU box."""
import numpy as np
import pytest

import dde_b200
from conftest import assert_bits_equal
from golden_util import cases, load

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", cases("dopri"))
def test_dopri_golden(name):
    g = load(name)
    got = dde_b200.dopri_batch(g["y0"], g["times"], g["model"], g["parms"], n_out=g["n_out"],
                               tcrit=g["tcrit"], **g["kw"])
    np.testing.assert_array_equal(got.code, g["code"])
    np.testing.assert_array_equal(got.statistics, g["statistics"])
    assert_bits_equal(got.t_last, g["t_last"], "t_last")
    assert_bits_equal(got.step_size, g["step_size"], "step_size")
    assert_bits_equal(got.y, g["y"], "y")
    if g["n_out"]:
        assert_bits_equal(got.output, g["output"], "output")


@pytest.mark.parametrize("name", cases("difeq"))
def test_difeq_golden(name):
    g = load(name)
    y0 = g["y0"]
    parms = g["parms"]
    b = dde_b200.DifeqBatch(g["model"], y0.shape[1], y0.shape[0],
                            n_parms=0 if parms is None else np.atleast_2d(parms).shape[1],
                            n_out=g["n_out"], n_history=int(g["n_history"]))
    pa = None if parms is None else np.ascontiguousarray(np.atleast_2d(parms))
    b.upload(np.ascontiguousarray(y0), pa)
    b.set_steps(g["steps"], bool(g["return_initial"]))
    b.run()
    y, out, code, y_final = b.download()
    b.close()
    np.testing.assert_array_equal(code, g["code"])
    assert_bits_equal(y, g["y"], "y")
    ok = g["code"] == 1
    assert_bits_equal(y_final[ok], g["y_final"][ok], "y_final")
    if g["n_out"]:
        assert_bits_equal(out, g["output"], "output")
