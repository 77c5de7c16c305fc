"""The C-ABI shared library: it loads, exports every symbol include/dde_b200.h
declares, its host-only entry points behave like the reference's glue, and
without a GPU every solver entry point fails loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import dde_b200
from dde_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "dde_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dde_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), "libdde_b200.so does not export %s" % name
    # and the ctypes table binds exactly the header's list
    assert sorted(n for n, _, _ in _lib.SYMBOLS) == names


PRODUCT_MODELS = {"lorenz", "rossler", "mackey_glass", "sir_delay", "seir_age"}
FIXTURE_MODELS = {"seir", "seir_int", "exponential", "exponential1", "exponential2", "exponential4", "linear", "delayed_growth", "delayed_growth_vec",
                  "delayed_growth_all", "mg_wrong_lag", "logistic", "additive", "fibonacci", "fib_out", "fib_all",
                  "fib_vec", "exponential_wide", "delayed_growth_wide"}


def test_model_plugin_boundary():
    """The product library holds the BASELINE.json workloads only; the reference's fixture
    models come from a model library built out of tree against include/ and loaded at run
    time (R/common.R:89-98, inst/include/dde/dde.c:10-78).  Checked in a fresh process:
    before / after loading the plug-in."""
    import subprocess
    import sys
    plug = ROOT / "tests" / "models" / "libdde_test_models.so"
    if not plug.exists():
        pytest.skip("tests/models/libdde_test_models.so not built (tests/models/build.sh)")
    code = ("import dde_b200, json; a = sorted(dde_b200.models()); "
            "dde_b200.load_model_library(%r); b = sorted(dde_b200.models()); "
            "print(json.dumps([a, b]))" % str(plug))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True,
                         cwd=str(ROOT)).stdout
    import json
    before, after = json.loads(out.strip().splitlines()[-1])
    assert set(before) == PRODUCT_MODELS
    assert set(after) == PRODUCT_MODELS | FIXTURE_MODELS
    # the plug-in is built from the public headers alone
    recipe = (ROOT / "tests" / "models" / "build.sh").read_text()
    assert "-I../../include" in recipe and "csrc" not in recipe


def test_reference_example_compiles_with_only_the_decoration_injected(tmp_path):
    """tools/decorate_model.py changes nothing but the DDE_MODEL_FN prefix of the function
    definitions, and the recipe turns the reference's example into a loadable model library
    (cross-compiled here; run on the GPU by test_reference_example_model_as_plugin)."""
    import subprocess
    import sys
    src = Path("/root/reference/inst/examples/seir.c")
    if not src.exists():
        pytest.skip("no reference tree")
    sys.path.insert(0, str(ROOT / "tools"))
    import decorate_model
    text = src.read_text()
    dec = decorate_model.decorate(text)
    assert dec.replace("DDE_MODEL_FN ", "") == text
    assert dec.count("DDE_MODEL_FN ") == 2  # seir, seir_output
    subprocess.run(["bash", str(ROOT / "tests" / "models" / "build_ref_example.sh")], check=True)
    plug = ROOT / "tests" / "models" / "_ref" / "libdde_ref_example.so"
    code = ("import dde_b200; dde_b200.load_model_library(%r); "
            "assert 'seir_reference_example' in dde_b200.models()" % str(plug))
    subprocess.run([sys.executable, "-c", code], check=True, cwd=str(ROOT))


def test_model_header_is_source_compatible():
    # the eight history functions of inst/include/dde/dde.h:3-13, same signatures
    text = (ROOT / "include" / "dde" / "dde.h").read_text()
    for proto in ("double ylag_1(double t, size_t i);", "void ylag_all(double t, double *y);",
                  "void ylag_vec(double t, const size_t *idx, size_t nidx, double *y);",
                  "void ylag_vec_int(double t, const int *idx, size_t nidx, double *y);",
                  "double yprev_1(int step, size_t i);", "void yprev_all(int step, double *y);",
                  "void yprev_vec(int step, const size_t *idx, size_t nidx, double *y);",
                  "void yprev_vec_int(int step, const int *idx, size_t nidx, double *y);"):
        assert proto in text


def test_control_defaults():
    lib = _lib.load()
    for method, name in ((0, "dopri5"), (1, "dopri853")):
        ctl = _lib.DopriControl()
        lib.dde_dopri_control_default(C.byref(ctl), method)
        # src/dopri.c:64-87, R/dopri.R:293-310
        assert (ctl.atol, ctl.rtol) == (1e-6, 1e-6)
        assert ctl.step_max_n == 100000 and ctl.stiff_check == 0 and ctl.n_history == 0
        assert ctl.method == method and ctl.return_initial == 1 and ctl.grow_history == 0
        assert ctl.step_size_initial == 0.0 and ctl.step_size_min_allow == 0


def test_error_strings_are_the_reference_ones():
    # src/r_dopri.c:264-297
    assert dde_b200.strerror(-1) == "Initialisation failure: Beginning and end times are the same"
    assert dde_b200.strerror(-2) == "Initialisation failure: Times have inconsistent sign"
    assert dde_b200.strerror(-3, 1.5) == "Integration failure: too many steps (at t = 1.50000)"
    assert dde_b200.strerror(-4, 2.0) == "Integration failure: step size too small (at t = 2.00000)"
    assert dde_b200.strerror(-5, 2.0) == "Integration failure: step size vanished (at t = 2.00000)"
    assert dde_b200.strerror(-6, 3.0, 0) == \
        "Integration failure: can't use ylag in model with no history"
    assert dde_b200.strerror(-6, 3.0, 10) == \
        "Integration failure: did not find time in history (at t = 3.00000)"
    assert dde_b200.strerror(-7, 0.25) == "Integration failure: problem became stiff (at t = 0.25000)"
    assert dde_b200.strerror(-8) == "non-finite derivative at initial time"  # src/dopri.c:333
    assert "did not find step" in dde_b200.strerror(-9)  # src/difeq.c:268
    assert dde_b200.strerror(1) == "OK"


def test_model_registry():
    names = dde_b200.models()
    for m in ("lorenz", "rossler", "mackey_glass", "seir", "seir_int", "exponential",
              "delayed_growth", "logistic", "fibonacci", "sir_delay"):
        assert m in names
    lib = _lib.load()
    kind, n, npar, nout = C.c_int32(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert lib.dde_model_info(b"mackey_glass", C.byref(kind), C.byref(n), C.byref(npar),
                              C.byref(nout)) == 0
    assert (kind.value & 1) and n.value == 1 and npar.value == 3
    assert lib.dde_model_info(b"no_such_model", None, None, None, None) != 0
    assert "unknown model" in _lib.last_error()


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(dde_b200.DdeError, match="no CUDA device|CUDA"):
        dde_b200.dopri_batch([[10.0, 1.0, 1.0]], [0.0, 1.0], "lorenz", [10.0, 28.0, 8.0 / 3.0])
    with pytest.raises(dde_b200.DdeError, match="no CUDA device|CUDA"):
        dde_b200.difeq([0.1], 5, "logistic", [1.5])


def test_argument_validation_matches_the_reference():
    # these are rejected before any device work
    with pytest.raises(dde_b200.DdeError, match="At least two times"):
        dde_b200.dopri_batch([[1.0]], [0.0], "exponential", [[1.0]])
    with pytest.raises(dde_b200.DdeError, match="must be one of"):
        dde_b200.dopri_batch([[1.0]], [0.0, 1.0], "exponential", [[1.0]], method="rk4")
    with pytest.raises(dde_b200.DdeError, match="n_history must be at least 2"):
        dde_b200.difeq([1.0], 5, "fib_out", None, n_out=1, n_history=1)
    with pytest.raises(dde_b200.DdeError, match="steps must be positive"):
        dde_b200.difeq([0.1], [-1, 3], "logistic", [1.5])
    with pytest.raises(dde_b200.DdeError):
        dde_b200.dopri_batch([[1.0]], [0.0, 1.0], "no_such_model", [[1.0]])
    # and the product path has no CPU fallback: without a device it fails loudly
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(dde_b200.DdeError, match="no CUDA device"):
            dde_b200.dopri_batch([[1.0]], [0.0, 1.0], "exponential", [[1.0]], grow_history=True)


def test_generated_tableau_is_in_sync():
    """include/dde_b200/coop_tableau_dp8.inc (the block-per-trajectory kernel's dop853 stage
    sums as straight-line code) is what tools/gen_coop_tableau.py makes of the tables in
    dopri_coop_kernel.cuh"""
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    subprocess.run([sys.executable, str(root / "tools" / "gen_coop_tableau.py"), "--check"],
                   check=True)
