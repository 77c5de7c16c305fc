"""Loading the committed reference fixtures (tests/golden/*.npz, written by
tools/gen_golden.py from the unmodified reference)."""
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def _coerce(v):
    if v in ("True", "False"):
        return v == "True"
    try:
        return int(v)
    except ValueError:
        pass
    try:
        return float(v)
    except ValueError:
        return v


def cases(kind):
    return sorted(p.stem for p in GOLDEN.glob(kind + "_*.npz"))


def load(name):
    z = np.load(GOLDEN / (name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["model"] = str(d["model"])
    d["parms"] = d["parms"] if bool(d["has_parms"]) else None
    if str(d["kind"]) == "dopri":
        d["tcrit"] = d["tcrit"] if d["tcrit"].size else None
        d["kw"] = {str(k): _coerce(str(v)) for k, v in zip(d["kw_keys"], d["kw_vals"])}
    d["n_out"] = int(d["n_out"])
    d["output"] = d["output"] if d["n_out"] > 0 else None
    return d
