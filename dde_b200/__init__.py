"""dde_b200 -- B200-native batched integrator behind mrc-ide/dde's solver API.

Thousands to millions of independent trajectories of one compiled model are
advanced concurrently by hand-written sm_100a FP64 kernels: Dormand-Prince
dopri5 / dop853 with per-trajectory adaptive steps, dense output,
per-trajectory HBM ring-buffer history for delay look-ups (`ylag`),
`dopri_interpolate`, and the `difeq` / `difeq_replicate` iteration with
`yprev` look-back.  See DESIGN.md and include/dde_b200.h.
"""
from .api import (DOPRI_5, DOPRI_853, OK_COMPLETE, DdeError, DifeqResult, DopriBatchResult,  # noqa: F401
                  DopriHistory, DopriResult, difeq, difeq_replicate, dopri, dopri5, dopri853,
                  dopri_batch, dopri_continue, dopri_interpolate, models, strerror)
from .batch import DifeqBatch, DopriBatch, DopriShards  # noqa: F401
from ._lib import load_model_library  # noqa: F401

__all__ = [
    "dopri", "dopri5", "dopri853", "dopri_batch", "dopri_continue", "dopri_interpolate", "difeq",
    "difeq_replicate", "DopriBatch", "DopriShards", "DifeqBatch", "load_model_library", "models", "strerror", "DdeError",
]
