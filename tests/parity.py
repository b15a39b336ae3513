"""The parity metric shared by the CPU-side and GPU-side tests.

FP64 bar (BASELINE.json north_star): 1e-9 relative to the oracle's robustness.  Robustness is a
difference u - p of quantities of the size of the field, so "relative" is taken against
max(|r|, 1e-3 * (|g0| + 1)): a robustness that happens to be ~0 is compared at 1e-3 of the field
scale (0.3 K for heat at 300 K, 1e-3 mm for waves) instead of against itself.
"""
import numpy as np

TOL_FP64 = 1e-9
TOL_FP32 = {"heat": 2e-3, "wave": 5e-2}   # FP32 mode is reported separately at its own tolerance


def scale_of(ref, batch):
    return np.maximum(np.abs(ref), 1e-3 * (np.abs(batch.g0) + 1.0))


def rel_err(got, ref, batch):
    return np.abs(got - ref) / scale_of(ref, batch)
