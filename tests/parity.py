"""The parity metric shared by the CPU-side and GPU-side tests.

PARITY UNPINNED: everything here is CUDA == oracle-v0.  oracle-v0 == the authors' femformal is pinned by nothing: the
reference ships neither the arithmetic nor one labelled value; the golden fixture under
tests/golden/ is produced by this repo's own oracle and guards against drift, it is not external evidence.

FP64 bar (BASELINE.json north_star): 1e-9 relative to the oracle's robustness.  A robustness is a
difference u - p of two quantities of the size of the field, so "relative" needs a floor for the
samples whose robustness happens to be ~0: it is compared against
    max(|r|, 1e-3 * F),   F = |g0| + 1 + max|input| * L / min(E)      (the field scale)
F is the boundary value plus the static response of the rod to the peak input (a temperature rise for
heat, a tip displacement for waves: ~1-5 mm on the synthetic wave set, where g0 = 0).

Why the field scale and not a constant: the kernels evaluate K u in assembled form on scaled unknowns
(DESIGN.md section 4), the oracle in element-flux form.  Both are exact algebra, but per step the
assembled form rounds at eps * |u| and the flux form at eps * |du|; the undamped Newmark scheme carries
that noise forward (growth ~ n^1.5), so after ~800 wave steps the two legitimately differ by up to
~4e-12 of the FIELD AMPLITUDE (measured against an 80-bit build of the oracle: oracle 1e-12, kernels
6e-11 at the 99.9th percentile of this metric).  Heat is dissipative: 5e-13.
"""
import numpy as np

TOL_FP64 = 1e-9
TOL_FP32 = {"heat": 2e-3, "wave": 5e-2}   # FP32 mode is reported separately at its own tolerance


def field_scale(batch):
    peak = np.abs(batch.knot_y).max(axis=0) if batch.knot_y.shape[0] else 0.0
    return np.abs(batch.g0) + 1.0 + peak * batch.L / np.minimum(batch.E_in, batch.E_out)


def scale_of(ref, batch):
    return np.maximum(np.abs(ref), 1e-3 * field_scale(batch))


def rel_err(got, ref, batch):
    """The floored metric the assertions use (see the module docstring)."""
    return np.abs(got - ref) / scale_of(ref, batch)


PURE_REL_MIN = 1e-6      # |r| below this is excluded from the pure-relative figure (a quotient by ~0 says nothing)


def pure_rel_err(got, ref):
    """|got - ref| / |ref| with NO floor -- the literal reading of the north star's "1e-9 relative" -- over the samples
    with |ref| >= PURE_REL_MIN; reported next to the floored metric so the two readings cannot be confused.  Returns
    (max, number of samples it covers)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    keep = np.abs(ref) >= PURE_REL_MIN
    if not keep.any():
        return 0.0, 0
    return float((np.abs(got - ref)[keep] / np.abs(ref)[keep]).max()), int(keep.sum())
