"""Sampling search for the boundary input of one specification.

In the reference the control is synthesised by the Gurobi MILP inside control.femformal (call site
test/PDEcontrol/evaluation/eval/eval_robustness_wrapper.py:16) and the dataset labels are that optimum; the MILP
is out of scope here.  This module is its sampling surrogate on the GPU: generations of candidate piecewise-linear
inputs are scored by the controls family (one specification, many controls, arg-max on the device) and the
sampling distribution is re-centred on the elite of each generation (cross-entropy method), so that
`eval_robustness_helper(program, nl, control=None)`-style callers can ask for a best-found robustness that is
comparable with the labels at the reference's own 1e-3 match rule (eval_script.py:25-30).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import native
from .problem import Problem, ProblemBatch, default_knot_times

# C11 magnitudes of the reference's case studies: heat flux in [0, 1e6], wave force in [-5e3, 5e3]
DEFAULT_BOUNDS = {"heat": (0.0, 1.0e6), "wave": (-5.0e3, 5.0e3)}


def search_control(problem: Problem, knot_t: Optional[np.ndarray] = None, bounds: Optional[Tuple[float, float]] = None,
                   population: int = 16384, generations: int = 8, elite_frac: float = 0.02, seed: int = 0, device: int = 0,
                   conv=None):
    """Best-found PWL control for `problem`.

    Returns dict(knot_t, knot_y, robustness, history, evaluations, kernel_ms): `history` is the best robustness after
    every generation; it never decreases because the incumbent is re-evaluated with each generation."""
    if knot_t is None:
        knot_t = default_knot_times(problem)
    knot_t = np.asarray(knot_t, dtype=np.float64)
    lo, hi = bounds if bounds is not None else DEFAULT_BOUNDS[problem.kind]
    nk = knot_t.size
    spec = ProblemBatch.from_problems([problem], kmax=max(nk, 1))
    ev = native.ControlsEvaluator(spec, knot_t, capacity=population, device=device, conv=conv)
    rng = np.random.Generator(np.random.PCG64(seed))
    mean = np.full(nk, 0.5 * (lo + hi))
    std = np.full(nk, 0.5 * (hi - lo))
    best_y, best_r = None, -np.inf
    history, kernel_ms = [], 0.0
    n_elite = max(2, int(round(elite_frac * population)))
    try:
        for g in range(generations):
            if g == 0:
                y = rng.uniform(lo, hi, size=(nk, population))
            else:
                y = np.clip(mean[:, None] + std[:, None] * rng.standard_normal((nk, population)), lo, hi)
            if best_y is not None:
                y[:, 0] = best_y                      # keep the incumbent in the population
            res = ev.eval(y)
            kernel_ms += res["kernel_ms"]
            r = np.where(res["status"] == 0, res["robustness"], -np.inf)
            if res["best_index"] >= 0 and res["best_value"] > best_r:
                best_r, best_y = res["best_value"], y[:, res["best_index"]].copy()
            elite = np.argpartition(-r, n_elite - 1)[:n_elite]
            mean = y[:, elite].mean(axis=1)
            std = np.maximum(y[:, elite].std(axis=1), 1e-3 * (hi - lo))
            history.append(best_r)
    finally:
        ev.close()
    return {"knot_t": knot_t, "knot_y": best_y, "robustness": best_r, "history": history,
            "evaluations": population * generations, "kernel_ms": kernel_ms}
