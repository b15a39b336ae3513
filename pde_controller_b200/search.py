"""Sampling search for the boundary input of one specification.

In the reference the control is synthesised by the Gurobi MILP inside control.femformal (call site
test/PDEcontrol/evaluation/eval/eval_robustness_wrapper.py:16) and the dataset labels are that optimum; the MILP
is out of scope here.  This module is its sampling surrogate on the GPU: generations of candidate piecewise-linear
inputs are scored by the controls family (one specification, many controls, arg-max on the device) and the
sampling distribution is re-centred on the elite of each generation (cross-entropy method).

`search_controls` runs the search for a LIST of problems (a few host threads, one CUDA stream each: the C calls
release the GIL, so the host side of one search overlaps the kernels of the others) and is what the drop-in helpers
call for a sample without an explicit control (`eval_script.eval_robustness_batch(controls="search")`,
`simulate_gt --control search`), so that the number returned under the reference's function names is, as in the
reference, the robustness under a synthesised input -- comparable with the labels at the reference's own 1e-3 match
rule (eval_script.py:25-30) to the extent a sampling search reaches the MILP optimum.
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import native
from .problem import Problem, ProblemBatch, default_knot_times

# C11 magnitudes of the reference's case studies: heat flux in [0, 1e6], wave force in [-5e3, 5e3]
DEFAULT_BOUNDS = {"heat": (0.0, 1.0e6), "wave": (-5.0e3, 5.0e3)}
_DEFAULT_MESH = {"heat": 30, "wave": 20}          # meshes the controls family is instantiated for


class _BatchScorer:
    """Scores candidate controls of a specification on a mesh the controls family does not cover: the specification
    is replicated once per candidate and goes through the batch kernels (chunk family), arg-max on the host."""

    def __init__(self, spec: ProblemBatch, knot_t, device, conv):
        self.spec, self.knot_t, self.device, self.conv = spec, knot_t, device, conv

    def eval(self, y, stream=0):
        nk, m = y.shape
        b = ProblemBatch(m, max(nk, 1), self.spec.pmax, self.spec.tmax)
        for name in b._names:
            if name in ("knot_t", "knot_y", "n_knots"):
                continue
            src, dst = getattr(self.spec, name), getattr(b, name)
            if src.ndim == 1:
                dst[:] = src[0]
            else:
                dst[:, :] = src[:, 0:1]
        b.n_knots[:] = nk
        b.knot_t[:nk, :] = self.knot_t[:, None]
        b.knot_y[:nk, :] = y
        out = native.evaluate_batch(b, device=self.device, conv=self.conv, stream=stream, timing=True)
        r = np.where(out["status"] == 0, out["robustness"], -np.inf)
        bi = int(np.argmax(r)) if m else -1
        ok = bi >= 0 and np.isfinite(r[bi])
        return {"robustness": out["robustness"], "status": out["status"], "best_index": bi if ok else -1,
                "best_value": float(r[bi]) if ok else -np.inf, "kernel_ms": float(out["timing"].total_ms)}

    def close(self):
        pass


def search_control(problem: Problem, knot_t: Optional[np.ndarray] = None, bounds: Optional[Tuple[float, float]] = None,
                   population: int = 16384, generations: int = 8, elite_frac: float = 0.02, seed=0, device: int = 0,
                   conv=None, stream: int = 0):
    """Best-found PWL control for `problem`.

    Returns dict(knot_t, knot_y, robustness, history, evaluations, kernel_ms): `history` is the best robustness after
    every generation; it never decreases because the incumbent is re-evaluated with each generation.  The first
    generation always holds the ZERO input (clipped into the bounds), so the result is never worse than it.
    Deterministic for a given seed."""
    if knot_t is None:
        knot_t = default_knot_times(problem)
    knot_t = np.asarray(knot_t, dtype=np.float64)
    lo, hi = bounds if bounds is not None else DEFAULT_BOUNDS[problem.kind]
    nk = knot_t.size
    population = max(int(population), 4)
    spec = ProblemBatch.from_problems([problem], kmax=max(nk, 1))
    if problem.n_elem == _DEFAULT_MESH.get(problem.kind):
        ev = native.ControlsEvaluator(spec, knot_t, capacity=population, device=device, conv=conv)
    else:
        ev = _BatchScorer(spec, knot_t, device, conv)
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    mean = np.full(nk, 0.5 * (lo + hi))
    std = np.full(nk, 0.5 * (hi - lo))
    best_y, best_r = None, -np.inf
    history, kernel_ms = [], 0.0
    n_elite = max(2, int(round(elite_frac * population)))
    try:
        for g in range(generations):
            if g == 0:
                y = rng.uniform(lo, hi, size=(nk, population))
                y[:, 1] = np.clip(0.0, lo, hi)        # the zero input competes from the start
                y[:, 2] = lo
                y[:, 3] = hi
            else:
                y = np.clip(mean[:, None] + std[:, None] * rng.standard_normal((nk, population)), lo, hi)
            if best_y is not None:
                y[:, 0] = best_y                      # keep the incumbent in the population
            res = ev.eval(y, stream=stream)
            kernel_ms += res["kernel_ms"]
            r = np.where(res["status"] == 0, res["robustness"], -np.inf)
            if res["best_index"] >= 0 and res["best_value"] > best_r:
                best_r, best_y = res["best_value"], y[:, res["best_index"]].copy()
            elite = np.argpartition(-r, n_elite - 1)[:n_elite]
            elite = elite[np.isfinite(r[elite])]
            if elite.size >= 2:
                mean = y[:, elite].mean(axis=1)
                std = np.maximum(y[:, elite].std(axis=1), 1e-3 * (hi - lo))
            history.append(best_r)
    finally:
        ev.close()
    return {"knot_t": knot_t, "knot_y": best_y, "robustness": best_r, "history": history,
            "evaluations": population * generations, "kernel_ms": kernel_ms}


def search_controls(problems: Sequence[Problem], population: int = 4096, generations: int = 6, elite_frac: float = 0.02,
                    seed: int = 0, device: int = 0, conv=None, bounds: Optional[Tuple[float, float]] = None,
                    workers: int = 4) -> List[dict]:
    """search_control for every problem of a list; problem i uses the seed (seed, i), so the result does not depend
    on the number of workers.  A problem the search cannot run for (a specification the controls set-up rejects)
    yields knot_y None and robustness -inf: the caller evaluates it as it would without a control and reports that
    sample's own status."""
    n = len(problems)
    out: List[Optional[dict]] = [None] * n
    if n == 0:
        return []
    L = native.lib()
    workers = max(1, min(int(workers), n))
    errors: List[BaseException] = []

    def work(w):
        stream = L.pdeb_stream_create(device) if workers > 1 else 0
        try:
            for i in range(w, n, workers):
                try:
                    out[i] = search_control(problems[i], bounds=bounds, population=population, generations=generations,
                                            elite_frac=elite_frac, seed=(seed, i), device=device, conv=conv,
                                            stream=stream or 0)
                except native.NativeError as e:
                    if native.device_count() < 1:
                        raise
                    out[i] = {"knot_t": default_knot_times(problems[i]), "knot_y": None, "robustness": -np.inf,
                              "history": [], "evaluations": 0, "kernel_ms": 0.0, "error": str(e)}
        except BaseException as e:        # surfaced on the calling thread
            errors.append(e)
        finally:
            if stream:
                L.pdeb_stream_sync(stream)
                L.pdeb_stream_destroy(stream)

    if workers == 1:
        work(0)
    else:
        ths = [threading.Thread(target=work, args=(w,)) for w in range(workers)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
    if errors:
        raise errors[0]
    return out  # type: ignore[return-value]
