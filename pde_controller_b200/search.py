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


ENGINES = ("device", "host")
_MAX_DEVICE_KNOTS = 32          # PDEB_CTL_KMAX
_MAX_DEVICE_ELITE = 2048        # PDEB_CEM_MAX_ELITE
_MAX_DEVICE_GENERATIONS = 64    # PDEB_CEM_MAX_GEN


_M64 = (1 << 64) - 1


def _mix64(z: int) -> int:
    """splitmix64 finaliser: pdeb_mix64 of csrc/pdeb_controls.cu."""
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _seed64(seed) -> int:
    """An integer seed IS the 64-bit generator seed; anything else goes through numpy's SeedSequence."""
    if isinstance(seed, (int, np.integer)):
        return int(seed) & _M64
    return int(np.random.SeedSequence(seed).generate_state(1, dtype=np.uint64)[0])


def sample_seed(seed, i: int) -> int:
    """Generator seed of problem i of a sweep: what pdeb_controls_search_batch derives for sample i."""
    return _mix64(_seed64(seed) ^ _mix64(int(i)))


def search_control(problem: Problem, knot_t: Optional[np.ndarray] = None, bounds: Optional[Tuple[float, float]] = None,
                   population: int = 16384, generations: int = 8, elite_frac: float = 0.02, seed=0, device: int = 0,
                   conv=None, stream: int = 0, engine: str = "device", evaluator=None):
    """Best-found PWL control for `problem`.

    Returns dict(knot_t, knot_y, robustness, history, evaluations, kernel_ms): `history` is the best robustness after
    every generation; it never decreases because the incumbent is re-evaluated with each generation.  The first
    generation always holds the ZERO input (clipped into the bounds), so the result is never worse than it.
    Deterministic for a given seed.

    engine: "device" (default) runs the whole search in ONE library call (pdeb_controls_search: candidates drawn, scored
    and ranked on the GPU, one launch per generation); "host" draws and ranks with numpy and calls the device once per
    generation (the round-1 form; also what meshes outside the controls family use, whatever `engine` says).  The two
    draw different candidates (different generators), so their results differ like two seeds do.
    evaluator: a ControlsEvaluator of capacity >= population to re-target (`respec`) instead of creating one: a sweep over
    many programs keeps one per thread (search_controls)."""
    if engine not in ENGINES:
        raise ValueError(f"engine must be one of {ENGINES}")
    if knot_t is None:
        knot_t = default_knot_times(problem)
    knot_t = np.asarray(knot_t, dtype=np.float64)
    lo, hi = bounds if bounds is not None else DEFAULT_BOUNDS[problem.kind]
    nk = knot_t.size
    population = max(int(population), 4)
    spec = ProblemBatch.from_problems([problem], kmax=max(nk, 1))
    family = problem.n_elem == _DEFAULT_MESH.get(problem.kind)
    n_elite = max(2, int(round(elite_frac * population)))
    if (engine == "device" and family and nk <= _MAX_DEVICE_KNOTS and n_elite <= _MAX_DEVICE_ELITE
            and 1 <= generations <= _MAX_DEVICE_GENERATIONS):
        own = evaluator is None
        ev = native.ControlsEvaluator(spec, knot_t, capacity=population, device=device, conv=conv) if own else evaluator
        try:
            if not own:
                ev.respec(spec, knot_t, conv=conv)
            res = ev.search(population, generations, elite_frac, lo, hi, seed=_seed64(seed), stream=stream)
        finally:
            if own:
                ev.close()
        return {"knot_t": knot_t, "knot_y": res["knot_y"], "robustness": res["robustness"], "history": res["history"],
                "evaluations": population * generations, "kernel_ms": res["kernel_ms"], "engine": "device"}
    if family:
        ev = native.ControlsEvaluator(spec, knot_t, capacity=population, device=device, conv=conv)
    else:
        ev = _BatchScorer(spec, knot_t, device, conv)
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
    mean = np.full(nk, 0.5 * (lo + hi))
    std = np.full(nk, 0.5 * (hi - lo))
    best_y, best_r = None, -np.inf
    history, kernel_ms = [], 0.0
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
            "evaluations": population * generations, "kernel_ms": kernel_ms, "engine": "host"}


def search_controls(problems: Sequence[Problem], population: int = 4096, generations: int = 6, elite_frac: float = 0.02,
                    seed: int = 0, device: int = 0, conv=None, bounds: Optional[Tuple[float, float]] = None,
                    workers: int = 0, engine: str = "device") -> List[dict]:
    """search_control for every problem of a list; problem i uses the seed sample_seed(seed, i), so the result does not
    depend on the number of workers and equals `search_control(problems[i], seed=sample_seed(seed, i))`.  A problem the
    search cannot run for (a specification the controls set-up rejects) yields knot_y None and robustness -inf: the
    caller evaluates it as it would without a control and reports that sample's own status.

    engine "device": ONE library call for the whole list (pdeb_controls_search_batch: `workers` host threads inside the
    library, 0 = all, each with a stream and a controls handle it re-targets from program to program; nothing per program
    in Python), then the per-problem route below for whatever is outside the controls family.  engine "host": the
    per-problem route for everything (Python threads, numpy draws; at most 8 workers)."""
    n = len(problems)
    out: List[Optional[dict]] = [None] * n
    if n == 0:
        return []
    if engine not in ENGINES:
        raise ValueError(f"engine must be one of {ENGINES}")
    L = native.lib()
    todo = list(range(n))
    population = max(int(population), 4)
    n_elite = max(2, int(round(elite_frac * population)))
    if engine == "device" and n_elite <= _MAX_DEVICE_ELITE and 1 <= generations <= _MAX_DEVICE_GENERATIONS:
        from .problem import default_knot_counts, fill_default_knot_times
        spec = ProblemBatch.from_problems(list(problems), kmax=1)
        K = default_knot_counts(spec.kind, spec.T)
        fam = np.array([p.n_elem == _DEFAULT_MESH.get(p.kind) for p in problems]) & (K <= _MAX_DEVICE_KNOTS)
        idx = np.nonzero(fam)[0]
        if idx.size:
            wide = ProblemBatch(n, int(K[idx].max()), spec.pmax, spec.tmax)
            for name in wide._names:
                if name in ("knot_t", "knot_y", "n_knots"):
                    continue
                src, dst = getattr(spec, name), getattr(wide, name)
                dst[...] = src
            fill_default_knot_times(wide, idx, K[idx])
            b4 = list(DEFAULT_BOUNDS["heat"]) + list(DEFAULT_BOUNDS["wave"])
            if bounds is not None:
                b4 = [bounds[0], bounds[1], bounds[0], bounds[1]]
            res = native.search_batch(wide, samples=idx, population=population, generations=generations, elite_frac=elite_frac,
                                      bounds=b4, seed=_seed64(seed), device=device, conv=conv, nthreads=workers, timing=True)
            per = res["kernel_ms"] / idx.size
            for i in idx.tolist():
                k = int(wide.n_knots[i])
                ok = bool(res["found"][i])
                r = float(res["best"][i]) if ok else -np.inf
                out[i] = {"knot_t": wide.knot_t[:k, i].copy(), "knot_y": wide.knot_y[:k, i].copy() if ok else None,
                          "robustness": r, "history": [r], "evaluations": population * generations, "kernel_ms": per,
                          "engine": "device"}
            todo = [i for i in todo if out[i] is None]
    if not todo:
        return out  # type: ignore[return-value]
    workers = max(1, min(int(workers) if workers > 0 else 4, 8, len(todo)))
    errors: List[BaseException] = []

    def work(w):
        stream = L.pdeb_stream_create(device) if workers > 1 else 0
        try:
            for i in todo[w::workers]:
                try:
                    out[i] = search_control(problems[i], bounds=bounds, population=population, generations=generations,
                                            elite_frac=elite_frac, seed=sample_seed(seed, i), device=device, conv=conv,
                                            stream=stream or 0, engine=engine)
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
