"""Drop-in for the robustness part of the reference's eval_script.py, backed by the CUDA path.

Mirrors test/PDEcontrol/evaluation/eval/eval_script.py:43-102 of the reference: same function
names, argument meaning and return shape.  There the helper spawns
`conda run -n pdecontrol python eval_robustness_wrapper.py <nl> <program>` per sample
(eval_script.py:48-59) and parses one JSON line {"robustness": r, "runtime": t}
(eval_robustness_wrapper.py:18-21).  Here the program text is parsed (no exec) and evaluated on
the GPU; `eval_robustness_batch` does a whole list in ONE call, which is what replaces the
process pool of simulate_gt.py:62-89.

Differences that are deliberate and documented in INTEGRATION.md:
  * the reference returns the robustness under the input femformal SYNTHESISES with Gurobi
    (eval_robustness_wrapper.py:16); the MILP is out of scope, so the control is an argument here:
      - an explicit piecewise-linear input `(knot_times, knot_values)` -> robustness under it;
      - `"search"` (the default of the reference-named helpers): the sampling surrogate of the synthesis
        (search.search_controls, controls family on the GPU) finds an input and the robustness UNDER THAT
        input is returned -- the closest stand-in for the reference's number this tier allows;
      - `"zero"`: zero input, on request only (never silently under the reference's function names).
    The mode is a parameter (`configure(control_mode=...)`, `controls=`) and callers record it.
  * `runtime` is the batch wall time divided by the batch size.
"""
from __future__ import annotations

import json
import operator
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi, native
from .frontend import ParseError, parse_many, parse_program  # noqa: F401
from .iou import eval_IOU, eval_IOU_DPO, eval_IOU_batch  # noqa: F401  (eval_script.py:229-281 of the reference, on the GPU)
from .problem import Problem, ProblemBatch

_DEVICE = 0
_DEVICES = None           # several devices of the box for the batched call (None: the single device above)
_MULTI_EVAL_MIN = 131072  # batches below this are evaluated on one device even when several are configured (the searches are not)
_PRECISION = 64
_CONV = None
_CONTROL_MODE = "search"
_SEARCH = {"population": 4096, "generations": 6, "seed": 0}
CONTROL_MODES = ("search", "zero")


def configure(device: int = 0, precision: int = 64, conventions=None, control_mode: Optional[str] = None,
              search: Optional[dict] = None, devices: Optional[Sequence[int]] = None):
    """Process-wide defaults of the helpers below: device ordinal, 64/32-bit time loop, convention flags, what a
    sample WITHOUT an explicit control is evaluated under (`control_mode`: "search" | "zero") and the search budget
    (`search`: population, generations, seed, elite_frac, bounds).  `devices`: a list of device ordinals for the batched
    call -- the batch is cut into cost-balanced shards (pdeb_eval_batch_multi) and the searches are dealt out over them;
    None (the default) leaves everything on `device`."""
    global _DEVICE, _DEVICES, _PRECISION, _CONV, _CONTROL_MODE, _SEARCH
    _DEVICE, _PRECISION, _CONV = device, precision, conventions
    _DEVICES = [int(d) for d in devices] if devices else None
    if control_mode is not None:
        if control_mode not in CONTROL_MODES:
            raise ValueError(f"control_mode must be one of {CONTROL_MODES}")
        _CONTROL_MODE = control_mode
    if search is not None:
        _SEARCH = {**_SEARCH, **search}


def control_mode() -> str:
    return _CONTROL_MODE


def is_correct(pred, ans, prec=1e-3):
    """eval_script.py:25-30 of the reference, verbatim semantics."""
    import math
    if isinstance(pred, (int, float)) and isinstance(ans, (int, float)):
        return math.isclose(pred, ans, rel_tol=prec)
    if ans == 'timeout':
        return True
    return pred == ans


def eval_robustness_batch(programs: Sequence[str], natural_language: Optional[Sequence[str]] = None,
                          controls=None, device: Optional[int] = None,
                          precision: Optional[int] = None, conventions=None, return_status: bool = False,
                          parse_workers: Optional[int] = None, return_controls: bool = False, search: Optional[dict] = None,
                          devices: Optional[Sequence[int]] = None, control_arrays=None):
    """Evaluate many programs in one GPU call (long lists are parsed by worker processes: frontend.parse_many).

    controls: a list with one entry per program -- `(knot_times, knot_values)` or None -- or one of the mode strings
    "search" / "zero" for every program, or None (= the process-wide mode, `configure(control_mode=)`, "search" unless
    changed).  An entry that is None is evaluated under the mode: "search" runs the sampling surrogate of the
    reference's control synthesis per program (search.search_controls) and evaluates under the input it found,
    "zero" evaluates under zero input.  `search` overrides the search budget for this call.

    control_arrays: the bulk form of `controls` for callers that hold their inputs as arrays -- `(knot_t [n, kmax],
    knot_y [n, kmax], n_knots [n])`, n_knots[i] < 0 meaning "no control for program i" (the mode applies); no Python object
    per program is made.  Give one of `controls` (list form) or `control_arrays`, not both.

    -> (robustness list, runtime list[, status array][, controls used]).  A program that does not parse, or a
    simulation that fails, yields robustness -100 (the failure label of simulate_gt.py:59);
    nothing raises per sample.
    """
    t0 = time.perf_counter()
    n = len(programs)
    mode = _CONTROL_MODE
    if isinstance(controls, str):
        if controls not in CONTROL_MODES:
            raise ValueError(f"controls must be a list or one of {CONTROL_MODES}")
        mode, controls = controls, None
    arr = None
    if control_arrays is not None:
        if controls is not None:
            raise ValueError("give controls or control_arrays, not both")
        at, ay, ak = (np.asarray(a) for a in control_arrays)
        if at.ndim != 2 or at.shape != ay.shape or at.shape[0] != n or ak.shape != (n,):
            raise ValueError("control_arrays = (knot_t [n, kmax], knot_y [n, kmax], n_knots [n])")
        ak = ak.astype(np.int64)
        if (ak > at.shape[1]).any():
            raise ValueError("control_arrays: n_knots exceeds the arrays' second dimension")
        arr = (np.ascontiguousarray(at, dtype=np.float64), np.ascontiguousarray(ay, dtype=np.float64), ak)
    controls = list(controls) if controls is not None else [None] * n
    if len(controls) != n:
        raise ValueError("controls must have one entry per program")
    dev = _DEVICE if device is None else device
    devs = [int(d) for d in devices] if devices else (_DEVICES if device is None else None)
    if devs:
        dev = devs[0]
    conv = conventions if conventions is not None else _CONV
    status = np.zeros(n, dtype=np.int32)
    rob = np.full(n, abi.FAILED_ROBUSTNESS, dtype=np.float64)
    # ---- front-end: the native template matcher takes what is the canonical template (every record of the reference,
    # ~1 us per program, no Python object), the Python front-end the rest (frontend.parse_many: worker processes)
    from . import frontend as _fe
    try:
        wide, accepted = _fe.parse_native(programs)
    except (native.NativeError, OSError):
        wide, accepted = None, np.zeros(n, dtype=bool)
    rest = np.nonzero(~accepted)[0]
    rest_problems: List[Optional[Problem]] = []
    if rest.size:
        as_text = [programs[i] for i in rest]
        as_text = [t.decode("utf-8", "replace") if isinstance(t, (bytes, bytearray)) else t for t in as_text]
        for prob, _msg in parse_many(as_text, workers=parse_workers):
            rest_problems.append(prob)
    bad = [int(i) for i, p in zip(rest, rest_problems) if p is None]
    status[bad] = abi.ST_PARSE
    ok_rest = [(int(i), p) for i, p in zip(rest, rest_problems) if p is not None]
    live = np.ones(n, dtype=bool)
    live[bad] = False
    if live.any():
        # ---- one batch of all n samples (a program that did not parse keeps a place and is masked out afterwards) ----
        pmax = int(max([int(wide.n_pred[accepted].max(initial=1)) if wide is not None else 1] + [len(p.predicates) for _, p in ok_rest] + [1]))
        tmax = int(max([2 * pmax - 1] + [len(p.tree) for _, p in ok_rest]))
        # controls: no per-sample numpy call (a list of n (times, values) pairs goes through two concatenations)
        if arr is not None:
            given = arr[2] >= 0
        else:
            given = (np.fromiter((c is not None for c in controls), dtype=bool, count=n) if None in controls
                     else np.ones(n, dtype=bool))
        need = np.nonzero(~given & live)[0]
        searching = need.size > 0 and mode == "search"
        k_need = np.zeros(0, dtype=np.int64)
        if searching:
            # default knots (C11) of the samples that will be searched: they size the batch's knot planes
            from .problem import default_knot_counts
            kind_all = wide.kind.copy() if wide is not None else np.zeros(n, dtype=np.uint8)
            T_all = wide.T.copy() if wide is not None else np.zeros(n)
            for i, p in ok_rest:
                kind_all[i], T_all[i] = (abi.KIND_HEAT if p.kind == "heat" else abi.KIND_WAVE), p.T
            k_need = default_knot_counts(kind_all[need], T_all[need])
            # more default knots than the controls family takes (a horizon beyond 31 time units): not searched, the sample
            # is evaluated without a control and reports its own status
            k_need = np.where(k_need <= 32, k_need, 0)
        have = np.nonzero(given)[0]
        cnt = np.zeros(n, dtype=np.int64)
        flat_t = flat_y = None
        if have.size and arr is not None:
            len_t = len_y = arr[2][have]
            km_in = np.arange(arr[0].shape[1])[None, :] < len_t[:, None]
            flat_t, flat_y = arr[0][have][km_in], arr[1][have][km_in]          # sample-major, like the concatenations below
            cnt[have] = len_t
        elif have.size:
            sel = controls if have.size == n else [controls[i] for i in have.tolist()]
            ts = list(map(operator.itemgetter(0), sel))
            ys = list(map(operator.itemgetter(1), sel))
            try:
                len_t, len_y = list(map(len, ts)), list(map(len, ys))
                same = len_t == len_y
                len_t = np.array(len_t, dtype=np.int64)
                len_y = len_t if same else np.array(len_y, dtype=np.int64)
                flat_t = np.concatenate(ts, dtype=np.float64) if len_t.sum() else np.zeros(0)
                flat_y = np.concatenate(ys, dtype=np.float64) if len_y.sum() else np.zeros(0)
                if flat_t.ndim != 1 or flat_y.ndim != 1:
                    raise ValueError("multi-dimensional entries")
            except (TypeError, ValueError):
                # scalars, nested or multi-dimensional entries: the slow, forgiving route
                ts = [np.asarray(t, dtype=np.float64).ravel() for t in ts]
                ys = [np.asarray(y, dtype=np.float64).ravel() for y in ys]
                len_t = np.array([t.size for t in ts], dtype=np.int64)
                len_y = np.array([y.size for y in ys], dtype=np.int64)
                flat_t = np.concatenate(ts) if ts else np.zeros(0)
                flat_y = np.concatenate(ys) if ys else np.zeros(0)
            if not np.array_equal(len_t, len_y):
                raise ValueError("control knot times and values differ in length")
            cnt[have] = len_t
        kmax = int(max(cnt.max(initial=0), k_need.max(initial=0), 1))
        batch = ProblemBatch(n, kmax, pmax, tmax)
        if wide is not None:
            for name in batch._names:
                if name in ("knot_t", "knot_y", "n_knots"):
                    continue
                src, dst = getattr(wide, name), getattr(batch, name)
                if src.ndim == 1:
                    dst[:] = src
                else:
                    dst[...] = src[: dst.shape[0]]
        for i, p in ok_rest:
            batch.set_problem(i, p, None)
        # a program that did not parse: a harmless place-holder (its results are overwritten below)
        for i in bad:
            batch.n_pred[i] = 0
        batch.n_knots[:] = cnt
        if flat_t is not None and flat_t.size:
            km = np.arange(kmax)[None, :] < cnt[:, None]
            batch.knot_t.T[km] = flat_t
            batch.knot_y.T[km] = flat_y
        if searching:
            # the sampling surrogate of the reference's synthesis, for every sample without a control: ONE library call
            # (pdeb_controls_search_batch) writes the inputs it finds straight into the batch's knot planes
            from . import search as _search
            from .problem import fill_default_knot_times
            budget = {**_SEARCH, **(search or {})}
            engine = budget.pop("engine", "device")
            b2 = budget.pop("bounds", None)
            b4 = list(_search.DEFAULT_BOUNDS["heat"]) + list(_search.DEFAULT_BOUNDS["wave"]) if b2 is None else [b2[0], b2[1], b2[0], b2[1]]
            fill_default_knot_times(batch, need, k_need)
            fam = (((batch.kind[need] == abi.KIND_HEAT) & (batch.n_elem[need] == 30)) |
                   ((batch.kind[need] == abi.KIND_WAVE) & (batch.n_elem[need] == 20))) & (k_need >= 2)
            done = k_need < 2
            if engine == "device" and fam.any():
                skw = dict(population=budget.get("population", 4096), generations=budget.get("generations", 6),
                           elite_frac=budget.get("elite_frac", 0.02), bounds=b4, seed=_search._seed64(budget.get("seed", 0)),
                           conv=conv, nthreads=budget.get("workers", 0))
                todo = need[fam]
                if devs and len(devs) > 1 and todo.size >= 2 * len(devs):
                    # one Python thread per device, each with its slice of the samples (the library call releases the GIL and
                    # writes only its own samples' columns); a sample's generator stream depends on its index alone
                    import threading
                    found_all = np.zeros(n, dtype=bool)
                    errs = []

                    def _shard(d, part):
                        try:
                            found_all[part] = native.search_batch(batch, samples=part, device=d, **skw)["found"][part]
                        except BaseException as e:          # surfaced on the calling thread
                            errs.append(e)
                    ths = [threading.Thread(target=_shard, args=(d, todo[q::len(devs)])) for q, d in enumerate(devs)]
                    for t in ths:
                        t.start()
                    for t in ths:
                        t.join()
                    if errs:
                        raise errs[0]
                    sr = {"found": found_all}
                else:
                    sr = native.search_batch(batch, samples=todo, device=dev, **skw)
                done[fam] = True
                lost = need[fam][~sr["found"][need[fam]]]
                batch.n_knots[lost] = 0                  # nothing found (an unusable specification): its own status is reported
            rest_i = need[~done]
            if rest_i.size:                              # other meshes, or the host engine on request: problem by problem
                probs = [batch.get_problem(int(i))[0] for i in rest_i]
                kw = {k: v for k, v in budget.items() if k in ("population", "generations", "elite_frac", "seed", "workers")}
                found = _search.search_controls(probs, device=dev, conv=conv, bounds=b2, engine=engine, **kw)
                for i, f in zip(rest_i.tolist(), found):
                    if f["knot_y"] is None:
                        batch.n_knots[i] = 0
                    else:
                        k = len(f["knot_t"])
                        batch.n_knots[i] = k
                        batch.knot_t[:k, i], batch.knot_y[:k, i] = f["knot_t"], f["knot_y"]
            if return_controls:
                for i in need.tolist():
                    k = int(batch.n_knots[i])
                    if k > 0:
                        controls[i] = (batch.knot_t[:k, i].copy(), batch.knot_y[:k, i].copy())
        # the number returned is ALWAYS the batch kernels' robustness under the control in `controls` (found or given)
        prec = _PRECISION if precision is None else precision
        if devs and len(devs) > 1 and n >= _MULTI_EVAL_MIN:
            # worth it for large batches only: 8,192 programs are 40 us of kernel time, less than starting the shard threads
            out = native.evaluate_batch_multi(batch, len(devs), devices=devs, precision=prec, conv=conv)
        else:
            out = native.evaluate_batch(batch, device=dev, precision=prec, conv=conv)
        rob[live] = out["robustness"][live]
        status[live] = out["status"][live]
    if return_controls and arr is not None:
        for i in np.nonzero(arr[2] >= 0)[0].tolist():
            controls[i] = (arr[0][i, : arr[2][i]].copy(), arr[1][i, : arr[2][i]].copy())
    runtime = (time.perf_counter() - t0) / max(n, 1)
    res = (rob.tolist(), [runtime] * n)
    if return_status:
        res = res + (status,)
    if return_controls:
        res = res + (controls,)
    return res


def eval_robustness_helper(robustness, natural_language, verbose=False, control=None):
    """Same call as the reference's helper (its first argument is the PROGRAM text, eval_script.py:43).

    control: `(knot_times, knot_values)`, "search", "zero", or None = the process-wide mode ("search" unless
    `configure(control_mode=...)` says otherwise: the reference's number is the robustness under a SYNTHESISED input,
    eval_robustness_wrapper.py:16, never under zero input).
    Raises RuntimeError when the program cannot be evaluated, as the reference does on a non-zero
    exit of the wrapper process (eval_script.py:63-65); callers turn that into the -100 label.
    """
    rob, rt, st = eval_robustness_batch([robustness], [natural_language], control if isinstance(control, str) else [control],
                                        return_status=True)
    if verbose:
        print(json.dumps({"robustness": rob[0], "runtime": rt[0]}))
    if st[0] != abi.OK:
        raise RuntimeError(f"Command failed with return code {int(st[0])}")
    return rob[0], rt[0]


def eval_robustness(item, pred_key='predicted_python', nl_key='nl', control=None):
    pred = item[pred_key].replace('\x00', '')
    nl = item[nl_key].replace('\x00', '')
    return eval_robustness_helper(pred, nl, control=control)


def eval_robustness_gt(item, nl_key='nl', python_key='python', robustness_key='robustness', time_key='time', control=None):
    return eval_robustness_helper(item[python_key], item[nl_key], control=control)


def eval_robustness_DPO(item, pred_key='predicted_intermediate_python', nl_key='nl', control=None):
    pred = item[pred_key].replace('\x00', '')
    nl = item[nl_key].replace('\x00', '')
    return eval_robustness_helper(pred, nl, control=control)


def wrapper_main(argv=None, out=None):
    """The wire format of eval_robustness_wrapper.py:13-21: argv = [nl_in_prompt, llm_output] -> EXACTLY one JSON line
    {"robustness": r, "runtime": t} on stdout and nothing else (the reference's caller json.loads the whole of stdout,
    eval_script.py:67-72); a failure exits non-zero with the message on stderr (what :63-65 turns into RuntimeError).
    The executable twin is pde_controller_b200/eval_robustness_wrapper.py.  The mode of a call without a control is the
    process-wide one; PDEC_CONTROL_MODE=zero|search in the environment overrides it for a subprocess."""
    import os
    import sys
    argv = sys.argv[1:] if argv is None else argv
    out = sys.stdout if out is None else out
    if len(argv) < 2:
        sys.stderr.write("usage: eval_robustness_wrapper.py <natural language> <program>\n")
        return 2
    env_mode = os.environ.get("PDEC_CONTROL_MODE")
    try:
        robustness, runtime = eval_robustness_helper(argv[1], argv[0], control=env_mode if env_mode in CONTROL_MODES else None)
    except Exception as e:                     # non-zero exit, nothing on stdout
        sys.stderr.write(f"{type(e).__name__}: {e}\n")
        return 1
    out.write(json.dumps({"robustness": robustness, "runtime": runtime}) + "\n")
    return 0
