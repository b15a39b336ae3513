"""Drop-in for the robustness part of the reference's eval_script.py, backed by the CUDA path.

Mirrors test/PDEcontrol/evaluation/eval/eval_script.py:43-102 of the reference: same function
names, argument meaning and return shape.  There the helper spawns
`conda run -n pdecontrol python eval_robustness_wrapper.py <nl> <program>` per sample
(eval_script.py:48-59) and parses one JSON line {"robustness": r, "runtime": t}
(eval_robustness_wrapper.py:18-21).  Here the program text is parsed (no exec) and evaluated on
the GPU; `eval_robustness_batch` does a whole list in ONE call, which is what replaces the
process pool of simulate_gt.py:62-89.

Differences that are deliberate and documented in INTEGRATION.md:
  * the control input is an explicit argument (`control=`); the reference synthesises it with
    Gurobi inside femformal, which is out of scope.  None means zero input.
  * `runtime` is the batch wall time divided by the batch size.
"""
from __future__ import annotations

import json
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi, native
from .frontend import ParseError, parse_many, parse_program  # noqa: F401
from .problem import Problem, ProblemBatch

_DEVICE = 0
_PRECISION = 64
_CONV = None


def configure(device: int = 0, precision: int = 64, conventions=None):
    """Process-wide defaults of the helpers below (device ordinal, 64/32-bit time loop, convention flags)."""
    global _DEVICE, _PRECISION, _CONV
    _DEVICE, _PRECISION, _CONV = device, precision, conventions


def is_correct(pred, ans, prec=1e-3):
    """eval_script.py:25-30 of the reference, verbatim semantics."""
    import math
    if isinstance(pred, (int, float)) and isinstance(ans, (int, float)):
        return math.isclose(pred, ans, rel_tol=prec)
    if ans == 'timeout':
        return True
    return pred == ans


def eval_robustness_batch(programs: Sequence[str], natural_language: Optional[Sequence[str]] = None,
                          controls: Optional[Sequence] = None, device: Optional[int] = None,
                          precision: Optional[int] = None, conventions=None, return_status: bool = False,
                          parse_workers: Optional[int] = None):
    """Evaluate many programs in one GPU call (long lists are parsed by worker processes: frontend.parse_many).

    -> (robustness list, runtime list[, status array]).  A program that does not parse, or a
    simulation that fails, yields robustness -100 (the failure label of simulate_gt.py:59);
    nothing raises per sample.
    """
    t0 = time.perf_counter()
    n = len(programs)
    controls = list(controls) if controls is not None else [None] * n
    if len(controls) != n:
        raise ValueError("controls must have one entry per program")
    problems: List[Optional[Problem]] = []
    status = np.zeros(n, dtype=np.int32)
    for i, (prob, _msg) in enumerate(parse_many(programs, workers=parse_workers)):
        problems.append(prob)
        if prob is None:
            status[i] = abi.ST_PARSE
    live = [i for i in range(n) if problems[i] is not None]
    rob = np.full(n, abi.FAILED_ROBUSTNESS, dtype=np.float64)
    if live:
        batch = ProblemBatch.from_problems([problems[i] for i in live], [controls[i] for i in live])
        out = native.evaluate_batch(batch, device=_DEVICE if device is None else device,
                                    precision=_PRECISION if precision is None else precision,
                                    conv=conventions if conventions is not None else _CONV)
        rob[live] = out["robustness"]
        status[live] = out["status"]
    runtime = (time.perf_counter() - t0) / max(n, 1)
    res = ([float(r) for r in rob], [runtime] * n)
    return res + (status,) if return_status else res


def eval_robustness_helper(robustness, natural_language, verbose=False, control=None):
    """Same call as the reference's helper (its first argument is the PROGRAM text, eval_script.py:43).

    Raises RuntimeError when the program cannot be evaluated, as the reference does on a non-zero
    exit of the wrapper process (eval_script.py:63-65); callers turn that into the -100 label.
    """
    rob, rt, st = eval_robustness_batch([robustness], [natural_language], [control], return_status=True)
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


def wrapper_main(argv=None):
    """The wire format of eval_robustness_wrapper.py:13-21: argv = [nl, program] -> one JSON line."""
    import sys
    argv = sys.argv[1:] if argv is None else argv
    robustness, runtime = eval_robustness_helper(argv[1], argv[0])
    print(json.dumps({"robustness": robustness, "runtime": runtime}))
