"""Golden fixtures: the 14 in-tree example records of the reference, as numbers, under 3 fixed controls.

tests/golden/records_v0.json was produced by tests/golden/make_golden.py in the build container
(the reference has no golden OUTPUTS for this path; see that script's header).
"""
import glob
import json
import os
from dataclasses import asdict

import numpy as np
import pytest

from parity import TOL_FP64, field_scale
from pde_controller_b200 import Predicate, Problem, ProblemBatch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "records_v0.json")))["records"]
REF = "/root/reference"


def problem_of(rec):
    d = dict(rec["problem"])
    d["predicates"] = [Predicate(**q) for q in d["predicates"]]
    d["tree"] = [tuple(t) for t in d["tree"]]
    return Problem(**d)


def golden_batch():
    probs, ctrls, expect, clauses = [], [], [], []
    for rec in GOLD:
        p = problem_of(rec)
        for name, c in rec["controls"].items():
            probs.append(p)
            ctrls.append(None if c is None else (np.array(c[0]), np.array(c[1])))
            expect.append(rec["robustness"][name])
            clauses.append(rec["clauses"][name])
    return ProblemBatch.from_problems(probs, ctrls), np.array(expect), clauses


def test_fixture_shape():
    assert len(GOLD) == 14
    kinds = sorted(r["problem"]["kind"] for r in GOLD)
    assert kinds.count("heat") == 7 and kinds.count("wave") == 7


def test_oracle_reproduces_golden(oracle_lib):
    b, expect, clauses = golden_batch()
    rob, st, cl = oracle_lib.eval_batch(b, want_clauses=True)
    assert (st == 0).all()
    np.testing.assert_allclose(rob, expect, rtol=1e-12, atol=0)
    for s, c in enumerate(clauses):
        np.testing.assert_allclose(cl[: len(c), s], c, rtol=1e-12, atol=0)


def test_numpy_restatement_reproduces_golden():
    from oracle import pde_oracle_np as onp
    for rec in GOLD:
        p = problem_of(rec)
        for name, c in rec["controls"].items():
            ctrl = None if c is None else (np.array(c[0]), np.array(c[1]))
            r, _ = onp.evaluate(p, ctrl)
            want = rec["robustness"][name]
            assert abs(r - want) <= 1e-10 * max(abs(want), 1e-3 * (abs(p.g0) + 1))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_frontend_parses_reference_records_to_the_fixture_numbers():
    from pde_controller_b200.frontend import parse_program
    by_src = {r["source"]: r for r in GOLD}
    seen = 0
    for f in sorted(glob.glob(os.path.join(REF, "utils/few_shot_prompts/examples/*/examples.jsonl"))):
        for line_no, line in enumerate(open(f), 1):
            rec = json.loads(line)
            for key in sorted(k for k in rec if "python" in k):
                src = f"{os.path.relpath(f, REF)}:{line_no}:{key}"
                p = parse_program(rec[key])
                got = json.loads(json.dumps(asdict(p)))
                assert got == by_src[src]["problem"], src
                seen += 1
    assert seen == len(GOLD)


@pytest.mark.gpu
def test_cuda_path_reproduces_golden():
    from pde_controller_b200 import native
    b, expect, clauses = golden_batch()
    out = native.evaluate_batch(b, want_clauses=True)
    assert (out["status"] == 0).all()
    scale = np.maximum(np.abs(expect), 1e-3 * field_scale(b))
    assert (np.abs(out["robustness"] - expect) / scale).max() <= TOL_FP64
