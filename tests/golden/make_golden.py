"""Generate tests/golden/records_v0.json (run in the build container, where /root/reference exists).

The reference ships NO golden outputs for this path: its 14 example records
(utils/few_shot_prompts/examples/*/examples.jsonl) carry inputs only.  This script
  1. parses those records with the front-end and stores the NUMBERS of each problem (not the text),
  2. evaluates each under three fixed controls with the C oracle (v0 conventions),
so that the fixture pins (a) the front-end on the reference's own inputs and (b) the oracle /
CUDA path against drift.  It does not pin the oracle to femformal: parity stays "unpinned".
"""
import glob
import json
import os
import sys
from dataclasses import asdict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from pde_controller_b200 import ProblemBatch, default_knot_times  # noqa: E402
from pde_controller_b200.frontend import parse_program  # noqa: E402

REF = "/root/reference/utils/few_shot_prompts/examples"


def controls_for(p, idx):
    t = default_knot_times(p)
    amp = 1.0e6 if p.kind == "heat" else 5.0e3
    rng = np.random.default_rng(1000 + idx)
    rnd = rng.uniform(0.0, amp, len(t)) if p.kind == "heat" else rng.uniform(-amp, amp, len(t))
    return {"zero": None, "ramp": (t, np.linspace(0.0, 0.6 * amp, len(t))), "seeded": (t, rnd)}


def main():
    records = []
    idx = 0
    for f in sorted(glob.glob(os.path.join(REF, "*", "examples.jsonl"))):
        for line_no, line in enumerate(open(f), 1):
            rec = json.loads(line)
            for key in sorted(k for k in rec if "python" in k):
                p = parse_program(rec[key])
                ctr = controls_for(p, idx)
                names = list(ctr)
                b = ProblemBatch.from_problems([p] * len(names), [ctr[n] for n in names])
                rob, st, cl = oracle.eval_batch(b, want_clauses=True)
                assert (st == 0).all()
                records.append({
                    "source": f"{os.path.relpath(f, '/root/reference')}:{line_no}:{key}",
                    "problem": asdict(p),
                    "sstl": rec.get("sstl"),            # the record's LaTeX twin (data), for the sstl cross-check tests
                    "controls": {n: (None if ctr[n] is None else [list(map(float, ctr[n][0])), list(map(float, ctr[n][1]))])
                                 for n in names},
                    "robustness": {n: float(rob[i]) for i, n in enumerate(names)},
                    "clauses": {n: [float(x) for x in cl[: len(p.predicates), i]] for i, n in enumerate(names)},
                })
                idx += 1
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "records_v0.json")
    with open(out, "w") as fh:
        json.dump({"conventions": "v0", "generator": "tests/golden/make_golden.py", "records": records}, fh, indent=1)
    print(f"wrote {len(records)} records to {out}")


if __name__ == "__main__":
    main()
