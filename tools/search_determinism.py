"""Repeats a sweep of searches under heavy concurrency and compares every result with the single-thread run (development tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_controller_b200 import search, synth  # noqa: E402

n = 4096
src = synth.make_batch(n, "mixed", seed=3)
probs = [src.get_problem(i)[0] for i in range(n)]
ref = [o["robustness"] for o in search.search_controls(probs, workers=1)]
for rep in range(4):
    for w in (16, 32):
        out = search.search_controls(probs, workers=w)
        r = [o["robustness"] for o in out]
        bad = [i for i, (a, b) in enumerate(zip(r, ref)) if a != b]
        print(f"rep {rep} threads={w}: {len(bad)} of {n} differ {bad[:5]}", flush=True)
