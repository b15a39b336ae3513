"""Throughput of the search mode behind the drop-in call (development tool; run on the GPU box).
  python tools/search_rate.py [n_programs] [workers]"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_controller_b200 import eval_script as es, search, synth  # noqa: E402
from pde_controller_b200.frontend import render_program  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
workers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
src = synth.make_batch(n, "heat", seed=0)
probs = [src.get_problem(i)[0] for i in range(n)]
texts = [render_program(p) for p in probs]
es.eval_robustness_batch(texts[:8], controls="search")
t0 = time.perf_counter()
rob, _ = es.eval_robustness_batch(texts, controls="search")
dt = time.perf_counter() - t0
print(f"eval_robustness_batch(search): {n / dt:.1f} programs/s ({1e3 * dt / n:.2f} ms per program), positive {sum(r > 0 for r in rob)}/{n}")
for w in (1, workers):
    t0 = time.perf_counter()
    out = search.search_controls(probs, workers=w)
    dt = time.perf_counter() - t0
    print(f"search_controls workers={w}: {n / dt:.1f} programs/s; kernel_ms per program {sum(o['kernel_ms'] for o in out) / n:.3f}")
pr = cProfile.Profile(); pr.enable(); search.search_controls(probs[:64], workers=1); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
