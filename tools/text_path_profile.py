"""Where the text-level end-to-end call spends its time (development tool; run on the GPU box).
  python tools/text_path_profile.py [n_programs]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_controller_b200 import eval_script as es, frontend, native, synth  # noqa: E402
from pde_controller_b200.frontend import render_program  # noqa: E402

m = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
src = synth.make_batch(m, "heat", seed=0, uniform_knots=True)
probs, ctrls = zip(*(src.get_problem(i) for i in range(m)))
texts = [render_program(p) for p in probs]
ctrls = list(ctrls)
es.eval_robustness_batch(texts[:64], controls=ctrls[:64])
for rep in range(3):
    t0 = time.perf_counter()
    es.eval_robustness_batch(texts, controls=ctrls)
    print(f"call {rep}: {m / (time.perf_counter() - t0):.4g} programs/s")
t0 = time.perf_counter(); wide, acc = frontend.parse_native(texts); t1 = time.perf_counter()
print(f"parse_native {1e3 * (t1 - t0):.1f} ms, accepted {int(acc.sum())}")
b = synth.make_batch(m, "heat", seed=0)
t0 = time.perf_counter(); out = native.evaluate_batch(b); t1 = time.perf_counter()
print(f"evaluate_batch {1e3 * (t1 - t0):.1f} ms")
t0 = time.perf_counter(); out = native.evaluate_batch(b); t1 = time.perf_counter()
print(f"evaluate_batch again {1e3 * (t1 - t0):.1f} ms")
pr = cProfile.Profile(); pr.enable(); es.eval_robustness_batch(texts, controls=ctrls); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
