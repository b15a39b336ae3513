"""Device time of one generation of the search against the plain scoring of the same population (development tool)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_controller_b200 import ProblemBatch, default_knot_times, native, synth  # noqa: E402

src = synth.make_batch(64, "heat", seed=0)
for i in (0, 7, 21):
    p = src.get_problem(i)[0]
    kt = default_knot_times(p)
    spec = ProblemBatch.from_problems([p], kmax=kt.size)
    for P in (4096, 16384):
        ev = native.ControlsEvaluator(spec, kt, capacity=P)
        y = np.random.default_rng(0).uniform(0, 1e6, (kt.size, P))
        ev.eval(y)
        t_eval = min(ev.eval(y)["kernel_ms"] for _ in range(5))
        t1 = min(ev.search(P, 1, 0.02, 0.0, 1e6, seed=1)["kernel_ms"] for _ in range(5))
        t6 = min(ev.search(P, 6, 0.02, 0.0, 1e6, seed=1)["kernel_ms"] for _ in range(5))
        print(f"sample {i} T={p.T:.1f} P={P}: score only {1e3*t_eval:.0f} us (incl. arg-max launch), one generation {1e3*t1:.0f} us, six {1e3*t6:.0f} us")
        ev.close()
