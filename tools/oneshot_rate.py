"""Throughput of the ONE-SHOT entry (pdeb_eval_batch: host arrays in, results out) against the batch size, from ordinary
(pageable) numpy memory and from a page-locked slab (development tool, GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_controller_b200 import native, synth  # noqa: E402

for n in (65536, 262144, 1048576):
    b = synth.make_batch(n, "heat", seed=0, uniform_knots=True)
    hp = native.pinned_batch(b.n, b.kmax, b.pmax, b.tmax, uniform_knots=True)
    hp.slab[:] = b.slab
    for name, batch in (("pageable", b), ("page-locked slab", hp)):
        native.evaluate_batch(batch)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); out = native.evaluate_batch(batch); ts.append(time.perf_counter() - t0)
        print(f"n={n:8d} {name:17s}: {n / min(ts):.4g} evals/s ({1e3 * min(ts):.2f} ms)", flush=True)
