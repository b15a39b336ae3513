"""Search-mode throughput of pdeb_controls_search_batch against the number of host threads (development tool, GPU box)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_controller_b200 import search, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
src = synth.make_batch(n, "heat", seed=0)
probs = [src.get_problem(i)[0] for i in range(n)]
search.search_controls(probs[:64], workers=4)
for rep in range(2):
    for w in (1, 2, 4, 8, 12, 16, 24, 32):
        t0 = time.perf_counter(); out = search.search_controls(probs, workers=w); dt = time.perf_counter() - t0
        print(f"rep {rep} threads={w}: {n/dt:.0f} programs/s, kernel_ms/program {sum(o['kernel_ms'] for o in out)/n:.2f}", flush=True)
