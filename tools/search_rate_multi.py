"""Search-mode throughput of the drop-in call over 1..N devices of the box (development tool, GPU box).
  python tools/search_rate_multi.py [n_programs]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pde_controller_b200 import eval_script as es, native, synth  # noqa: E402
from pde_controller_b200.frontend import render_program  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
src = synth.make_batch(n, "heat", seed=0)
texts = [render_program(src.get_problem(i)[0]) for i in range(n)]
nd = native.device_count()
ref = None
k = 1
while k <= nd:
    devs = list(range(k))
    es.eval_robustness_batch(texts[: 64 * k], controls="search", devices=devs)
    t0 = time.perf_counter()
    rob, _ = es.eval_robustness_batch(texts, controls="search", devices=devs)
    dt = time.perf_counter() - t0
    same = "" if ref is None else f", equal to one device: {rob == ref}"
    ref = rob if ref is None else ref
    t0 = time.perf_counter()
    es.eval_robustness_batch(texts, controls="zero", devices=devs)
    dz = time.perf_counter() - t0
    print(f"devices={k}: search {n / dt:.0f} programs/s, zero-input {n / dz:.0f} programs/s{same}", flush=True)
    k *= 2
