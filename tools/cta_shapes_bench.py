"""Times the kernel families that can take the refined heat mesh (development tool; run under gpurun).

  python tools/cta_shapes_bench.py [lib.so ...]

PDEB_CTA = 0: chunk family (warp per trajectory, R = 32); 1: CTA family 4 warps x 8 rows; 2: CTA family 2 warps x 16 rows.
Each library given (default: the in-tree one) is loaded in a fresh process."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def one():
    sys.path.insert(0, ROOT)
    import numpy as np
    from pde_controller_b200 import native, synth
    b = synth.make_batch(4096, "heat_refined", seed=0)
    out, ref = {}, None
    for mode in os.environ.get("CTA_MODES", "0,1,2").split(","):
        os.environ["PDEB_CTA"] = mode
        r = native.ResidentBatch(b)
        for _ in range(2):
            r.eval(64, timing=True)
        ms = sorted(r.eval(64, timing=True).kernel_ms for _ in range(5))
        got = r.fetch()["robustness"]
        r.close()
        ref = got if ref is None else ref
        out[mode] = {"ms": round(ms[2], 2), "evals_per_s": round(4096 / (ms[2] * 1e-3)), "max_abs_diff": float(np.abs(got - ref).max())}
    print(json.dumps(out))


if __name__ == "__main__":
    if os.environ.get("CTA_CHILD"):
        one()
    else:
        libs = sys.argv[1:] or [os.path.join(ROOT, "pde_controller_b200", "libpdeb200.so")]
        for lib in libs:
            p = subprocess.run([sys.executable, __file__], env=dict(os.environ, PDEB_LIB=lib, CTA_CHILD="1"), capture_output=True, text=True)
            print(os.path.basename(lib), p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-300:], flush=True)
