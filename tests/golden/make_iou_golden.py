"""Golden vectors for the IoU boolean-grid metric, produced by the REAL reference.

Runs the reference's own `eval_IOU` (test/PDEcontrol/evaluation/eval/eval_script.py:233-281, with `check_conditions`
:437-487 and the G/F bit-array classes :159-219) UNMODIFIED on pairs of sSTL strings and writes
tests/golden/iou_v0.json.  The reference imports two packages the build container does not have (`bitarray`,
`editdistance`); tests/golden/_ref_stubs/ holds minimal stand-ins with exactly the operations the IoU code uses.
The reference evaluates a Python `eval` per grid cell (~1.4 M per clause): one clause takes ~100 s, so the pairs run in
worker processes (about ten minutes on 8 cores).  Build-container only: the GPU box has no /root/reference; the tests
read the committed JSON.

    python tests/golden/make_iou_golden.py
"""
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_EVAL = "/root/reference/test/PDEcontrol/evaluation/eval"

C = {
    "h1": "F_[[1.8288, 4.6769]] (\\forall x \\in [5, 97]  (u(x) - (0.0771 \\cdot x + 326.154) > 0))",
    "h1_thr": "F_[[1.8288, 4.6769]] (\\forall x \\in [5, 97]  (u(x) - (0.0771 \\cdot x + 326.154) > 25))",
    "h1_G": "G_[[1.8288, 4.6769]] (\\forall x \\in [5, 97]  (u(x) - (0.0771 \\cdot x + 326.154) > 0))",
    "h1_lt": "F_[[1.8288, 4.6769]] (\\forall x \\in [5, 97]  (u(x) - (0.0771 \\cdot x + 326.154) < 0))",
    "h1_pos": "F_[[1.0, 3.5]] (\\forall x \\in [20.7, 60.2]  (u(x) - (0.0771 \\cdot x + 326.154) > 0))",
    "w1": "F_[[0.09, 0.192]] (\\forall x \\in [32712, 42454]  (u(x) - (-4.692e-05 \\cdot x + 1.3255) > 0))",
    "w1_pos": "F_[[0.09, 0.192]] (\\forall x \\in [30000, 45000.5]  (u(x) - (-4.692e-05 \\cdot x + 1.3255) > 0))",
    "g2": "G_[[0.5, 1.5]] (\\forall x \\in [10, 40]  (u(x) - (0.25 \\cdot x + 290.5) > 0)) \\land G_[[2.0, 3.0]] (\\forall x \\in [50, 90]  (u(x) - (-0.031 \\cdot x + 310.0) < 0))",
    "g2_shift": "G_[[0.7, 1.5]] (\\forall x \\in [10, 40]  (u(x) - (0.25 \\cdot x + 290.5) > 0)) \\land G_[[2.0, 3.25]] (\\forall x \\in [50, 90]  (u(x) - (-0.031 \\cdot x + 310.0) < 0))",
    "g2_or": "G_[[0.5, 1.5]] (\\forall x \\in [10, 40]  (u(x) - (0.25 \\cdot x + 290.5) > 0)) \\lor G_[[2.0, 3.0]] (\\forall x \\in [50, 90]  (u(x) - (-0.031 \\cdot x + 310.0) < 0))",
    "fg": "F_[[0.5, 1.5]] (\\forall x \\in [10, 40]  (u(x) - (0.25 \\cdot x + 290.5) > 0)) \\land G_[[2.0, 3.0]] (\\forall x \\in [50, 90]  (u(x) - (-0.031 \\cdot x + 310.0) < 0))",
    "gf": "G_[[0.5, 1.5]] (\\forall x \\in [10, 40]  (u(x) - (0.25 \\cdot x + 290.5) > 0)) \\land F_[[2.0, 3.0]] (\\forall x \\in [50, 90]  (u(x) - (-0.031 \\cdot x + 310.0) < 0))",
    "eq": "G_[[0.5, 1.5]] (\\forall x \\in [10, 40]  (u(x) - (0.0 \\cdot x + 1.0) = 0))",
    "eq2": "G_[[0.5, 1.5]] (\\forall x \\in [10, 12]  (u(x) - (0.0 \\cdot x + 0.0) = -500))",
    "div": "G_[[0.5, 2.5]] (\\forall x \\in [8, 64]  (u(x) - (x/(4) + 290.5) > 0))",
    "mul": "G_[[0.5, 2.5]] (\\forall x \\in [8, 64]  (u(x) - (0.25 \\cdot x + 290.5) > 0))",
    "nop": "G_[[0.5, 2.5]] (\\forall x \\in [8, 64]  (u(x) > 300))",
    "nop2": "G_[[0.5, 2.0]] (\\forall x \\in [8, 64]  (u(x) >= 280))",
    "br": "(G_[[0.5, 1.0]] (\\forall x \\in [10, 40]  (u(x) - (0.25 \\cdot x + 290.5) > 0)) \\land G_[[1.5, 2.0]] (\\forall x \\in [50, 90]  (u(x) - (-0.031 \\cdot x + 310.0) < 0))) \\lor (G_[[2.5, 3.0]] (\\forall x \\in [10, 40]  (u(x) - (0.1 \\cdot x + 300.0) > 0)) \\land G_[[3.5, 4.0]] (\\forall x \\in [50, 90]  (u(x) - (0.0 \\cdot x + 280.0) > 0)))",
    "br_flat": "G_[[0.5, 1.0]] (\\forall x \\in [10, 40]  (u(x) - (0.25 \\cdot x + 290.5) > 0)) \\land G_[[1.5, 2.0]] (\\forall x \\in [50, 90]  (u(x) - (-0.031 \\cdot x + 310.0) < 0)) \\lor G_[[2.5, 3.0]] (\\forall x \\in [10, 40]  (u(x) - (0.1 \\cdot x + 300.0) > 0))",
    "g3": "G_[[0.2, 1.1]] (\\forall x \\in [12, 12]  (u(x) - (0.1138 \\cdot x + 284.19) > 0)) \\land G_[[0.1, 1.1]] (\\forall x \\in [48.0, 63.0]  (u(x) - (-0.0667 \\cdot x + 278.9) < 0)) \\land F_[[0.3, 0.9]] (\\forall x \\in [70, 95]  (u(x) - (0.05 \\cdot x + 270.0) > 0))",
}
PAIRS = [("h1", "h1_thr"), ("h1", "h1_G"), ("h1", "h1_lt"), ("h1", "h1_pos"), ("w1", "w1_pos"), ("g2", "g2_shift"),
         ("g2", "g2_or"), ("fg", "gf"), ("g2", "fg"), ("eq", "eq2"), ("div", "mul"), ("nop", "nop2"), ("mul", "nop"),
         ("br", "br_flat"), ("g3", "g2"), ("h1", "bad")]
C["bad"] = "this is not a formula"


def one(pair):
    sys.path.insert(0, os.path.join(HERE, "_ref_stubs"))
    sys.path.insert(0, REF_EVAL)
    import eval_script as ref
    gt, pred = pair
    t0 = time.time()
    try:
        v = ref.eval_IOU({"sstl": C[gt], "predicted_sstl": C[pred]})
    except Exception as e:                     # what the reference does with this pair: an exception, recorded as such
        v = {"raises": type(e).__name__}
    return {"gt": gt, "pred": pred, "sstl": C[gt], "predicted_sstl": C[pred], "iou": v, "seconds": round(time.time() - t0, 1)}


if __name__ == "__main__":
    with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
        out = pool.map(one, PAIRS, chunksize=1)
    with open(os.path.join(HERE, "iou_v0.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_iou_golden.py: the reference's own eval_IOU (eval_script.py:233-281), unmodified",
                   "pairs": out}, f, indent=1)
    for r in out:
        print(r["gt"], r["pred"], r["iou"], r["seconds"])
