"""IoU boolean-grid metric: oracle against the REAL reference's outputs (CPU), CUDA path against both (GPU).

tests/golden/iou_v0.json was produced by the reference's own eval_IOU (eval_script.py:233-281), unmodified, in the build
container (tests/golden/make_iou_golden.py).  This is the one sub-path of the repository whose parity is PINNED by the
reference itself: integer / boolean work, so the bar is exact equality."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "iou_v0.json")))["pairs"]


def want_of(rec):
    return rec["iou"]


@pytest.mark.parametrize("rec", GOLD, ids=lambda r: f"{r['gt']}-{r['pred']}")
def test_oracle_reproduces_the_reference(rec):
    from oracle import iou_oracle
    want = want_of(rec)
    if isinstance(want, dict):                                  # the reference raised behind its parser
        with pytest.raises(Exception):
            iou_oracle.eval_iou(rec["sstl"], rec["predicted_sstl"])
        return
    got = iou_oracle.eval_iou(rec["sstl"], rec["predicted_sstl"])
    assert got == want                                          # exact: a ratio of two bit counts (or "failed")


def test_formula_structure_follows_the_reference():
    from pde_controller_b200 import iou
    a = "G_[[0, 1]] (\\forall x \\in [1, 2]  (u(x) - (1 \\cdot x + 0) > 0))"
    cl, j, t = iou.split_formula(f"({a} \\land {a}) \\lor ({a} \\land {a})")
    assert (j, t) == (["and", "or", "and"], [[0, 1], [2, 3]])
    cl, j, t = iou.split_formula(f"{a} \\land {a} \\lor {a}")
    assert (j, t) == (["and", "or"], [0, 1, 2])                 # no precedence: folded left to right
    c = iou.parse_clause("(" + a.replace("> 0", "\\leq 3"))
    assert (c["op"], c["thr"], c["p0"], c["p1"], c["kind"], c["coef"]) == ("<=", 3.0, 1, 2, "mul", 1.0)
    c = iou.parse_clause("F_[[0.5, 2]] (\\forall x \\in [8.9, 64.2]  (u(x) - (-x/(4) - 2.5) >= 1))")
    assert (c["temporal_op"], c["p0"], c["p1"], c["kind"], c["coef"], c["neg"], c["c"], c["op"]) == ("F", 8, 64, "div", 4.0, True, -2.5, ">=")


def test_without_a_device_the_metric_fails_loudly():
    from pde_controller_b200 import iou, native
    if native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(native.NativeError):
        iou.eval_IOU({"sstl": GOLD[0]["sstl"], "predicted_sstl": GOLD[0]["predicted_sstl"]})
    assert iou.eval_IOU({"sstl": "nonsense", "predicted_sstl": "nonsense"}) == "failed"     # parse failure: no device needed


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_exactly():
    from pde_controller_b200 import iou
    items = [{"sstl": r["sstl"], "predicted_sstl": r["predicted_sstl"]} for r in GOLD]
    got = iou.eval_IOU_batch(items)
    for r, g in zip(GOLD, got):
        want = want_of(r)
        if isinstance(want, dict):
            assert g == "failed", (r["gt"], r["pred"])
        else:
            assert g == want, (r["gt"], r["pred"], g, want)
    # one at a time, through the reference-named call, and the DPO key names
    assert iou.eval_IOU(items[0]) == want_of(GOLD[0])
    assert iou.eval_IOU_DPO({"w_sstl": items[1]["sstl"], "predicted_intermediate_sstl": items[1]["predicted_sstl"]}) == want_of(GOLD[1])


@pytest.mark.gpu
def test_cuda_path_equals_oracle_on_generated_pairs():
    """More formulas than the reference could be run on in reasonable time: CUDA against the (pinned) oracle, exactly."""
    import numpy as np
    from oracle import iou_oracle
    from pde_controller_b200 import iou
    rng = np.random.default_rng(7)

    def clause():
        top = "GF"[rng.integers(2)]
        t0 = round(float(rng.uniform(0, 3)), 3)
        t1 = round(t0 + float(rng.uniform(0.05, 2)), 3)
        a = int(rng.integers(0, 60))
        b = a + int(rng.integers(1, 40))
        m, c = round(float(rng.uniform(-1, 1)), 4), round(float(rng.uniform(-300, 400)), 3)
        form = [f"{m} \\cdot x + {c}", f"x/({int(rng.integers(2, 9))}) + {c}", f"x \\cdot {m} + {c}", f"{m}x + {c}"][rng.integers(4)]
        op = ["<", ">", "<=", ">=", "="][rng.integers(5)]
        thr = int(rng.integers(-50, 50))
        return f"{top}_[[{t0}, {t1}]] (\\forall x \\in [{a}, {b}]  (u(x) - ({form}) {op} {thr}))"

    def formula():
        n = int(rng.integers(1, 4))
        out = clause()
        for _ in range(n - 1):
            out += [" \\land ", " \\lor "][rng.integers(2)] + clause()
        return out

    items = [{"sstl": formula(), "predicted_sstl": formula()} for _ in range(24)]
    got = iou.eval_IOU_batch(items)
    for it, g in zip(items, got):
        try:
            want = iou_oracle.eval_iou(it["sstl"], it["predicted_sstl"])
        except ValueError:
            want = "failed"
        assert g == want, (it, g, want)


@pytest.mark.gpu
def test_iou_throughput_is_recorded():
    """Not an assertion on speed: records pairs/s of the batched call next to the reference's ~60-250 s per pair."""
    import time
    from pde_controller_b200 import iou
    items = [{"sstl": r["sstl"], "predicted_sstl": r["predicted_sstl"]} for r in GOLD if not isinstance(r["iou"], (dict, str))] * 16
    iou.eval_IOU_batch(items[:4])
    t0 = time.perf_counter()
    out = iou.eval_IOU_batch(items)
    dt = time.perf_counter() - t0
    assert len(out) == len(items)
    ref_s = sum(r["seconds"] for r in GOLD if not isinstance(r["iou"], (dict, str)))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/iou_throughput.json", "w") as f:
        json.dump({"pairs": len(items), "seconds": dt, "pairs_per_s": len(items) / dt,
                   "reference_seconds_for_the_15_golden_pairs_one_core_each": ref_s}, f)
