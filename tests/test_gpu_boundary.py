"""The reference-shaped Python surface (eval_script mirror, simulate_gt driver) on the GPU."""
import json

import numpy as np
import pytest

from parity import TOL_FP64, rel_err
from pde_controller_b200 import ProblemBatch, synth
from pde_controller_b200.frontend import render_program

pytestmark = pytest.mark.gpu


def programs(n, wl, seed):
    b = synth.make_batch(n, wl, seed=seed)
    probs, ctrls = zip(*(b.get_problem(s) for s in range(n)))
    return b, [render_program(p) for p in probs], list(ctrls)


def test_eval_robustness_batch_matches_oracle(oracle_lib):
    from pde_controller_b200 import eval_script as es
    for wl in ("heat", "wave"):
        b, texts, ctrls = programs(200, wl, 31)
        texts[7] = "this is not a program"
        texts[11] = texts[11].replace("dt = ", "dt = 0 * ")
        rob, rt, st = es.eval_robustness_batch(texts, ["nl"] * len(texts), ctrls, return_status=True)
        ref, ref_st = oracle_lib.eval_batch(b)
        assert st[7] == 1 and rob[7] == -100.0 and st[11] == 1 and rob[11] == -100.0
        keep = np.ones(len(texts), bool)
        keep[[7, 11]] = False
        assert (st[keep] == ref_st[keep]).all()
        assert rel_err(np.array(rob), ref, b)[keep].max() <= TOL_FP64
        assert len(rt) == len(texts) and all(t > 0 for t in rt)


def test_reference_shaped_helpers(oracle_lib):
    from pde_controller_b200 import eval_script as es
    b, texts, ctrls = programs(3, "heat", 32)
    ref, _ = oracle_lib.eval_batch(b)
    r, t = es.eval_robustness_helper(texts[0], "some natural language", control=ctrls[0])
    assert isinstance(r, float) and isinstance(t, float) and abs(r - ref[0]) <= 1e-9 * max(abs(ref[0]), 0.3)
    item = {"python": texts[1], "predicted_python": texts[1] + "\x00", "predicted_intermediate_python": texts[1], "nl": "x\x00"}
    for fn in (es.eval_robustness, es.eval_robustness_gt, es.eval_robustness_DPO):
        r, _ = fn(item, control=ctrls[1])
        assert abs(r - ref[1]) <= 1e-9 * max(abs(ref[1]), 0.3)
    with pytest.raises(RuntimeError):          # the reference raises on a non-zero exit of the wrapper process
        es.eval_robustness_helper("N = ", "nl")
    assert es.is_correct(1.0005, 1.0) and not es.is_correct(1.01, 1.0) and es.is_correct(3.0, "timeout")


def test_simulate_gt_driver(tmp_path, oracle_lib):
    from pde_controller_b200 import simulate_gt
    b, texts, ctrls = programs(40, "wave", 33)
    d = tmp_path / "wave_nc_test"
    d.mkdir()
    with open(d / "test.jsonl", "w") as f:
        for i, (txt, c) in enumerate(zip(texts, ctrls)):
            rec = {"nl": f"sample {i}", "sstl": "G_[[0, 1]] (...)", "python": txt,
                   "control": {"t": list(map(float, c[0])), "y": list(map(float, c[1]))}}
            f.write(json.dumps(rec) + "\n")
    args = simulate_gt.build_parser().parse_args(["--data_dir", str(d), "--max_num_examples", "16", "--eval_robustness"])
    m = simulate_gt.main(args)
    out = [json.loads(l) for l in open(str(d) + "_16/test.jsonl")]
    assert len(out) == 16 and m["n_samples"] == 16
    assert abs(m["gt positive robustness rate"] + m["gt negative robustness rate"] + m["gt failed robustness rate"] - 1) < 1e-12
    ref, _ = oracle_lib.eval_batch(b)
    for rec in out:
        i = int(rec["nl"].split()[1])
        assert abs(rec["robustness"] - ref[i]) <= 1e-9 * max(abs(ref[i]), 1e-3)
        assert list(rec.keys())[-2:] == ["robustness", "time"]
    assert json.load(open(str(d) + "_16/metrics.json")) == m
