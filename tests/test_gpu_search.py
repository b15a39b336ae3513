"""Search mode behind the drop-in call, the wrapper executable, and the round-2 boundary fixes, on the GPU.

The reference returns the robustness under a SYNTHESISED input (eval_robustness_wrapper.py:16); the MILP is out of
scope, `controls="search"` is its sampling surrogate.  What can be tested without femformal: the search result is never
below the zero-input result, it is reproducible under a seed, and the number returned is the oracle's robustness under
the input that is returned with it.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from parity import TOL_FP64, rel_err
from pde_controller_b200 import ProblemBatch, default_conventions, native, synth
from pde_controller_b200.frontend import render_program
from test_golden import GOLD, problem_of

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUDGET = {"population": 2048, "generations": 4, "seed": 5}


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert native.device_count() >= 1, "no CUDA device: the -m gpu tests must run on the B200 box"


def test_search_mode_on_the_14_reference_records(oracle_lib):
    from pde_controller_b200 import eval_script as es
    probs = [problem_of(r) for r in GOLD]
    texts = [render_program(p) for p in probs]
    zero, _, st0 = es.eval_robustness_batch(texts, controls="zero", return_status=True)
    found, _, st1, ctrls = es.eval_robustness_batch(texts, controls="search", return_status=True, return_controls=True,
                                                     search=BUDGET)
    again, _ = es.eval_robustness_batch(texts, controls="search", search=BUDGET)
    assert (st0 == 0).all() and (st1 == 0).all()
    assert found == again                                  # reproducible under a seed, whatever the thread schedule
    for z, f in zip(zero, found):
        assert f >= z - 1e-9 * max(abs(z), 1.0)            # the zero input is a candidate of the first generation
    # the number returned is the robustness UNDER the returned input: check it against the oracle on that input
    b = ProblemBatch.from_problems(probs, ctrls)
    ref, ref_st = oracle_lib.eval_batch(b)
    assert (ref_st == 0).all()
    assert rel_err(np.array(found), ref, b).max() <= TOL_FP64
    assert sum(f > z + 1e-6 for z, f in zip(zero, found)) >= 3     # and the search does improve on zero input


def test_default_mode_of_the_reference_named_helpers_is_search(oracle_lib):
    from pde_controller_b200 import eval_script as es
    assert es.control_mode() == "search"
    es.configure(search=BUDGET)
    p = problem_of(GOLD[0])
    text = render_program(p)
    r_search, _ = es.eval_robustness_helper(text, "nl")
    r_zero, _ = es.eval_robustness_helper(text, "nl", control="zero")
    item = {"python": text, "nl": "nl"}
    assert es.eval_robustness_gt(item)[0] == r_search      # same seed -> same number
    b0 = ProblemBatch.from_problems([p], [None])
    assert abs(r_zero - oracle_lib.eval_batch(b0)[0][0]) <= 1e-9 * max(abs(r_zero), 0.3)
    assert r_search >= r_zero - 1e-9
    es.configure(control_mode="zero")
    try:
        assert es.eval_robustness_helper(text, "nl")[0] == r_zero
    finally:
        es.configure(control_mode="search")


def test_search_on_a_mesh_outside_the_controls_family(oracle_lib):
    """N = 12 goes through the batch kernels with a replicated specification (search._BatchScorer)."""
    from pde_controller_b200 import search
    p = problem_of(GOLD[0])
    p.n_elem = 12
    out = search.search_control(p, population=256, generations=3, seed=2)
    b = ProblemBatch.from_problems([p], [(out["knot_t"], out["knot_y"])])
    ref, st = oracle_lib.eval_batch(b)
    assert st[0] == 0 and abs(out["robustness"] - ref[0]) <= 1e-9 * max(abs(ref[0]), 0.3)
    assert all(b >= a for a, b in zip(out["history"], out["history"][1:]))


def test_simulate_gt_control_modes(tmp_path):
    from pde_controller_b200 import simulate_gt
    texts = [render_program(problem_of(r)) for r in GOLD[:6]]
    d = tmp_path / "heat_nc_test"
    d.mkdir()
    with open(d / "test.jsonl", "w") as f:
        for i, t in enumerate(texts):
            f.write(json.dumps({"nl": f"sample {i}", "sstl": "", "python": t}) + "\n")
    res = {}
    for mode in ("zero", "search"):
        args = simulate_gt.build_parser().parse_args(["--data_dir", str(d), "--max_num_examples", "6", "--eval_robustness",
                                                      "--control", mode, "--search_population", "1024", "--search_generations", "3"])
        m = simulate_gt.main(args)
        assert m["control_mode"] == mode
        recs = [json.loads(l) for l in open(str(d) + "_6/test.jsonl")]
        res[mode] = {r["nl"]: r for r in recs}
        assert json.load(open(str(d) + "_6/metrics.json"))["control_mode"] == mode
    for k, rz in res["zero"].items():
        rs = res["search"][k]
        assert rs["robustness"] >= rz["robustness"] - 1e-9
        assert "control" in rs and len(rs["control"]["t"]) == len(rs["control"]["y"]) and "control" not in rz


def test_wrapper_executable_prints_exactly_one_json_object():
    """eval_robustness_wrapper.py:13-21 of the reference: argv = [nl, program]; stdout is parsed whole by json.loads."""
    text = render_program(problem_of(GOLD[3]))
    exe = os.path.join(ROOT, "pde_controller_b200", "eval_robustness_wrapper.py")
    env = dict(os.environ, PDEC_CONTROL_MODE="zero")
    p = subprocess.run([sys.executable, exe, "some natural language", text], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr
    out = json.loads(p.stdout)                       # the whole of stdout is ONE JSON object
    assert set(out) == {"robustness", "runtime"} and p.stdout.count("\n") == 1
    assert isinstance(out["robustness"], float) and out["runtime"] > 0
    want = GOLD[3]["robustness"]["zero"]
    assert abs(out["robustness"] - want) <= 1e-9 * max(abs(want), 0.3)
    bad = subprocess.run([sys.executable, exe, "nl", "N = "], capture_output=True, text=True, env=env, timeout=300)
    assert bad.returncode != 0 and bad.stdout == "" and bad.stderr


def test_wrapper_main_in_process(capsys):
    from pde_controller_b200 import eval_script as es
    text = render_program(problem_of(GOLD[8]))
    es.configure(control_mode="zero")
    try:
        assert es.wrapper_main(["nl", text]) == 0
    finally:
        es.configure(control_mode="search")
    out = capsys.readouterr().out
    assert json.loads(out)["robustness"] == pytest.approx(GOLD[8]["robustness"]["zero"], rel=1e-9)
    assert es.wrapper_main(["nl", "not a program"]) == 1 and capsys.readouterr().out == ""


@pytest.mark.parametrize("wl", ["heat", "wave"])
@pytest.mark.parametrize("bad", [np.inf, np.nan])
def test_nonfinite_input_at_a_late_knot_is_flagged(oracle_lib, wl, bad):
    """A trajectory that turns inf / NaN part-way through a window is status NONFINITE / -100 on both sides (the
    running maxima of the kernels drop NaN margins: the state itself is tested).  A non-finite knot BEHIND the last
    observed step is not an error."""
    b = synth.make_batch(256, wl, seed=21)
    last = b.n_knots - 1
    hit = np.arange(b.n) % 2 == 0
    for s in np.nonzero(hit)[0]:
        b.knot_y[last[s], s] = bad
        b.pred_tlo[0, s], b.pred_thi[0, s] = 0.0, b.T[s]          # clause 0 observes the whole run
    got = native.evaluate_batch(b, want_clauses=True)
    ref, ref_st = oracle_lib.eval_batch(b)
    assert (got["status"] == ref_st).all()
    assert (ref_st[hit] == 2).all() and (got["robustness"][hit] == -100.0).all()
    clean = synth.make_batch(256, wl, seed=21)            # the parity scale is formed from finite knot values
    assert rel_err(got["robustness"], ref, clean)[~hit].max() <= TOL_FP64
    # ... and behind the last observed step it is harmless
    c = synth.make_batch(256, wl, seed=22)
    for s in range(c.n):
        c.knot_y[c.n_knots[s] - 1, s] = bad
        for j in range(c.n_pred[s]):
            c.pred_thi[j, s] = min(c.pred_thi[j, s], 0.5 * c.T[s])
            c.pred_tlo[j, s] = min(c.pred_tlo[j, s], c.pred_thi[j, s])
    got = native.evaluate_batch(c)
    ref, ref_st = oracle_lib.eval_batch(c)
    assert (got["status"] == ref_st).all()
    ok = ref_st == 0
    assert ok.sum() > 200 and rel_err(got["robustness"], ref, synth.make_batch(256, wl, seed=22))[ok].max() <= TOL_FP64


def test_precision_32_is_refused_for_meshes_without_an_fp32_kernel():
    b = synth.make_batch(64, "heat", seed=4)
    b.n_elem[:] = 48                                     # chunk family: FP64 only
    with pytest.raises(native.NativeError, match="precision 32"):
        native.evaluate_batch(b, precision=32)
    r = native.ResidentBatch(b)
    with pytest.raises(native.NativeError, match="precision 32"):
        r.eval(32)
    r.eval(64)
    r.close()


@pytest.mark.parametrize("kw", [dict(), dict(steps_rule=1, window_rule=1), dict(steps_rule=2)])
def test_plan_orders_by_the_callers_conventions(kw):
    """The order is formed under the step / window rules the caller evaluates with: warps hold trajectories whose last
    observed step (under THOSE rules) falls into one bucket of 8 steps, longest first."""
    b = synth.make_batch(8192, "heat", seed=9)
    b.dt[:] = 0.0437                                     # windows no longer fall on the grid: the rules differ
    conv = default_conventions(**kw)
    r = native.ResidentBatch(b)
    r.eval(64, conv)
    order = r.plan_order()
    r.close()
    assert sorted(order.tolist()) == list(range(b.n))
    q = b.pred_thi / b.dt
    a1 = np.trunc(q) if conv.window_rule == 1 else np.floor(q + 1e-9)
    live = np.arange(b.pmax)[:, None] < b.n_pred[None, :]
    S = b.steps(conv.steps_rule)
    last = np.minimum(np.where(live, a1, 0).max(axis=0), S).astype(np.int64)
    bucket = np.maximum(last, 1)[order] // 8
    assert (np.diff(bucket) <= 0).all()                  # non-increasing buckets along the order


def _spec_of(p):
    from pde_controller_b200 import default_knot_times
    kt = default_knot_times(p)
    return ProblemBatch.from_problems([p], kmax=max(kt.size, 1)), kt


@pytest.mark.parametrize("rec", [0, 9])
def test_device_search_one_generation_equals_numpy_ranking(rec):
    """pdeb_cem_update_kernel against numpy on the very candidates the device drew: elite = the n best (ties by lowest
    index), mean / standard deviation of their knot values, incumbent = first best candidate."""
    p = problem_of(GOLD[rec])
    spec, kt = _spec_of(p)
    lo, hi = (0.0, 1.0e6) if p.kind == "heat" else (-5.0e3, 5.0e3)
    P, frac = 3000, 0.05
    ev = native.ControlsEvaluator(spec, kt, capacity=4096)
    try:
        res = ev.search(P, 1, frac, lo, hi, seed=77)
        y = ev.fetch_knots(P)                                    # generation 0, as drawn on the device
        out = ev.eval(None, m=P)                                 # the same candidates scored again
    finally:
        ev.close()
    assert y.shape == (kt.size, P) and (y >= lo).all() and (y <= hi).all()
    assert (y[:, 1] == min(max(0.0, lo), hi)).all() and (y[:, 2] == lo).all() and (y[:, 3] == hi).all()
    r = np.where((out["status"] == 0) & np.isfinite(out["robustness"]), out["robustness"], -np.inf)
    n_el = max(2, int(round(frac * P)))
    order = np.lexsort((np.arange(P), -r))[:n_el]                # value descending, index ascending
    order = order[np.isfinite(r[order])]
    assert order.size >= 2
    mean, std = y[:, order].mean(axis=1), np.maximum(y[:, order].std(axis=1), 1e-3 * (hi - lo))
    assert np.allclose(res["mean"], mean, rtol=1e-12, atol=0) and np.allclose(res["std"], std, rtol=1e-10, atol=0)
    bi = int(np.argmax(r))
    assert res["robustness"] == r[bi] and np.array_equal(res["knot_y"], y[:, bi]) and res["history"] == [r[bi]]
    assert res["launches"] == 1                                  # draw, score and rank of a generation are ONE launch
    # uniform draws: every knot's sample mean is near the middle of the bounds
    assert np.all(np.abs(y[:, 4:].mean(axis=1) - 0.5 * (lo + hi)) < 0.05 * (hi - lo))


def test_device_search_properties_on_the_reference_records(oracle_lib):
    """Never below the zero input, deterministic for a seed, monotone history, and the number returned is the oracle's
    robustness under the input returned with it."""
    from pde_controller_b200 import search
    probs = [problem_of(r) for r in GOLD]
    kw = dict(population=2048, generations=5, elite_frac=0.03)
    a = search.search_controls(probs, seed=3, workers=4, **kw)
    b = search.search_controls(probs, seed=3, workers=1, **kw)
    c = search.search_controls(probs, seed=4, workers=4, **kw)
    single = search.search_control(probs[5], seed=search.sample_seed(3, 5), **kw)
    assert all(x["engine"] == "device" for x in a)
    assert len(single["history"]) == 5 and all(h2 >= h1 for h1, h2 in zip(single["history"], single["history"][1:]))
    assert [x["robustness"] for x in a] == [x["robustness"] for x in b]          # whatever the thread schedule
    assert all(np.array_equal(x["knot_y"], y["knot_y"]) for x, y in zip(a, b))
    assert single["robustness"] == a[5]["robustness"] and np.array_equal(single["knot_y"], a[5]["knot_y"])
    assert any(x["robustness"] != y["robustness"] for x, y in zip(a, c))          # another seed, other candidates
    ctrls = [(x["knot_t"], x["knot_y"]) for x in a]
    batch = ProblemBatch.from_problems(probs, ctrls)
    ref, st = oracle_lib.eval_batch(batch)
    assert (st == 0).all()
    assert rel_err(np.array([x["robustness"] for x in a]), ref, batch).max() <= TOL_FP64
    zero, st0 = oracle_lib.eval_batch(ProblemBatch.from_problems(probs, [None] * len(probs)))
    for x, z, p in zip(a, zero, probs):
        lo, hi = search.DEFAULT_BOUNDS[p.kind]
        assert x["robustness"] >= z - 1e-9 * max(abs(z), 1.0)
        assert all(h2 >= h1 for h1, h2 in zip(x["history"], x["history"][1:])) and x["history"][-1] == x["robustness"]
        assert (x["knot_y"] >= lo).all() and (x["knot_y"] <= hi).all()
    # the device engine is as good a searcher as the host one it replaces (same budget, other candidates)
    host = search.search_controls(probs, seed=3, workers=2, engine="host", **kw)
    dev_r, host_r = np.array([x["robustness"] for x in a]), np.array([x["robustness"] for x in host])
    assert all(x["engine"] == "host" for x in host)
    scale = np.maximum(np.abs(host_r), 1.0)
    assert np.median((dev_r - host_r) / scale) >= -0.05, (dev_r, host_r)


def test_device_search_on_an_unusable_specification_and_bad_arguments():
    from pde_controller_b200 import search
    p = problem_of(GOLD[0])
    spec, kt = _spec_of(p)
    ev = native.ControlsEvaluator(spec, kt, capacity=512)
    try:
        with pytest.raises(native.NativeError):
            ev.search(1024, 2, 0.05, 0.0, 1.0, seed=1)            # population above the capacity
        with pytest.raises(native.NativeError):
            ev.search(256, 0, 0.05, 0.0, 1.0, seed=1)
        with pytest.raises(native.NativeError):
            ev.search(256, 2, 0.05, 1.0, 0.0, seed=1)             # lo > hi
        bad = problem_of(GOLD[0])
        bad.T = -1.0
        spec_bad = ProblemBatch.from_problems([bad], kmax=max(kt.size, 1))
        ev.respec(spec_bad, kt)
        res = ev.search(256, 2, 0.05, 0.0, 1.0e6, seed=1)
        assert res["knot_y"] is None and res["robustness"] == -np.inf
        ev.respec(spec, kt)                                        # and the handle is usable again
        assert ev.search(256, 2, 0.05, 0.0, 1.0e6, seed=1)["knot_y"] is not None
    finally:
        ev.close()
    out = search.search_controls([bad], population=256, generations=2)
    assert out[0]["knot_y"] is None


def test_search_through_the_drop_in_call_is_the_batch_search(oracle_lib):
    """eval_robustness_batch(controls="search") = ONE pdeb_controls_search_batch call over the samples without a control;
    sample i draws from the generator stream of index i, so the list API, the batch API and the drop-in call agree."""
    from pde_controller_b200 import eval_script as es, search
    src = synth.make_batch(96, "mixed", seed=8)
    probs = [src.get_problem(i)[0] for i in range(96)]
    texts = [render_program(p) for p in probs]
    budget = {"population": 1024, "generations": 3, "seed": 11}
    given = {4: (np.array([0.0, 1.0]), np.array([2.0e5, 3.0e5])), 50: ([0.0, 0.5], [10.0, -10.0])}
    ctrl = [given.get(i) for i in range(96)]
    rob, _, st, used = es.eval_robustness_batch(texts, controls=ctrl, return_status=True, return_controls=True, search=budget)
    lst = search.search_controls(probs, **budget)
    assert (st == 0).all()
    for i in range(96):
        if i in given:
            assert np.array_equal(np.asarray(used[i][1], dtype=float), np.asarray(given[i][1], dtype=float))
        else:
            assert np.array_equal(used[i][1], lst[i]["knot_y"]) and np.array_equal(used[i][0], lst[i]["knot_t"])
            assert abs(rob[i] - lst[i]["robustness"]) <= 1e-9 * max(abs(rob[i]), 1.0)    # controls family vs batch kernels
    b = ProblemBatch.from_problems(probs, used)
    ref, ref_st = oracle_lib.eval_batch(b)
    assert (ref_st == 0).all() and rel_err(np.array(rob), ref, b).max() <= TOL_FP64
    # the thread count of the library's pool does not change anything
    one = search.search_controls(probs, workers=1, **budget)
    assert [x["robustness"] for x in one] == [x["robustness"] for x in lst]


def test_drop_in_call_sharded_over_devices_equals_one_device(tmp_path):
    """eval_robustness_batch(devices=[...]) / simulate_gt --devices: cost-balanced evaluation shards (pdeb_eval_batch_multi) and
    the searches dealt out over the devices; a sample's search depends on its index alone, so the numbers are those of one
    device.  Two real devices when the box has them, the same device twice otherwise."""
    from pde_controller_b200 import eval_script as es, simulate_gt
    devs = [0, 1] if native.device_count() >= 2 else [0, 0]
    src = synth.make_batch(80, "mixed", seed=13)
    texts = [render_program(src.get_problem(i)[0]) for i in range(80)]
    budget = {"population": 512, "generations": 3, "seed": 2}
    one, _, st1, c1 = es.eval_robustness_batch(texts, controls="search", return_status=True, return_controls=True, search=budget)
    two, _, st2, c2 = es.eval_robustness_batch(texts, controls="search", return_status=True, return_controls=True, search=budget,
                                               devices=devs)
    assert one == two and (st1 == st2).all()
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(c1, c2))
    z1, _ = es.eval_robustness_batch(texts, controls="zero")
    old_min = es._MULTI_EVAL_MIN
    es._MULTI_EVAL_MIN = 16                     # small batches stay on one device by default: force the sharded evaluation
    try:
        z2, _ = es.eval_robustness_batch(texts, controls="zero", devices=devs)
        two_b, _ = es.eval_robustness_batch(texts, controls="search", search=budget, devices=devs)
    finally:
        es._MULTI_EVAL_MIN = old_min
    assert z1 == z2 and two_b == one
    assert simulate_gt._device_list("all") == list(range(native.device_count()))
    assert simulate_gt._device_list("0,0") == [0, 0] and simulate_gt._device_list("1") == [0] and simulate_gt._device_list(None) is None
    d = tmp_path / "set"
    d.mkdir()
    with open(d / "test.jsonl", "w") as f:
        for i, t in enumerate(texts[:24]):
            f.write(json.dumps({"nl": f"s{i}", "sstl": "", "python": t}) + "\n")
    base = ["--data_dir", str(d), "--max_num_examples", "24", "--eval_robustness", "--search_population", "512", "--search_generations", "3"]
    m1 = simulate_gt.main(simulate_gt.build_parser().parse_args(base))
    r1 = [json.loads(l)["robustness"] for l in open(str(d) + "_24/test.jsonl")]
    m2 = simulate_gt.main(simulate_gt.build_parser().parse_args(base + ["--devices", ",".join(map(str, devs))]))
    r2 = [json.loads(l)["robustness"] for l in open(str(d) + "_24/test.jsonl")]
    assert r1 == r2 and m1 == m2


def test_control_arrays_with_missing_entries_are_searched_like_the_list_form(oracle_lib):
    from pde_controller_b200 import eval_script as es
    src = synth.make_batch(48, "heat", seed=17)
    probs, ctrls = zip(*(src.get_problem(i) for i in range(48)))
    texts = [render_program(p).encode() for p in probs]                 # programs as bytes, for good measure
    ctrls = [None if i % 3 == 0 else c for i, c in enumerate(ctrls)]
    kmax = max(len(c[0]) for c in ctrls if c is not None)
    T, Y, K = np.zeros((48, kmax)), np.zeros((48, kmax)), np.full(48, -1)
    for i, c in enumerate(ctrls):
        if c is not None:
            K[i] = len(c[0]); T[i, : K[i]] = c[0]; Y[i, : K[i]] = c[1]
    budget = {"population": 512, "generations": 3, "seed": 9}
    a = es.eval_robustness_batch(texts, controls=ctrls, return_status=True, return_controls=True, search=budget)
    b = es.eval_robustness_batch(texts, control_arrays=(T, Y, K), return_status=True, return_controls=True, search=budget)
    assert a[0] == b[0] and (a[2] == 0).all() and (b[2] == 0).all()
    assert all(np.array_equal(x[1], y[1]) and np.array_equal(x[0], y[0]) for x, y in zip(a[3], b[3]))
    batch = ProblemBatch.from_problems(list(probs), b[3])
    ref, st = oracle_lib.eval_batch(batch)
    assert (st == 0).all() and rel_err(np.array(b[0]), ref, batch).max() <= TOL_FP64


def test_sweep_is_deterministic_under_heavy_concurrency():
    """32 library threads x 3 searches in flight each (more streams than hardware queues, handles created while other
    searches run): every result equals the single-thread run.  Regression for the ticket counter of the fused generation
    kernel, which a default-stream memset left unordered against the first search of a new handle."""
    from pde_controller_b200 import search
    src = synth.make_batch(1536, "mixed", seed=31)
    probs = [src.get_problem(i)[0] for i in range(1536)]
    kw = dict(population=1024, generations=4, seed=5)
    ref = search.search_controls(probs, workers=1, **kw)
    for _ in range(2):
        got = search.search_controls(probs, workers=32, **kw)
        assert [x["robustness"] for x in got] == [x["robustness"] for x in ref]
        assert all(np.array_equal(a["knot_y"], b["knot_y"]) for a, b in zip(got, ref))
    native.controls_pool_clear()                 # the handles the sweeps left for the next one; a sweep after it starts afresh
    again = search.search_controls(probs[:64], workers=4, **kw)
    assert [x["robustness"] for x in again] == [x["robustness"] for x in ref[:64]]
