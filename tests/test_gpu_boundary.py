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


def test_packed_upload_equals_planar_upload(oracle_lib):
    """pdeb_resident_upload_packed: the ragged transfer form unpacked on the device gives bit-identical results, for
    every list shape (empty lists, eight clauses, explicit and uniform knots, sizes around the scan block of 1,024)."""
    import numpy as np
    from pde_controller_b200 import PackedBatch, native, synth
    from test_gpu_parity import _stress_batch, _wide_batch
    cases = [synth.make_batch(n, wl, seed=5, uniform_knots=uk)
             for wl, uk, n in (("heat", True, 65536), ("wave", False, 3000), ("mixed", False, 1025), ("heat", False, 1),
                               ("heat", True, 1023), ("wave", True, 1024))]
    cases += [_wide_batch("heat", 2500, 43), _stress_batch("wave", 2048, 44), synth.make_batch(40, "heat_refined", seed=2)]
    for b in cases:
        b.n_knots[: min(7, b.n)] = 0
        ref = native.evaluate_batch(b, want_clauses=True)
        res = native.ResidentBatch(b)
        pk = native.pinned_packed(b.n, b.kmax, b.pmax, b.tmax, uniform_knots=b.uniform_knots).pack(b)
        # the image still holds the planar upload: overwrite it with a different batch first, then restore it packed
        other = synth.make_batch(b.n, "heat" if int(b.n_elem[0]) == 30 else ("wave" if int(b.n_elem[0]) == 20 else "heat_refined"),
                                 seed=99, uniform_knots=b.uniform_knots)
        if (other.kmax, other.pmax, other.tmax) == (b.kmax, b.pmax, b.tmax):
            res.upload(other)
        res.upload_packed(pk)
        res.eval()
        got = res.fetch(want_clauses=True)
        res.close()
        assert np.array_equal(got["status"], ref["status"])
        assert np.array_equal(got["robustness"], ref["robustness"])
        live = np.arange(b.pmax)[:, None] < b.n_pred[None, :]
        assert np.array_equal(got["clause"][live], ref["clause"][live])


def test_packed_encodings_on_device_are_bit_exact():
    """Arrays stored as float / as one value / as f64 (with -0.0, an off-float value) decode to the same image bits:
    results equal the planar upload bitwise, per clause."""
    import numpy as np
    from pde_controller_b200 import abi, native, synth
    b = synth.make_batch(4096, "heat", seed=12)
    b.g0[::7] = np.nextafter(b.g0[::7], 1e9)           # g0 is no longer float-exact: stays f64
    b.x_a[5] = -0.0
    b.E_in[:] = 1.0e6                                   # one value for all
    pk = native.pinned_packed(b.n, b.kmax, b.pmax, b.tmax).pack(b)
    assert pk.enc["g0"] == abi.ENC_F64 and pk.enc["E_in"] == abi.ENC_ONE and pk.enc["L"] == abi.ENC_F32
    ref = native.evaluate_batch(b, want_clauses=True)
    res = native.ResidentBatch(b)
    res.upload(synth.make_batch(b.n, "heat", seed=77))  # something else in the image first
    res.upload_packed(pk)
    res.eval()
    got = res.fetch(want_clauses=True)
    res.close()
    assert np.array_equal(got["robustness"], ref["robustness"]) and np.array_equal(got["status"], ref["status"])
    live = np.arange(b.pmax)[:, None] < b.n_pred[None, :]
    assert np.array_equal(got["clause"][live], ref["clause"][live])


def test_packed_upload_rejects_inconsistent_totals():
    import ctypes as C
    from pde_controller_b200 import native, synth
    b = synth.make_batch(300, "heat", seed=6)
    res = native.ResidentBatch(b)
    pk = native.pinned_packed(b.n, b.kmax, b.pmax, b.tmax).pack(b)
    cp = pk.as_c()
    cp.total_knots -= 1
    rc = native.lib().pdeb_resident_upload_packed(res._h, C.addressof(cp), None)
    assert rc == -1 and b"totals" in native.lib().pdeb_last_error()
    res.close()


def test_block_pipeline_packed_matches_planar():
    import numpy as np
    from pde_controller_b200 import native, synth
    b = synth.make_batch(70001, "mixed", seed=8)
    want = native.evaluate_batch(b)
    for packed in (False, True):
        pipe = native.BlockPipeline(b.n, b.kmax, b.pmax, b.tmax, block=16384, depth=3, packed=packed)
        for g in range(3):
            pipe.fill(b, g)
            pipe.submit(g)
        for g in range(3):
            rob, st = pipe.wait(g)
            assert np.array_equal(rob, want["robustness"]) and np.array_equal(st, want["status"])
        if packed:
            assert pipe.h2d_bytes < 0.8 * sum(x.slab_bytes for x in pipe.blocks[0])
        pipe.close()


@pytest.mark.gpu
def test_device_pci_bus_id_and_host_binding():
    """pdeb_device_pci_bus_id names the device the way sysfs does; bind_host_near never leaves the process without CPUs."""
    import os
    import re
    from pde_controller_b200 import native

    bus = native.device_pci_bus_id(0)
    assert re.fullmatch(r"[0-9a-f]{4,8}:[0-9a-f]{2}:[0-9a-f]{2}\.[0-9a-f]", bus), bus
    with pytest.raises(native.NativeError):
        native.device_pci_bus_id(native.device_count())
    before = os.sched_getaffinity(0)
    try:
        info = native.bind_host_near(0)
        assert info["bus_id"] == bus
        now = os.sched_getaffinity(0)
        assert now and now <= before and (info["bound"] or now == before)
    finally:
        os.sched_setaffinity(0, before)
