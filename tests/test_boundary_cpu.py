"""Host-side contract of the drop-in boundary that can be checked without a GPU."""
import os
import subprocess
import sys

import pytest

from pde_controller_b200 import eval_script as es, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_control_mode_is_explicit_and_validated():
    assert es.control_mode() == "search"          # the reference's number is under a synthesised input, never zero input
    with pytest.raises(ValueError):
        es.configure(control_mode="given")
    with pytest.raises(ValueError):
        es.eval_robustness_batch(["N = 1"], controls="none")
    with pytest.raises(ValueError):
        es.eval_robustness_batch(["N = 1", "N = 2"], controls=[None])


def test_unparsable_programs_never_reach_the_device():
    rob, rt, st = es.eval_robustness_batch(["not a program", ""], return_status=True)
    assert rob == [-100.0, -100.0] and list(st) == [1, 1] and len(rt) == 2


def test_wrapper_without_a_device_fails_loudly_and_prints_nothing_on_stdout():
    """eval_script.py:63-65 of the reference turns a non-zero exit into RuntimeError; a CPU fallback does not exist."""
    if native.device_count() > 0:
        pytest.skip("a CUDA device is present")
    from test_frontend import HEAT
    exe = os.path.join(ROOT, "pde_controller_b200", "eval_robustness_wrapper.py")
    p = subprocess.run([sys.executable, exe, "nl", HEAT], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, PDEC_CONTROL_MODE="zero"))
    assert p.returncode != 0 and p.stdout == "" and "NativeError" in p.stderr
    p = subprocess.run([sys.executable, exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 2 and p.stdout == ""


def test_bind_host_near_uses_the_gpu_numa_node(tmp_path, monkeypatch):
    """native.bind_host_near: PCI bus id -> sysfs numa_node -> cpulist of that node -> sched_setaffinity; nothing changes
    on boxes that hide the information (node -1, no sysfs entry, no device)."""
    import os
    from pde_controller_b200 import native

    assert native._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert native._cpulist("") == set()
    before = os.sched_getaffinity(0)
    keep = sorted(before)[: max(1, len(before) // 2)]
    bus = "0000:1b:00.0"
    dev_dir = tmp_path / "bus" / "pci" / "devices" / bus
    dev_dir.mkdir(parents=True)
    node_dir = tmp_path / "devices" / "system" / "node" / "node1"
    node_dir.mkdir(parents=True)
    (node_dir / "cpulist").write_text(",".join(str(c) for c in keep) + ",4093-4095\n")    # CPUs outside the process's set are dropped
    monkeypatch.setattr(native, "device_pci_bus_id", lambda device: bus)
    try:
        (dev_dir / "numa_node").write_text("-1\n")
        info = native.bind_host_near(0, sysfs=str(tmp_path))
        assert info == {"bus_id": bus, "node": -1, "cpus": 0, "bound": False} and os.sched_getaffinity(0) == before
        (dev_dir / "numa_node").write_text("1\n")
        info = native.bind_host_near(0, sysfs=str(tmp_path))
        assert info["bound"] and info["node"] == 1 and info["cpus"] == len(keep)
        assert os.sched_getaffinity(0) == set(keep)
        os.sched_setaffinity(0, before)
        info = native.bind_host_near(0, sysfs=str(tmp_path / "absent"))
        assert not info["bound"] and os.sched_getaffinity(0) == before
    finally:
        os.sched_setaffinity(0, before)


def test_default_knots_vectorised_equal_the_per_problem_rule():
    """problem.fill_default_knot_times / default_knot_counts (what the batch search and eval_robustness_batch use) against
    default_knot_times, sample by sample, including horizons that round half to even and a non-positive one."""
    import numpy as np
    from pde_controller_b200 import ProblemBatch, default_knot_times, synth
    from pde_controller_b200.problem import default_knot_counts, fill_default_knot_times

    for wl in ("heat", "wave"):
        b = synth.make_batch(200, wl, seed=2)
        b.T[:6] = [6.5, 7.5, 8.49999, 0.2, 5.5, 13.5]
        w = ProblemBatch(200, 16, b.pmax, b.tmax)
        for name in w._names:
            if name not in ("knot_t", "knot_y", "n_knots"):
                getattr(w, name)[...] = getattr(b, name)
        w.knot_y[:] = 7.0
        idx = np.arange(0, 200, 3)
        K = fill_default_knot_times(w, idx)
        for q, i in enumerate(idx):
            kt = default_knot_times(b.get_problem(int(i))[0])
            assert K[q] == kt.size == w.n_knots[i] and np.array_equal(w.knot_t[: kt.size, i], kt)
            assert not w.knot_t[kt.size:, i].any() and not w.knot_y[:, i].any()
        untouched = np.setdiff1d(np.arange(200), idx)
        assert (w.n_knots[untouched] == 0).all() and (w.knot_y[:, untouched] == 7.0).all()
    k = default_knot_counts(np.array([0, 0, 0, 1], dtype=np.uint8), np.array([-3.0, np.nan, 1e30, 123.0]))
    assert k[0] == 2 and k[1] == 2 and k[2] > 32 and k[3] == 5


def test_search_seeds_are_splitmix64():
    """search.sample_seed mirrors pdeb_mix64 of csrc/pdeb_controls.cu: the first outputs of splitmix64 are published values."""
    from pde_controller_b200 import search
    assert search._mix64(0) == 0xE220A8397B1DCDAF
    assert search._mix64(0x9E3779B97F4A7C15) == 0x6E789E6AA1B965F4          # second output of the generator seeded with 0
    assert search._seed64(5) == 5 and search._seed64(-1) == (1 << 64) - 1 and 0 <= search._seed64((3, 4)) < (1 << 64)
    assert search.sample_seed(7, 0) != search.sample_seed(7, 1) != search.sample_seed(8, 1)
    assert search.sample_seed(7, 3) == search._mix64(7 ^ search._mix64(3))
