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
