#!/usr/bin/env python
"""Executable twin of the reference's test/PDEcontrol/evaluation/eval/eval_robustness_wrapper.py:13-21.

    python eval_robustness_wrapper.py <nl_in_prompt> <llm_output>

prints exactly ONE JSON object {"robustness": r, "runtime": t} on stdout (the reference's caller parses the whole of
stdout with json.loads, eval_script.py:67-72) and exits 0; on failure it prints nothing on stdout, the reason on
stderr, and exits non-zero (eval_script.py:63-65 raises RuntimeError on that).  Where the reference's file imports
`control.femformal.eval_robustness` (py2 + Gurobi), this one calls the CUDA path; there is no CPU fallback.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from pde_controller_b200.eval_script import wrapper_main
    return wrapper_main(sys.argv[1:])


if __name__ == "__main__":
    sys.exit(main())
