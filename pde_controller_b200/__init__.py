"""B200-native utility evaluation for PDE-Controller: 1D heat / wave FEM simulation + STL robustness.

Public surface (mirrors the reference's test/PDEcontrol/evaluation/eval/eval_script.py:43-102):
    eval_robustness_helper, eval_robustness, eval_robustness_gt, eval_robustness_DPO
plus the batched forms the GPU path is built for:
    eval_robustness_batch, ProblemBatch, evaluate_batch
"""
from .abi import default_conventions  # noqa: F401
from .problem import PackedBatch, Predicate, Problem, ProblemBatch, default_knot_times  # noqa: F401

__all__ = ["PackedBatch", "Predicate", "Problem", "ProblemBatch", "default_knot_times", "default_conventions"]
