"""B200-native utility evaluation for PDE-Controller: 1D heat / wave FEM simulation + STL robustness.

Public surface (mirrors the reference's test/PDEcontrol/evaluation/eval/eval_script.py:43-102):
    eval_robustness_helper, eval_robustness, eval_robustness_gt, eval_robustness_DPO
plus the batched forms the GPU path is built for:
    eval_robustness_batch, ProblemBatch, evaluate_batch
"""
import os as _os

# A sweep of searches (pdeb_controls_search_batch) keeps about 24 streams busy from its host threads; with CUDA's default of 8
# hardware queues their streams alias and serialise.  Only has an effect if this import precedes the first CUDA call of the
# process, and never overrides the caller's own setting.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .abi import default_conventions  # noqa: E402,F401
from .problem import PackedBatch, Predicate, Problem, ProblemBatch, default_knot_times  # noqa: E402,F401

__all__ = ["PackedBatch", "Predicate", "Problem", "ProblemBatch", "default_knot_times", "default_conventions"]
