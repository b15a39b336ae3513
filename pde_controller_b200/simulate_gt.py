"""Batched ground-truth labelling: the reference's simulate_gt.py driver with ONE GPU call instead of a process pool.

Mirrors test/PDEcontrol/evaluation/infer/simulate_gt.py:92-191 of the reference: same CLI flags (the ones that
concern this path), same sampling (`random.seed(42)`, `random.sample`), same output files
(`{data_dir}_{max_num_examples}/test.jsonl` with `robustness` and `time` added, `metrics.json` with the
three rates).  `evaluate_future` (simulate_gt.py:62-89: one future per sample, 300 s timeout) is replaced by
`eval_robustness_batch`.  A sample whose program cannot be evaluated is labelled -100, exactly what
simulate_gt.py:55-59 / :144-147 count as "failed"; there are no timeouts any more.

Controls: the reference's labels are the robustness under the MILP-synthesised input (out of scope here).  A record
that carries an input in the optional `control_key` field ({"t": [...], "y": [...]}) is evaluated under it; every other
record under `--control`: "search" (default: the sampling surrogate of the synthesis, search.search_controls; the
label is the robustness under the input found, and the input is written back into the record) or "zero" (zero
input, on request).  The mode is recorded in metrics.json ("control_mode").
"""
from __future__ import annotations

import argparse
import json
import os
import random

from .eval_script import eval_robustness_batch


def evaluate_future(eval_fn, tasks, _timeout=300, python_key="python", nl_key="nl", control_key="control",
                    control_mode=None, write_back=False, search=None, devices=None):
    """Drop-in for simulate_gt.evaluate_future: -> (robustness tuple, time tuple, timeout_cnt).

    `eval_fn` is accepted for signature compatibility and ignored: the whole task list goes to the GPU at once.
    control_mode: "search" | "zero" | None (= eval_script's process-wide mode) for the records without an input of
    their own; write_back stores the input a record was evaluated under in task[control_key]; devices: the GPUs of the
    box to shard over (the reference's pool of cpu_count // 2 worker processes, simulate_gt.py:64-65, becomes one shard
    per device).
    """
    from . import eval_script
    controls = []
    for t in tasks:
        c = t.get(control_key)
        controls.append((c["t"], c["y"]) if isinstance(c, dict) and "t" in c and "y" in c else None)
    mode = control_mode or eval_script.control_mode()
    if mode not in eval_script.CONTROL_MODES:
        raise ValueError(f"control mode must be one of {eval_script.CONTROL_MODES}")
    prev = eval_script.control_mode()
    eval_script.configure(eval_script._DEVICE, eval_script._PRECISION, eval_script._CONV, control_mode=mode,
                          devices=eval_script._DEVICES)
    try:
        rob, rt, used = eval_robustness_batch([t[python_key] for t in tasks], [t.get(nl_key, "") for t in tasks], controls,
                                              return_controls=True, search=search, devices=devices)
    finally:
        eval_script.configure(eval_script._DEVICE, eval_script._PRECISION, eval_script._CONV, control_mode=prev,
                              devices=eval_script._DEVICES)
    if write_back:
        for t, c in zip(tasks, used):
            if c is not None and control_key not in t:
                t[control_key] = {"t": [float(x) for x in c[0]], "y": [float(x) for x in c[1]]}
    return tuple(rob), tuple(rt), 0


def _device_list(spec):
    """--devices: "all", a count ("4" = devices 0..3) or a comma list ("0,2,5"); None / "" = the configured device."""
    if not spec:
        return None
    from . import native
    if spec == "all":
        return list(range(native.device_count()))
    if "," in spec:
        return [int(x) for x in spec.split(",") if x.strip()]
    return list(range(int(spec)))


def main(args):
    random.seed(42)
    test_data = []
    fname = "validation.jsonl" if args.infer_on_train_set else "test.jsonl"
    with open(os.path.join(args.data_dir, fname)) as fin:
        for line in fin:
            example = json.loads(line)
            example["python"] = example[args.python_key].strip()
            example["sstl"] = example.get(args.stl_key)
            example["nl"] = example.get(args.nl_key, "")
            robustness = example.get(args.robustness_key, None)
            if robustness is not None:
                example["robustness"] = robustness
            test_data.append(example)
    if args.max_num_examples and len(test_data) > args.max_num_examples:
        test_data = random.sample(test_data, args.max_num_examples)
    if not test_data:
        print("Ending. There was no data to test.")
        return None
    args.save_dir = os.path.join(args.data_dir + "_" + str(args.max_num_examples))
    os.makedirs(args.save_dir, exist_ok=True)

    metrics = {"n_samples": len(test_data)}
    if args.eval_robustness:
        rob, times, timeouts = evaluate_future(None, test_data, control_key=args.control_key, control_mode=args.control,
                                               write_back=True, devices=_device_list(args.devices),
                                               search={"population": args.search_population, "generations": args.search_generations})
        metrics["control_mode"] = args.control
        for item, r, t in zip(test_data, rob, times):
            if "robustness" not in item:
                item["robustness"] = r
                item["time"] = t
        pos = sum(1 for it in test_data if it["robustness"] > 0)
        failed = sum(1 for it in test_data if it["robustness"] == -100)
        neg = sum(1 for it in test_data if it["robustness"] < 0 and it["robustness"] != -100)
        if pos + neg + failed != len(test_data):
            bad = [it["robustness"] for it in test_data if it["robustness"] == 0]
            raise ValueError(f"gt_r = {bad[0] if bad else 'nan'}")       # simulate_gt.py:148-149: exact 0 is an error
        metrics["gt positive robustness rate"] = pos / len(test_data)
        metrics["gt negative robustness rate"] = neg / len(test_data)
        metrics["gt failed robustness rate"] = failed / len(test_data)
        print(f"Timeout count >>> output eval robustness = {timeouts}", flush=True)

    with open(os.path.join(args.save_dir, fname), "w") as fout:
        for item in test_data:
            for key in ("sstl", "python", "robustness", "time"):       # same key order as simulate_gt.py:167-178
                if key in item:
                    item[key] = item.pop(key)
            json.dump(item, fout, ensure_ascii=True)
            fout.write("\n")
    with open(os.path.join(args.save_dir, "metrics.json"), "w") as fout:
        json.dump(metrics, fout, indent=4)
    return metrics


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--data_dir", type=str, required=True)
    p.add_argument("--python_key", type=str, default="python")
    p.add_argument("--stl_key", type=str, default="sstl")
    p.add_argument("--nl_key", type=str, default="nl")
    p.add_argument("--robustness_key", type=str, default="robustness")
    p.add_argument("--control_key", type=str, default="control")
    p.add_argument("--control", type=str, default="search", choices=["search", "zero"],
                   help="what a record without an input of its own is evaluated under (module docstring)")
    p.add_argument("--devices", type=str, default=None,
                   help='GPUs of the box to shard the batch over: "all", a count, or a comma list (default: one device)')
    p.add_argument("--search_population", type=int, default=4096)
    p.add_argument("--search_generations", type=int, default=6)
    p.add_argument("--max_num_examples", type=int, default=None)
    p.add_argument("--infer_on_train_set", action="store_true")
    p.add_argument("--eval_robustness", action="store_true")
    p.add_argument("--gpus", type=str, default=None)
    return p


if __name__ == "__main__":
    a, _ = build_parser().parse_known_args()
    if a.gpus is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = a.gpus
    main(a)
