#!/usr/bin/env python
"""bench.py — heat/wave 1D simulate + STL-robustness evaluations per second on B200.

One "step" = one pass of the hot path over one batch of synthetic samples.
Default workload = BASELINE.json configs[1]: 1D heat, 65,536 synthetic specs/controls at the
default mesh (N=30, dt=0.05), FP64, one B200.  Under torchrun every rank evaluates its own shard
of the same size (weak scaling, no data-path collective; optional NCCL gather of the results).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload heat|wave|mixed|heat_refined]
                  [--batch B] [--precision 64|32] [--gather] [--impl reference]

Prints ONE JSON line (rank 0).  --impl reference times the CPU oracle (the reference's own
femformal path is not in its snapshot and cannot run here) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "heat/wave 1D sim+STL-robustness evals/sec"
UNIT = "evals/s"
WORKLOAD_TEXT = {
    "heat": "1D heat, default mesh N=30 dt=0.05, synthetic specs/controls, fused STL robustness",
    "wave": "1D wave (Newmark beta=1/4 gamma=1/2), default mesh N=20 dt=0.0025, synthetic specs, fused STL robustness",
    "mixed": "mixed 50% heat (N=30) + 50% wave (N=20) synthetic eval set, sorted by kind",
    "heat_refined": "1D heat, refined mesh N=1024, 10,000 steps, synthetic specs/controls",
    "controls_heat": "ONE 1D heat spec (default mesh) against many candidate PWL controls, arg-max on the device (search mode)",
    "controls_wave": "ONE 1D wave spec (default mesh) against many candidate PWL controls, arg-max on the device (search mode)",
}
DEFAULT_BATCH = {"heat": 65536, "wave": 65536, "mixed": 65536, "heat_refined": 4096, "controls_heat": 1048576,
                 "controls_wave": 1048576}


L2_NOTE = "GPU arm: flushed (256 MiB write) between timed steps; CPU arm: not applicable"


def common_config(args, world):
    """The `config` object is THE SAME in both arms (same keys, same values): what the driver compares."""
    return {"workload": WORKLOAD_TEXT[args.workload], "batch_per_gpu": args.batch, "global_batch": args.batch * world,
            "precision": args.precision, "seed": args.seed, "l2": L2_NOTE}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_profile():
    """Numbers measured on B200 by the tools of this repo and committed under profiles/ (never invented here):
    fp64 peak from tools/probe/fp64_probe.cu, per-launch DRAM traffic of the dominant kernel from one
    `ncu --set full` capture of this same command (B200_PROFILING.md recipe)."""
    p = os.path.join(ROOT, "profiles", "measured_b200.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def fp64_view(batch, copies, kernel_ms, prof, wave_form="a"):
    """The compute view of the path (it is FP64-latency/issue bound, not HBM bound): executed FP64 flops (FMA = 2) and
    FP64 instructions of `copies` evaluations of `batch` per launch against the measured DFMA rate of the GPU."""
    flops, instr = batch.algorithmic_flops("executed", wave_form=wave_form)
    survey, _ = batch.algorithmic_flops("survey")
    peak = prof.get("fp64_peak_tflops")
    ach = flops * copies / (kernel_ms * 1e-3) / 1e12
    out = {"bound": "fp64 pipe", "achieved": ach, "unit": "TFLOP/s", "peak": peak, "frac": ach / peak if peak else None,
           "pipe_slots_frac": (instr * copies / (kernel_ms * 1e-3)) / (peak * 1e12 / 2.0) if peak else None,
           "peak_source": "measured DFMA rate, tools/probe/fp64_probe.cu (profiles/measured_b200.json)" if peak else None,
           "executed_flops_per_launch": flops * copies, "survey_convention_flops_per_launch": survey * copies,
           "note": "flops of the algorithm as implemented (DESIGN.md 4: 8 per row and step for heat, 20 for waves (15 in the "
                   "momentum form of the controls family), 4 per clause node); pipe_slots_frac counts FP64 instructions, an add or multiply holds the pipe like an FMA"}
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                "hw_power_brake_slowdown": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
            }
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop_evt.is_set():
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(get(h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report, do not invent numbers
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.error:
            out["error"] = self.error
        return out


def nvml_index(local_rank):
    """NVML index of this rank's GPU (CUDA_VISIBLE_DEVICES may renumber the CUDA ordinals)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    ids = [v.strip() for v in vis.split(",") if v.strip()]
    if local_rank < len(ids) and ids[local_rank].isdigit():
        return int(ids[local_rank])
    return local_rank


def run_reference(args, rank, world):
    """CPU arm: the FP64 oracle (C, OpenMP over samples) on all host cores, same workload/metric."""
    if rank != 0:
        return
    import oracle
    from pde_controller_b200 import synth
    oracle.build()
    nthreads = os.cpu_count() or 1
    cfg = common_config(args, world)          # before the bounded sample below may shrink args.batch
    sample_note = f"{args.batch} samples per step"
    if args.workload.startswith("controls_"):
        # bounded sample of the same workload: the first 65,536 candidate controls of rank 0 for the same specification
        wl = args.workload.split("_")[1]
        spec = synth.make_batch(8, wl, seed=args.seed)
        nk = int(spec.n_knots[3])
        lo, hi = (0.0, 1e6) if wl == "heat" else (-5e3, 5e3)
        m = min(args.batch, 65536)
        y = np.random.Generator(np.random.PCG64([args.seed, 0])).uniform(lo, hi, size=(nk, args.batch))[:, :m]
        batch = synth.replicate_spec(spec, 3, np.linspace(0.0, float(spec.T[3]), nk), y)
        sample_note = f"{m} of the {args.batch} candidate controls per step"
        args.batch = m
    else:
        batch = synth.make_batch(args.batch, args.workload, seed=args.seed, total=args.batch)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        oracle.eval_batch(batch, nthreads=nthreads)
        t1 = time.perf_counter()
        if i >= args.warmup:
            times.append(t1 - t0)
    total = sum(times)
    value = args.batch * len(times) / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": cfg,
        "detail": {"note": "reference femformal (Python 2 + Gurobi) is absent from the snapshot; this is the CPU oracle port "
                           "(oracle/pde_oracle.c, OpenMP over samples); like the kernels it simulates up to the last "
                           "observed step of each sample: both arms do the same stepping work",
                   "samples_per_step": args.batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": f"{sample_note}, {len(times)} steps, OpenMP over samples, stops at the last observed step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, batch, gpu_robustness=None, gpu_status=None):
    """Bounded CPU sample on rank 0 at N=1: the oracle over the whole batch, best of 2.  The oracle is the checker: its
    results for this very batch are compared with the ones the GPU produced in the timed runs (`parity`)."""
    import oracle
    oracle.build()
    nthreads = os.cpu_count() or 1
    best = None
    ref = ref_st = None
    for _ in range(2):
        t0 = time.perf_counter()
        ref, ref_st = oracle.eval_batch(batch, nthreads=nthreads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out = {"value": batch.n / best, "unit": UNIT, "cores": nthreads, "kind": "port",
           "sample": f"the same {batch.n}-sample batch, best of 2 passes, C oracle -O3 + OpenMP over samples, each sample "
                     "simulated up to its last observed step (the work the kernels do)"}
    if gpu_robustness is not None and args.precision == 64:
        # the parity metric of tests/parity.py: |gpu - oracle| / max(|oracle|, 1e-3 * field scale); pure relative next to it
        peak = np.abs(batch.knot_y).max(axis=0) if batch.knot_y.shape[0] else 0.0
        field = np.abs(batch.g0) + 1.0 + peak * batch.L / np.minimum(batch.E_in, batch.E_out)
        err = np.abs(gpu_robustness - ref) / np.maximum(np.abs(ref), 1e-3 * field)
        big = np.abs(ref) >= 1e-6
        out["parity"] = {"samples": int(batch.n), "status_equal": bool((gpu_status == ref_st).all()),
                         "max_rel_err_floored": float(err.max()), "tolerance": 1e-9,
                         "max_rel_err_pure": float((np.abs(gpu_robustness - ref)[big] / np.abs(ref)[big]).max()) if big.any() else 0.0,
                         "note": "GPU results of the timed runs against the oracle on the same batch; the oracle itself is unpinned (DESIGN.md 2)"}
    out["context"] = cpu_context(batch, nthreads)
    return out


def cpu_context(batch, nthreads):
    """The two context figures of BASELINE.md section 3 (neither is the baseline): the numpy restatement on one core
    (closest in spirit to femformal's numpy time loops) and the ceiling of the reference's process-per-sample design
    (eval_script.py:48-59 spawns `conda run ... python` per sample from cpu_count // 2 workers, simulate_gt.py:64-65):
    the spawn rate of an EMPTY interpreter times that worker count."""
    import subprocess
    from oracle import pde_oracle_np
    ctx = {}
    try:
        # the restatement solves a DENSE system every step: bounded to the default meshes and to ~2 s
        if int(batch.n_elem[: min(batch.n, 12)].max(initial=0)) > 64:
            ctx["numpy_port_1core"] = {"skipped": "dense per-step solves: default meshes only"}
        else:
            k, t0 = 0, time.perf_counter()
            while k < min(batch.n, 12) and time.perf_counter() - t0 < 2.0:
                p, ctl = batch.get_problem(k)
                pde_oracle_np.evaluate(p, ctl)
                k += 1
            ctx["numpy_port_1core"] = {"value": k / (time.perf_counter() - t0), "unit": UNIT, "sample": f"first {k} samples"}
    except Exception as e:  # context only: never fails the bench
        ctx["numpy_port_1core"] = {"error": repr(e)}
    try:
        k = 10
        t0 = time.perf_counter()
        for _ in range(k):
            subprocess.run([sys.executable, "-c", "pass"], check=True)
        per = (time.perf_counter() - t0) / k
        workers = max(nthreads // 2, 1)
        ctx["process_per_sample_floor"] = {"value": workers / per, "unit": UNIT, "workers": workers,
                                           "sample": f"{k} spawns of an empty interpreter, {per * 1e3:.1f} ms each"}
    except Exception as e:
        ctx["process_per_sample_floor"] = {"error": repr(e)}
    return ctx


def run_controls(args, rank, local_rank, world):
    """Search mode: every rank scores its own args.batch candidate controls for the same specification."""
    import torch
    import torch.distributed as dist
    from pde_controller_b200 import native, synth
    wl = args.workload.split("_")[1]
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    spec = synth.make_batch(8, wl, seed=args.seed)
    s = 3
    T, nk = float(spec.T[s]), int(spec.n_knots[s])
    knot_t = np.linspace(0.0, T, nk)
    lo, hi = (0.0, 1e6) if wl == "heat" else (-5e3, 5e3)
    m = args.batch
    pin = native.PinnedBuffer(8 * nk * m)
    y = pin.array.view(np.float64).reshape(nk, m)
    y[:] = np.random.Generator(np.random.PCG64([args.seed, rank])).uniform(lo, hi, size=(nk, m))
    ev = native.ControlsEvaluator(spec, knot_t, capacity=m, sample=s, device=local_rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    first = ev.eval(y)
    for _ in range(args.warmup):
        ev.eval(None, m=m, want_robustness=False)
    sampler = ClockSampler(nvml_index(local_rank))
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    kernel_ms, launches = 0.0, 0
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        r = ev.eval(None, m=m, want_robustness=False)       # candidates resident in HBM; CUDA events inside the library
        kernel_ms += r["kernel_ms"]
        launches += r["launches"]
    barrier()
    wall1 = time.perf_counter()
    clocks = sampler.stop()
    t_dev = torch.tensor([kernel_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    total = m * world
    value = total * args.steps / (total_ms * 1e-3)
    # end to end: candidates from pinned host memory every step, robustness + status + arg-max back on the host.
    # The candidates go through TWO evaluators (halves) on two streams, each driven by its own host thread (the C call
    # releases the GIL), so the upload of one half overlaps the kernel of the other: a stream of candidate sets is
    # bound by max(copy, kernel) instead of their sum.
    L = native.lib()
    mh = (m + 1) // 2
    bounds = [(0, mh), (mh, m)] if m > 1 else [(0, m)]
    evs, pins, ys, streams, outs = [], [], [], [], []
    for lo_, hi_ in bounds:
        pb_ = native.PinnedBuffer(8 * nk * (hi_ - lo_))
        yy = pb_.array.view(np.float64).reshape(nk, hi_ - lo_)
        yy[:] = y[:, lo_:hi_]
        pr_, ps_ = native.PinnedBuffer(8 * (hi_ - lo_)), native.PinnedBuffer(4 * (hi_ - lo_))   # page-locked results
        pins += [pb_, pr_, ps_]; ys.append(yy)
        outs.append({"robustness": pr_.array.view(np.float64), "status": ps_.array.view(np.int32)})
        evs.append(native.ControlsEvaluator(spec, knot_t, capacity=hi_ - lo_, sample=s, device=local_rank))
        streams.append(L.pdeb_stream_create(local_rank))
    last = [None] * len(bounds)

    def half(i, reps):
        for _ in range(reps):
            last[i] = evs[i].eval(ys[i], stream=streams[i], out=outs[i])

    for i in range(len(bounds)):
        half(i, 1)
    barrier()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=half, args=(i, args.steps)) for i in range(len(bounds))]
    for t_ in ths:
        t_.start()
    for t_ in ths:
        t_.join()
    torch.cuda.synchronize()
    dt_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
    e2e = total * args.steps / float(dt_.item())
    e2e_rob = np.concatenate([x["robustness"] for x in last])
    if not np.array_equal(e2e_rob, first["robustness"]):
        raise SystemExit("bench: repeated evaluation of the same candidates differs")
    best = max(last, key=lambda x: x["best_value"])
    if best["best_value"] != first["best_value"]:
        raise SystemExit("bench: arg-max over the halves differs from the arg-max over all candidates")
    for e_, st_ in zip(evs, streams):
        e_.close()
        L.pdeb_stream_destroy(st_)
    if rank == 0:
        peak, peak_src = peaks()
        prof = measured_profile()
        steps_sim = int(max(np.floor(spec.pred_thi[: int(spec.n_pred[s]), s] / spec.dt[s] + 1e-9)))
        sub = synth.replicate_spec(spec, s, knot_t, np.zeros((nk, 1)))
        alg_bytes = m * (8 * nk + 12)                      # per candidate: its knot values in, robustness + status out
        kms = total_ms / args.steps
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": common_config(args, world),
            "detail": {"candidates_per_gpu": m, "knots": nk, "sim_steps": steps_sim, "n_pred": int(spec.n_pred[s]),
                       "e2e_overlap": "two halves of the candidates on two streams, one host thread each",
                       "wall_s_timed_region": wall1 - wall0},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(8 * nk * m), "d2h_bytes_per_step": int(12 * m)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes / (kms * 1e-3) / 1e9 / peak,
                         "traffic": (prof.get("dram_bytes_per_launch", {}) or {}).get(f"{args.workload}:{m}:64"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes},
        }
        # the controls family steps waves in momentum form under the default conventions (pdeb_controls.cu)
        line["roofline_fp64"] = fp64_view(sub, m, kms, prof, wave_form="momentum")
        print(json.dumps(line), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="heat", choices=list(WORKLOAD_TEXT))
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per step")
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--gather", action="store_true", help="NCCL all_gather of the per-sample results after each step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-block", type=int, default=65536, help="samples per block of the end-to-end pipeline")
    ap.add_argument("--explicit-knots", action="store_true", help="carry per-sample knot times instead of uniform knots")
    ap.add_argument("--no-text", action="store_true", help="skip the text-level end-to-end leg (e2e_text)")
    ap.add_argument("--text-programs", type=int, default=32768, help="programs of the text-level leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the device-timed rates of the other BASELINE configs")
    ap.add_argument("--e2e-planar-only", action="store_true", help="skip the two packed end-to-end legs (long-running workloads)")
    ap.add_argument("--no-numa-bind", action="store_true", help="leave the placement of this rank's host thread and page-locked buffers to the OS")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = DEFAULT_BATCH[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.gpus != world and rank == 0:
        sys.stderr.write(f"bench: --gpus {args.gpus} but WORLD_SIZE={world}: N>1 runs are launched with torch.distributed.run "
                         f"(one rank per GPU); this process reports n_gpus={world}\n")
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pde_controller_b200 import native, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    if args.workload.startswith("controls_"):
        run_controls(args, rank, local_rank, world)
        return
    torch.cuda.set_device(local_rank)
    dev = local_rank
    # one process per GPU: run on the CPUs of the NUMA node the GPU hangs off BEFORE any page-locked buffer is allocated,
    # so that this rank's host -> device copies stay off the socket interconnect (native.bind_host_near); the text-level and CPU legs
    # below get the original CPU set back
    all_cpus = os.sched_getaffinity(0)
    numa = {"bound": False, "note": "--no-numa-bind"} if args.no_numa_bind else native.bind_host_near(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream().cuda_stream

    # this rank's shard of the global synthetic set (weak scaling: args.batch samples per GPU)
    total = args.batch * world
    # compact form: the synthetic controls sit on uniform knots (C11), so knot times are formed on the device
    # (knot_t == NULL in the C ABI) instead of crossing PCIe; --explicit-knots sends them as per-sample arrays
    uk = not args.explicit_knots
    # mixed: every rank takes its 1/N slice of each kind (contiguous ranges of the kind-sorted set would leave the
    # last ranks with waves only, three times the work of the first ones)
    src = synth.make_batch(args.batch, args.workload, seed=args.seed, start=rank * args.batch, total=total, uniform_knots=uk,
                           kind_slice=(rank, world) if args.workload == "mixed" and world > 1 else None)
    host = native.pinned_batch(src.n, src.kmax, src.pmax, src.tmax, uniform_knots=uk)
    host.slab[:] = src.slab
    res = native.ResidentBatch(host, device=dev, stream=stream)
    rob_dev, _, st_dev = res.device_ptrs()
    pin_rob, pin_st = native.PinnedBuffer(8 * host.n), native.PinnedBuffer(4 * host.n)      # page-locked result buffers
    out_host = {"robustness": pin_rob.array.view(np.float64), "status": pin_st.array.view(np.int32)}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    class _Dev:  # zero-copy view of the library's result buffer for the optional NCCL gather
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}
    rob_t = torch.as_tensor(_Dev(rob_dev, host.n), device="cuda") if args.gather and world > 1 else None
    gathered = torch.empty(total, dtype=torch.float64, device="cuda") if rob_t is not None else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- kernels with inputs resident in HBM ("value") ----
    launches_per_step = None
    for _ in range(args.warmup):
        tm = res.eval(args.precision, stream=stream, timing=True)
        launches_per_step = tm.launches
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sampler = ClockSampler(nvml_index(local_rank))
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                      # evict the batch from L2 (outside the per-step event pair)
        ev0[i].record()
        res.eval(args.precision, stream=stream)
        if rob_t is not None:
            dist.all_gather_into_tensor(gathered, rob_t)
        ev1[i].record()
    barrier()
    wall1 = time.perf_counter()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    value = total * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with HOST buffers ("e2e") ----
    # The caller holds a PdebBatch (include/pdeb200.h) in page-locked host memory.  Every step the whole batch goes
    # host -> H2D -> plan -> kernels -> D2H (robustness + status) as args.e2e_block-sample blocks, E2E_DEPTH generations
    # of buffers on as many streams (native.BlockPipeline); a step ends when the last byte of its results is in host memory.
    #   e2e            PLANAR route: the PdebBatch slab itself crosses PCIe (one copy per block), nothing is prepared
    #                  outside the timed region.  This is the headline: what a caller of pdeb_resident_upload pays.
    #   e2e_packed     packed transfer form WITH its host-side producer (pdeb_pack_batch) inside the timed region
    #   e2e_prepacked  packed transfer form produced OUTSIDE the timed region (a producer that writes PdebPackedBatch
    #                  itself, e.g. a data-set loader): the PCIe-lean route, reported for the multi-GPU curve only
    E2E_DEPTH = 3

    def e2e_route(packed, pack_inside):
        pipe = native.BlockPipeline(src.n, src.kmax, src.pmax, src.tmax, block=args.e2e_block, device=dev, depth=E2E_DEPTH,
                                    uniform_knots=uk, packed=packed)
        for g in range(E2E_DEPTH):
            pipe.fill(src, g)

        def loop(depth):
            """K steps; depth 1 = wait for every step before the next, depth d = steps i+1 .. i+d-1 are enqueued while
            step i runs (their buffers are the other generations): the host->device copy engine never waits for the host."""
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                g = i % depth
                if i >= depth:
                    pipe.wait(g)                      # results of step i-depth are in host memory: they are read here
                pipe.submit(g, args.precision, pack=pack_inside)
            for g in range(depth):
                pipe.wait(g)
            torch.cuda.synchronize()
            dt_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt_, op=dist.ReduceOp.MAX)
            return total * args.steps / float(dt_.item()), float(dt_.item())

        for g in range(E2E_DEPTH):                    # every generation of buffers once (also: results for the check below)
            pipe.run(args.precision, gen=g)
        v1, _ = loop(1)
        vd, secs = loop(E2E_DEPTH)
        ok = all(np.array_equal(pipe.robustness[g], out_host["robustness"]) for g in range(E2E_DEPTH))
        out = {"value": vd, "unit": UNIT, "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(12 * host.n),
               "depth1_value": v1, "h2d_gbs_per_rank": pipe.h2d_bytes * args.steps / secs / 1e9}
        pipe.close()
        return out, ok

    res.eval(args.precision, stream=stream)
    res.fetch(stream=stream, out=out_host)
    e2e, ok_a = e2e_route(packed=False, pack_inside=False)
    e2e["route"] = "planar PdebBatch slab, one H2D copy per block; nothing prepared outside the timed region"
    e2e_packed, e2e_prepacked, ok_b, ok_c = None, None, True, True
    if not args.e2e_planar_only:
        e2e_packed, ok_b = e2e_route(packed=True, pack_inside=True)
        e2e_packed["route"] = "pdeb_pack_batch (host, inside the timed region) -> packed H2D -> device-side unpack"
        e2e_prepacked, ok_c = e2e_route(packed=True, pack_inside=False)
        e2e_prepacked["route"] = "PdebPackedBatch produced outside the timed region -> packed H2D -> device-side unpack"
    if not (ok_a and ok_b and ok_c):
        raise SystemExit("bench: pipelined end-to-end results differ from the resident evaluation")

    st = out_host["status"]
    n_fail = int((st != 0).sum())
    # the ONE-SHOT entry (pdeb_eval_batch: a synchronous call with host arrays, what eval_robustness_batch makes): from
    # ordinary pageable memory and from the page-locked slab, best of 3 -- the cost of not pipelining, next to `e2e`
    one_shot = {}
    if rank == 0:
        for name, hb in (("pageable", src), ("page_locked", host)):
            native.evaluate_batch(hb, device=dev, precision=args.precision)
            best = None
            for _ in range(3):
                t0 = time.perf_counter()
                o1 = native.evaluate_batch(hb, device=dev, precision=args.precision)
                dt_ = time.perf_counter() - t0
                best = dt_ if best is None else min(best, dt_)
            if not np.array_equal(o1["robustness"], out_host["robustness"]):
                raise SystemExit("bench: one-shot results differ from the resident evaluation")
            one_shot[name] = {"value": src.n / best, "unit": UNIT, "ms_per_call": best * 1e3}
    os.sched_setaffinity(0, all_cpus)
    if world > 1:
        nb_ = torch.tensor([1.0 if numa.get("bound") else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(nb_, op=dist.ReduceOp.SUM)
        numa["ranks_bound"] = int(nb_.item())

    # ---- the reference's real call surface: program TEXT in, robustness out (eval_script.py:43-79) ----
    e2e_text = None
    if rank == 0 and not args.no_text:
        from pde_controller_b200 import eval_script as es
        from pde_controller_b200.frontend import render_program
        m = min(args.text_programs, src.n)
        probs, ctrls = zip(*(src.get_problem(i) for i in range(m)))
        texts = [render_program(p) for p in probs]
        es.configure(device=dev, precision=args.precision)
        ctrls = list(ctrls)
        es.eval_robustness_batch(texts[:64], controls=ctrls[:64])                  # imports, regex compilation
        t0 = time.perf_counter()
        rob_text, _ = es.eval_robustness_batch(texts, controls=ctrls)              # first call of this size: page-locked mirror,
        first = time.perf_counter() - t0                                           # device image and host pages are allocated here
        if not np.array_equal(np.array(rob_text), out_host["robustness"][:m]):
            raise SystemExit("bench: text-level results differ from the resident evaluation")
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            es.eval_robustness_batch(texts, controls=ctrls)
        secs = (time.perf_counter() - t0) / reps
        # the same call WITHOUT controls: the default mode of the drop-in helpers, a searched input per program
        # (pdeb_controls_search_batch: 4,096 candidates x 6 generations each, drawn, scored and ranked on the device)
        ms_ = min(2048, m)
        es.eval_robustness_batch(texts[:64], controls="search")
        t0 = time.perf_counter()
        rob_s, _ = es.eval_robustness_batch(texts[:ms_], controls="search")
        secs_s = time.perf_counter() - t0
        search_leg = {"value": ms_ / secs_s, "unit": "programs/s", "programs": ms_, "candidates_per_program": 4096 * 6,
                      "positive": int(sum(r > 0 for r in rob_s)),
                      "positive_under_the_given_controls": int((out_host["robustness"][:ms_] > 0).sum())}
        # and with the controls handed over as arrays (control_arrays): what a caller that holds them in bulk pays
        kmax_ = max(len(c[0]) for c in ctrls)
        T_, Y_, K_ = np.zeros((m, kmax_)), np.zeros((m, kmax_)), np.zeros(m, dtype=np.int64)
        for i_, c_ in enumerate(ctrls):
            K_[i_] = len(c_[0]); T_[i_, : K_[i_]] = c_[0]; Y_[i_, : K_[i_]] = c_[1]
        es.eval_robustness_batch(texts, control_arrays=(T_, Y_, K_))
        t0 = time.perf_counter()
        for _ in range(reps):
            rob_a, _ = es.eval_robustness_batch(texts, control_arrays=(T_, Y_, K_))
        secs_a = (time.perf_counter() - t0) / reps
        if not np.array_equal(np.array(rob_a), out_host["robustness"][:m]):
            raise SystemExit("bench: text-level results (control arrays) differ from the resident evaluation")
        e2e_text = {"value": m / secs, "unit": "programs/s", "programs": m, "cores": os.cpu_count(), "calls_timed": reps,
                    "control_arrays_value": m / secs_a, "search_mode": search_leg,
                    "first_call_value": m / first,
                    "route": "eval_robustness_batch(program strings, given controls): native template matcher "
                             "(pdeb_parse_programs, host threads; anything else -> Python front-end) -> PdebBatch -> "
                             "pdeb_eval_batch -> robustness list"}

    # ---- the other single-GPU configurations of BASELINE.json, device-timed like `value` (configs 3, 5, 4) ----
    configs = {}
    if not args.no_configs and args.workload == "heat":
        for name, wl, nb, steps_c in (("configs[2] wave 65536", "wave", 65536, 5), ("configs[4] mixed 65536 per GPU", "mixed", 65536, 5),
                                      ("configs[3] refined heat 1024x10000, 16384 per GPU", "heat_refined", 16384, 2)):
            bsrc = synth.make_batch(nb, wl, seed=args.seed, start=rank * nb, total=nb * world, uniform_knots=True,
                                    kind_slice=(rank, world) if wl == "mixed" and world > 1 else None)
            r2 = native.ResidentBatch(bsrc, device=dev, stream=stream)
            r2.eval(64, stream=stream, timing=True)
            ms = 0.0
            for _ in range(steps_c):
                flush.zero_()
                ms += r2.eval(64, stream=stream, timing=True).kernel_ms
            r2.close()
            t2 = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            configs[name] = {"value": nb * world * steps_c / (float(t2.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t2.item()) / steps_c,
                             "batch_per_gpu": nb, "steps": steps_c, "dtype": "f64",
                             "timing": "CUDA events inside pdeb_resident_eval, inputs resident, L2 flushed, max over ranks"}

    if not args.no_configs and args.workload == "heat" and args.precision == 64:
        # north star: "an FP32 mode reported separately at its own tolerance" (tests: 2e-3 heat / 5e-2 wave by the parity metric)
        res.eval(32, stream=stream, timing=True)
        ms = 0.0
        for _ in range(5):
            flush.zero_()
            ms += res.eval(32, stream=stream, timing=True).kernel_ms
        t2 = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        configs["heat 65536, FP32 time loop (reported separately)"] = {
            "value": total * 5 / (float(t2.item()) * 1e-3), "unit": UNIT, "ms_per_step": float(t2.item()) / 5, "batch_per_gpu": args.batch,
            "steps": 5, "dtype": "f32", "tolerance": "2e-3 of the parity metric (tests/test_gpu_parity.py::test_fp32_mode_reported_at_its_own_tolerance)"}
        res.eval(args.precision, stream=stream)             # leave the FP64 results in the image

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = host.algorithmic_bytes()            # 105 + 16K + 50P + 3(2P-1) per evaluation
        kernel_ms = total_ms / args.steps
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        prof = measured_profile()
        traffic = (prof.get("dram_bytes_per_launch", {}) or {}).get(f"{args.workload}:{args.batch}:{args.precision}")
        fp64 = fp64_view(host, 1, kernel_ms, prof)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
            "config": common_config(args, world),
            "detail": {"gather": bool(rob_t is not None), "failed_samples": n_fail,
                       "knots": "uniform (formed on device)" if uk else "explicit", "e2e_block": args.e2e_block,
                       "e2e_pipeline_depth": E2E_DEPTH, "wall_s_timed_region": wall1 - wall0, "numa": numa,
                       "one_shot_call": one_shot},
            "e2e": e2e,
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "fused kernel is FP64-issue bound by design (state never leaves registers); see DESIGN.md"},
            "roofline_fp64": fp64,
        }
        if e2e_packed is not None:
            line["e2e_packed"], line["e2e_prepacked"] = e2e_packed, e2e_prepacked
        if e2e_text is not None:
            line["e2e_text"] = e2e_text
        if configs:
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, src, out_host["robustness"].copy(), out_host["status"].copy())
        print(json.dumps(line), flush=True)
    res.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
