"""Summarise an .ncu-rep (read here, no GPU needed): key metrics, stall reasons, opcode mix.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25]
"""
import collections
import csv
import io
import subprocess
import sys


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 22
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "sass__inst_executed_shared_loads", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    for k in keys:
        if k in m:
            print(f"{k:78s} {m[k][1]} {m[k][0]}")
    print("-- stall reasons (warps per issue-active cycle)")
    st = [(h, float(v)) for h, (u, v) in m.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for h, v in sorted(st, key=lambda x: -x[1])[:10]:
        print(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:.3f}")
    rows = ncu_csv(rep, "source")
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, samp = collections.Counter(), collections.Counter()
    n_all = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        sass = r[ix["Source"]].strip().split()
        op = (sass[1] if sass[0].startswith("@") else sass[0]).split(".")[0]
        n = int(r[ix["Instructions Executed"]])
        tot[op] += n
        samp[op] += int(r[ix["# Samples"]])
        n_all += n
    print(f"-- opcode mix ({n_all} warp instructions)")
    for op, n in tot.most_common(top):
        print(f"   {op:10s} {n:12d} {100.0 * n / n_all:5.1f}%   samples {samp[op]}")


if __name__ == "__main__":
    main()
