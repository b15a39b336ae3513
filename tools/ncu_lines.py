"""Per-source-line attribution of an .ncu-rep: instructions, samples, and selected opcodes by line.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep [--ops IMAD,FSEL] [--top 30]
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    ops = sys.argv[sys.argv.index("--ops") + 1].split(",") if "--ops" in sys.argv else []
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, cur, fname = None, None, ""
    per = collections.defaultdict(collections.Counter)
    line_inst, line_samp, src = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r or r[0] == "Function Name":
            continue
        i_inst, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
        if r[0].strip().isdigit():
            cur = (fname, int(r[0]))
            src[cur] = r[1].strip()[:100]
            try:
                line_inst[cur] += int(r[i_inst])
                line_samp[cur] += int(r[i_s])
            except ValueError:
                pass
        else:
            sass = r[3].strip().split()
            if not sass:
                continue
            op = (sass[1] if sass[0].startswith("@") else sass[0]).split(".")[0]
            try:
                per[op][cur] += int(r[i_inst])
            except ValueError:
                pass
    tot = sum(line_inst.values())
    tots = sum(line_samp.values())
    print(f"total warp instructions {tot}, samples {tots}")
    for k, n in line_inst.most_common(top):
        print(f"{k[0]}:{k[1]:4d} inst {100 * n / tot:5.1f}%  samp {100 * line_samp[k] / max(tots, 1):5.1f}% | {src[k]}")
    for op in ops:
        t = sum(per[op].values())
        print(f"-- {op}: {t}")
        for k, n in per[op].most_common(8):
            print(f"     {k[0]}:{k[1]:4d} {n:11d} {100 * n / max(t, 1):5.1f}% | {src.get(k, '')}")


if __name__ == "__main__":
    main()
