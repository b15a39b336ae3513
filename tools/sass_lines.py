"""Static SASS instruction counts per source line of one kernel of the built library (no GPU needed).

  python tools/sass_lines.py 'pdeb_reg_kernelILi0ELi30Ed' [--so pde_controller_b200/libpdeb200.so] [--lines 200-260,355-380] [--ops]

Uses `nvdisasm -g` on the embedded cubins (the library is built with -lineinfo).  In the fully unrolled time
loop a static count is the per-step dynamic count, which is what the kernels are tuned by.
"""
import collections
import glob
import os
import re
import subprocess
import sys
import tempfile


def main():
    pat = sys.argv[1]
    so = sys.argv[sys.argv.index("--so") + 1] if "--so" in sys.argv else os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pde_controller_b200", "libpdeb200.so")
    ranges = []
    if "--lines" in sys.argv:
        for part in sys.argv[sys.argv.index("--lines") + 1].split(","):
            a, _, b = part.partition("-")
            ranges.append((int(a), int(b or a)))
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
        text = ""
        for cub in sorted(glob.glob(os.path.join(td, "*.cubin"))):
            out = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
            if pat in out:
                text = out
                break
    if not text:
        sys.exit(f"no kernel matching {pat!r}")
    per_line = collections.Counter()
    per_line_ops = collections.defaultdict(collections.Counter)
    ops = collections.Counter()
    inside, cur, in_group = False, ("?", 0), False
    for ln in text.splitlines():
        if ln.startswith(".text."):
            inside = pat in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            # a group of consecutive tags lists the innermost location first, then its "inlined at" callers
            if not in_group:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            in_group = True
            continue
        in_group = False
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            op = m.group(1)
            per_line[cur] += 1
            per_line_ops[cur][op] += 1
            ops[op] += 1
    total = sum(per_line.values())
    print(f"{total} static instructions")
    if "--ops" in sys.argv:
        print("  ".join(f"{o}:{c}" for o, c in ops.most_common(16)))
    items = per_line.items()
    if ranges:
        items = [(k, v) for k, v in items if any(a <= k[1] <= b for a, b in ranges)]
        print("selected lines:", sum(v for _, v in items))
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    for (f, l), c in sorted(items, key=lambda kv: -kv[1])[:top]:
        print(f"  {f}:{l:4d}  {c:5d}   " + " ".join(f"{o}:{n}" for o, n in per_line_ops[(f, l)].most_common(6)))


if __name__ == "__main__":
    main()
