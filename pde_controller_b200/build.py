"""Build the CUDA library in-tree: pde_controller_b200/libpdeb200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libpdeb200.so")
SOURCES = ["pdeb_api.cu", "pdeb_kernels.cu", "pdeb_chunk_heat.cu", "pdeb_chunk_wave.cu", "pdeb_cta_heat.cu", "pdeb_controls.cu", "pdeb_iou.cu", "pdeb_frontend.cpp"]
HEADERS = ["pdeb_core.h", "pdeb_launch.h", "pdeb_chunk.cuh", "pdeb_cta.cuh", os.path.join("..", "..", "include", "pdeb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: build an experimental variant (tools/variant_bench.py) next to the default library."""
    global_so = SO if out is None else out
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and os.path.exists(global_so) and os.path.getmtime(global_so) >= _newest(deps):
        return global_so
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    for s in srcs:
        # a variant's objects live next to the variant (one directory to delete), the default library's next to the sources
        o = (os.path.splitext(s)[0] + ".o" if global_so == SO
             else os.path.join(os.path.dirname(global_so), os.path.basename(os.path.splitext(s)[0]) + "." + os.path.basename(global_so) + ".o"))
        objs.append(o)
        # NVCC_EXTRA: extra compiler flags of an experimental variant (tools/variant_bench.py), e.g. "-Xptxas --opt-level=4"
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("NVCC_EXTRA", "").split() + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", global_so] + objs + ["-lcudart"]
    subprocess.run(link, check=True)
    if global_so != SO:
        for o in objs:              # a variant's objects are not reused: 20 MB each would swamp the snapshot sent to the GPU box
            os.remove(o)
    return global_so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
