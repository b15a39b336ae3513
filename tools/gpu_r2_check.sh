# round-2 GPU check: tests, default bench (both arms).  usage: bash tools/gpu_r2_check.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
tail -15 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_heat.json 2> gpurun_out/${tag}_bench_heat.err
tail -3 gpurun_out/${tag}_bench_heat.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_heat.json", "gpurun_out/${tag}_bench_reference.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"],
              {k: "%.4g" % v["value"] for k, v in d.items() if k.startswith("e2e_") and isinstance(v, dict)},
              {k: "%.4g" % v["value"] for k, v in d.get("configs", {}).items()})
    except Exception as e:
        print(f, "unreadable:", e)
PY
