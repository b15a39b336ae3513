# Round evidence on one B200: tests, every workload's bench line, launch list, ncu captures.  usage: bash tools/gpu_round_run.sh <tag>
tag=${1:-r02}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_heat.json 2> gpurun_out/${tag}_bench_heat.err
for w in wave mixed heat_refined controls_heat controls_wave; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-text --no-configs > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
done
timeout 240 python bench.py --precision 32 --workload heat --steps 20 --warmup 3 --no-cpu-baseline --no-text --no-configs > gpurun_out/${tag}_bench_heat_fp32.json 2>/dev/null
timeout 240 python bench.py --precision 32 --workload wave --steps 20 --warmup 3 --no-cpu-baseline --no-text --no-configs > gpurun_out/${tag}_bench_wave_fp32.json 2>/dev/null
timeout 240 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>/dev/null
[ -n "$SKIP_NCU" ] || timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_heat.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-text --no-configs > gpurun_out/${tag}_ncu_launches.log 2>&1
for w in heat wave heat_refined; do
  [ -n "$SKIP_NCU" ] || timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pdeb_(reg|chunk)_kernel' --launch-skip 3 --launch-count 1 -o gpurun_out/${tag}_$w python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --no-text --no-configs > gpurun_out/${tag}_ncu_$w.log 2>&1
done
cat gpurun_out/${tag}_pytest.log
