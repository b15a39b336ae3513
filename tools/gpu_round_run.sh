set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/h6_pytest.log
for w in heat wave mixed heat_refined controls_heat controls_wave; do
  timeout 240 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/h6_bench_$w.json 2> gpurun_out/h6_bench_$w.err
done
timeout 240 python bench.py --precision 32 --workload heat --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/h6_bench_heat_fp32.json 2>/dev/null
timeout 240 python bench.py --precision 32 --workload wave --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/h6_bench_wave_fp32.json 2>/dev/null
timeout 240 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h6_bench_reference.json 2>/dev/null
[ -n "$SKIP_NCU" ] || timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h6_launches_heat.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/h6_ncu_launches.log 2>&1
for w in heat wave heat_refined; do
  [ -n "$SKIP_NCU" ] || timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pdeb_(reg|chunk)_kernel' --launch-skip 3 --launch-count 1 -o gpurun_out/h6_$w python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/h6_ncu_$w.log 2>&1
done
for w in controls_heat controls_wave; do
  [ -n "$SKIP_NCU" ] || timeout 300 ncu --set full --clock-control none --import-source on -k regex:pdeb_controls_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/h6_$w python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/h6_ncu_$w.log 2>&1
done
cat gpurun_out/h6_pytest.log
