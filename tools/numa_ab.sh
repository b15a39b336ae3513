#!/bin/bash
# 8-GPU box: topology as the driver and the kernel see it, then the end-to-end legs with and without NUMA binding of the ranks.
# usage (on the GPU box): tools/numa_ab.sh <tag> [ngpu]
tag=${1:-numa}; n=${2:-8}
out=gpurun_out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)" > $out/${tag}_lscpu.txt
for d in /sys/bus/pci/devices/*; do
  if [ -f $d/vendor ] && [ "$(cat $d/vendor)" = "0x10de" ]; then echo "$(basename $d) numa_node=$(cat $d/numa_node) class=$(cat $d/class)"; fi
done > $out/${tag}_sysfs.txt 2>&1
for mode in bind nobind; do
  flag=""; [ $mode = nobind ] && flag="--no-numa-bind"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n \
    --steps 20 --warmup 3 --no-text --no-configs --no-cpu-baseline $flag > $out/${tag}_${mode}.json 2> $out/${tag}_${mode}.err
  python - <<PY
import json
try:
    l = json.loads(open("$out/${tag}_${mode}.json").read().strip().splitlines()[-1])
    print("$mode", "value %.4g" % l["value"], "e2e %.4g" % l["e2e"]["value"], "GB/s/rank %.1f" % l["e2e"]["h2d_gbs_per_rank"],
          "prepacked %.4g" % l["e2e_prepacked"]["value"], "packed %.4g" % l["e2e_packed"]["value"], l["detail"]["numa"])
except Exception as e:
    print("$mode failed", e)
PY
done
