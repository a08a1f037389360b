#!/bin/bash
# 1 -> N scaling records of one workload on one box: tools/scale_run.sh <workload> <tag> "<N list>" [extra bench args]
# (each N is launched exactly as the driver launches it; 240 s per leg: on an 8-GPU box a hung leg costs 8x its wall time)
w=$1; tag=$2; ns=$3; shift 3
mkdir -p gpurun_out
port=29600
for n in $ns; do
  port=$((port+1))
  out=gpurun_out/${tag}_${w}_n${n}
  if [ "$n" = "1" ]; then
    timeout 240 python bench.py --gpus 1 --workload $w --steps 20 --warmup 5 --no-cpu-baseline "$@" 2> $out.err | grep "^{" > $out.json
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $w --steps 20 --warmup 5 --no-cpu-baseline "$@" 2> $out.err | grep "^{" > $out.json
  fi
  python - $out.json <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = l.get("e2e") or {}
    print("%s n=%d %s: %.4g %s  %.3f ms/step  e2e %s" % (sys.argv[1], l["n_gpus"], l["scaling"], l["value"], l["unit"], l["ms_per_step"],
          ("%.4g (%.3f ms)" % (e["value"], e.get("ms_per_step", 0))) if e else None))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
