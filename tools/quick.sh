#!/bin/bash
# quick GPU check: parity tests + kernel table of the bench (used during optimisation)
python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=line -x 2>&1 | cut -c1-300 | grep -vE "^ +\+" | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("value %.4g edges/s  ms/step %.3f  e2e %s  fma %.3f launches %d" % (l["value"], l["ms_per_step"], l["e2e"] and "%.4g" % l["e2e"]["value"], l["fma"]["frac"], l["gpu_launches"]))
for k,v in l["kernels"].items(): print("  %-24s %3d  %.3f ms  %.1f%%" % (k, v["launches"], v["ms_per_step"], 100*v["share"]))
PY
tail -3 gpurun_out/bench_q.err
