#!/bin/bash
# wide path: primitive tests + parity + C4 bench kernel table
timeout 600 python -m pytest tests/test_gpu_wide_ops.py tests/test_gpu_wide_parity.py -q -m gpu --tb=short -x 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --workload ${WL:-c4} --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err
tail -3 gpurun_out/bench_w.err
python - <<'PY'
import json
l=json.loads(open("gpurun_out/bench_w.json").read().strip().splitlines()[-1])
print("value %.4g edges/s ms/step %.2f tensor frac %.3f (gemm-only %.3f) e2e %s launches %d" % (l["value"], l["ms_per_step"], l["tensor"]["frac"], l["roofline"]["frac"], l["e2e"] and "%.4g"%l["e2e"]["value"], l["gpu_launches"]))
for k,v in l["kernels"].items(): print("  %-24s %3d  %.3f ms  %.1f%%" % (k, v["launches"], v["ms_per_step"], 100*v["share"]))
PY
