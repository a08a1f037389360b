mkdir -p gpurun_out
timeout 120 python bench.py --steps 20 --warmup 3 --graphs 32 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r02_bench_g32.json 2> gpurun_out/r02_bench_g32.err
timeout 120 python bench.py --steps 20 --warmup 3 --graphs 64 --no-e2e --no-cpu-baseline --no-extras --no-profile > gpurun_out/r02_bench_g64.json 2> gpurun_out/r02_bench_g64.err
timeout 300 python -m pytest tests -x -q -m gpu -k "wide" -s 2>&1 | grep -E "passed|failed|worst|top |Error|c4-shape" | cut -c1-260 > gpurun_out/r02_pytest_wide3.txt
