mkdir -p gpurun_out
PFS_EDGE_TC=1 timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2 | cut -c1-300 > gpurun_out/r02_pytest_tc4.txt
PFS_EDGE_TC=1 timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --no-graph > gpurun_out/r02_bench_tc4.json 2> gpurun_out/r02_bench_tc4.err
