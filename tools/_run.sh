mkdir -p gpurun_out
python -m pytest tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest3.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_a.json 2> gpurun_out/r02_bench_c3_a.err
