mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train_step.py -x -q -m gpu -s 2>&1 | grep -E "trajectory vs|passed|failed|^E|loss curve|ref " | cut -c1-250 > gpurun_out/r02_pytest_n2.txt
timeout 120 python bench.py --workload train --steps 50 --warmup 5 > gpurun_out/r02_bench_train.json 2> gpurun_out/r02_bench_train.err
