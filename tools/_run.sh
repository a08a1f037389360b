mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_step.py tests/test_gpu_loss.py tests/test_gpu_dp.py -x -q -m gpu 2>&1 | tail -4 | cut -c1-300 > gpurun_out/r02_pytest_m1.txt
PFS_NODE_MMA=0 timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "module_parity or full_size" 2>&1 | tail -2 | cut -c1-300 >> gpurun_out/r02_pytest_m1.txt
timeout 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_m1.json 2> gpurun_out/r02_bench_m1.err
