mkdir -p gpurun_out
python -m pytest tests/test_gpu_wide_shard.py -x -q -m gpu 2>&1 | grep -E "Error|assert|err" | head -20 > gpurun_out/r02_pytest_n2b.txt
PFS_WIDE_PREC=none python -m pytest tests/test_gpu_wide_shard.py -x -q -m gpu 2>&1 | grep -E "Error|assert|err|passed|failed" | head -20 >> gpurun_out/r02_pytest_n2b.txt
