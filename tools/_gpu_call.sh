timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "full_size" > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -5 gpurun_out/pytest_full.log
