timeout 600 python -m pytest tests/test_amr.py -x -q -m gpu > gpurun_out/pytest_amr.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_amr.log
tail -15 gpurun_out/pytest_amr.log
