timeout 600 python -m pytest tests/test_interp.py tests/test_boundary.py tests/test_integral.py -x -q -m gpu > gpurun_out/pytest_n4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n4.log
tail -15 gpurun_out/pytest_n4.log
