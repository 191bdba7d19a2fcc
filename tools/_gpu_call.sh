set -x
python -m pytest tests/test_boundary.py tests/test_integral.py -x -q -m gpu > gpurun_out/pytest_n2n3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n2n3.log
python tools/time_quad_bc.py > gpurun_out/time_quad_bc.log 2>&1
tail -5 gpurun_out/pytest_n2n3.log; cat gpurun_out/time_quad_bc.log
