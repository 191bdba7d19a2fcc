timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/check_fd_stream.py 100 11 2>&1 | tail -1
timeout 200 python tools/check_fft_strided.py 60 1 2>&1 | tail -1
timeout 200 python tools/check_dense_general.py 100 1 2>&1 | tail -1
