timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 200 python tools/check_fd_stream.py 200 9 2>&1 | tail -2
python tools/run_fd.py 512 512 1024 2 2 3 > gpurun_out/plain_fdrow.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'fd_row|fd_march' -s 2 -c 1 -f -o gpurun_out/prof_fd_row python tools/run_fd.py 512 512 1024 2 2 3 > gpurun_out/ncu_fdrow.log 2>&1
echo done
