timeout 300 python tools/check_fd_stream.py 400 10 2>&1 | tail -3
python tools/time_fd_fused.py 512 > gpurun_out/time_fd_fused2.log 2>&1
cat gpurun_out/time_fd_fused2.log
